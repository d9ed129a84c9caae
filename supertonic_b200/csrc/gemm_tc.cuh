// tcgen05 / TMEM / TMA GEMM for the dense contractions (pointwise projections, attention projections)
// of the Supertonic forward pass — hand-written for sm_100a.
//
//   out[M,N] = epilogue( A[M,K] · W[N,K]^T ),   fp32 accumulate in TMEM.
//
// Arithmetic: "bf16x3" — each fp32 operand v is carried as a split pair (hi = bf16(v), lo = bf16(v-hi))
// and the product is formed as A_hi·W_hi + A_lo·W_hi + A_hi·W_lo with three kind::f16 MMAs per K-slice
// (bf16 x bf16 products are exact in fp32; the dropped lo·lo term is 2^-16 relative). Measured effect on the
// full surrogate at 20 Euler steps: max-abs latent error 1e-5 vs 9e-4 for single-pass TF32 and 7e-3 for plain
// bf16 (DESIGN.md "precision"); the north-star bound is 1e-3. The vocoder's GEMMs (bound: waveform SNR >= 40 dB) instead run the kF16
// instantiation: one fp16 value per operand element, one MMA per K-slice, half the operand bytes (~68 dB against the oracle).
//
// Structure (persistent over 128 x BN output tiles, 320 threads, one CTA per SM):
//   warp 0   : TMA producer   — cp.async.bulk.tensor.2d, 128B-swizzled K-major tiles, STAGES-deep mbarrier ring
//   warp 1   : MMA issuer     — one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16);
//                               tcgen05.commit releases smem stages and finally signals the accumulator barrier
//   warps 2-9: epilogue       — tcgen05.ld 32x32b.x16 (lane = row) -> per-warp smem staging -> re-read with
//                               lanes along the row, so that residual reads and output stores are coalesced;
//                               fused bias / GELU(erf) / layer-scale / residual / mask / Euler update;
//                               fp32 or split-bf16 stores. Two TMEM accumulator buffers: the epilogue of tile i
//                               overlaps the MMAs of tile i+1.
//
// Operand traffic: the kernel is bound by L2 -> shared-memory operand bytes, not by the tensor pipe (measured: profiles/r1b_*,
// r1c_gemm_sweep.txt): a 128 x BN tile re-reads its A rows for every N tile and its W rows for every M tile. Thread-block clusters
// with TMA multicast of the shared A / W slices were built and measured in round 1 (2x1 / 1x2 / 2x2: 5-100 % slower on every hot
// shape — multicast halves L2 reads but every SM still ingests its full tiles, and the cluster runs in lock-step) and removed in
// round 2; the two-SM form that does help the big shapes is gemm2_tc.cuh.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace stc {
namespace tc {

constexpr int BM = 128;        // UMMA_M (cta_group::1)
constexpr int BK = 64;         // bf16 elements per 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;   // two warps per TMEM lane quarter, each takes half of the tile's columns
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int EPI_CHUNK = 16;                  // accumulator columns moved TMEM -> smem -> global per step
constexpr int EPI_PITCH = EPI_CHUNK + 4;       // floats per staged row: 16-byte aligned, conflict-free for v4 accesses
constexpr int EPI_BYTES = EPI_WARPS * 32 * EPI_PITCH * 4;
constexpr int EPI_BS = 256;                    // floats per warp: bias [0,128) and layer-scale [128,256) of the warp's columns
constexpr int EPI_BS_BYTES = EPI_WARPS * EPI_BS * 4;

template <int BN> struct Tile {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int W_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
    static constexpr int STAGES = (196608 / STAGE_BYTES) > 6 ? 6 : (196608 / STAGE_BYTES);
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;                       // mbarriers + TMEM slot
    static constexpr int EPI_OFF = BAR_OFF + 256;                              // epilogue staging
    static constexpr int BS_OFF = EPI_OFF + EPI_BYTES;                          // per-warp bias / layer-scale slices
    static constexpr int SMEM_BYTES = BS_OFF + EPI_BS_BYTES + 1024 /*align slack*/;
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
    static constexpr int TMEM_COLS = 2 * BN;        // two accumulator buffers (BN in {64,128,256} -> power of two)
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
STC_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

STC_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
STC_DEVINL void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
STC_DEVINL void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
STC_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
STC_DEVINL void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
STC_DEVINL void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
STC_DEVINL bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
STC_DEVINL void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
STC_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
STC_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
STC_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
STC_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
STC_DEVINL void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
STC_DEVINL void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor: K-major tile, 128-byte swizzle, 8-row atoms 1024 B apart
// (bit layout: cute/arch/mma_sm100_desc.hpp UMMA::SmemDescriptor — start>>4 [0,14), LBO>>4 [16,30),
//  SBO>>4 [32,46), version=1 [46,48), layout_type [61,64) with SWIZZLE_128B = 2).
STC_DEVINL uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)0 << 16;                     // LBO: unused for swizzled K-major
    d |= (uint64_t)(1024 >> 4) << 32;           // SBO
    d |= (uint64_t)1 << 46;                     // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                     // SWIZZLE_128B
    return d;
}
// Instruction descriptor (UMMA::InstrDescriptor): c=F32 [4,6)=1, a=BF16 [7,10)=1, b=BF16 [10,13)=1,
// a/b K-major (bits 15,16 = 0), N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {           // a = b = F16 (format 0)
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Params {
    int M, N, K;
    Epilogue ep;
    float* out_f32;                 // [M, ldo] when !split
    __nv_bfloat16* out_hi;          // [M, ldo] when split
    __nv_bfloat16* out_lo;
    int ldo;
    int split;
    int cm, cn;                     // (two-SM form: cm = 2; unused by the one-SM kernel)
    long long* trace;               // debug (stc_debug_gemm, STC_GEMM_TRACE=1): clock64() stamps of block 0's producer / MMA warps
};

// erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7) with MUFU reciprocal / exp2 against ~40 instructions for erff() + IEEE
// division. With z = |x|/sqrt2 and E = erf(z) in [0, 1], the graphs' (x * (erf(x/sqrt2) + 1)) * 0.5 is h + |h| * E, h = x/2 (erf is
// odd): 14 instructions — the constants carry the 1/sqrt2 and log2(e) factors, the sign needs no copysign.
STC_DEVINL float gelu_erf_mufu(float x) {
    float t; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678f, fabsf(x), 1.0f)));
    float poly = fmaf(t, 1.061405429f, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    poly *= t;
    float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((x * x) * (-0.5f * 1.4426950408889634f)));
    const float erf_abs = fmaf(-poly, e, 1.0f);
    const float h = 0.5f * x;
    return fmaf(fabsf(h), erf_abs, h);
}

// The same function on two values with the packed fp32 FMAs of sm_100 (FFMA2 / FMUL2: two independent IEEE operations per
// instruction, so each lane is bit-identical to gelu_erf_mufu) — the GELU epilogues are issue bound (the vocoder's pw1: 68 us against
// 42 us of MMAs), and this form issues ~9.5 instead of ~14 instructions per element; the two MUFUs per element stay scalar. The
// polynomial runs with negated coefficients (fma(-a, b, -c) = -fma(a, b, c) exactly), which saves the negation before the last FMA.
STC_DEVINL float2 gelu_erf_mufu2(float2 x) {
    const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
    const float2 d = __ffma2_rn(make_float2(0.3275911f * 0.70710678f, 0.3275911f * 0.70710678f), ax, make_float2(1.0f, 1.0f));
    float2 t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(d.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(d.y));
    float2 np = __ffma2_rn(t, make_float2(-1.061405429f, -1.061405429f), make_float2(1.453152027f, 1.453152027f));
    np = __ffma2_rn(np, t, make_float2(-1.421413741f, -1.421413741f));
    np = __ffma2_rn(np, t, make_float2(0.284496736f, 0.284496736f));
    np = __ffma2_rn(np, t, make_float2(-0.254829592f, -0.254829592f));
    np = __fmul2_rn(np, t);                                                     // = -poly
    const float kk = -0.5f * 1.4426950408889634f;
    const float2 a2 = __fmul2_rn(__fmul2_rn(x, x), make_float2(kk, kk));
    float2 e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(a2.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(a2.y));
    const float2 erf_abs = __ffma2_rn(np, e, make_float2(1.0f, 1.0f));
    const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
    return __ffma2_rn(make_float2(fabsf(h.x), fabsf(h.y)), erf_abs, h);
}

// GELU(erf) with ONE special-function instruction per element, for the fp16-output epilogue of the vocoder's pw1 (which is bound by
// the MUFU pipe with the two-MUFU form of gemm_tc.cuh: ncu xu 51 % of peak, `mio` throttle on every MUFU): erf by Abramowitz-Stegun
// 7.1.28, erf(z) = 1 - (1 + a1 z + ... + a6 z^6)^-16 (|err| <= 3e-7), the 1/sqrt2 powers folded into the coefficients, the 16th power
// as four packed squarings, the reciprocal on the MUFU. With h = x/2: (x (erf(x/sqrt2) + 1)) / 2 = (h + |h|) - |h| r  (h + |h| is
// exact: x or 0). Max |error| against the exact function 7e-7 in fp32 (4.7e-7 for the two-MUFU form; tests/test_gelu_forms.py) — below
// half an fp16 ulp of the output from |y| = 2e-3 up, the same few 1e-7 in absolute terms in the negative tail; p^16 overflows to +inf for |x| > ~40 and the reciprocal returns 0 there, which is the right limit.
STC_DEVINL float2 gelu_erf_rcp2(float2 x) {
    const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
    constexpr float c1 = 0.0705230784f * 0.70710678f, c2 = 0.0422820123f * 0.5f, c3 = 0.0092705272f * 0.35355339f,
                    c4 = 0.0001520143f * 0.25f, c5 = 0.0002765672f * 0.17677670f, c6 = 0.0000430638f * 0.125f;
    float2 q = __ffma2_rn(ax, make_float2(c6, c6), make_float2(c5, c5));
    q = __ffma2_rn(q, ax, make_float2(c4, c4));
    q = __ffma2_rn(q, ax, make_float2(c3, c3));
    q = __ffma2_rn(q, ax, make_float2(c2, c2));
    q = __ffma2_rn(q, ax, make_float2(c1, c1));
    q = __ffma2_rn(q, ax, make_float2(1.0f, 1.0f));
    q = __fmul2_rn(q, q); q = __fmul2_rn(q, q); q = __fmul2_rn(q, q); q = __fmul2_rn(q, q);
    float2 r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(q.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(q.y));
    const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
    const float2 ah = __fmul2_rn(ax, make_float2(0.5f, 0.5f));
    return __ffma2_rn(make_float2(-ah.x, -ah.y), r, __fadd2_rn(h, ah));
}

// v (fp32 x2) -> packed bf16x2 hi and lo with v ~= hi + lo
STC_DEVINL void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));            // upper half <- first source
    const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}

STC_DEVINL uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
STC_DEVINL uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
STC_DEVINL uint32_t cluster_count_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
STC_DEVINL void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
STC_DEVINL void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// kRope: the rotary-embedding epilogue (Q / K projections) lives in its own instantiation — its sincosf slow path costs
// registers, a stack frame and unrolling in every epilogue it is compiled into (the vocoder GEMMs lost 10 % to it).
// kF16: single-pass fp16 arithmetic for the vocoder (the default there, DESIGN.md "precision"): map_a_hi / map_w_hi describe fp16
// [rows, K] operands; a stage holds 128 K-elements — K sub-block 0 in the slots of the bf16 hi halves, sub-block 1 in the slots of the
// lo halves — and issues 8 MMAs per 128 K-elements instead of 24 (a sub-block that lies wholly beyond K is neither loaded nor issued).
template <int BN, bool kRope = false, bool kF16 = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                   const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                   const Params p) {
    using T = Tile<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // SWIZZLE_128B needs 1024-B alignment
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));           // (same offset in every CTA of the cluster)
    const uint32_t bar_base = smem_base + T::BAR_OFF;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (T::STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * T::STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * T::STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * T::STAGES + 4);
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + T::BAR_OFF + 8 * (2 * T::STAGES + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int KSTAGE = kF16 ? 2 * BK : BK;                                  // K-elements per stage
    const int num_kb = (p.K + KSTAGE - 1) / KSTAGE;
    const int n_tiles = (p.N + BN - 1) / BN;
    const int m_tiles = (p.M + BM - 1) / BM;
    const int num_ct = m_tiles * n_tiles;          // tile ct: m tile ct / n_tiles, n tile ct % n_tiles (n fastest: concurrent CTAs share A rows in L2)
    const int ct0 = (int)blockIdx.x, ct_step = (int)gridDim.x;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo);
        tma_prefetch_desc(&map_w_hi); tma_prefetch_desc(&map_w_lo);
        for (int s = 0; s < T::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, T::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        // ===== TMA producer =====
        // Under programmatic dependent launch (kernels.cuh) the W halves of the first STAGES stages are requested BEFORE the dependency
        // wait — weights do not depend on the predecessor — so only the A halves are outstanding when it returns.
        if (elect_one()) {
            auto load = [&](int ct, int kb, uint32_t kbc, bool first_pass, bool w_part, bool a_part) {
                const int m0 = (ct / n_tiles) * BM, n0 = (ct % n_tiles) * BN;
                const int s = kbc % T::STAGES;
                const uint32_t st = smem_base + s * T::STAGE_BYTES;
                if (!first_pass) mbar_wait(empty_bar(s), ((kbc / T::STAGES) & 1) ^ 1);
                if constexpr (kF16) {
                    const int k0 = kb * KSTAGE;
                    const bool two = k0 + BK < p.K;
                    if (w_part) {
                        mbar_expect_tx(full_bar(s), two ? T::STAGE_BYTES : T::STAGE_BYTES / 2);
                        tma_load_2d(st + 2 * T::A_BYTES, &map_w_hi, full_bar(s), k0, n0);
                        if (two) tma_load_2d(st + 2 * T::A_BYTES + T::W_BYTES, &map_w_hi, full_bar(s), k0 + BK, n0);
                    }
                    if (a_part) {
                        tma_load_2d(st, &map_a_hi, full_bar(s), k0, m0);
                        if (two) tma_load_2d(st + T::A_BYTES, &map_a_hi, full_bar(s), k0 + BK, m0);
                    }
                } else {
                    if (w_part) {
                        mbar_expect_tx(full_bar(s), T::STAGE_BYTES);
                        tma_load_2d(st + 2 * T::A_BYTES, &map_w_hi, full_bar(s), kb * BK, n0);
                        tma_load_2d(st + 2 * T::A_BYTES + T::W_BYTES, &map_w_lo, full_bar(s), kb * BK, n0);
                    }
                    if (a_part) {
                        tma_load_2d(st, &map_a_hi, full_bar(s), kb * BK, m0);
                        tma_load_2d(st + T::A_BYTES, &map_a_lo, full_bar(s), kb * BK, m0);
                    }
                }
            };
            uint32_t kbc = 0;
            for (int ct = ct0; ct < num_ct && kbc < (uint32_t)T::STAGES; ct += ct_step)
                for (int kb = 0; kb < num_kb && kbc < (uint32_t)T::STAGES; ++kb, ++kbc) load(ct, kb, kbc, true, true, false);
            pdl_wait();
            kbc = 0;
            for (int ct = ct0; ct < num_ct; ct += ct_step)
                for (int kb = 0; kb < num_kb; ++kb, ++kbc) {
                    const bool first_pass = kbc < (uint32_t)T::STAGES;
                    load(ct, kb, kbc, first_pass, !first_pass, true);
                }
        }
        __syncwarp();
        pdl_wait();
    } else if (warp == 1) {
        // ===== MMA issuer =====
        pdl_wait();
        constexpr uint32_t idesc = kF16 ? make_idesc_f16(BM, BN) : make_idesc_bf16(BM, BN);
        uint32_t kbc = 0, it = 0;
        for (int ct = ct0; ct < num_ct; ct += ct_step, ++it) {
            const uint32_t ab = it & 1, aph = (it >> 1) & 1;
            mbar_wait(tempty_bar(ab), aph ^ 1);                 // epilogue has drained this accumulator buffer
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + ab * BN;
            for (int kb = 0; kb < num_kb; ++kb, ++kbc) {
                const int s = kbc % T::STAGES;
                const uint32_t ph = (kbc / T::STAGES) & 1;
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t st = smem_base + s * T::STAGE_BYTES;
                    const uint64_t a_hi = make_smem_desc(st), a_lo = make_smem_desc(st + T::A_BYTES);
                    const uint64_t w_hi = make_smem_desc(st + 2 * T::A_BYTES), w_lo = make_smem_desc(st + 2 * T::A_BYTES + T::W_BYTES);
                    if constexpr (kF16) {
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                            umma_bf16(tmem_d, a_hi + adv, w_hi + adv, idesc, (kb | k) != 0);
                        }
                        if (kb * KSTAGE + BK < p.K) {
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k) {
                                const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                                umma_bf16(tmem_d, a_lo + adv, w_lo + adv, idesc, 1);   // K sub-block 1 (the "lo" slots)
                            }
                        }
                    } else {
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);       // 32 B per K-slice inside the swizzle row
                        umma_bf16(tmem_d, a_lo + adv, w_hi + adv, idesc, (kb | k) != 0);
                        umma_bf16(tmem_d, a_hi + adv, w_lo + adv, idesc, 1);
                        umma_bf16(tmem_d, a_hi + adv, w_hi + adv, idesc, 1);
                    }
                    }
                    umma_commit(empty_bar(s));
                    if (kb == num_kb - 1) umma_commit(tfull_bar(ab));   // accumulator complete
                }
                __syncwarp();
            }
        }
        pdl_trigger_late();                  // every MMA of this CTA is issued: its last epilogue is what is left
    } else {
        // ===== epilogue warps 2..9: TMEM lane quarter = warp % 4, column half = (warp-2)/4 =====
        pdl_wait();                             // bias / residual / mask reads and every store below come after the predecessor
        const int q = warp & 3, half = (warp - 2) >> 2;
        constexpr int COLS_PER_WARP = BN / 2;
        float* stg = reinterpret_cast<float*>(smem_gen + T::EPI_OFF) + (warp - 2) * 32 * EPI_PITCH;
        float* bs = reinterpret_cast<float*>(smem_gen + T::BS_OFF) + (warp - 2) * EPI_BS;
        const int sub = lane >> 2, cq = (lane & 3) * 4;         // phase 2: 8 rows x 4 float4 per warp instruction
        const float* resid = static_cast<const float*>(p.ep.resid);
        uint32_t it = 0;
        for (int ct = ct0; ct < num_ct; ct += ct_step, ++it) {
            const int m0 = (ct / n_tiles) * BM + q * 32, n0 = (ct % n_tiles) * BN + half * COLS_PER_WARP;
            const uint32_t ab = it & 1, aph = (it >> 1) & 1;
            float mk[4], pos[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = m0 + i * 8 + sub;
                mk[i] = (p.ep.mask && row < p.M) ? __ldg(p.ep.mask + row) : 1.f;
                pos[i] = 0.f;
                if (kRope && p.ep.rope_freqs && row < p.M) {   // rotary position of this output row
                    const int b = find_seq(p.ep.rope_off, p.ep.rope_B, row);
                    if (b >= 0) {
                        pos[i] = (float)(row - __ldg(p.ep.rope_off + b));
                        if (p.ep.rope_len) pos[i] = pos[i] / __ldg(p.ep.rope_len + b);
                    }
                }
            }
            // bias / layer-scale of this warp's columns -> shared memory while the tile's MMAs still run: with the operand ring in
            // place the SM has no L1, and a __ldg inside the chunk loop is an L2 round trip on the epilogue's critical path
            __syncwarp();
#pragma unroll
            for (int c = lane; c < COLS_PER_WARP; c += 32) {
                const int col = n0 + c;
                bs[c] = (p.ep.bias && col < p.N) ? __ldg(p.ep.bias + col) : 0.f;
                bs[128 + c] = (p.ep.scale && col < p.N) ? __ldg(p.ep.scale + col) : 1.f;
            }
            __syncwarp();
            mbar_wait(tfull_bar(ab), aph);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < COLS_PER_WARP; c += EPI_CHUNK) {
                uint32_t r[16];
                float4 rs[4];                                   // residual rows of this chunk: in flight under the TMEM load + staging
                if (resid && n0 + c + cq < p.N) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int row = m0 + i * 8 + sub;
                        if (row < p.M) rs[i] = *reinterpret_cast<const float4*>(resid + (size_t)row * p.ldo + n0 + c + cq);
                    }
                }
                __syncwarp();                                   // tcgen05.ld is .sync.aligned; staging of the previous chunk is consumed
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * BN + half * COLS_PER_WARP + c), r);
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<uint4*>(stg + lane * EPI_PITCH + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
                __syncwarp();
                const int col = n0 + c + cq;
                if (col < p.N) {                                // N % 4 == 0 is checked on the host
                    const float4 bias = *reinterpret_cast<const float4*>(bs + c + cq), scale = *reinterpret_cast<const float4*>(bs + 128 + c + cq);
                    float f0 = 0.f, f1 = 0.f;
                    if (kRope && p.ep.rope_freqs) {             // this float4 = rotary pairs i, i+1 of its head
                        const int i2 = (col % p.ep.rope_dh) >> 1;
                        f0 = __ldg(p.ep.rope_freqs + i2); f1 = __ldg(p.ep.rope_freqs + i2 + 1);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rl = i * 8 + sub, row = m0 + rl;
                        if (row >= p.M) continue;
                        float4 v = *reinterpret_cast<const float4*>(stg + rl * EPI_PITCH + cq);
                        v.x += bias.x; v.y += bias.y; v.z += bias.z; v.w += bias.w;
                        if (kRope && p.ep.rope_freqs) {
                            float s0, c0, s1, c1;
                            sincosf(pos[i] * f0, &s0, &c0); sincosf(pos[i] * f1, &s1, &c1);
                            v = make_float4(v.x * c0 - v.y * s0, v.x * s0 + v.y * c0, v.z * c1 - v.w * s1, v.z * s1 + v.w * c1);
                        }
                        if (p.ep.gelu) { v.x = gelu_erf_mufu(v.x); v.y = gelu_erf_mufu(v.y); v.z = gelu_erf_mufu(v.z); v.w = gelu_erf_mufu(v.w); }
                        if (p.ep.scale) { v.x *= scale.x; v.y *= scale.y; v.z *= scale.z; v.w *= scale.w; }
                        const size_t o = (size_t)row * p.ldo + col;
                        if (resid) { v.x += rs[i].x; v.y += rs[i].y; v.z += rs[i].z; v.w += rs[i].w; }
                        if (p.ep.mask) { v.x *= mk[i]; v.y *= mk[i]; v.z *= mk[i]; v.w *= mk[i]; }
                        if (p.split && !p.out_lo) {             // single fp16 operand of the next GEMM
                            *reinterpret_cast<uint2*>(p.out_hi + o) = make_uint2(pack_f16x2(v.x, v.y), pack_f16x2(v.z, v.w));
                        } else if (p.split) {
                            uint2 hi, lo;
                            split_pair(v.x, v.y, hi.x, lo.x); split_pair(v.z, v.w, hi.y, lo.y);
                            *reinterpret_cast<uint2*>(p.out_hi + o) = hi;
                            *reinterpret_cast<uint2*>(p.out_lo + o) = lo;
                        } else {
                            *reinterpret_cast<float4*>(p.out_f32 + o) = v;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(ab));
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, T::TMEM_COLS);
}

}  // namespace tc
}  // namespace stc
