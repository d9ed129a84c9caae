// tcgen05 / TMEM / TMA GEMM for the dense contractions (pointwise projections, attention projections)
// of the Supertonic forward pass — hand-written for sm_100a.
//
//   out[M,N] = epilogue( A[M,K] · W[N,K]^T ),   fp32 accumulate in TMEM.
//
// Arithmetic: "bf16x3" — each fp32 operand v is carried as a split pair (hi = bf16(v), lo = bf16(v-hi))
// and the product is formed as A_hi·W_hi + A_lo·W_hi + A_hi·W_lo with three kind::f16 MMAs per K-slice
// (bf16 x bf16 products are exact in fp32; the dropped lo·lo term is 2^-16 relative). Measured effect on the
// full surrogate at 20 Euler steps: max-abs latent error 1e-5 vs 9e-4 for single-pass TF32 and 7e-3 for plain
// bf16 (DESIGN.md "precision"); the north-star bound is 1e-3.
//
// Structure (one 128 x BN output tile per CTA, 192 threads):
//   warp 0   : TMA producer   — cp.async.bulk.tensor.2d, 128B-swizzled K-major tiles, STAGES-deep mbarrier ring
//   warp 1   : MMA issuer     — one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16);
//                               tcgen05.commit releases smem stages and finally signals the accumulator barrier
//   warps 2-5: epilogue       — tcgen05.ld 32x32b.x32 (lane = row, 32 columns per load), fused
//                               bias / GELU(erf) / layer-scale / residual / mask / Euler update, then
//                               fp32 or split-bf16 stores.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
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

template <int BN> struct Tile {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int W_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
    static constexpr int STAGES = (196608 / STAGE_BYTES) > 6 ? 6 : (196608 / STAGE_BYTES);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int TMEM_COLS = 2 * BN;        // two accumulator buffers (BN in {64,128} -> power of two)
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

struct Params {
    int M, N, K;
    Epilogue ep;
    float* out_f32;                 // [M, ldo] when !split
    __nv_bfloat16* out_hi;          // [M, ldo] when split
    __nv_bfloat16* out_lo;
    int ldo;
    int split;
};

// Fused epilogue for 32 consecutive columns of one row held in registers.
template <bool kFull>
STC_DEVINL void epilogue_store(const Params& p, float (&v)[32], int row, int col0, float mk) {
    const float* resid = static_cast<const float*>(p.ep.resid);
    const size_t o = (size_t)row * p.ldo + col0;
    if (kFull) {
        if (p.ep.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 b = __ldg(reinterpret_cast<const float4*>(p.ep.bias + col0 + j));
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
        }
        if (p.ep.gelu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]);
        }
        if (p.ep.scale) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 s = __ldg(reinterpret_cast<const float4*>(p.ep.scale + col0 + j));
                v[j] *= s.x; v[j + 1] *= s.y; v[j + 2] *= s.z; v[j + 3] *= s.w;
            }
        }
        if (resid) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 s = *reinterpret_cast<const float4*>(resid + o + j);
                v[j] += s.x; v[j + 1] += s.y; v[j + 2] += s.z; v[j + 3] += s.w;
            }
        }
        if (p.ep.mask) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= mk;
        }
        if (p.split) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    __nv_bfloat16 h0 = __float2bfloat16_rn(v[j + 2 * t]), h1 = __float2bfloat16_rn(v[j + 2 * t + 1]);
                    __nv_bfloat16 l0 = __float2bfloat16_rn(v[j + 2 * t] - __bfloat162float(h0));
                    __nv_bfloat16 l1 = __float2bfloat16_rn(v[j + 2 * t + 1] - __bfloat162float(h1));
                    hi[t] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                    lo[t] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                }
                *reinterpret_cast<uint4*>(p.out_hi + o + j) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(p.out_lo + o + j) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(p.out_f32 + o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
    } else {
        // ragged N edge: scalar path (static indexing keeps v[] in registers)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (col0 + j < p.N) {
                float x = v[j];
                if (p.ep.bias) x += p.ep.bias[col0 + j];
                if (p.ep.gelu) x = gelu_erf_fast(x);
                if (p.ep.scale) x *= p.ep.scale[col0 + j];
                if (resid) x += resid[o + j];
                if (p.ep.mask) x *= mk;
                if (p.split) {
                    __nv_bfloat16 h = __float2bfloat16_rn(x);
                    p.out_hi[o + j] = h;
                    p.out_lo[o + j] = __float2bfloat16_rn(x - __bfloat162float(h));
                } else p.out_f32[o + j] = x;
            }
        }
    }
}

// Persistent over output tiles (tile = blockIdx.x + i*gridDim.x, n fastest). Two TMEM accumulator buffers:
// the epilogue of tile i overlaps the MMAs of tile i+1.
template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                   const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                   const Params p) {
    using T = Tile<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // SWIZZLE_128B needs 1024-B alignment
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + T::STAGES * T::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (T::STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * T::STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * T::STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * T::STAGES + 4);
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + T::STAGES * T::STAGE_BYTES + 8 * (2 * T::STAGES + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (p.K + BK - 1) / BK;
    const int n_tiles = (p.N + BN - 1) / BN;
    const int m_tiles = (p.M + BM - 1) / BM;
    const int num_tiles = n_tiles * m_tiles;

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
        if (elect_one()) {
            uint32_t kbc = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
                for (int kb = 0; kb < num_kb; ++kb, ++kbc) {
                    const int s = kbc % T::STAGES;
                    const uint32_t ph = (kbc / T::STAGES) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1);
                    const uint32_t st = smem_base + s * T::STAGE_BYTES;
                    mbar_expect_tx(full_bar(s), T::STAGE_BYTES);
                    tma_load_2d(st, &map_a_hi, full_bar(s), kb * BK, m0);
                    tma_load_2d(st + T::A_BYTES, &map_a_lo, full_bar(s), kb * BK, m0);
                    tma_load_2d(st + 2 * T::A_BYTES, &map_w_hi, full_bar(s), kb * BK, n0);
                    tma_load_2d(st + 2 * T::A_BYTES + T::W_BYTES, &map_w_lo, full_bar(s), kb * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
        uint32_t kbc = 0, it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
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
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);       // 32 B per K-slice inside the swizzle row
                        umma_bf16(tmem_d, a_lo + adv, w_hi + adv, idesc, (kb | k) != 0);
                        umma_bf16(tmem_d, a_hi + adv, w_lo + adv, idesc, 1);
                        umma_bf16(tmem_d, a_hi + adv, w_hi + adv, idesc, 1);
                    }
                    umma_commit(empty_bar(s));                          // frees the smem stage when these MMAs retire
                    if (kb == num_kb - 1) umma_commit(tfull_bar(ab));   // accumulator complete
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue warps 2..9: TMEM lane quarter = warp % 4, column half = (warp-2)/4 =====
        const int q = warp & 3, half = (warp - 2) >> 2;
        constexpr int COLS_PER_WARP = BN / 2;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
            const uint32_t ab = it & 1, aph = (it >> 1) & 1;
            const int row = m0 + q * 32 + lane;
            const bool row_ok = row < p.M;
            const float mk = (p.ep.mask && row_ok) ? __ldg(p.ep.mask + row) : 1.f;
            mbar_wait(tfull_bar(ab), aph);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < COLS_PER_WARP; c += 32) {
                const int c0 = half * COLS_PER_WARP + c;
                uint32_t r[32];
                __syncwarp();                                   // tcgen05.ld is .sync.aligned: reconverge first
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * BN + c0), r);
                const int col0 = n0 + c0;
                if (row_ok && col0 < p.N) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    if (col0 + 32 <= p.N) epilogue_store<true>(p, v, row, col0, mk);
                    else epilogue_store<false>(p, v, row, col0, mk);
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
