// Two-SM ("CTA pair", cta_group::2) form of the split-bf16 tcgen05 GEMM for the large vocoder projections (sm_100a).
//
//   out[M,N] = epilogue( A[M,K] · W[N,K]^T ),   fp32 accumulate in TMEM, three kind::f16 MMAs per K-slice (gemm_tc.cuh).
//
// Why: the one-SM kernel is bound by SHARED-MEMORY bandwidth, not by the tensor pipe: per K-slice of 16 its three
// 128 x 256 MMAs read (a_lo, w_hi), (a_hi, w_lo), (a_hi, w_hi) = 36 KB in 384 tensor cycles (94 B/clk) while TMA writes the
// next stage into the same shared memory (62 B/clk) — 156 B/clk against the SM's 128 B/clk. A CTA pair computes a 256 x 256 tile
// with ONE instruction stream: each CTA holds its own 128 A rows and only HALF of the W tile (128 of the 256 weight rows), the
// tensor cores of both SMs read both halves. Per SM that is 24 KB of operand reads and 16 KB of TMA fill per K-slice
// (64 + 43 = 107 B/clk), and a stage is 64 KB instead of 96 KB.
//
// Protocol (cluster of 2 along x; rank 0 = leader):
//   warp 0 (both CTAs) : TMA producer — own A rows + own half of W into local smem with cp.async.bulk.tensor ... cta_group::2,
//                        completing on the LEADER's full barrier (which expects the bytes of both CTAs)
//   warp 1 (leader)    : issues tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16); tcgen05.commit ... multicast::cluster
//                        releases the stage in both CTAs and signals both CTAs' accumulator-full barriers
//   warps 2-17 (both)  : epilogue of the CTA's own 128 rows out of its own TMEM (same fused epilogue as the one-SM kernel);
//                        the accumulator buffer is handed back on the leader's barrier (remote mbarrier arrive from the peer)
#pragma once
#include "gemm_tc.cuh"

namespace stc {
namespace tc2 {

using namespace tc;

constexpr int BN = 256, HALF = BN / 2;
constexpr int A_BYTES = BM * BK * 2;                 // 16 KB: 128 rows x 64 bf16 (hi or lo)
constexpr int W_BYTES = HALF * BK * 2;               // 16 KB: this CTA's 128 weight rows x 64 bf16 (hi or lo)
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
constexpr int STAGES = 3;
constexpr int EPIW = 16;                             // epilogue warps: four per TMEM lane quarter, 64 accumulator columns each —
                                                     // with eight the GELU epilogue of a K = 512 tile (17k cycles, latency bound at two
                                                     // warps per scheduler) outlasts its MMAs (12k)
constexpr int THREADS = 64 + 32 * EPIW;
constexpr int STG_BYTES = 32 * EPI_CHUNK * 4;        // per-warp staging: 32 rows x 16 floats, 16-byte chunks XOR-swizzled (no padding)
constexpr int BAR_OFF = STAGES * STAGE_BYTES;
constexpr int EPI_OFF = BAR_OFF + 256;
constexpr int SMEM_BYTES = EPI_OFF + EPIW * STG_BYTES + 1024;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
constexpr int TMEM_COLS = 2 * BN;

STC_DEVINL uint32_t mapa_rank(uint32_t local, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank)); return r;
}
STC_DEVINL void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
STC_DEVINL void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
STC_DEVINL void umma2_commit(uint32_t bar) {         // arrives on the barrier at this offset in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
STC_DEVINL void tmem_alloc2(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
STC_DEVINL void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
STC_DEVINL void st_global_256(void* p, const uint32_t* v) {       // one full 32-byte sector per lane (sm_100: STG.256)
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
// TMEM -> registers without the wait, and the wait as a separate step that carries the destination registers as in/out operands, so
// that no consumer of them can be scheduled above it: the load of chunk c + 1 is in flight under the math of chunk c.
STC_DEVINL void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
STC_DEVINL void tmem_ld_wait16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}

// (The same lane = row form for the fp32 + residual epilogue of pw2 — 256-bit loads of the residual row, 256-bit stores — was
// measured slower than the staged one: vocoder 1.82 -> 1.91 ms; a warp-wide 32-byte access to 32 different rows costs 32 L2 requests.)

// Hands a drained TMEM accumulator back to the leader's MMA warp: the remote arrive in its default form (what CUTLASS's
// ClusterBarrier::arrive(cta_id) issues; SASS: a bare SYNCS.ARRIVE). The tcgen05.ld reads it orders are complete (tcgen05.wait::ld +
// tcgen05.fence::before_thread_sync). The explicit .release.cluster form used before compiles to MEMBAR.ALL.GPU + the arrive and made
// the warp wait for every global store of its tile to be acknowledged first (9 % of the warp samples of the fp16 pw1 were stall_membar).
STC_DEVINL void mbar_arrive_cluster(uint32_t bar_cluster) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}

// kF16: single-pass fp16 operands (gemm_tc.cuh): a stage holds 128 K-elements, 8 MMAs instead of 24 per 128 K-elements.
template <bool kF16 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm2_bf16x3_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                    const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                    const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));           // same offset in both CTAs of the pair
    const uint32_t bar_base = smem_base + BAR_OFF;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + BAR_OFF + 8 * (2 * STAGES + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    constexpr int KSTAGE = kF16 ? 2 * BK : BK;
    const int num_kb = (p.K + KSTAGE - 1) / KSTAGE;
    const int n_tiles = (p.N + BN - 1) / BN;
    const int m_pairs = ((p.M + BM - 1) / BM + 1) / 2;
    const int num_ct = m_pairs * n_tiles;
    const int ct0 = (int)cluster_id_x(), ct_step = (int)cluster_count_x();

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo);
        tma_prefetch_desc(&map_w_hi); tma_prefetch_desc(&map_w_lo);
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * EPIW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc2(tmem_slot, TMEM_COLS);          // one warp of EACH CTA, same warp id, same slot offset
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                 // the peer's barriers exist before anything is signalled at them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    pdl_wait();

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own A rows, own half of the W tile; bytes complete on the leader's barrier =====
        if (elect_one()) {
            uint32_t kbc = 0;
            for (int ct = ct0; ct < num_ct; ct += ct_step) {
                const int m0 = ((ct / n_tiles) * 2 + rank) * BM, n0 = (ct % n_tiles) * BN + rank * HALF;
                for (int kb = 0; kb < num_kb; ++kb, ++kbc) {
                    const int s = kbc % STAGES;
                    const uint32_t ph = (kbc / STAGES) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1);                         // the pair's MMAs on this stage have retired
                    if (p.trace && blockIdx.x == 0 && kbc < 48) p.trace[64 + kbc] = clock64();
                    const uint32_t st = smem_base + s * STAGE_BYTES;
                    const uint32_t fb = mapa_rank(full_bar(s), 0);
                    if constexpr (kF16) {
                        const int k0 = kb * KSTAGE;
                        const bool two = k0 + BK < p.K;
                        if (rank == 0) mbar_expect_tx(full_bar(s), two ? 2 * STAGE_BYTES : STAGE_BYTES);
                        tma_load_2d_2sm(st, &map_a_hi, fb, k0, m0);
                        tma_load_2d_2sm(st + 2 * A_BYTES, &map_w_hi, fb, k0, n0);
                        if (two) {
                            tma_load_2d_2sm(st + A_BYTES, &map_a_hi, fb, k0 + BK, m0);
                            tma_load_2d_2sm(st + 2 * A_BYTES + W_BYTES, &map_w_hi, fb, k0 + BK, n0);
                        }
                        continue;
                    }
                    if (rank == 0) mbar_expect_tx(full_bar(s), 2 * STAGE_BYTES);
                    tma_load_2d_2sm(st, &map_a_hi, fb, kb * BK, m0);
                    tma_load_2d_2sm(st + A_BYTES, &map_a_lo, fb, kb * BK, m0);
                    tma_load_2d_2sm(st + 2 * A_BYTES, &map_w_hi, fb, kb * BK, n0);
                    tma_load_2d_2sm(st + 2 * A_BYTES + W_BYTES, &map_w_lo, fb, kb * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (rank == 0) {
            constexpr uint32_t idesc = kF16 ? make_idesc_f16(2 * BM, BN) : make_idesc_bf16(2 * BM, BN);
            uint32_t kbc = 0, it = 0;
            for (int ct = ct0; ct < num_ct; ct += ct_step, ++it) {
                const uint32_t ab = it & 1, aph = (it >> 1) & 1;
                mbar_wait(tempty_bar(ab), aph ^ 1);             // both CTAs' epilogues have drained this accumulator buffer
                tc_fence_after();
                if (p.trace && blockIdx.x == 0 && it < 8 && lane == 0) p.trace[120 + it] = clock64();
                const uint32_t tmem_d = tmem_base + ab * BN;
                for (int kb = 0; kb < num_kb; ++kb, ++kbc) {
                    const int s = kbc % STAGES;
                    const uint32_t ph = (kbc / STAGES) & 1;
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    if (p.trace && blockIdx.x == 0 && kbc < 48 && lane == 0) p.trace[kbc] = clock64();
                    if (elect_one()) {
                        const uint32_t st = smem_base + s * STAGE_BYTES;
                        const uint64_t a_hi = make_smem_desc(st), a_lo = make_smem_desc(st + A_BYTES);
                        const uint64_t w_hi = make_smem_desc(st + 2 * A_BYTES), w_lo = make_smem_desc(st + 2 * A_BYTES + W_BYTES);
                        if constexpr (kF16) {
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k) {
                                const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                                umma2_bf16(tmem_d, a_hi + adv, w_hi + adv, idesc, (kb | k) != 0);
                            }
                            if (kb * KSTAGE + BK < p.K) {
#pragma unroll
                                for (int k = 0; k < BK / UMMA_K; ++k) {
                                    const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                                    umma2_bf16(tmem_d, a_lo + adv, w_lo + adv, idesc, 1);
                                }
                            }
                        } else {
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                            umma2_bf16(tmem_d, a_lo + adv, w_hi + adv, idesc, (kb | k) != 0);
                            umma2_bf16(tmem_d, a_hi + adv, w_lo + adv, idesc, 1);
                            umma2_bf16(tmem_d, a_hi + adv, w_hi + adv, idesc, 1);
                        }
                        }
                        umma2_commit(empty_bar(s));
                        if (kb == num_kb - 1) umma2_commit(tfull_bar(ab));
                    }
                    __syncwarp();
                }
            }
        }
        pdl_trigger_late();                  // (the peer's idle MMA warp releases at once: the leader's release is the pair's)
    } else {
        // ===== epilogue warps 2..17 (both CTAs): TMEM lane quarter = warp % 4, column quarter = (warp-2)/4 =====
        const int q = warp & 3, part = (warp - 2) >> 2;
        constexpr int COLS_PER_WARP = BN / (EPIW / 4);
        float* stg = reinterpret_cast<float*>(smem_gen + EPI_OFF + (warp - 2) * STG_BYTES);
        const int sub = lane >> 2, cq = lane & 3;               // phase 2: 8 rows x 4 float4 per warp instruction
        const float* resid = static_cast<const float*>(p.ep.resid);
        // fp16 operand out, bias + GELU only: the direct form below
        const bool direct = kF16 && p.split && !p.out_lo && p.ep.gelu && p.ep.bias && !p.ep.scale && !p.ep.mask && !resid && p.N % 64 == 0 && p.ldo % 16 == 0;
        // direct form: the warp's 64 bias values of EVERY column tile go to its staging slot once (2 KB = 8 tiles: N <= 2048)
        const bool bias_all = kF16 && direct && n_tiles * COLS_PER_WARP * (int)sizeof(float) <= STG_BYTES;
        if (bias_all) {
            for (int i = lane; i < n_tiles * COLS_PER_WARP; i += 32) {
                const int col = (i / COLS_PER_WARP) * BN + part * COLS_PER_WARP + i % COLS_PER_WARP;
                stg[i] = col < p.N ? __ldg(p.ep.bias + col) : 0.f;
            }
            __syncwarp();
        }
        uint32_t it = 0;
        for (int ct = ct0; ct < num_ct; ct += ct_step, ++it) {
            const int m0 = ((ct / n_tiles) * 2 + rank) * BM + q * 32, n0 = (ct % n_tiles) * BN + part * COLS_PER_WARP;
            const uint32_t ab = it & 1, aph = (it >> 1) & 1;
            float mk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = m0 + i * 8 + sub;
                mk[i] = (p.ep.mask && row < p.M) ? __ldg(p.ep.mask + row) : 1.f;
            }
            if (kF16 && direct) {
                // bias -> GELU -> fp16 straight out of TMEM: lane = row, 16 consecutive columns = 32 contiguous bytes per lane and chunk
                // (no shared-memory transpose; this epilogue, not the MMAs, bounds the K = 512 tiles of the vocoder's pw1). The bias
                // of every column tile sits in the warp's staging slot since the kernel's start (bias_all; otherwise it is fetched
                // here, an L2 round trip per tile: with 226 KB of shared memory in use there is no L1).
                const float* bs = stg + (bias_all ? (ct % n_tiles) * COLS_PER_WARP : 0);
                if (!bias_all) {
                    __syncwarp();
                    stg[lane] = n0 < p.N ? __ldg(p.ep.bias + n0 + lane) : 0.f;
                    stg[32 + lane] = n0 < p.N ? __ldg(p.ep.bias + n0 + 32 + lane) : 0.f;
                    __syncwarp();
                }
                mbar_wait(tfull_bar(ab), aph);
                tc_fence_after();
                const int row = m0 + lane;
                __nv_bfloat16* orow = p.out_hi + (size_t)row * p.ldo + n0;
                const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * BN + part * COLS_PER_WARP);
                uint32_t ra[16], rb[16];
                tmem_ld16_issue(tcol, ra);
                auto chunk = [&](uint32_t (&r)[16], uint32_t (&nxt)[16], int c) {
                    tmem_ld_wait16(r);
                    if (c + 16 < COLS_PER_WARP) tmem_ld16_issue(tcol + c + 16, nxt);
                    else {                                     // the whole accumulator slice is in registers: hand the buffer back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(mapa_rank(tempty_bar(ab), 0));
                    }
                    uint32_t o[8];
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 b = *reinterpret_cast<const float4*>(bs + c + j);                      // broadcast read
                        const float2 g0 = gelu_erf_rcp2(__fadd2_rn(make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), make_float2(b.x, b.y)));
                        const float2 g1 = gelu_erf_rcp2(__fadd2_rn(make_float2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])), make_float2(b.z, b.w)));
                        o[j / 2] = pack_f16x2(g0.x, g0.y); o[j / 2 + 1] = pack_f16x2(g1.x, g1.y);
                    }
                    if (row < p.M && n0 + c < p.N) st_global_256(orow + c, o);
                };
                static_assert(COLS_PER_WARP == 64, "four chunks of 16 columns");
                chunk(ra, rb, 0); chunk(rb, ra, 16); chunk(ra, rb, 32); chunk(rb, ra, 48);
                continue;
            }
            mbar_wait(tfull_bar(ab), aph);
            tc_fence_after();
            // bias / layer-scale of chunk c + 1 and the residual rows of chunk c are requested before the TMEM load of chunk c: there is
            // no L1 beside 226 KB of shared memory, every __ldg is an L2 round trip
            float4 bias_n = make_float4(0.f, 0.f, 0.f, 0.f), scale_n = make_float4(1.f, 1.f, 1.f, 1.f);
            if (n0 + cq * 4 < p.N) {
                if (p.ep.bias) bias_n = __ldg(reinterpret_cast<const float4*>(p.ep.bias + n0 + cq * 4));
                if (p.ep.scale) scale_n = __ldg(reinterpret_cast<const float4*>(p.ep.scale + n0 + cq * 4));
            }
#pragma unroll 1
            for (int c = 0; c < COLS_PER_WARP; c += EPI_CHUNK) {
                uint32_t r[16];
                const float4 bias = bias_n, scale = scale_n;
                if (c + EPI_CHUNK < COLS_PER_WARP && n0 + c + EPI_CHUNK + cq * 4 < p.N) {
                    if (p.ep.bias) bias_n = __ldg(reinterpret_cast<const float4*>(p.ep.bias + n0 + c + EPI_CHUNK + cq * 4));
                    if (p.ep.scale) scale_n = __ldg(reinterpret_cast<const float4*>(p.ep.scale + n0 + c + EPI_CHUNK + cq * 4));
                }
                float4 rs[4];
                if (resid && n0 + c + cq * 4 < p.N) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int row = m0 + i * 8 + sub;
                        if (row < p.M) rs[i] = *reinterpret_cast<const float4*>(resid + (size_t)row * p.ldo + n0 + c + cq * 4);
                    }
                }
                __syncwarp();
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * BN + part * COLS_PER_WARP + c), r);
                // lane = row: its four 16-byte chunks go to slots j ^ ((row >> 1) & 3) — conflict-free for the row-wise writes
                // here and for the 8-rows-by-4-chunks reads below
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(stg + lane * EPI_CHUNK + 4 * (j ^ ((lane >> 1) & 3))) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                __syncwarp();
                const int col = n0 + c + cq * 4;
                if (col < p.N) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rl = i * 8 + sub, row = m0 + rl;
                        if (row >= p.M) continue;
                        float4 v = *reinterpret_cast<const float4*>(stg + rl * EPI_CHUNK + 4 * (cq ^ ((rl >> 1) & 3)));
                        v.x += bias.x; v.y += bias.y; v.z += bias.z; v.w += bias.w;
                        if (p.ep.gelu) { v.x = gelu_erf_mufu(v.x); v.y = gelu_erf_mufu(v.y); v.z = gelu_erf_mufu(v.z); v.w = gelu_erf_mufu(v.w); }
                        if (p.ep.scale) { v.x *= scale.x; v.y *= scale.y; v.z *= scale.z; v.w *= scale.w; }
                        const size_t o = (size_t)row * p.ldo + col;
                        if (resid) { v.x += rs[i].x; v.y += rs[i].y; v.z += rs[i].z; v.w += rs[i].w; }
                        if (p.ep.mask) { v.x *= mk[i]; v.y *= mk[i]; v.z *= mk[i]; v.w *= mk[i]; }
                        if (p.split && !p.out_lo) {
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
            if (lane == 0) mbar_arrive_cluster(mapa_rank(tempty_bar(ab), 0));      // the leader's MMA warp waits for both CTAs
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                 // neither CTA leaves (or frees TMEM) while the pair still computes / signals
    tc_fence_after();
    if (warp == 1) tmem_dealloc2(tmem_base, TMEM_COLS);
}

}  // namespace tc2
}  // namespace stc
