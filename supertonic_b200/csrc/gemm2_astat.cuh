// A-stationary form of the two-SM fp16 GEMM for the vocoder's pw1 (M = 27.7k rows, N = 2048, K = 512; bias + GELU -> fp16 operand):
//
//   out[M,N] = fp16( GELU( A[M,K] . W[N,K]^T + bias ) ),   single-pass fp16 operands, fp32 accumulation in TMEM (sm_100a)
//
// Why: with 256 x 256 tiles the kernel of gemm2_tc.cuh pulls 512 KB of operands from L2 per tile pair = 128 FLOP per byte; ncu shows
// 460 MB of L2 -> SM traffic per launch = 7.9 TB/s while it runs at ~1000 TFLOP/s — it is bound by the L2 -> SM fabric, not by the
// tensor pipe (43 % busy) and, after the one-MUFU GELU, not by its epilogue. K = 512 is small enough for a CTA's 128 A rows to STAY in
// shared memory (128 KB): a cluster takes a unit of (row-tile pair, NG = 4 consecutive column tiles), loads A once per unit and
// streams only its halves of the four W tiles through a five-slot ring of 64-element K blocks -> 160 KB instead of 256 KB per SM
// and tile, 205 FLOP per byte.
//
// Protocol (cluster of 2 along x; rank 0 = leader), as in gemm2_tc.cuh but with separate barriers for the resident A blocks:
//   warp 0 (both CTAs) : TMA — for the unit's first tile the A block kb right before the W block kb (a_empty[kb] -> a_full[kb], both
//                        CTAs' bytes complete on the LEADER's barrier), then W blocks only (w_empty[s] -> w_full[s])
//   warp 1 (leader)    : tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16) x 4 per K block; tcgen05.commit ... multicast::cluster
//                        frees the W slot in both CTAs, on the unit's LAST tile also the A block (so the next unit's A streams in
//                        under the last tile's MMAs), after the last K block signals the accumulator
//   warps 2-17 (both)  : epilogue of the CTA's own 128 rows: TMEM -> bias -> GELU (one MUFU) -> fp16, 16-column chunks with the next
//                        chunk's tcgen05.ld in flight; the warp's bias values of the unit's four column tiles are staged per unit
#pragma once
#include "gemm2_tc.cuh"

namespace stc {
namespace tc2a {

using namespace tc;
using tc2::mapa_rank; using tc2::tma_load_2d_2sm; using tc2::umma2_bf16; using tc2::umma2_commit; using tc2::tmem_alloc2; using tc2::tmem_dealloc2;
using tc2::st_global_256; using tc2::mbar_arrive_cluster; using tc2::tmem_ld16_issue; using tc2::tmem_ld_wait16;

constexpr int BN = 256, HALF = 128;
constexpr int KBLK = 64;                               // K elements per block (one 128-byte swizzle row of fp16)
constexpr int BLK_BYTES = BM * KBLK * 2;               // 16 KB: 128 rows (A rows, or this CTA's half of the W tile) x 64 fp16
constexpr int MAX_KB = 8;                              // K <= 512
constexpr int WST = 5;                                 // W ring slots
constexpr int NG = 4;                                  // column tiles per unit
constexpr int EPIW = 16;
constexpr int THREADS = 64 + 32 * EPIW;
constexpr int COLS_PER_WARP = BN / (EPIW / 4);         // 64
constexpr int STG_BYTES = NG * COLS_PER_WARP * 4;      // per-warp bias staging: 1 KB
constexpr int A_OFF = 0, W_OFF = MAX_KB * BLK_BYTES, BAR_OFF = W_OFF + WST * BLK_BYTES;
constexpr int N_BARS = 2 * MAX_KB + 2 * WST + 4;
constexpr int EPI_OFF = BAR_OFF + 256;
constexpr int SMEM_BYTES = EPI_OFF + EPIW * STG_BYTES + 1024;
static_assert(8 * N_BARS + 8 <= 256, "barrier block");
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
constexpr int TMEM_COLS = 2 * BN;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm2_f16_astat_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));           // same offset in both CTAs of the pair
    const uint32_t bar_base = smem_base + BAR_OFF;
    auto a_full = [&](int kb) { return bar_base + 8u * kb; };
    auto a_empty = [&](int kb) { return bar_base + 8u * (MAX_KB + kb); };
    auto w_full = [&](int s) { return bar_base + 8u * (2 * MAX_KB + s); };
    auto w_empty = [&](int s) { return bar_base + 8u * (2 * MAX_KB + WST + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * MAX_KB + 2 * WST + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * MAX_KB + 2 * WST + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * N_BARS;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + BAR_OFF + 8 * N_BARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int num_kb = (p.K + KBLK - 1) / KBLK;                               // <= MAX_KB (checked on the host)
    const int n_tiles = (p.N + BN - 1) / BN;
    const int n_groups = (n_tiles + NG - 1) / NG;
    const int m_pairs = ((p.M + BM - 1) / BM + 1) / 2;
    const int num_units = m_pairs * n_groups;
    const int u0 = (int)cluster_id_x(), u_step = (int)cluster_count_x();

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_w);
        for (int i = 0; i < 2 * MAX_KB + 2 * WST + 2; ++i) mbar_init(bar_base + 8u * i, 1);
        for (int a = 0; a < 2; ++a) mbar_init(tempty_bar(a), 2 * EPIW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc2(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    pdl_wait();

    if (warp == 0) {
        // ===== TMA producer (both CTAs) =====
        if (elect_one()) {
            uint32_t wc = 0, ui = 0;
            if (p.trace && blockIdx.x == 0) p.trace[64] = clock64();
            for (int u = u0; u < num_units; u += u_step, ++ui) {
                const int m0 = ((u / n_groups) * 2 + rank) * BM;
                const int t0 = (u % n_groups) * NG, t1 = min(n_tiles, t0 + NG);
                for (int t = t0; t < t1; ++t) {
                    const int n0 = t * BN + rank * HALF;
                    for (int kb = 0; kb < num_kb; ++kb, ++wc) {
                        if (t == t0) {
                            mbar_wait(a_empty(kb), (ui & 1) ^ 1);            // the previous unit's last tile has read this A block
                            if (rank == 0) mbar_expect_tx(a_full(kb), 2 * BLK_BYTES);
                            tma_load_2d_2sm(smem_base + A_OFF + kb * BLK_BYTES, &map_a, mapa_rank(a_full(kb), 0), kb * KBLK, m0);
                        }
                        const int s = wc % WST;
                        mbar_wait(w_empty(s), ((wc / WST) & 1) ^ 1);
                        if (rank == 0) mbar_expect_tx(w_full(s), 2 * BLK_BYTES);
                        tma_load_2d_2sm(smem_base + W_OFF + s * BLK_BYTES, &map_w, mapa_rank(w_full(s), 0), kb * KBLK, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (rank == 0) {
            constexpr uint32_t idesc = make_idesc_f16(2 * BM, BN);
            uint32_t wc = 0, ui = 0, it = 0;
            for (int u = u0; u < num_units; u += u_step, ++ui) {
                const int t0 = (u % n_groups) * NG, t1 = min(n_tiles, t0 + NG);
                for (int t = t0; t < t1; ++t, ++it) {
                    const uint32_t ab = it & 1, aph = (it >> 1) & 1;
                    const bool tr = p.trace && blockIdx.x == 0 && it < 16 && lane == 0;      // debug stamps (stc_debug_gemm, STC_GEMM_TRACE=1)
                    if (tr) p.trace[3 * it] = clock64();
                    mbar_wait(tempty_bar(ab), aph ^ 1);         // both CTAs' epilogues have drained this accumulator buffer
                    tc_fence_after();
                    if (tr) p.trace[3 * it + 1] = clock64();
                    const uint32_t tmem_d = tmem_base + ab * BN;
                    for (int kb = 0; kb < num_kb; ++kb, ++wc) {
                        if (t == t0) mbar_wait(a_full(kb), ui & 1);
                        const int s = wc % WST;
                        mbar_wait(w_full(s), (wc / WST) & 1);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t a_d = make_smem_desc(smem_base + A_OFF + kb * BLK_BYTES);
                            const uint64_t w_d = make_smem_desc(smem_base + W_OFF + s * BLK_BYTES);
#pragma unroll
                            for (int k = 0; k < KBLK / UMMA_K; ++k) {
                                const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                                umma2_bf16(tmem_d, a_d + adv, w_d + adv, idesc, (kb | k) != 0);
                            }
                            umma2_commit(w_empty(s));
                            if (t == t1 - 1) umma2_commit(a_empty(kb));
                            if (kb == num_kb - 1) umma2_commit(tfull_bar(ab));
                        }
                        __syncwarp();
                    }
                    if (tr) p.trace[3 * it + 2] = clock64();
                }
            }
        }
        pdl_trigger_late();
    } else {
        // ===== epilogue warps 2..17 (both CTAs): TMEM lane quarter = warp % 4, column quarter = (warp - 2) / 4 =====
        const int q = warp & 3, part = (warp - 2) >> 2;
        float* stg = reinterpret_cast<float*>(smem_gen + EPI_OFF + (warp - 2) * STG_BYTES);
        uint32_t it = 0;
        for (int u = u0; u < num_units; u += u_step) {
            const int m0 = ((u / n_groups) * 2 + rank) * BM + q * 32;
            const int t0 = (u % n_groups) * NG, t1 = min(n_tiles, t0 + NG);
            __syncwarp();                                       // every lane is done with the previous unit's bias values
            for (int i = lane; i < (t1 - t0) * COLS_PER_WARP; i += 32) {
                const int col = (t0 + i / COLS_PER_WARP) * BN + part * COLS_PER_WARP + i % COLS_PER_WARP;
                stg[i] = col < p.N ? __ldg(p.ep.bias + col) : 0.f;
            }
            __syncwarp();
            const int row = m0 + lane;
            for (int t = t0; t < t1; ++t, ++it) {
                const int n0 = t * BN + part * COLS_PER_WARP;
                const uint32_t ab = it & 1, aph = (it >> 1) & 1;
                const float* bs = stg + (t - t0) * COLS_PER_WARP;
                const bool tr = p.trace && blockIdx.x == 0 && warp == 2 && it < 16 && lane == 0;
                if (tr) p.trace[65 + 3 * it] = clock64();
                mbar_wait(tfull_bar(ab), aph);
                tc_fence_after();
                if (tr) p.trace[66 + 3 * it] = clock64();
                __nv_bfloat16* orow = p.out_hi + (size_t)row * p.ldo + n0;
                const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * BN + part * COLS_PER_WARP);
                // (all 64 columns in one tcgen05.ld round trip, the buffer handed back at once, was measured SLOWER: 62.3 against 55.8 us)
                // (measured and not kept, profiles/r2an_pw1_epilogue.txt: all 64 columns in one tcgen05.ld round trip with the buffer handed
                // back at once 62.3 us; the next chunk's load fenced in front of this chunk's math 56.0 us; the bias values one chunk
                // ahead in registers 58.2 us — against 55.5 us for this form. The chunk's ~112 packed fp32 operations occupy the FMA
                // pipe for two cycles each: 3.6k of the 5.4k cycles a tile's epilogue takes per scheduler.)
                uint32_t ra[16], rb[16];
                tmem_ld16_issue(tcol, ra);
                auto chunk = [&](uint32_t (&r)[16], uint32_t (&nxt)[16], int c) {
                    tmem_ld_wait16(r);
                    if (c + 16 < COLS_PER_WARP) tmem_ld16_issue(tcol + c + 16, nxt);
                    else {                                      // the accumulator slice is in registers: hand the buffer back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(mapa_rank(tempty_bar(ab), 0));
                    }
                    uint32_t o[8];
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 b = *reinterpret_cast<const float4*>(bs + c + j);
                        const float2 g0 = gelu_erf_rcp2(__fadd2_rn(make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), make_float2(b.x, b.y)));
                        const float2 g1 = gelu_erf_rcp2(__fadd2_rn(make_float2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])), make_float2(b.z, b.w)));
                        o[j / 2] = pack_f16x2(g0.x, g0.y); o[j / 2 + 1] = pack_f16x2(g1.x, g1.y);
                    }
                    if (row < p.M && n0 + c < p.N) st_global_256(orow + c, o);
                };
                chunk(ra, rb, 0); chunk(rb, ra, 16); chunk(ra, rb, 32); chunk(rb, ra, 48);
                if (tr) p.trace[67 + 3 * it] = clock64();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                 // neither CTA leaves (or frees TMEM) while the pair still computes / signals
    tc_fence_after();
    if (warp == 1) tmem_dealloc2(tmem_base, TMEM_COLS);
}

}  // namespace tc2a
}  // namespace stc
