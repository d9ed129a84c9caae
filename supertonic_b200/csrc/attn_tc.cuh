// tcgen05 / TMEM multi-head attention core for packed variable-length sequences (sm_100a):
//   O = softmax(Q K^T * scale  [keys >= Nk masked]) V      per (sequence b, head h), head dim 64, up to 320 keys.
// Used for the latent->text and latent->style cross-attentions of the vector estimator (total_step x per utterance,
// reference call site cpp/helper.cpp:620-647) and the text encoder's self / style attention.
//
// Arithmetic: the same split-bf16 ("bf16x3") scheme as the GEMMs — Q, K, V and the probabilities P are carried as
// (hi, lo) bf16 pairs and each product is hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM; the softmax itself
// (max, exp, sum, normalisation) is fp32 in registers. Exact two-pass softmax: the whole score row (<= 320 keys) stays
// in TMEM, so the row maximum is known before the first exponential and no rescaling is needed.
//
// One CTA = 128 queries of one (b, h). 192 threads:
//   warp 0   : TMA   — Q tile, all K blocks (64 keys each), all V^T blocks; 128B-swizzled K-major boxes
//   warp 1   : MMA   — S_j = Q K_j^T (N = 64 per block, K = 64 -> 4 slices x 3 MMAs) for every block, then, as the
//                      softmax warps hand over P_j (double-buffered in the shared memory the K blocks occupied),
//                      O += P_j V_j (A = P_j [128 x 64 keys], B = V^T_j [64 d x 64 keys])
//   warps 2-5: softmax — lane = query row: TMEM -> registers, max, exp2, sum, split to bf16 pairs, swizzled st.shared of P_j;
//                      finally O / sum -> split-bf16 (or fp32) rows of the output.
// Operands: Q/K as [rows, H*64] bf16 pairs written by the projection GEMMs (rotary embedding fused into their epilogue),
// V transposed per (b, h) to [64, keys padded to 64] by v_prep_kernel so that the PV product sees a K-major B operand.
#pragma once
#include "gemm_tc.cuh"

namespace stc {
namespace attn {

constexpr int DH = 64;            // head dim
constexpr int BQ = 128;           // queries per CTA (UMMA M)
constexpr int KB = 64;            // keys per block (one 128-byte swizzle row of P / V^T)
constexpr int MAX_BLOCKS = 5;     // up to 320 keys (a 300-byte chunk is ~310 tokens, cpp/helper.cpp:698)
constexpr int NUM_THREADS = 192;
constexpr int Q_BYTES = BQ * DH * 2;                   // 16 KB per half
constexpr int KBLK_BYTES = KB * DH * 2;                // 8 KB per half per block (K block and V^T block alike)
constexpr int P_BYTES = BQ * KB * 2;                   // 16 KB per half
constexpr int OFF_Q = 0;                               // Q hi, Q lo
constexpr int OFF_K = 2 * Q_BYTES;                     // K hi[5], K lo[5]; later P buffers: {hi, lo} x 2
constexpr int K_REGION = 2 * MAX_BLOCKS * KBLK_BYTES;  // 80 KB >= 4 * P_BYTES
constexpr int OFF_V = OFF_K + K_REGION;                // V^T hi[5], lo[5]
constexpr int OFF_BAR = OFF_V + K_REGION;
constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
constexpr int TMEM_COLS = 512;                         // S: 5 x 64 columns, O: 64 columns at 320
constexpr int O_COL = MAX_BLOCKS * KB;
static_assert(4 * P_BYTES <= K_REGION, "P double buffer must fit in the K region");
// Layout per maximum key-block count. MAXB = 1 (<= 64 keys: the 50-key style attentions) needs 80 KB of shared memory and 128
// TMEM columns, so TWO CTAs share an SM and the ~200 (sequence, head, query tile) CTAs of a batch run in one wave instead of two.
template <int MAXB> struct Lay {
    static constexpr int K_REGION = MAXB == 1 ? 2 * P_BYTES : (2 * MAXB * KBLK_BYTES > 4 * P_BYTES ? 2 * MAXB * KBLK_BYTES : 4 * P_BYTES);
    static constexpr int V_REGION = 2 * MAXB * KBLK_BYTES;
    static constexpr int OFF_V = OFF_K + K_REGION;
    static constexpr int OFF_BAR = OFF_V + V_REGION;
    static constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
    static constexpr int TMEM_COLS = MAXB == 1 ? 128 : 512;
    static constexpr int O_COL = MAXB * KB;
    static constexpr int CTAS_PER_SM = MAXB == 1 ? 2 : 1;
};
static_assert(Lay<MAX_BLOCKS>::SMEM_BYTES == SMEM_BYTES && Lay<MAX_BLOCKS>::O_COL == O_COL, "the 320-key layout is unchanged");

struct Params {
    const int* qoff; const int* koff; const int* kcnt;       // packed row offsets [B+1]; valid key count or null
    int heads;
    float scale_log2e;                                        // softmax scale * log2(e)
    float* out_f32; __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;   // [Rq, heads*64]
    int split;
};

template <int MAXB>
__global__ void __launch_bounds__(NUM_THREADS, Lay<MAXB>::CTAS_PER_SM)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
                    const __grid_constant__ CUtensorMap map_k_hi, const __grid_constant__ CUtensorMap map_k_lo,
                    const __grid_constant__ CUtensorMap map_v_hi, const __grid_constant__ CUtensorMap map_v_lo,
                    const Params p) {
    // (pre-wait, kernels.cuh: the sequence offsets, the K and V^T blocks were produced at least two kernels ago — the K / V projections
    // come before the Q projection, hoisted ones long before — so set-up and the K / V loads overlap the Q projection's tail)
    using namespace tc;
    using L = Lay<MAXB>;
    constexpr int OFF_V = L::OFF_V, OFF_BAR = L::OFF_BAR, TMEM_COLS = L::TMEM_COLS, O_COL = L::O_COL, MAX_BLOCKS = MAXB;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
    const int qbase = __ldg(p.qoff + b), Nq = __ldg(p.qoff + b + 1) - qbase;
    if (q0 >= Nq) return;                                     // block-uniform, before any barrier / TMEM use
    const int kbase = __ldg(p.koff + b);
    int Nk = __ldg(p.koff + b + 1) - kbase;
    if (p.kcnt) Nk = min(Nk, __ldg(p.kcnt + b));
    const int nblk = (Nk + KB - 1) / KB;                      // 1..MAX_BLOCKS (checked on the host)

    const uint32_t bar = smem_base + OFF_BAR;
    const uint32_t bar_qk = bar, bar_v = bar + 8, bar_s = bar + 16, bar_o = bar + 24;
    auto bar_p = [&](int i) { return bar + 32 + 8u * i; };         // P_i written (128 softmax threads arrive)
    auto bar_pfree = [&](int i) { return bar + 48 + 8u * i; };     // P_i consumed by the MMAs (tcgen05.commit)
    const uint32_t tmem_slot = bar + 64;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + OFF_BAR + 64);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_q_hi); tma_prefetch_desc(&map_q_lo); tma_prefetch_desc(&map_k_hi);
        tma_prefetch_desc(&map_k_lo); tma_prefetch_desc(&map_v_hi); tma_prefetch_desc(&map_v_lo);
        mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(bar_p(i), 4); mbar_init(bar_pfree(i), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(bar_qk, 2 * Q_BYTES + 2 * nblk * KBLK_BYTES);
            for (int j = 0; j < nblk; ++j) {
                tma_load_2d(smem_base + OFF_K + j * KBLK_BYTES, &map_k_hi, bar_qk, h * DH, kbase + j * KB);
                tma_load_2d(smem_base + OFF_K + (MAX_BLOCKS + j) * KBLK_BYTES, &map_k_lo, bar_qk, h * DH, kbase + j * KB);
            }
            mbar_expect_tx(bar_v, 2 * nblk * KBLK_BYTES);
            const int vrow = (b * p.heads + h) * DH;
            for (int j = 0; j < nblk; ++j) {
                tma_load_2d(smem_base + OFF_V + j * KBLK_BYTES, &map_v_hi, bar_v, j * KB, vrow);
                tma_load_2d(smem_base + OFF_V + (MAX_BLOCKS + j) * KBLK_BYTES, &map_v_lo, bar_v, j * KB, vrow);
            }
            pdl_wait();                                       // Q is the predecessor's output
            tma_load_2d(smem_base + OFF_Q, &map_q_hi, bar_qk, h * DH, qbase + q0);
            tma_load_2d(smem_base + OFF_Q + Q_BYTES, &map_q_lo, bar_qk, h * DH, qbase + q0);
        }
        __syncwarp();
        pdl_wait(); pdl_trigger_light();
    } else if (warp == 1) {
        pdl_wait(); pdl_trigger_light();
        constexpr uint32_t idesc = make_idesc_bf16(BQ, KB);           // M = 128, N = 64 for both products
        mbar_wait(bar_qk, 0);
        tc_fence_after();
        if (elect_one()) {
            const uint64_t q_hi = make_smem_desc(smem_base + OFF_Q), q_lo = make_smem_desc(smem_base + OFF_Q + Q_BYTES);
            for (int j = 0; j < nblk; ++j) {
                const uint64_t k_hi = make_smem_desc(smem_base + OFF_K + j * KBLK_BYTES);
                const uint64_t k_lo = make_smem_desc(smem_base + OFF_K + (MAX_BLOCKS + j) * KBLK_BYTES);
                const uint32_t d = tmem_base + j * KB;
#pragma unroll
                for (int k = 0; k < DH / UMMA_K; ++k) {
                    const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                    umma_bf16(d, q_lo + adv, k_hi + adv, idesc, k != 0);
                    umma_bf16(d, q_hi + adv, k_lo + adv, idesc, 1);
                    umma_bf16(d, q_hi + adv, k_hi + adv, idesc, 1);
                }
            }
            umma_commit(bar_s);                                       // all scores complete; the K blocks are dead
        }
        __syncwarp();
        mbar_wait(bar_v, 0);
        for (int j = 0; j < nblk; ++j) {
            const int i = j & 1;
            mbar_wait(bar_p(i), (j >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t p_hi = make_smem_desc(smem_base + OFF_K + (2 * i) * P_BYTES);
                const uint64_t p_lo = make_smem_desc(smem_base + OFF_K + (2 * i + 1) * P_BYTES);
                const uint64_t v_hi = make_smem_desc(smem_base + OFF_V + j * KBLK_BYTES);
                const uint64_t v_lo = make_smem_desc(smem_base + OFF_V + (MAX_BLOCKS + j) * KBLK_BYTES);
                const uint32_t d = tmem_base + O_COL;
#pragma unroll
                for (int k = 0; k < KB / UMMA_K; ++k) {
                    const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                    umma_bf16(d, p_lo + adv, v_hi + adv, idesc, (j | k) != 0);
                    umma_bf16(d, p_hi + adv, v_lo + adv, idesc, 1);
                    umma_bf16(d, p_hi + adv, v_hi + adv, idesc, 1);
                }
                umma_commit(bar_pfree(i));
                if (j == nblk - 1) umma_commit(bar_o);
            }
            __syncwarp();
        }
    } else {
        // ===== softmax warps 2..5: query row r = 32 * (warp % 4) + lane =====
        pdl_wait(); pdl_trigger_light();
        const int quarter = warp & 3, r = quarter * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
        mbar_wait(bar_s, 0);
        tc_fence_after();
        // pass 1: row maximum over the valid keys
        float mx = -INFINITY;
        for (int c = 0; c < Nk; c += 32) {
            uint32_t v[32];
            __syncwarp();
            tmem_ld32(trow + c, v);
#pragma unroll
            for (int t = 0; t < 32; ++t) if (c + t < Nk) mx = fmaxf(mx, __uint_as_float(v[t]));
        }
        const float mxs = mx * p.scale_log2e;
        // pass 2: P_j = exp2(s*scale*log2e - max*scale*log2e), row sum, split, swizzled store
        float sum = 0.f;
        const uint32_t prow_off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
        for (int j = 0; j < nblk; ++j) {
            const int i = j & 1;
            if (j >= 2) mbar_wait(bar_pfree(i), ((j >> 1) - 1) & 1);          // MMAs of block j-2 have read this buffer
            uint8_t* p_hi = smem_gen + OFF_K + (2 * i) * P_BYTES + prow_off;
            uint8_t* p_lo = smem_gen + OFF_K + (2 * i + 1) * P_BYTES + prow_off;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t v[32];
                __syncwarp();
                tmem_ld32(trow + j * KB + half * 32, v);
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {                              // 8 keys = one 16-byte chunk
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int key = j * KB + half * 32 + c8 * 8 + 2 * t;
                        float e0 = exp2f(fmaf(__uint_as_float(v[c8 * 8 + 2 * t]), p.scale_log2e, -mxs));
                        float e1 = exp2f(fmaf(__uint_as_float(v[c8 * 8 + 2 * t + 1]), p.scale_log2e, -mxs));
                        e0 = key < Nk ? e0 : 0.f; e1 = key + 1 < Nk ? e1 : 0.f;
                        sum += e0 + e1;
                        split_pair(e0, e1, hi[t], lo[t]);
                    }
                    const int chunk = (half * 4 + c8) ^ (r & 7);              // 128B swizzle: chunk index XOR row-in-atom
                    *reinterpret_cast<uint4*>(p_hi + chunk * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(p_lo + chunk * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy stores -> visible to the MMA (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_p(i));
        }
        // epilogue: O / sum
        mbar_wait(bar_o, 0);
        tc_fence_after();
        const float inv = sum > 0.f ? 1.0f / sum : 0.f;
        const bool row_ok = q0 + r < Nq;
        const size_t o = ((size_t)qbase + q0 + r) * (size_t)(p.heads * DH) + (size_t)h * DH;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t v[32];
            __syncwarp();
            tmem_ld32(trow + O_COL + half * 32, v);
            if (row_ok) {
                if (p.split) {
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            split_pair(__uint_as_float(v[c8 * 8 + 2 * t]) * inv, __uint_as_float(v[c8 * 8 + 2 * t + 1]) * inv, hi[t], lo[t]);
                        *reinterpret_cast<uint4*>(p.out_hi + o + half * 32 + c8 * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4*>(p.out_lo + o + half * 32 + c8 * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < 32; t += 4)
                        *reinterpret_cast<float4*>(p.out_f32 + o + half * 32 + t) =
                            make_float4(__uint_as_float(v[t]) * inv, __uint_as_float(v[t + 1]) * inv, __uint_as_float(v[t + 2]) * inv,
                                        __uint_as_float(v[t + 3]) * inv);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---- operand preparation -----------------------------------------------------------------------------------------
// Q and K arrive as split-bf16 [rows, heads*64] straight from their projection GEMMs (rotary embedding fused into the
// GEMM epilogue, gemm_tc.cuh); only V needs a pass of its own:
// V: fp32 [Rk, heads*64] packed rows -> V^T (hi, lo) [(b*heads + h)*64 + d][ldk], keys >= Nk_b zero-filled up to ldk.
// Block = (64-key tile, head, sequence): coalesced reads along d, transposed through shared memory, coalesced writes along keys.
__global__ void __launch_bounds__(256)
v_prep_kernel(const float* __restrict__ v, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
              const int* __restrict__ koff, int heads, int ldk) {
    pdl_wait(); pdl_trigger_light();
    __shared__ float tile[KB][DH + 1];
    const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * KB;
    const int kbase = __ldg(koff + b), Nk = __ldg(koff + b + 1) - kbase;
    for (int i = threadIdx.x; i < KB * DH; i += blockDim.x) {
        const int kk = i / DH, d = i % DH;
        tile[kk][d] = (k0 + kk < Nk) ? v[((size_t)kbase + k0 + kk) * heads * DH + (size_t)h * DH + d] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < DH * (KB / 2); i += blockDim.x) {
        const int d = i / (KB / 2), kk = (i % (KB / 2)) * 2;
        uint32_t ph, pl;
        tc::split_pair(tile[kk][d], tile[kk + 1][d], ph, pl);
        const size_t o = ((size_t)(b * heads + h) * DH + d) * ldk + k0 + kk;
        *reinterpret_cast<uint32_t*>(hi + o) = ph;
        *reinterpret_cast<uint32_t*>(lo + o) = pl;
    }
}

}  // namespace attn
}  // namespace stc
