// Fused ConvNeXt MLP on tcgen05 for the C = 256, H = 1024 blocks of the vector estimator and the text encoder (sm_100a):
//
//     x[rows, C]  <-  ( x + gamma * ( GELU( a W1 + b1 ) W2 + b2 ) ) * mask          a = LayerNorm(dwconv(x)) as split-bf16
//
// replaces ORT's MatMul + Add -> Erf-GELU -> MatMul + Add -> Mul -> Add -> Mul(mask) node chain of every ConvNeXt block
// (vector_estimator.onnx / text_encoder.onnx, run at reference cpp/helper.cpp:552, 643). The hidden activations never leave the SM.
//
// ONE kernel for every row count ("stream" form). A 128-row tile is shared by `nslice` CTAs; CTA (tile, slice) owns a contiguous
// range of the 1024 hidden units (a multiple of 64: the sixteen 64-unit blocks are dealt out as evenly as possible, so nslice
// need not divide 16) and walks it in CHUNKS of 128 (or a final 64) hidden units, flash-attention style:
//
//     S_c[128 x w]  = a[128 x 256] . W1[chunk c, :]^T              (TMEM buffer c % 2, 128 columns)
//     P_c           = split-bf16( GELU(S_c + b1) )                  written by the epilogue warps IN PLACE over S_c in TMEM
//     O[128 x 256] += P_c[128 x w] . W2[:, chunk c]^T               (TMEM columns 256..511; A operand read from TMEM)
//
// The MMA warp issues S_{c+1} before O_c, so the GELU of chunk c runs under the MMAs of S_{c+1} and the tensor pipe never waits for
// the epilogue warps in steady state. The host picks the plan (model.cu mlp_plan: slices per tile, single CTAs or pairs, from a
// cost model fitted to tools/mlp_sweep.py): <= 9 tiles -> 16 slices of 64 units (the batch-1 latency path), 19..37 -> 4 slices
// (148 CTAs at 37 tiles), 38..49 -> 3 slices in ONE wave (the 256-unit form took two waves there: 20 -> 35 us per block),
// 50..74 -> 2, >= 75 tiles -> 1 slice = the whole hidden layer per CTA (pair). Every CTA writes its partial O (fp32, by TMA) into
// slot `slice` of the scratch tensor; mlp_reduce_kernel / mlp_reduce_post_kernel add the slots in slice order (deterministic) and
// apply b2 / layer-scale / residual / mask (+ what follows the block in the graph).
//
// Warps: 0 = TMA producer (a-tile K blocks interleaved with the first weight units, then weight units of 128 (64) rows x 64 K,
// hi + lo, through a 3-slot ring), 1 = MMA issuer, 2..9 = epilogue (TMEM lane quarter = warp % 4, column half = (warp - 2) / 4).
// The last chunk's O units run output-half-major, so the first 128 output columns are final (and are being stored) while the
// tensor pipe still accumulates the other 128.
//
// PAIR form (convnext_mlp_stream2_kernel, cta_group::2): the one-CTA form is paced by the shared-memory port (an S unit of 12 MMAs
// reads 96 KB of operands while TMA fills 32 KB: 1 000 cycles for 768 cycles of MMAs) and pulls 640 KB into each SM per block at
// 4 slices. Two row tiles that share a hidden slice run as a CTA pair with ONE instruction stream (M = 256 MMAs issued by the leader):
// each CTA keeps its own `a` tile, its own S / P / O in its own TMEM, and only HALF of every weight unit (64 of the 128 weight rows,
// 6-slot ring of 16 KB) — 384 KB per CTA, ~795 cycles per unit. The pair's barriers: TMA of both CTAs completes on the leader's full
// barriers, tcgen05.commit multicasts to both CTAs, the peer's epilogue warps arrive remotely on the leader's P barriers. An odd last
// row tile runs the one-CTA form inside the same launch (its cluster's two CTAs take two of its hidden slices), so 37 tiles x 4 slices
// is still 148 CTAs in one wave; where it does not add a wave it is padded to a pair instead (the phantom CTA computes on zero rows
// and stores nothing).
#pragma once
#include "mlp_tc.cuh"
#include "gemm2_tc.cuh"

namespace stc {
namespace mlp {

constexpr int ST_OFF_B1 = OFF_BAR + 512;                    // two b1 slices of 128 floats (chunk parity)
constexpr int ST_SMEM_BYTES = ST_OFF_B1 + 1024 + 1024;      // + alignment slack
static_assert(ST_SMEM_BYTES <= 232448, "shared memory budget");

struct StreamParams {
    int M;                  // rows
    int nslice;             // CTAs per 128-row tile (1..16)
    int npairs;             // pair kernel only: row-tile pairs; a row tile after them (odd count, not padded) runs the one-CTA form
    const float* b1;        // [H]
    long long* trace;       // debug (stc_debug_mlp with STC_MLP_TRACE=1): clock64() stamps of CTA 0
};

// mbarrier wait with a bound: a protocol error in this kernel must surface as a launch failure, not as a hung GPU
// (try_wait suspends the thread for a hardware-defined time slice per attempt, so the bound is seconds, never reached otherwise)
STC_DEVINL void mbar_wait_b(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && spins > (1u << 24)) __trap();
    }
}

// hidden range of slice c out of s: 64-unit blocks [c*16/s, (c+1)*16/s)
STC_DEVINL void stream_range(int c, int s, int& u0, int& u1) { u0 = (c * 16) / s; u1 = ((c + 1) * 16) / s; }

// The ring's unit order, shared by the producer and the MMA warp: f(kind, chunk, a, b) with kind 0 = S unit (a = K block of C),
// kind 1 = O unit (a = 64-unit sub-block of the chunk, b = output half).
template <typename F>
STC_DEVINL void stream_units(int nchunks, int nblk64, F&& f) {
    for (int kb = 0; kb < C / BK; ++kb) f(0, 0, kb, 0);
    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks)
            for (int kb = 0; kb < C / BK; ++kb) f(0, c + 1, kb, 0);
        const int nj = nblk64 - 2 * c >= 2 ? 2 : 1;
        if (c + 1 < nchunks) {
            for (int j = 0; j < nj; ++j)
                for (int nh = 0; nh < 2; ++nh) f(1, c, j, nh);
        } else {
            for (int nh = 0; nh < 2; ++nh)
                for (int j = 0; j < nj; ++j) f(1, c, j, nh);
        }
    }
}

STC_DEVINL void umma2_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

struct StreamMaps {
    const CUtensorMap *a_hi, *a_lo;          // [rows, C], box 128 rows x 64
    const CUtensorMap *w1_hi, *w1_lo;        // [H, C], box = this CTA's weight rows of a 128-unit chunk (128; pair: 64) x 64
    const CUtensorMap *w1_hi_s, *w1_lo_s;    // the same for a final 64-unit chunk (one-CTA form only)
    const CUtensorMap *w2_hi, *w2_lo;        // [C, H], box = this CTA's output rows of a half (128; pair: 64) x 64
    const CUtensorMap *part;                 // [nslice * tiles * 128, C] fp32, box 32 x 32
};

// One CTA's share of a block: row tile `tile`, hidden slice `slice`. PAIR: this CTA is rank `rank` of a CTA pair (cluster of two
// along x) that owns row tiles (tile - rank, tile - rank + 1); hidden slices are whole 128-unit chunks there (the eight chunks dealt out
// as evenly as possible: 3 slices = 2 + 3 + 3 chunks). The partition of the hidden layer may differ from tile to tile — the reduce kernel
// only adds a row's nslice partials — so an odd last tile keeps the one-CTA form's 64-unit granularity.
template <bool PAIR>
STC_DEVINL void stream_body(const StreamMaps& mp, const StreamParams& p, const int tile, const int slice, const int rank, uint8_t* smem_raw) {
    using namespace tc;
    constexpr int NS = PAIR ? 2 * SLOTS : SLOTS;            // ring slots
    constexpr int UNIT_B = PAIR ? KBLK : UNIT;              // bytes per slot: this CTA's weight rows x 64 K, hi + lo
    constexpr int HALF_B = UNIT_B / 2;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;        // same offset in both CTAs of a pair
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar = smem_base + OFF_BAR, bar2 = bar + 16u * NS;
    auto full_bar = [&](int s) { return bar + 8u * s; };
    auto empty_bar = [&](int s) { return bar + 8u * (NS + s); };
    auto bar_ak = [&](int kb) { return bar2 + 8u * kb; };                  // a-tile K block kb landed (pair: both CTAs')
    auto bar_s = [&](int buf) { return bar2 + 32 + 8u * buf; };            // S chunk complete in TMEM buffer buf
    auto bar_p = [&](int buf, int j) { return bar2 + 48 + 8u * (2 * buf + j); };  // P sub-block j of buffer buf written (pair: by both CTAs)
    auto bar_o = [&](int half) { return bar2 + 80 + 8u * half; };          // output half final
    const uint32_t tmem_slot = bar2 + 96;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + OFF_BAR + 16 * NS + 96);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = tile * BM;
    int u0, u1;
    if constexpr (PAIR) { u0 = 2 * ((slice * 8) / p.nslice); u1 = 2 * (((slice + 1) * 8) / p.nslice); }      // whole 128-unit chunks
    else stream_range(slice, p.nslice, u0, u1);
    const int nblk64 = u1 - u0, nchunks = (nblk64 + 1) >> 1, h0 = u0 * 64;
    auto chunk_w = [&](int c) { return nblk64 - 2 * c >= 2 ? 128 : 64; };
#define STC_STRACE(idx) do { if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[idx] = clock64(); } while (0)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(mp.a_hi); tma_prefetch_desc(mp.a_lo); tma_prefetch_desc(mp.w1_hi); tma_prefetch_desc(mp.w1_lo);
        if (!PAIR) { tma_prefetch_desc(mp.w1_hi_s); tma_prefetch_desc(mp.w1_lo_s); }
        tma_prefetch_desc(mp.w2_hi); tma_prefetch_desc(mp.w2_lo);
        tma_prefetch_desc(mp.part);
        for (int s = 0; s < NS; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int kb = 0; kb < C / BK; ++kb) mbar_init(bar_ak(kb), 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_s(b), 1); mbar_init(bar_p(b, 0), (PAIR ? 2 : 1) * EPI_WARPS); mbar_init(bar_p(b, 1), (PAIR ? 2 : 1) * EPI_WARPS); mbar_init(bar_o(b), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) { if constexpr (PAIR) tc2::tmem_alloc2(tmem_slot, 512); else tmem_alloc(tmem_slot, 512); }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();             // the peer's barriers exist before anything is signalled at them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    if (warp == 2) STC_STRACE(0);
    pdl_wait();

    if (warp == 0) {
        // ===== TMA producer (pair: both CTAs — own a-tile, own half of every weight unit; bytes complete on the LEADER's barriers) =====
        if (elect_one()) {
            auto load_a = [&](int kb) {
                const uint32_t d_hi = smem_base + OFF_X + kb * KBLK, d_lo = smem_base + OFF_X + (C / BK + kb) * KBLK;
                if constexpr (PAIR) {
                    if (rank == 0) mbar_expect_tx(bar_ak(kb), 4 * KBLK);
                    const uint32_t rb = tc2::mapa_rank(bar_ak(kb), 0);
                    tc2::tma_load_2d_2sm(d_hi, mp.a_hi, rb, kb * BK, m0);
                    tc2::tma_load_2d_2sm(d_lo, mp.a_lo, rb, kb * BK, m0);
                } else {
                    mbar_expect_tx(bar_ak(kb), 2 * KBLK);
                    tma_load_2d(d_hi, mp.a_hi, bar_ak(kb), kb * BK, m0);
                    tma_load_2d(d_lo, mp.a_lo, bar_ak(kb), kb * BK, m0);
                }
            };
            int u = 0;
            load_a(0);
            stream_units(nchunks, nblk64, [&](int kind, int c, int a, int b) {
                const int s = u % NS;
                mbar_wait_b(empty_bar(s), ((u / NS) & 1) ^ 1);
                const uint32_t dst = smem_base + OFF_RING + s * UNIT_B;
                if (p.trace && blockIdx.x == 0 && u < 16) p.trace[40 + u] = clock64();
                if constexpr (PAIR) {
                    if (rank == 0) mbar_expect_tx(full_bar(s), 2 * UNIT_B);
                    const uint32_t fb = tc2::mapa_rank(full_bar(s), 0);
                    if (kind == 0) {        // W1[this CTA's 64 hidden rows of chunk c, K block a of C]
                        const int row = h0 + c * 128 + rank * 64;
                        tc2::tma_load_2d_2sm(dst, mp.w1_hi, fb, a * BK, row);
                        tc2::tma_load_2d_2sm(dst + HALF_B, mp.w1_lo, fb, a * BK, row);
                    } else {                // W2[this CTA's 64 output rows of half b, hidden K block]
                        tc2::tma_load_2d_2sm(dst, mp.w2_hi, fb, h0 + c * 128 + a * BK, b * 128 + rank * 64);
                        tc2::tma_load_2d_2sm(dst + HALF_B, mp.w2_lo, fb, h0 + c * 128 + a * BK, b * 128 + rank * 64);
                    }
                } else {
                    const uint32_t fb = full_bar(s);
                    if (kind == 0) {        // W1[hidden rows of chunk c, K block a of C]
                        const int w = chunk_w(c), row = h0 + c * 128;
                        mbar_expect_tx(fb, 2 * w * BK * 2);
                        tma_load_2d(dst, w == 128 ? mp.w1_hi : mp.w1_hi_s, fb, a * BK, row);
                        tma_load_2d(dst + HALF_B, w == 128 ? mp.w1_lo : mp.w1_lo_s, fb, a * BK, row);
                    } else {                // W2[output rows b*128.., hidden K block]
                        mbar_expect_tx(fb, UNIT_B);
                        tma_load_2d(dst, mp.w2_hi, fb, h0 + c * 128 + a * BK, b * 128);
                        tma_load_2d(dst + HALF_B, mp.w2_lo, fb, h0 + c * 128 + a * BK, b * 128);
                    }
                }
                // the rest of the a-tile queues between the first weight units: the first MMAs need 32 KB of `a` + one unit
                if (u < C / BK - 1) load_a(u + 1);
                ++u;
            });
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer (pair: the leader's warp issues for both CTAs) =====
        if (!PAIR || rank == 0) {
            int u = 0;
            stream_units(nchunks, nblk64, [&](int kind, int c, int a, int b) {
                const int s = u % NS, buf = c & 1;
                const uint32_t par = (uint32_t)((c >> 1) & 1);
                if (kind == 0 && c == 0) mbar_wait_b(bar_ak(a), 0);                       // a-tile K block a has landed
                if (kind == 1 && (c + 1 < nchunks ? b == 0 : true)) {
                    // P sub-block a of chunk c is in TMEM (non-last chunks reach it first with b == 0; the last chunk's half-major
                    // order reaches every sub-block once per half — a second wait on a completed phase returns at once)
                    mbar_wait_b(bar_p(buf, a), par);
                }
                mbar_wait_b(full_bar(s), (u / NS) & 1);
                tc_fence_after();
                if (p.trace && blockIdx.x == 0 && lane == 0 && u < 16) p.trace[8 + u] = clock64();
                if (elect_one()) {
                    const uint32_t st = smem_base + OFF_RING + s * UNIT_B;
                    const uint64_t w_hi = make_smem_desc(st), w_lo = make_smem_desc(st + HALF_B);
                    auto commit = [&](uint32_t b_) { if constexpr (PAIR) tc2::umma2_commit(b_); else umma_commit(b_); };
                    if (kind == 0) {
                        const uint32_t xk = smem_base + OFF_X + a * KBLK;
                        const uint64_t a_hi = make_smem_desc(xk), a_lo = make_smem_desc(xk + (C / BK) * KBLK);
                        const uint32_t d = tmem_base + buf * 128;
                        const uint32_t idesc = PAIR ? make_idesc_bf16(2 * BM, 128) : chunk_w(c) == 128 ? make_idesc_bf16(BM, 128) : make_idesc_bf16(BM, 64);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                            if constexpr (PAIR) {
                                tc2::umma2_bf16(d, a_lo + adv, w_hi + adv, idesc, (a | k) != 0);
                                tc2::umma2_bf16(d, a_hi + adv, w_lo + adv, idesc, 1);
                                tc2::umma2_bf16(d, a_hi + adv, w_hi + adv, idesc, 1);
                            } else {
                                umma_bf16(d, a_lo + adv, w_hi + adv, idesc, (a | k) != 0);
                                umma_bf16(d, a_hi + adv, w_lo + adv, idesc, 1);
                                umma_bf16(d, a_hi + adv, w_hi + adv, idesc, 1);
                            }
                        }
                        commit(empty_bar(s));
                        if (a == C / BK - 1) commit(bar_s(buf));                         // S chunk c complete
                    } else {
                        constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 2 * BM : BM, 128);
                        const uint32_t d = tmem_base + 256 + b * 128;
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                            const uint32_t p_hi = tmem_base + buf * 128 + a * BK + k * UMMA_K, p_lo = p_hi + UMMA_K / 2;
                            if constexpr (PAIR) {
                                umma2_bf16_ts(d, p_lo, w_hi + adv, idesc, (c | a | k) != 0);
                                umma2_bf16_ts(d, p_hi, w_lo + adv, idesc, 1);
                                umma2_bf16_ts(d, p_hi, w_hi + adv, idesc, 1);
                            } else {
                                umma_bf16_ts(d, p_lo, w_hi + adv, idesc, (c | a | k) != 0);
                                umma_bf16_ts(d, p_hi, w_lo + adv, idesc, 1);
                                umma_bf16_ts(d, p_hi, w_hi + adv, idesc, 1);
                            }
                        }
                        commit(empty_bar(s));
                        if (c + 1 == nchunks) {
                            const int nj = nblk64 - 2 * c >= 2 ? 2 : 1;
                            if (a == nj - 1) commit(bar_o(b));                           // output half b is final
                        }
                    }
                }
                __syncwarp();
                ++u;
            });
        }
        pdl_trigger_late();                  // every MMA of this CTA (pair) is issued: what is left is the drain and the store of the partials
    } else {
        // ===== epilogue 1, per chunk: P = split(GELU(S + b1)), in place in TMEM =====
        constexpr int PARTS = EPI_WARPS / 4;                                      // column parts per TMEM lane quarter
        const int q = warp & 3, part = (warp - 2) >> 2;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        float* b1s = reinterpret_cast<float*>(smem_gen + ST_OFF_B1);
        const int et = (int)threadIdx.x - 64;                                    // 0 .. 32 * EPI_WARPS - 1
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
            const int buf = c & 1, w = chunk_w(c);
            // this chunk's b1 slice -> shared memory while S is still being accumulated (no L1 next to 226 KB of shared memory: a
            // __ldg inside the loop below would be an L2 round trip on the CTA's serial chain). Slot buf was last read two chunks ago.
            if (et < w) b1s[buf * 128 + et] = __ldg(p.b1 + h0 + c * 128 + et);
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");       // the epilogue warps only
            mbar_wait_b(bar_s(buf), (uint32_t)((c >> 1) & 1));
            tc_fence_after();
            if (warp == 2) STC_STRACE(24 + (c < 4 ? c : 3));
#pragma unroll 1
            for (int j = 0; j < w / BK; ++j) {
#pragma unroll
                for (int g = 0; g < BK / (16 * PARTS); ++g) {
                    const int col = j * BK + (part * (BK / (16 * PARTS)) + g) * 16;
                    uint32_t v[16], o[16];
                    __syncwarp();
                    tmem_ld16(trow + buf * 128 + col, v);
                    const float* b1 = b1s + buf * 128 + col;
#pragma unroll
                    for (int t = 0; t < 8; t += 2) {
                        const float4 bv = *reinterpret_cast<const float4*>(b1 + 2 * t);                  // broadcast
                        const float2 g0 = gelu_erf_mufu2(__fadd2_rn(make_float2(__uint_as_float(v[2 * t]), __uint_as_float(v[2 * t + 1])), make_float2(bv.x, bv.y)));
                        const float2 g1 = gelu_erf_mufu2(__fadd2_rn(make_float2(__uint_as_float(v[2 * t + 2]), __uint_as_float(v[2 * t + 3])), make_float2(bv.z, bv.w)));
                        split_pair(g0.x, g0.y, o[t], o[8 + t]);
                        split_pair(g1.x, g1.y, o[t + 1], o[8 + t + 1]);
                    }
                    tmem_st16(trow + buf * 128 + col, o);
                }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (PAIR) tc2::mbar_arrive_cluster(tc2::mapa_rank(bar_p(buf, j), 0));      // the leader's MMA warp waits for both CTAs
                    else mbar_arrive(bar_p(buf, j));
                }
            }
            if (warp == 2) STC_STRACE(28 + (c < 4 ? c : 3));
        }
        // ===== partial O (256 columns) -> global scratch by TMA, output half by output half: 32 rows x 32 columns (one 128-byte
        //       swizzled row per lane) per store, two staging buffers per warp (in the dead a-tile) so that the copy of one box
        //       overlaps the TMEM read of the next =====
        const uint32_t stg = smem_base + OFF_X + (uint32_t)(warp - 2) * 8192u;
        const int mpad = ((p.M + BM - 1) / BM) * BM;
        const int grow = slice * mpad + m0 + q * 32;
        const bool phantom = m0 >= mpad;         // pair form, odd tile count padded to a whole pair: the second CTA computes on zero rows and stores nothing
        int it = 0;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            mbar_wait_b(bar_o(half), 0);         // every MMA up to the last one of this half has retired (so has every read of the a-tile)
            tc_fence_after();
            if (warp == 2) STC_STRACE(32 + half);
#pragma unroll 1
            for (int cc = 0; cc < 128 / PARTS; cc += 32, ++it) {
                const int col = half * 128 + part * (128 / PARTS) + cc;
                const uint32_t buf = stg + (uint32_t)(it & 1) * 4096u;
                if (it >= 2) { if (lane == 0) bulk_wait_read<1>(); __syncwarp(); }        // the store that last read this buffer has drained
                uint32_t v[32];
                tmem_ld32(trow + 256 + col, v);
#pragma unroll
                for (int ch = 0; ch < 8; ++ch)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf + (uint32_t)lane * 128u + (uint32_t)((ch ^ (lane & 7)) * 16)),
                                 "r"(v[4 * ch]), "r"(v[4 * ch + 1]), "r"(v[4 * ch + 2]), "r"(v[4 * ch + 3]) : "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0 && !phantom) { tma_store_2d(mp.part, buf, col, grow); bulk_commit(); }
            }
        }
        if (warp == 2) STC_STRACE(35);
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
        if (warp == 2) STC_STRACE(36);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) STC_STRACE(37);
    if constexpr (PAIR) cluster_sync_all();             // neither CTA leaves (or frees TMEM) while the pair still computes / signals
    tc_fence_after();
    if (warp == 2) STC_STRACE(34);
    if (warp == 1) { if constexpr (PAIR) tc2::tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
#undef STC_STRACE
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
convnext_mlp_stream_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                           const __grid_constant__ CUtensorMap map_w1_hi, const __grid_constant__ CUtensorMap map_w1_lo,
                           const __grid_constant__ CUtensorMap map_w1_hi64, const __grid_constant__ CUtensorMap map_w1_lo64,
                           const __grid_constant__ CUtensorMap map_w2_hi, const __grid_constant__ CUtensorMap map_w2_lo,
                           const __grid_constant__ CUtensorMap map_part, const StreamParams p) {
    extern __shared__ uint8_t smem_raw[];
    const StreamMaps mp{&map_a_hi, &map_a_lo, &map_w1_hi, &map_w1_lo, &map_w1_hi64, &map_w1_lo64, &map_w2_hi, &map_w2_lo, &map_part};
    stream_body<false>(mp, p, (int)(blockIdx.x / p.nslice), (int)(blockIdx.x % p.nslice), 0, smem_raw);
}

// Pair form: clusters [0, npairs * nslice) are CTA pairs (row tiles 2 i, 2 i + 1; hidden slice = cluster % nslice); the clusters after
// them are the odd last tile in the one-CTA form, two hidden slices per cluster. nslice <= 8 (a pair's slice is at least one 128-unit chunk).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
convnext_mlp_stream2_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                            const __grid_constant__ CUtensorMap map_w1_hi, const __grid_constant__ CUtensorMap map_w1_lo,
                            const __grid_constant__ CUtensorMap map_w1_hi64, const __grid_constant__ CUtensorMap map_w1_lo64,
                            const __grid_constant__ CUtensorMap map_w2_hi, const __grid_constant__ CUtensorMap map_w2_lo,
                            const __grid_constant__ CUtensorMap map_w2_hi64, const __grid_constant__ CUtensorMap map_w2_lo64,
                            const __grid_constant__ CUtensorMap map_part, const StreamParams p) {
    extern __shared__ uint8_t smem_raw[];
    const int cl = (int)tc::cluster_id_x(), rank = (int)tc::cluster_ctarank(), npc = p.npairs * p.nslice;
    if (cl < npc) {
        const StreamMaps mp{&map_a_hi, &map_a_lo, &map_w1_hi64, &map_w1_lo64, nullptr, nullptr, &map_w2_hi64, &map_w2_lo64, &map_part};
        stream_body<true>(mp, p, 2 * (cl / p.nslice) + rank, cl % p.nslice, rank, smem_raw);
    } else {
        const int e = (cl - npc) * 2 + rank;
        if (e >= p.nslice) return;
        const StreamMaps mp{&map_a_hi, &map_a_lo, &map_w1_hi, &map_w1_lo, &map_w1_hi64, &map_w1_lo64, &map_w2_hi, &map_w2_lo, &map_part};
        stream_body<false>(mp, p, 2 * p.npairs, e, 0, smem_raw);
    }
}

}  // namespace mlp
}  // namespace stc
