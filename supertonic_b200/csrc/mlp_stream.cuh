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
// the epilogue warps in steady state. The host picks nslice = floor(SMs / row tiles) (clamped to 1..16): 37 tiles -> 4 slices of 256
// units (148 CTAs), 38..49 tiles -> 3 slices of 320/320/384 in ONE wave (the 256-unit form took two waves there: 20 -> 35 us per
// block), 50..74 -> 2, >= 75 tiles -> 1 slice = the whole hidden layer per CTA (four times fewer CTAs than before), <= 9 tiles ->
// 16 slices of 64 (the batch-1 latency path). Every CTA writes its partial O (fp32, by TMA) into slot `slice` of the scratch
// tensor; mlp_reduce_kernel / mlp_reduce_post_kernel add the slots in slice order (deterministic) and apply b2 / layer-scale /
// residual / mask (+ what follows the block in the graph).
//
// Warps: 0 = TMA producer (a-tile K blocks interleaved with the first weight units, then weight units of 128 (64) rows x 64 K,
// hi + lo, through a 3-slot ring), 1 = MMA issuer, 2..9 = epilogue (TMEM lane quarter = warp % 4, column half = (warp - 2) / 4).
// The last chunk's O units run output-half-major, so the first 128 output columns are final (and are being stored) while the
// tensor pipe still accumulates the other 128.
#pragma once
#include "mlp_tc.cuh"

namespace stc {
namespace mlp {

constexpr int ST_OFF_B1 = OFF_BAR + 512;                    // two b1 slices of 128 floats (chunk parity)
constexpr int ST_SMEM_BYTES = ST_OFF_B1 + 1024 + 1024;      // + alignment slack
static_assert(ST_SMEM_BYTES <= 232448, "shared memory budget");

struct StreamParams {
    int M;                  // rows
    int nslice;             // CTAs per 128-row tile (1..16)
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

__global__ void __launch_bounds__(NUM_THREADS, 1)
convnext_mlp_stream_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                           const __grid_constant__ CUtensorMap map_w1_hi, const __grid_constant__ CUtensorMap map_w1_lo,
                           const __grid_constant__ CUtensorMap map_w1_hi64, const __grid_constant__ CUtensorMap map_w1_lo64,
                           const __grid_constant__ CUtensorMap map_w2_hi, const __grid_constant__ CUtensorMap map_w2_lo,
                           const __grid_constant__ CUtensorMap map_part, const StreamParams p) {
    using namespace tc;
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar = smem_base + OFF_BAR;
    auto full_bar = [&](int s) { return bar + 8u * s; };                   // 0..2
    auto empty_bar = [&](int s) { return bar + 24 + 8u * s; };             // 3..5
    auto bar_ak = [&](int kb) { return bar + 48 + 8u * kb; };              // a-tile K block kb landed
    auto bar_s = [&](int buf) { return bar + 80 + 8u * buf; };             // S chunk complete in TMEM buffer buf
    auto bar_p = [&](int buf, int j) { return bar + 96 + 8u * (2 * buf + j); };   // P sub-block j of buffer buf written
    auto bar_o = [&](int half) { return bar + 128 + 8u * half; };          // output half final
    const uint32_t tmem_slot = bar + 144;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + OFF_BAR + 144);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slice = (int)(blockIdx.x % p.nslice);
    const int m0 = (int)(blockIdx.x / p.nslice) * BM;
    int u0, u1;
    stream_range(slice, p.nslice, u0, u1);
    const int nblk64 = u1 - u0, nchunks = (nblk64 + 1) >> 1, h0 = u0 * 64;
    auto chunk_w = [&](int c) { return nblk64 - 2 * c >= 2 ? 128 : 64; };
#define STC_STRACE(idx) do { if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[idx] = clock64(); } while (0)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_w1_hi); tma_prefetch_desc(&map_w1_lo);
        tma_prefetch_desc(&map_w1_hi64); tma_prefetch_desc(&map_w1_lo64); tma_prefetch_desc(&map_w2_hi); tma_prefetch_desc(&map_w2_lo);
        tma_prefetch_desc(&map_part);
        for (int s = 0; s < SLOTS; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int kb = 0; kb < C / BK; ++kb) mbar_init(bar_ak(kb), 1);
        for (int b = 0; b < 2; ++b) { mbar_init(bar_s(b), 1); mbar_init(bar_p(b, 0), 8); mbar_init(bar_p(b, 1), 8); mbar_init(bar_o(b), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    if (warp == 2) STC_STRACE(0);
    pdl_wait();

    if (warp == 0) {
        if (elect_one()) {
            auto load_a = [&](int kb) {
                mbar_expect_tx(bar_ak(kb), 2 * KBLK);
                tma_load_2d(smem_base + OFF_X + kb * KBLK, &map_a_hi, bar_ak(kb), kb * BK, m0);
                tma_load_2d(smem_base + OFF_X + (C / BK + kb) * KBLK, &map_a_lo, bar_ak(kb), kb * BK, m0);
            };
            int u = 0;
            load_a(0);
            stream_units(nchunks, nblk64, [&](int kind, int c, int a, int b) {
                const int s = u % SLOTS;
                mbar_wait_b(empty_bar(s), ((u / SLOTS) & 1) ^ 1);
                const uint32_t dst = smem_base + OFF_RING + s * UNIT, fb = full_bar(s);
                if (p.trace && blockIdx.x == 0 && u < 16) p.trace[40 + u] = clock64();
                if (kind == 0) {            // W1[hidden rows of chunk c, K block a of C]
                    const int w = chunk_w(c), row = h0 + c * 128;
                    mbar_expect_tx(fb, 2 * w * BK * 2);
                    tma_load_2d(dst, w == 128 ? &map_w1_hi : &map_w1_hi64, fb, a * BK, row);
                    tma_load_2d(dst + KBLK, w == 128 ? &map_w1_lo : &map_w1_lo64, fb, a * BK, row);
                } else {                    // W2[output rows b*128.., hidden K block]
                    mbar_expect_tx(fb, UNIT);
                    tma_load_2d(dst, &map_w2_hi, fb, h0 + c * 128 + a * BK, b * 128);
                    tma_load_2d(dst + KBLK, &map_w2_lo, fb, h0 + c * 128 + a * BK, b * 128);
                }
                // the rest of the a-tile queues between the first weight units: the first MMAs need 32 KB of `a` + one unit
                if (u < C / BK - 1) load_a(u + 1);
                ++u;
            });
        }
        __syncwarp();
    } else if (warp == 1) {
        int u = 0;
        stream_units(nchunks, nblk64, [&](int kind, int c, int a, int b) {
            const int s = u % SLOTS, buf = c & 1;
            const uint32_t par = (uint32_t)((c >> 1) & 1);
            if (kind == 0 && c == 0) mbar_wait_b(bar_ak(a), 0);                       // a-tile K block a has landed
            if (kind == 1 && (c + 1 < nchunks ? b == 0 : true)) {
                // P sub-block a of chunk c is in TMEM (non-last chunks reach it first with b == 0; the last chunk's half-major
                // order reaches every sub-block once per half — a second wait on a completed phase returns at once)
                mbar_wait_b(bar_p(buf, a), par);
            }
            mbar_wait_b(full_bar(s), (u / SLOTS) & 1);
            tc_fence_after();
            if (p.trace && blockIdx.x == 0 && lane == 0 && u < 16) p.trace[8 + u] = clock64();
            if (elect_one()) {
                const uint32_t st = smem_base + OFF_RING + s * UNIT;
                const uint64_t w_hi = make_smem_desc(st), w_lo = make_smem_desc(st + KBLK);
                if (kind == 0) {
                    const uint32_t xk = smem_base + OFF_X + a * KBLK;
                    const uint64_t a_hi = make_smem_desc(xk), a_lo = make_smem_desc(xk + (C / BK) * KBLK);
                    const uint32_t d = tmem_base + buf * 128;
                    const uint32_t idesc = chunk_w(c) == 128 ? make_idesc_bf16(BM, 128) : make_idesc_bf16(BM, 64);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                        umma_bf16(d, a_lo + adv, w_hi + adv, idesc, (a | k) != 0);
                        umma_bf16(d, a_hi + adv, w_lo + adv, idesc, 1);
                        umma_bf16(d, a_hi + adv, w_hi + adv, idesc, 1);
                    }
                    umma_commit(empty_bar(s));
                    if (a == C / BK - 1) umma_commit(bar_s(buf));                    // S chunk c complete
                } else {
                    constexpr uint32_t idesc = make_idesc_bf16(BM, 128);
                    const uint32_t d = tmem_base + 256 + b * 128;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                        const uint32_t p_hi = tmem_base + buf * 128 + a * BK + k * UMMA_K, p_lo = p_hi + UMMA_K / 2;
                        umma_bf16_ts(d, p_lo, w_hi + adv, idesc, (c | a | k) != 0);
                        umma_bf16_ts(d, p_hi, w_lo + adv, idesc, 1);
                        umma_bf16_ts(d, p_hi, w_hi + adv, idesc, 1);
                    }
                    umma_commit(empty_bar(s));
                    if (c + 1 == nchunks) {
                        const int nj = nblk64 - 2 * c >= 2 ? 2 : 1;
                        if (a == nj - 1) umma_commit(bar_o(b));                      // output half b is final
                    }
                }
            }
            __syncwarp();
            ++u;
        });
    } else {
        // ===== epilogue 1, per chunk: P = split(GELU(S + b1)), in place in TMEM =====
        const int q = warp & 3, part = (warp - 2) >> 2;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        float* b1s = reinterpret_cast<float*>(smem_gen + ST_OFF_B1);
        const int et = (int)threadIdx.x - 64;                                    // 0..255
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
            const int buf = c & 1, w = chunk_w(c);
            // this chunk's b1 slice -> shared memory while S is still being accumulated (no L1 next to 226 KB of shared memory: a
            // __ldg inside the loop below would be an L2 round trip on the CTA's serial chain). Slot buf was last read two chunks ago.
            if (et < w) b1s[buf * 128 + et] = __ldg(p.b1 + h0 + c * 128 + et);
            asm volatile("bar.sync 1, 256;" ::: "memory");                       // the eight epilogue warps only
            mbar_wait_b(bar_s(buf), (uint32_t)((c >> 1) & 1));
            tc_fence_after();
            if (warp == 2) STC_STRACE(24 + (c < 4 ? c : 3));
#pragma unroll 1
            for (int j = 0; j < w / BK; ++j) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const int col = j * BK + part * 32 + g * 16;
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
                if (lane == 0) mbar_arrive(bar_p(buf, j));
            }
            if (warp == 2) STC_STRACE(28 + (c < 4 ? c : 3));
        }
        // ===== partial O (256 columns) -> global scratch by TMA, output half by output half: 32 rows x 32 columns (one 128-byte
        //       swizzled row per lane) per store, two staging buffers per warp (in the dead a-tile) so that the copy of one box
        //       overlaps the TMEM read of the next =====
        const uint32_t stg = smem_base + OFF_X + (uint32_t)(warp - 2) * 8192u;
        const int mpad = ((p.M + BM - 1) / BM) * BM;
        const int grow = slice * mpad + m0 + q * 32;
        int it = 0;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            mbar_wait_b(bar_o(half), 0);         // every MMA up to the last one of this half has retired (so has every read of the a-tile)
            tc_fence_after();
            if (warp == 2) STC_STRACE(32 + half);
#pragma unroll 1
            for (int cc = 0; cc < 64; cc += 32, ++it) {
                const int col = half * 128 + part * 64 + cc;
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
                if (lane == 0) { tma_store_2d(&map_part, buf, col, grow); bulk_commit(); }
            }
        }
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2) STC_STRACE(34);
    if (warp == 1) tmem_dealloc(tmem_base, 512);
#undef STC_STRACE
}

}  // namespace mlp
}  // namespace stc
