// Shared pieces of the fused ConvNeXt MLP (C = 256, H = 1024 blocks of the vector estimator and the text encoder, sm_100a):
// tile constants and shared-memory layout, the TMEM-operand MMA / tcgen05.st / TMA-store wrappers, and the reduce kernels that
// finish a block from the hidden-slice partial outputs. The tensor-core kernel itself is mlp_stream.cuh.
//
// History (measurements in DESIGN.md §5): a 4-CTA-cluster form with a DSMEM reduction, a 256-unit "split" form (+ N = 256 MMAs), "thin"
// 128 / 64-unit forms for small batches, a TMEM-operand "TS" form, an in-kernel producer of LayerNorm(dwconv(x)) and an in-kernel
// reduce were all built and measured in round 1; the stream form replaces the ones that won and the rest lost — none of them is
// in the tree any more.
#pragma once
#include "gemm_tc.cuh"

namespace stc {
namespace mlp {

constexpr int C = 256, H = 1024;                  // channels, hidden units
constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int EPI_WARPS = 8;                      // 16 warps were measured: the GELU of a 128-unit chunk takes 2 800 cycles either way (MUFU + issue bound:
                                                  // 2 MUFU and ~22 issue slots per element), profiles/r2t_mlp_sweep_epi16.txt
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int KBLK = BM * BK * 2;                 // 16 KB: 128 rows x 64 bf16, one half (hi or lo)
constexpr int X_BYTES = 2 * (C / BK) * KBLK;      // 128 KB: a-tile (hi k-blocks 0..3, lo k-blocks 0..3); later the store staging
constexpr int UNIT = 2 * KBLK;                    // 32 KB: 128 weight rows x 64 K, hi + lo
constexpr int SLOTS = 3;
constexpr int OFF_X = 0, OFF_RING = X_BYTES, OFF_BAR = OFF_RING + SLOTS * UNIT;

STC_DEVINL void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
STC_DEVINL void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
STC_DEVINL void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

STC_DEVINL void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
STC_DEVINL void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> STC_DEVINL void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// x <- ((p0 + p1 + p2 + p3 + b2) * gamma + x) * mask : the four hidden-slice partials in rank order (deterministic).
// One warp per row (lane owns 8 consecutive channels), so what FOLLOWS the block in the graphs can ride along instead of
// costing a launch of its own:
//   add_vec : x <- (x + add_vec) * mask           (time conditioning after a ConvNeXt block, cpp-side: vector_estimator graph)
//   out     : the next layer's tensor-core operand — LayerNorm(x) (pre-LN of an attention layer; ln_g / ln_b) or x itself
//             (ln_g == null; input of an output projection) as split-bf16
// WIDE (the thin forms: 8 or 16 hidden slices, small batches): the partial loads of up to eight slices are in flight together —
// those launches wait for L2 round trips, not for bandwidth (batch-1 latency 5.22 -> 4.96 ms); with four slices and 4 736 rows the
// extra registers cost occupancy instead (10.40 -> 10.86 ms/step), so the two-deep loop stays there.
template <bool WIDE>
__global__ void __launch_bounds__(256)
mlp_reduce_post_kernel(const float* __restrict__ partial, size_t slice, const float* __restrict__ b2, const float* __restrict__ gamma,
                  const float* __restrict__ mask, float* __restrict__ x, int M, const float* __restrict__ add_vec,
                  const float* __restrict__ ln_g, const float* __restrict__ ln_b, float eps,
                  __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, int nslice) {
    pdl_wait(); pdl_trigger_light();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + warp;
    if (row >= M) return;
    const int c0 = lane * 8;
    const size_t i = (size_t)row * C + c0;
    float y[8];
    {
        const float4 a0 = *reinterpret_cast<const float4*>(partial + i), a1 = *reinterpret_cast<const float4*>(partial + i + 4);
        y[0] = a0.x; y[1] = a0.y; y[2] = a0.z; y[3] = a0.w; y[4] = a1.x; y[5] = a1.y; y[6] = a1.z; y[7] = a1.w;
    }
    if constexpr (!WIDE) {
#pragma unroll 4
        for (int s = 1; s < nslice; ++s) {
            const float4 v0 = *reinterpret_cast<const float4*>(partial + s * slice + i), v1 = *reinterpret_cast<const float4*>(partial + s * slice + i + 4);
            y[0] += v0.x; y[1] += v0.y; y[2] += v0.z; y[3] += v0.w; y[4] += v1.x; y[5] += v1.y; y[6] += v1.z; y[7] += v1.w;
        }
    } else
    for (int s0 = 1; s0 < nslice; s0 += 8) {           // eight slices' loads in flight at a time, summed in slice order
        float4 q0[8], q1[8];
#pragma unroll
        for (int s = 0; s < 8; ++s)
            if (s0 + s < nslice) {
                q0[s] = *reinterpret_cast<const float4*>(partial + (s0 + s) * slice + i);
                q1[s] = *reinterpret_cast<const float4*>(partial + (s0 + s) * slice + i + 4);
            }
#pragma unroll
        for (int s = 0; s < 8; ++s)
            if (s0 + s < nslice) {
                y[0] += q0[s].x; y[1] += q0[s].y; y[2] += q0[s].z; y[3] += q0[s].w; y[4] += q1[s].x; y[5] += q1[s].y; y[6] += q1[s].z; y[7] += q1[s].w;
            }
    }
    const float mk = mask ? __ldg(mask + row) : 1.f;
    float bb[8], gg[8], rr[8];
    *reinterpret_cast<float4*>(bb) = __ldg(reinterpret_cast<const float4*>(b2 + c0)); *reinterpret_cast<float4*>(bb + 4) = __ldg(reinterpret_cast<const float4*>(b2 + c0 + 4));
    *reinterpret_cast<float4*>(gg) = __ldg(reinterpret_cast<const float4*>(gamma + c0)); *reinterpret_cast<float4*>(gg + 4) = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    *reinterpret_cast<float4*>(rr) = *reinterpret_cast<const float4*>(x + i); *reinterpret_cast<float4*>(rr + 4) = *reinterpret_cast<const float4*>(x + i + 4);
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = ((y[j] + bb[j]) * gg[j] + rr[j]) * mk;
    if (add_vec) {
        float tt[8];
        *reinterpret_cast<float4*>(tt) = __ldg(reinterpret_cast<const float4*>(add_vec + c0)); *reinterpret_cast<float4*>(tt + 4) = __ldg(reinterpret_cast<const float4*>(add_vec + c0 + 4));
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = (y[j] + tt[j]) * mk;
    }
    *reinterpret_cast<float4*>(x + i) = make_float4(y[0], y[1], y[2], y[3]);
    *reinterpret_cast<float4*>(x + i + 4) = make_float4(y[4], y[5], y[6], y[7]);
    if (!out_hi) return;
    if (ln_g) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += y[j];
        const float mean = warp_sum<float>(s) / (float)C;
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { y[j] -= mean; v += y[j] * y[j]; }
        const float inv = 1.0f / sqrtf(warp_sum<float>(v) / (float)C + eps);
        float g8[8], h8[8];
        *reinterpret_cast<float4*>(g8) = __ldg(reinterpret_cast<const float4*>(ln_g + c0)); *reinterpret_cast<float4*>(g8 + 4) = __ldg(reinterpret_cast<const float4*>(ln_g + c0 + 4));
        *reinterpret_cast<float4*>(h8) = __ldg(reinterpret_cast<const float4*>(ln_b + c0)); *reinterpret_cast<float4*>(h8 + 4) = __ldg(reinterpret_cast<const float4*>(ln_b + c0 + 4));
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = y[j] * inv * g8[j] + h8[j];
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) tc::split_pair(y[2 * t], y[2 * t + 1], hi[t], lo[t]);
    *reinterpret_cast<uint4*>(out_hi + i) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(out_lo + i) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// The plain reduce (no post-ops): one float4 per thread — twice the threads of the row-wise kernel above and ~3 us faster
// (4.2 vs 7.0 us at 4 736 rows, profiles/r1w_mlp_ncu_full_summary.txt), so blocks with nothing folded in keep this form.
template <bool WIDE>
__global__ void __launch_bounds__(256)
mlp_reduce_kernel(const float* __restrict__ partial, size_t slice, const float* __restrict__ b2, const float* __restrict__ gamma,
                  const float* __restrict__ mask, float* __restrict__ x, int M, int nslice) {
    pdl_wait(); pdl_trigger_light();
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= (size_t)M * C) return;
    const int row = (int)(i / C), col = (int)(i % C);
    float4 acc;
    if constexpr (WIDE) {
        // every partial's load is in flight before the first add (one L2 round trip for up to 16 hidden slices), summed in slice order
        float4 pv[16];
#pragma unroll
        for (int s = 0; s < 16; ++s) if (s < nslice) pv[s] = *reinterpret_cast<const float4*>(partial + s * slice + i);
        acc = pv[0];
#pragma unroll
        for (int s = 1; s < 16; ++s) if (s < nslice) { acc.x += pv[s].x; acc.y += pv[s].y; acc.z += pv[s].z; acc.w += pv[s].w; }
    } else {
        acc = *reinterpret_cast<const float4*>(partial + i);
#pragma unroll 4
        for (int s = 1; s < nslice; ++s) {
            const float4 v = *reinterpret_cast<const float4*>(partial + s * slice + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    const float4 b = __ldg(reinterpret_cast<const float4*>(b2 + col)), g = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 r = *reinterpret_cast<const float4*>(x + i);
    const float mk = mask ? __ldg(mask + row) : 1.f;
    acc.x = ((acc.x + b.x) * g.x + r.x) * mk; acc.y = ((acc.y + b.y) * g.y + r.y) * mk;
    acc.z = ((acc.z + b.z) * g.z + r.z) * mk; acc.w = ((acc.w + b.w) * g.w + r.w) * mk;
    *reinterpret_cast<float4*>(x + i) = acc;
}

}  // namespace mlp
}  // namespace stc
