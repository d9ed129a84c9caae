// Bandwidth / CUDA-core kernels of libsupertonic_cuda (sm_100a).
//
// Activation layout everywhere: channels-last rows, x[(b*N + n)*C + c]  (N = frames or tokens);
// masks are flat float [B*N] (the reference's [B,1,N] tensors are exactly that in memory,
// cpp/helper.cpp:740-757). Templated on T = float (TE / VE / vocoder) or double (duration
// predictor, evaluated in fp64 so the integer frame counts derived from it are reproducible —
// DESIGN.md "bit-exact durations").
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace stc {

#define STC_DEVINL __device__ __forceinline__

// Programmatic dependent launch (env STC_PDL, model.cu launch_k): with the launch attribute set a kernel may START while its
// predecessor still runs; pdl_wait() blocks until the predecessor grid has completed and its writes are visible. Without the
// attribute both instructions are no-ops. The rules every kernel of the library follows:
//   * nothing a predecessor may still be writing is read, and nothing it may still be reading is written, before pdl_wait();
//   * a kernel releases its dependents only AFTER its own pdl_wait() — light kernels right after it (pdl_trigger_light), the
//     tensor-core kernels when a CTA has issued its last MMA (pdl_trigger_late) — so while kernel n runs its pre-wait code, kernel
//     n-1 has passed its wait, i.e. kernel n-2 and everything before it are complete;
//   * hence pre-wait code may read what was produced at least TWO kernels earlier (weights, sequence offsets), never activations.
// c_pdl_mode: 2 = as described (default), 3 = nobody releases early (dependents start when the last CTA exits).
__constant__ int c_pdl_mode;
STC_DEVINL void pdl_trigger_light() { if (c_pdl_mode == 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
STC_DEVINL void pdl_trigger_late() { if (c_pdl_mode == 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
STC_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename T> STC_DEVINL T t_erf(T x);
template <> STC_DEVINL float t_erf<float>(float x) { return erff(x); }
template <> STC_DEVINL double t_erf<double>(double x) { return erf(x); }
template <typename T> STC_DEVINL T t_sqrt(T x);
template <> STC_DEVINL float t_sqrt<float>(float x) { return sqrtf(x); }
template <> STC_DEVINL double t_sqrt<double>(double x) { return sqrt(x); }
template <typename T> STC_DEVINL T t_exp(T x);
template <> STC_DEVINL float t_exp<float>(float x) { return expf(x); }
template <> STC_DEVINL double t_exp<double>(double x) { return exp(x); }

// exact (erf) GELU in the operation order the graphs use: (x * (erf(x / sqrt2) + 1)) * 0.5
template <typename T> STC_DEVINL T gelu_erf(T x) {
    const T rsq2 = (T)1.41421354f;   // float32(sqrt(2)) as stored in the graph
    return (x * (t_erf<T>(x / rsq2) + (T)1)) * (T)0.5;
}

// float path used inside the tensor-core epilogue: multiply by 1/sqrt2 instead of the IEEE division (which
// compiles to ~100 instructions); differs from the graph's Div by at most 1 ulp of the erf argument.
STC_DEVINL float gelu_erf_fast(float x) { return (x * (erff(x * 0.70710678f) + 1.0f)) * 0.5f; }

template <typename T> STC_DEVINL T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
STC_DEVINL float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- packed sequences ------------------------------------------------------------------------
// B variable-length sequences are stored back to back: rows of sequence b are [off[b], off[b+1]).
// A padded rectangle [B,N] is the special case off[b] = b*N (plus a 0/1 row mask). Rows >= off[B]
// (bucket padding of a captured CUDA graph) belong to no sequence.
STC_DEVINL int find_seq(const int* __restrict__ off, int B, int row) {
    if (row >= __ldg(off + B)) return -1;
    int lo = 0, hi = B;                 // invariant: off[lo] <= row < off[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= row) lo = mid; else hi = mid;
    }
    return lo;
}

// ---- GEMM operand stores ---------------------------------------------------------------------
// Operands of the tensor-core GEMMs are kept as split bf16 pairs: v ~= hi + lo, hi = bf16(v),
// lo = bf16(v - hi) (16 mantissa bits). Plain mode keeps T. A null `lo` selects the single fp16 form
// (vocoder GEMMs): `hi` then holds fp16 bits, converted with saturation.
struct SplitPtr { __nv_bfloat16* hi; __nv_bfloat16* lo; };

template <typename T> struct OutPlain {
    T* p;
    STC_DEVINL void store(size_t i, T v) const { p[i] = v; }
};
// fp32 x2 -> packed fp16x2 (first argument in the lower half), round to nearest, saturating at +-65504: the single-pass fp16
// operand form of the vocoder GEMMs (DESIGN.md "precision"; `lo == nullptr` in OutSplit / Act selects it, `hi` then holds fp16 bits)
STC_DEVINL uint32_t pack_f16x2(float a, float b) {
    uint32_t r; asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r;
}
struct OutSplit {
    __nv_bfloat16* hi; __nv_bfloat16* lo;
    STC_DEVINL void store(size_t i, float v) const {
        if (!lo) { reinterpret_cast<uint16_t*>(hi)[i] = (uint16_t)(pack_f16x2(v, 0.f) & 0xffffu); return; }
        __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[i] = h;
        lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
};

// ---- embedding: out[row,:] = emb[ids[b, t],:] * mask[row]  (Gather -> Transpose -> Mul) ---------
// rows are packed sequences (row = off[b] + t) over the caller's [B, T] id rectangle; mask (per row) may be null.
template <typename T>
__global__ void embed_kernel(const int64_t* __restrict__ ids, const float* __restrict__ emb,
                             const float* __restrict__ mask, T* __restrict__ out, int rows, int C, int V,
                             const int* __restrict__ off, int B, int Tn) {
    pdl_wait(); pdl_trigger_light();
    int row = blockIdx.x * blockDim.y + threadIdx.y;
    if (row >= rows) return;
    const int b = find_seq(off, B, row);
    if (b < 0) {                                       // bucket padding row
        for (int c = threadIdx.x; c < C; c += 32) out[(size_t)row * C + c] = (T)0;
        return;
    }
    int64_t id = ids[(size_t)b * Tn + (row - __ldg(off + b))];
    if (id < 0 || id >= V) id = 0;
    T m = mask ? (T)mask[row] : (T)1;
    for (int c = threadIdx.x; c < C; c += 32) out[(size_t)row * C + c] = (T)emb[(size_t)id * C + c] * m;
}

// ---- x[row,:] = (x[row,:] + v[b,:]) * mask[row]   (style add in DP, time conditioning in VE) -----
template <typename T>
__global__ void add_rowvec_mask_kernel(T* __restrict__ x, const T* __restrict__ v, const float* __restrict__ mask,
                                       int rows, int C, int vstride, const int* __restrict__ off, int B) {
    pdl_wait(); pdl_trigger_light();
    // vstride = C: one vector per sequence (row -> b through the packed offsets); vstride = 0: one vector for every row
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)rows * C) return;
    int row = (int)(i / C), c = (int)(i % C);
    int b = 0;
    if (vstride) { b = find_seq(off, B, row); if (b < 0) return; }
    T m = mask ? (T)mask[row] : (T)1;
    x[i] = (x[i] + v[(size_t)b * vstride + c]) * m;
}

// ---- depthwise conv1d (+bias) -> LayerNorm over C, one warp per row ------------------------------
// y[c] = b[c] + sum_k w[c,k] * x[n + k*dil - pad_left, c] (zero outside [0,N) of the same utterance);
// out = (y-mean)/sqrt(var+eps)*g + beta.   K == 0 selects plain LayerNorm (no conv).
template <typename T, int CPL, typename Out>
__global__ void __launch_bounds__(256)
dwconv_ln_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ wb,
                 const float* __restrict__ g, const float* __restrict__ beta, Out out,
                 int rows, const int* __restrict__ off, int B, int K, int dil, int pad_left, float eps) {
    pdl_wait(); pdl_trigger_light();
    constexpr int C = CPL * 32;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int row = blockIdx.x * (blockDim.x >> 5) + warp;
    if (row >= rows) return;
    int base = 0, n = 0, N = 0;
    if (K != 0) {
        int b = find_seq(off, B, row);
        if (b < 0) {                                   // bucket padding row: keep it finite
#pragma unroll
            for (int i = 0; i < CPL; ++i) out.store((size_t)row * C + lane + 32 * i, (T)0);
            return;
        }
        base = __ldg(off + b); n = row - base; N = __ldg(off + b + 1) - base;
    }
    T y[CPL];
    if (K == 0) {
#pragma unroll
        for (int i = 0; i < CPL; ++i) y[i] = x[(size_t)row * C + lane + 32 * i];
    } else {
#pragma unroll
        for (int i = 0; i < CPL; ++i) y[i] = (T)wb[lane + 32 * i];
        for (int k = 0; k < K; ++k) {
            int nn = n + k * dil - pad_left;
            if (nn < 0 || nn >= N) continue;
            const T* xr = x + ((size_t)base + nn) * C;
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                int c = lane + 32 * i;
                y[i] += (T)w[c * K + k] * xr[c];
            }
        }
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < CPL; ++i) s += y[i];
    T mean = warp_sum<T>(s) / (T)C;
    T v = 0;
#pragma unroll
    for (int i = 0; i < CPL; ++i) { y[i] -= mean; v += y[i] * y[i]; }
    T var = warp_sum<T>(v) / (T)C;
    T den = t_sqrt<T>(var + (T)eps);
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        int c = lane + 32 * i;
        out.store((size_t)row * C + c, y[i] / den * (T)g[c] + (T)beta[c]);
    }
}

// Vectorised float variant for C in {128, 256, 512}: lane owns CPL CONSECUTIVE channels (float4 loads of x and of the
// tap-major weights wT[K][C], 8/16-byte operand stores), one warp per row, 8 rows per block.
STC_DEVINL void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}
template <int CPL> STC_DEVINL void store_row_vec(const OutSplit& o, size_t i, const float (&y)[CPL]) {
    uint32_t hi[CPL / 2], lo[CPL / 2];
    if (!o.lo) {                // single fp16 operand
#pragma unroll
        for (int j = 0; j < CPL / 2; ++j) hi[j] = pack_f16x2(y[2 * j], y[2 * j + 1]);
        if constexpr (CPL == 4) *reinterpret_cast<uint2*>(o.hi + i) = make_uint2(hi[0], hi[1]);
        else {
#pragma unroll
            for (int j = 0; j < CPL / 8; ++j) *reinterpret_cast<uint4*>(o.hi + i + 8 * j) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < CPL / 2; ++j) split2(y[2 * j], y[2 * j + 1], hi[j], lo[j]);
    if constexpr (CPL == 4) {
        *reinterpret_cast<uint2*>(o.hi + i) = make_uint2(hi[0], hi[1]);
        *reinterpret_cast<uint2*>(o.lo + i) = make_uint2(lo[0], lo[1]);
    } else {
#pragma unroll
        for (int j = 0; j < CPL / 8; ++j) {
            *reinterpret_cast<uint4*>(o.hi + i + 8 * j) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
            *reinterpret_cast<uint4*>(o.lo + i + 8 * j) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
        }
    }
}
template <int CPL> STC_DEVINL void store_row_vec(const OutPlain<float>& o, size_t i, const float (&y)[CPL]) {
#pragma unroll
    for (int j = 0; j < CPL; j += 4) *reinterpret_cast<float4*>(o.p + i + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
}

template <int CPL, typename Out>
__global__ void __launch_bounds__(256)
dwconv_ln_vec_kernel(const float* __restrict__ x, const float* __restrict__ wT, const float* __restrict__ wb,
                     const float* __restrict__ g, const float* __restrict__ beta, Out out,
                     int rows, const int* __restrict__ off, int B, int K, int dil, int pad_left, float eps) {
    pdl_wait(); pdl_trigger_light();
    constexpr int C = CPL * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + warp;
    if (row >= rows) return;
    const int c0 = lane * CPL;
    float y[CPL];
    if (K == 0) {
#pragma unroll
        for (int j = 0; j < CPL; j += 4) {
            const float4 v = *reinterpret_cast<const float4*>(x + (size_t)row * C + c0 + j);
            y[j] = v.x; y[j + 1] = v.y; y[j + 2] = v.z; y[j + 3] = v.w;
        }
    } else {
        const int b = find_seq(off, B, row);
        if (b < 0) {                                   // bucket padding row: keep it finite
#pragma unroll
            for (int j = 0; j < CPL; ++j) y[j] = 0.f;
            store_row_vec<CPL>(out, (size_t)row * C + c0, y);
            return;
        }
        const int base = __ldg(off + b), n = row - base, N = __ldg(off + b + 1) - base;
#pragma unroll
        for (int j = 0; j < CPL; j += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(wb + c0 + j));
            y[j] = v.x; y[j + 1] = v.y; y[j + 2] = v.z; y[j + 3] = v.w;
        }
        // taps in groups of up to 4: all loads of a group are issued before the first FMA (independent 16-byte requests in
        // flight per lane: the input is L2-resident, so the kernel is bound by load latency, not bandwidth)
        for (int k0 = 0; k0 < K; k0 += 4) {
            float4 xv[4][CPL / 4];
            bool ok[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int nn = n + (k0 + t) * dil - pad_left;
                ok[t] = (k0 + t < K) && nn >= 0 && nn < N;
                const float* xr = x + ((size_t)base + (ok[t] ? nn : n)) * C + c0;
#pragma unroll
                for (int j = 0; j < CPL / 4; ++j) xv[t][j] = *reinterpret_cast<const float4*>(xr + 4 * j);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (!ok[t]) continue;
                const float* wr = wT + (size_t)(k0 + t) * C + c0;
#pragma unroll
                for (int j = 0; j < CPL / 4; ++j) {
                    const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + 4 * j));
                    y[4 * j] += wv.x * xv[t][j].x; y[4 * j + 1] += wv.y * xv[t][j].y;
                    y[4 * j + 2] += wv.z * xv[t][j].z; y[4 * j + 3] += wv.w * xv[t][j].w;
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < CPL; ++j) s += y[j];
    const float mean = warp_sum<float>(s) / (float)C;
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < CPL; ++j) { y[j] -= mean; v += y[j] * y[j]; }
    const float inv = 1.0f / sqrtf(warp_sum<float>(v) / (float)C + eps);      // see dwconv_ln_tile_kernel
#pragma unroll
    for (int j = 0; j < CPL; j += 4) {
        const float4 gv = __ldg(reinterpret_cast<const float4*>(g + c0 + j)), bv = __ldg(reinterpret_cast<const float4*>(beta + c0 + j));
        y[j] = y[j] * inv * gv.x + bv.x; y[j + 1] = y[j + 1] * inv * gv.y + bv.y;
        y[j + 2] = y[j + 2] * inv * gv.z + bv.z; y[j + 3] = y[j + 3] * inv * gv.w + bv.w;
    }
    store_row_vec<CPL>(out, (size_t)row * C + c0, y);
}

// Shared-memory tiled variant: a block owns R consecutive packed rows and stages the R + (K-1)*dil input rows its taps
// touch ONCE (coalesced float4), instead of re-reading every tap row through L2 (7 x 57 MB per vocoder block before).
// Lane owns the float4 chunks {lane + 32 j}: every shared/global access of a warp is one contiguous 512-byte run.
template <int CPL, typename Out>
__global__ void __launch_bounds__(256)
dwconv_ln_tile_kernel(const float* __restrict__ x, const float* __restrict__ wT, const float* __restrict__ wb,
                      const float* __restrict__ g, const float* __restrict__ beta, Out out,
                      int rows, const int* __restrict__ off, int B, int K, int dil, int pad_left, float eps, int R) {
    pdl_wait(); pdl_trigger_light();
    constexpr int C = CPL * 32, V = CPL / 4;
    extern __shared__ float4 tile4[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = blockIdx.x * R;
    const int span = (K - 1) * dil;
    const int w0 = r0 - pad_left, wrows = R + span;               // staged window of packed rows [w0, w0 + wrows)
    {   // cp.async fill: every 16-byte piece of the window is in flight at once (no register staging); rows outside
        // [0, rows) are zero-filled by a zero source size
        const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(tile4);
        for (int i = threadIdx.x; i < wrows * (C / 4); i += blockDim.x) {
            const int gr = w0 + i / (C / 4);
            const bool ok = gr >= 0 && gr < rows;
            const float* src = ok ? x + (size_t)gr * C + (i % (C / 4)) * 4 : x;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sbase + (uint32_t)i * 16u), "l"(src), "r"(ok ? 16 : 0) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    int b = -2;                                        // sequence of the previous row handled by this warp (-2: none yet)
    for (int rl = warp; rl < R; rl += (blockDim.x >> 5)) {
        const int row = r0 + rl;
        if (row >= rows) break;
        float4 y[V];
        if (b == -2) b = find_seq(off, B, row);
        else if (b >= 0) { while (b < B && row >= __ldg(off + b + 1)) ++b; if (b >= B) b = -1; }
        if (b < 0) {                                   // bucket padding row: keep it finite
#pragma unroll
            for (int j = 0; j < V; ++j) y[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            const int base = __ldg(off + b), n = row - base, N = __ldg(off + b + 1) - base;
#pragma unroll
            for (int j = 0; j < V; ++j) y[j] = __ldg(reinterpret_cast<const float4*>(wb) + lane + 32 * j);
            for (int k = 0; k < K; ++k) {
                const int nn = n + k * dil - pad_left;
                if (nn < 0 || nn >= N) continue;
                const float4* xr = tile4 + (size_t)(rl + k * dil) * (C / 4);
                const float4* wr = reinterpret_cast<const float4*>(wT + (size_t)k * C);
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const float4 xv = xr[lane + 32 * j], wv = __ldg(wr + lane + 32 * j);
                    y[j].x += wv.x * xv.x; y[j].y += wv.y * xv.y; y[j].z += wv.z * xv.z; y[j].w += wv.w * xv.w;
                }
            }
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < V; ++j) s += (y[j].x + y[j].y) + (y[j].z + y[j].w);
            const float mean = warp_sum<float>(s) / (float)C;
            float v = 0.f;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                y[j].x -= mean; y[j].y -= mean; y[j].z -= mean; y[j].w -= mean;
                v += (y[j].x * y[j].x + y[j].y * y[j].y) + (y[j].z * y[j].z + y[j].w * y[j].w);
            }
            // one reciprocal per row instead of C IEEE divisions (~10 instructions each in a kernel that is issue bound):
            // (y * (1/den)) * g + b differs from the graph's (y / den) * g + b by <= 1 ulp of the quotient
            const float inv = 1.0f / sqrtf(warp_sum<float>(v) / (float)C + eps);
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + lane + 32 * j);
                const float4 bv = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * j);
                y[j].x = y[j].x * inv * gv.x + bv.x; y[j].y = y[j].y * inv * gv.y + bv.y;
                y[j].z = y[j].z * inv * gv.z + bv.z; y[j].w = y[j].w * inv * gv.w + bv.w;
            }
        }
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float t[4] = {y[j].x, y[j].y, y[j].z, y[j].w};
            store_row_vec<4>(out, (size_t)row * C + 4 * (lane + 32 * j), t);
        }
    }
}

// Sliding-window variant: a thread owns FOUR channels (one float4) and walks a CHAIN of output rows r, r + dil, r + 2 dil, ...
// Along such a chain a dilated convolution is an undilated one, so the K tap rows of consecutive outputs overlap in K - 1 rows,
// which stay in registers: every input row is read from global memory once per chain (+ K - 1 halo rows), coalesced (the C/4
// threads of a group read one contiguous row), with no shared-memory staging and no re-read of taps or weights — the K tap
// weights and the bias live in registers for the whole chain (the tile kernel above issues an LDS + an LDG per tap and float4:
// ~1100 warp instructions per 512-channel row, which made it issue bound at 20 % of the HBM rate). Tap products use the packed
// fma.rn.f32x2 of sm_100 in the same k order as the other variants (bit-identical convolution). LayerNorm needs the whole
// row = the NW warps of a group: partial sums of U = 4 rows at a time go through a transposing butterfly (6 shuffles for four
// sums), shared memory and one named barrier per pass; two-pass (mean, then centred squares) like everywhere else.
// Grid: chain c of a tile of RT * dil rows has phase c % dil; 4 / NW groups (chains) per 128-thread block.
STC_DEVINL float warp_sum4(float v0, float v1, float v2, float v3, int lane) {
    // lane l ends with the warp total of value 2 * bit4(l) + bit3(l)
    const bool h16 = lane & 16, h8 = lane & 8;
    const float a0 = (h16 ? v2 : v0) + __shfl_xor_sync(0xffffffffu, h16 ? v0 : v2, 16);
    const float a1 = (h16 ? v3 : v1) + __shfl_xor_sync(0xffffffffu, h16 ? v1 : v3, 16);
    float b = (h8 ? a1 : a0) + __shfl_xor_sync(0xffffffffu, h8 ? a0 : a1, 8);
    b += __shfl_xor_sync(0xffffffffu, b, 4);
    b += __shfl_xor_sync(0xffffffffu, b, 2);
    b += __shfl_xor_sync(0xffffffffu, b, 1);
    return b;
}
template <int NW> STC_DEVINL void group_barrier(int grp) {
    if constexpr (NW == 1) __syncwarp();
    else if constexpr (NW == 4) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(32 * NW) : "memory");
}
template <int NW> STC_DEVINL float group_total(const float* p) {     // sum of the NW per-warp partials, same order in every thread
    if constexpr (NW == 4) { const float4 v = *reinterpret_cast<const float4*>(p); return (v.x + v.y) + (v.z + v.w); }
    else if constexpr (NW == 2) { const float2 v = *reinterpret_cast<const float2*>(p); return v.x + v.y; }
    else return p[0];
}

// RING: the rows of the next RING_D iterations are in flight as per-thread cp.async copies into a shared-memory ring (a thread
// reads back only what it copied itself: no barrier), 128 KB per SM instead of the 32 KB a register prefetch of one iteration
// keeps in flight — the long vocoder chains are bound by bytes in flight, the short VE / TE chains (one or two iterations)
// by latency and keep the register prefetch.
constexpr int RING_D = 4;
template <int NW, int K, bool RING, typename Out>
__global__ void __launch_bounds__(128)
dwconv_ln_slide_kernel(const float* __restrict__ x, const float* __restrict__ wT, const float* __restrict__ wb,
                       const float* __restrict__ g, const float* __restrict__ beta, Out out,
                       int rows, const int* __restrict__ off, int B, int dil, int pad_left, float eps, int RT) {
    constexpr int C = 128 * NW, GT = 32 * NW, GPB = 4 / NW, U = 4;
    __shared__ __align__(16) float red[2][GPB][U][NW];
    const int grp = threadIdx.x / GT, t = threadIdx.x % GT, wig = t >> 5, lane = threadIdx.x & 31;
    const int chain = blockIdx.x * GPB + grp;
    const int r_first = (chain / dil) * (RT * dil) + chain % dil;
    if (r_first >= rows) return;                       // the whole group leaves (barriers are per group)
    const float* xc = x + 4 * t;
    float2 w[K][2];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(wT + (size_t)k * C) + t);
        w[k][0] = make_float2(v.x, v.y); w[k][1] = make_float2(v.z, v.w);
    }
    const float4 bias = __ldg(reinterpret_cast<const float4*>(wb) + t);
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + t), bv = __ldg(reinterpret_cast<const float4*>(beta) + t);
    auto load_row = [&](int r) -> float4 {
        if (!(r >= 0 && r < rows)) return make_float4(0.f, 0.f, 0.f, 0.f);
        return *reinterpret_cast<const float4*>(xc + (size_t)r * C);
    };
    const int rw0 = r_first - pad_left;                // window slot j of output i holds row rw0 + (i + j) * dil
    float4 win[K - 1 + U], nxt[RING ? 1 : U];
    __shared__ float4 ring[RING ? (RING_D + 1) * U * 128 : 1];          // [slot][u][thread]: conflict-free 16-byte accesses
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring) + threadIdx.x * 16u;
    auto fetch = [&](int it) {                         // rows of iteration `it` -> ring slot it % (RING_D + 1); always one commit group
        if (it * U < RT) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int r = rw0 + (K - 1 + it * U + u) * dil;
                const bool ok = r >= 0 && r < rows;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(ring_s + (uint32_t)((it % (RING_D + 1)) * U + u) * 2048u),
                             "l"(ok ? xc + (size_t)r * C : xc), "r"(ok ? 16 : 0) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // tap weights, LayerNorm parameters and the sequence lookup (a binary search: five dependent L2 round trips) do not depend on the
    // predecessor kernel: under programmatic dependent launch they overlap its tail (kernels.cuh: pre-wait rules)
    int b = find_seq(off, B, r_first), lo = 0x7fffffff, hi = 0x7fffffff;
    if (b >= 0) { lo = __ldg(off + b); hi = __ldg(off + b + 1); }
    pdl_wait(); pdl_trigger_light();
    if constexpr (RING) {
#pragma unroll
        for (int it = 0; it < RING_D; ++it) fetch(it);
    }
#pragma unroll
    for (int j = 0; j < K - 1; ++j) win[j] = load_row(rw0 + j * dil);
    if constexpr (!RING) {
#pragma unroll
        for (int u = 0; u < U; ++u) nxt[u] = load_row(rw0 + (K - 1 + u) * dil);
    }
    for (int i0 = 0; i0 < RT; i0 += U) {
        if (r_first + i0 * dil >= rows) break;         // uniform in the group
        if constexpr (RING) {
            asm volatile("cp.async.wait_group %0;" ::"n"(RING_D - 1) : "memory");
            const int slot = (i0 / U) % (RING_D + 1);
#pragma unroll
            for (int u = 0; u < U; ++u) win[K - 1 + u] = ring[(slot * U + u) * 128 + threadIdx.x];
            fetch(i0 / U + RING_D);                    // into the slot consumed one iteration ago
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) win[K - 1 + u] = nxt[u];
            if (i0 + U < RT) {                         // next iteration's rows are in flight during this one's arithmetic
#pragma unroll
                for (int u = 0; u < U; ++u) nxt[u] = load_row(rw0 + (K - 1 + i0 + U + u) * dil);
            }
        }
        float2 y[U][2];
        bool pad[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int r = r_first + (i0 + u) * dil;
            if (b >= 0 && r >= hi) {                   // next sequence (empty ones are skipped); past the last one: padding
                do ++b; while (b < B && r >= __ldg(off + b + 1));
                if (b >= B) b = -1; else { lo = __ldg(off + b); hi = __ldg(off + b + 1); }
            }
            pad[u] = b < 0;
            y[u][0] = make_float2(bias.x, bias.y); y[u][1] = make_float2(bias.z, bias.w);
            const int first = r - pad_left;
            if (b < 0) {                               // bucket padding row: keep it finite
                y[u][0] = y[u][1] = make_float2(0.f, 0.f);
            } else if (first >= lo && first + (K - 1) * dil < hi) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    y[u][0] = __ffma2_rn(w[k][0], make_float2(win[u + k].x, win[u + k].y), y[u][0]);
                    y[u][1] = __ffma2_rn(w[k][1], make_float2(win[u + k].z, win[u + k].w), y[u][1]);
                }
            } else {                                   // a sequence edge: taps outside [lo, hi) are the zero padding
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int rk = first + k * dil;
                    if (rk >= lo && rk < hi) {
                        y[u][0] = __ffma2_rn(w[k][0], make_float2(win[u + k].x, win[u + k].y), y[u][0]);
                        y[u][1] = __ffma2_rn(w[k][1], make_float2(win[u + k].z, win[u + k].w), y[u][1]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < K - 1; ++j) win[j] = win[j + U];
        // LayerNorm over the C channels of each of the U rows
        float tot = warp_sum4((y[0][0].x + y[0][0].y) + (y[0][1].x + y[0][1].y), (y[1][0].x + y[1][0].y) + (y[1][1].x + y[1][1].y),
                              (y[2][0].x + y[2][0].y) + (y[2][1].x + y[2][1].y), (y[3][0].x + y[3][0].y) + (y[3][1].x + y[3][1].y), lane);
        if ((lane & 7) == 0) red[0][grp][lane >> 3][wig] = tot;
        group_barrier<NW>(grp);
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float mean = group_total<NW>(&red[0][grp][u][0]) / (float)C;
            const float2 nm = make_float2(-mean, -mean);
            y[u][0] = __fadd2_rn(y[u][0], nm); y[u][1] = __fadd2_rn(y[u][1], nm);
            v[u] = (y[u][0].x * y[u][0].x + y[u][0].y * y[u][0].y) + (y[u][1].x * y[u][1].x + y[u][1].y * y[u][1].y);
        }
        tot = warp_sum4(v[0], v[1], v[2], v[3], lane);
        if ((lane & 7) == 0) red[1][grp][lane >> 3][wig] = tot;
        group_barrier<NW>(grp);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int r = r_first + (i0 + u) * dil;
            if (r >= rows) continue;
            const float inv = 1.0f / sqrtf(group_total<NW>(&red[1][grp][u][0]) / (float)C + eps);     // see dwconv_ln_tile_kernel
            const float2 iv = make_float2(inv, inv);
            const float2 o0 = __ffma2_rn(__fmul2_rn(y[u][0], iv), make_float2(gv.x, gv.y), make_float2(bv.x, bv.y));
            const float2 o1 = __ffma2_rn(__fmul2_rn(y[u][1], iv), make_float2(gv.z, gv.w), make_float2(bv.z, bv.w));
            const float o[4] = {pad[u] ? 0.f : o0.x, pad[u] ? 0.f : o0.y, pad[u] ? 0.f : o1.x, pad[u] ? 0.f : o1.y};
            store_row_vec<4>(out, (size_t)r * C + 4 * t, o);
        }
    }
}

// ---- elementwise copy into operand format (split bf16 or plain) ----------------------------------
template <typename Out>
__global__ void convert_kernel(const float* __restrict__ x, Out out, size_t n) {
    pdl_wait(); pdl_trigger_light();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out.store(i, x[i]);
}

// ---- fp32/fp64 CUDA-core GEMM with the fused epilogue --------------------------------------------
// out[M,N] = epi( A[M,K] (T, row-major, lda) * W[K,N] (float, row-major) )
//   v = acc + bias[col]; if gelu: v = gelu(v); if scale: v *= scale[col]; if resid: v += resid[row,col];
//   if mask: v *= mask[row]
struct Epilogue {
    const float* bias = nullptr;    // [N]
    const float* scale = nullptr;   // [N] layer-scale gamma, or dt for the Euler update
    const void* resid = nullptr;    // [M,N] (T)
    const float* mask = nullptr;    // [M]
    int gelu = 0;
    // rotary embedding on the (bias-added) output, pairs in adjacent columns (tensor-core GEMM only; see rope_kernel)
    const float* rope_freqs = nullptr;   // [DH/2]; null = no rotation
    const int* rope_off = nullptr;       // packed row offsets [B+1] of the output rows
    const float* rope_len = nullptr;     // [B] sequence lengths (length-aware variant) or null
    int rope_B = 0, rope_dh = 64;
};

template <typename T, typename Out, int BM = 64>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const T* __restrict__ A, int lda, const float* __restrict__ W, Out out, int ldo,
                 int M, int N, int K, Epilogue ep) {
    pdl_wait(); pdl_trigger_light();
    constexpr int BN = 64, BK = 16, TM = BM / 16;
    __shared__ T As[BK][BM + 1];
    __shared__ T Ws[BK][BN + 1];
    int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;       // 16 x 16 threads, TM x 4 micro-tile
    int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    T acc[TM][4] = {};
    for (int k0 = 0; k0 < K; k0 += BK) {
        for (int i = threadIdx.x; i < BM * BK; i += 256) {
            int r = i / BK, c = i % BK;
            int gm = m0 + r, gk = k0 + c;
            As[c][r] = (gm < M && gk < K) ? A[(size_t)gm * lda + gk] : (T)0;
        }
        for (int i = threadIdx.x; i < BK * BN; i += 256) {
            int r = i / BN, c = i % BN;
            int gk = k0 + r, gn = n0 + c;
            Ws[r][c] = (gk < K && gn < N) ? (T)W[(size_t)gk * N + gn] : (T)0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            T a[TM], w[4];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
            for (int i = 0; i < 4; ++i) w[i] = Ws[k][tx * 4 + i];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * w[j];
        }
        __syncthreads();
    }
    const T* resid = static_cast<const T*>(ep.resid);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int gm = m0 + ty * TM + i;
        if (gm >= M) continue;
        T mk = ep.mask ? (T)ep.mask[gm] : (T)1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            T v = acc[i][j];
            if (ep.bias) v += (T)ep.bias[gn];
            if (ep.gelu) v = gelu_erf<T>(v);
            if (ep.scale) v *= (T)ep.scale[gn];
            if (resid) v += resid[(size_t)gm * ldo + gn];
            if (ep.mask) v *= mk;
            out.store((size_t)gm * ldo + gn, v);
        }
    }
}

// ---- sequence lengths from masks: len[b] = sum_n mask[b,n] ---------------------------------------
__global__ void mask_len_kernel(const float* __restrict__ mask, float* __restrict__ len, int* __restrict__ cnt, int N) {
    pdl_wait(); pdl_trigger_light();
    // len[b] = sum of the mask (what the graphs' ReduceSum sees); cnt[b] = 1 + index of the last non-zero entry
    int b = blockIdx.x;
    float s = 0.f; int last = 0;
    for (int n = threadIdx.x; n < N; n += 32) { float m = mask[(size_t)b * N + n]; s += m; if (m != 0.f) last = n + 1; }
    s = warp_sum<float>(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
    if (threadIdx.x == 0) { len[b] = s; cnt[b] = last; }
}

// ---- rotary embedding in place on [rows, heads*DH] ------------------------------------------------
// The graphs rotate the pairs (d, d + DH/2) ("rotate-half"). The library permutes the output columns of every rotary Q / K
// projection at load time so that such a pair sits in ADJACENT columns (2i, 2i+1) — Q.K^T is invariant under a common
// permutation of the head dimension — which lets the tensor-core GEMM epilogue rotate inside one float4. This kernel is the
// CUDA-core path's form of the same thing. pos = n (abs) or n / len[b] (length-aware RoPE); ang = pos * freqs[i].
__global__ void rope_kernel(float* __restrict__ x, const float* __restrict__ freqs, const float* __restrict__ len,
                            int rows, const int* __restrict__ off, int B, int heads, int DH, int normalise) {
    pdl_wait(); pdl_trigger_light();
    int half = DH / 2;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)rows * heads * half;
    if (i >= total) return;
    int d = (int)(i % half);
    int h = (int)((i / half) % heads);
    int row = (int)(i / ((size_t)half * heads));
    int b = find_seq(off, B, row);
    if (b < 0) return;
    float pos = (float)(row - __ldg(off + b));
    if (normalise) pos = pos / len[b];
    float ang = pos * freqs[d];
    float c = cosf(ang), s = sinf(ang);
    float2* p = reinterpret_cast<float2*>(x + (size_t)row * heads * DH + (size_t)h * DH) + d;
    const float2 t = *p;
    *p = make_float2(t.x * c - t.y * s, t.x * s + t.y * c);
}

// ---- multi-head attention core: O = softmax(Q K^T * scale + keymask) V ----------------------------
// Q [B*Nq, H*DH], K/V [B*Nk, H*DH] (already rotated), kmask [B*Nk] or null. One warp = QPW queries,
// keys streamed through shared memory in chunks of 32 with an online softmax (fp32).
template <int DH, typename Out>
__global__ void __launch_bounds__(128)
attention_kernel(const float* __restrict__ Q, const float* __restrict__ Kt, const float* __restrict__ Vt,
                 const float* __restrict__ kmask, Out out, const int* __restrict__ qoff, const int* __restrict__ koff,
                 const int* __restrict__ kcnt, int heads, float scale) {
    pdl_wait(); pdl_trigger_light();
    constexpr int QPW = 4, WARPS = 4, QT = QPW * WARPS, KC = 32, DPL = DH / 32;
    __shared__ float Ks[KC][DH + 1];
    __shared__ float Vs[KC][DH + 1];
    __shared__ float Qs[QT][DH];
    __shared__ float Ms[KC];
    int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QT;
    const int qbase = __ldg(qoff + b), Nq = __ldg(qoff + b + 1) - qbase;
    const int kbase = __ldg(koff + b);
    const int Nk = kcnt ? min(__ldg(koff + b + 1) - kbase, __ldg(kcnt + b)) : __ldg(koff + b + 1) - kbase;   // skip the masked tail
    if (q0 >= Nq) return;                                   // block-uniform
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int C = heads * DH;
    for (int i = threadIdx.x; i < QT * DH; i += 128) {
        int qi = i / DH, d = i % DH;
        int q = q0 + qi;
        Qs[qi][d] = (q < Nq) ? Q[((size_t)qbase + q) * C + h * DH + d] : 0.f;
    }
    float m[QPW], l[QPW], o[QPW][DPL];
#pragma unroll
    for (int i = 0; i < QPW; ++i) { m[i] = -INFINITY; l[i] = 0.f;
#pragma unroll
        for (int j = 0; j < DPL; ++j) o[i][j] = 0.f; }
    for (int k0 = 0; k0 < Nk; k0 += KC) {
        __syncthreads();
        for (int i = threadIdx.x; i < KC * DH; i += 128) {
            int kj = i / DH, d = i % DH;
            int k = k0 + kj;
            bool ok = k < Nk;
            size_t g = ((size_t)kbase + (ok ? k : 0)) * C + h * DH + d;
            Ks[kj][d] = ok ? Kt[g] : 0.f;
            Vs[kj][d] = ok ? Vt[g] : 0.f;
        }
        if (threadIdx.x < KC) {
            int k = k0 + threadIdx.x;
            Ms[threadIdx.x] = (k < Nk) ? (kmask ? kmask[(size_t)kbase + k] : 1.f) : 0.f;
        }
        __syncthreads();
        bool valid = Ms[lane] != 0.f;
#pragma unroll
        for (int qi = 0; qi < QPW; ++qi) {
            const float* qv = Qs[warp * QPW + qi];
            float s = 0.f;
#pragma unroll 16
            for (int d = 0; d < DH; ++d) s += qv[d] * Ks[lane][d];
            s = valid ? s * scale : -INFINITY;
            float mx = fmaxf(m[qi], warp_max(s));
            if (mx == -INFINITY) continue;              // nothing valid yet (warp-uniform)
            float p = valid ? expf(s - mx) : 0.f;
            float corr = expf(m[qi] - mx);              // m = -inf -> 0
            l[qi] = l[qi] * corr + warp_sum<float>(p);
            m[qi] = mx;
#pragma unroll
            for (int j = 0; j < DPL; ++j) o[qi][j] *= corr;
            for (int kj = 0; kj < KC; ++kj) {
                float pj = __shfl_sync(0xffffffffu, p, kj);
#pragma unroll
                for (int j = 0; j < DPL; ++j) o[qi][j] += pj * Vs[kj][lane + 32 * j];
            }
        }
    }
#pragma unroll
    for (int qi = 0; qi < QPW; ++qi) {
        int q = q0 + warp * QPW + qi;
        if (q >= Nq) continue;
        float inv = l[qi] > 0.f ? 1.f / l[qi] : 0.f;
#pragma unroll
        for (int j = 0; j < DPL; ++j)
            out.store(((size_t)qbase + q) * C + h * DH + lane + 32 * j, o[qi][j] * inv);
    }
}

// ---- layout changes at the API boundary -----------------------------------------------------------
// NCL [B,C,N] -> NLC [B,N,C] (and back), 32x32 shared-memory tile transpose
template <typename TI, typename TO>
__global__ void transpose_kernel(const TI* __restrict__ in, TO* __restrict__ out, int R, int Cc) {
    pdl_wait(); pdl_trigger_light();
    // in: [batch][R][Cc] -> out: [batch][Cc][R]
    __shared__ float tile[32][33];
    int b = blockIdx.z;
    int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const TI* ip = in + (size_t)b * R * Cc;
    TO* op = out + (size_t)b * R * Cc;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && c < Cc) ? (float)ip[(size_t)r * Cc + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, r = r0 + threadIdx.x;
        if (c < Cc && r < R) op[(size_t)c * R + r] = (TO)tile[threadIdx.x][i];
    }
}

// ---- noise * mask into the channels-last loop state (sampleNoisyLatent, cpp/helper.cpp:446-466) ---
// host-provided noise [B][D][ld] (reference draw order b,d,t) or Philox-keyed Gaussian.
STC_DEVINL uint32_t mix32(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31; return (uint32_t)(z >> 16);
}
__global__ void init_latent_kernel(const float* __restrict__ noise, int64_t ld, const uint64_t* __restrict__ seed_p,
                                   const float* __restrict__ mask, float* __restrict__ x, int rows,
                                   const int* __restrict__ off, int B, int D, const int* __restrict__ noise_index) {
    pdl_wait(); pdl_trigger_light();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)rows * D) return;
    int d = (int)(i % D);
    int row = (int)(i / D);
    int b = find_seq(off, B, row);
    if (b < 0) { x[i] = 0.f; return; }
    int l = row - __ldg(off + b);
    float v;
    if (noise) v = noise[((size_t)b * D + d) * ld + l];
    else {
        // utterance b draws stream noise_index[b] (default b): a request split over several calls / GPUs gets the noise of one call
        const uint64_t ub = noise_index ? (uint64_t)(uint32_t)__ldg(noise_index + b) : (uint64_t)b;
        uint64_t key = __ldg(seed_p) * 0x9E3779B97F4A7C15ull + (ub << 40) + ((uint64_t)d << 24) + (uint64_t)l;
        float u1 = (mix32(key) + 1.0f) * (1.0f / 4294967808.0f);          // (0,1]
        float u2 = mix32(key ^ 0xD1B54A32D192ED03ull) * (1.0f / 4294967296.0f);
        v = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
    }
    x[i] = mask ? v * mask[row] : v;
}

// latent mask from wav lengths: mask[b,l] = l < ceil(wav_len[b]/cs)  (getLatentMask, cpp/helper.cpp:759-770)
__global__ void latent_mask_kernel(const int64_t* __restrict__ wav_len, float* __restrict__ mask, int B, int L, int cs) {
    pdl_wait(); pdl_trigger_light();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * L) return;
    int b = i / L, l = i % L;
    int64_t ll = (wav_len[b] + cs - 1) / cs;
    mask[i] = (l < ll) ? 1.f : 0.f;
}

// ---- vocoder front-end: de-normalise, un-compress [B,L,f*ld] -> frames [B,f*L,ld], im2col for conv_in
// A[(b*fL + n), k*ld + c] = z[b, n - (K-1) + k, c] (causal, zero left pad);
// z[b, f*l + j, c] = lat[b,l, j*ld + c] * std[j*ld+c] + mean[j*ld+c]
template <typename Out>
__global__ void voc_im2col_kernel(const float* __restrict__ lat, const float* __restrict__ sd,
                                  const float* __restrict__ mean, Out out, int rows6, const int* __restrict__ off,
                                  int B, int f, int ld, int K, int lda) {
    pdl_wait(); pdl_trigger_light();
    // off = LATENT-frame offsets; output rows run at f x the latent rate: row6 in [f*off[b], f*off[b+1])
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int KW = K * ld;
    size_t total = (size_t)rows6 * lda;
    if (i >= total) return;
    int col = (int)(i % lda);
    int row6 = (int)(i / lda);
    float v = 0.f;
    if (col < KW) {
        int k = col / ld, c = col % ld;
        int b = find_seq(off, B, row6 / f);
        if (b >= 0) {
            int base = __ldg(off + b);
            int nn = row6 - f * base - (K - 1) + k;
            if (nn >= 0) {
                int l = nn / f, j = nn % f;
                int ch = j * ld + c;
                v = lat[((size_t)base + l) * (f * ld) + ch] * sd[ch] + mean[ch];
            }
        }
    }
    out.store(i, v);
}

// ---- duration head (fp64): dur[b] = sum_t mask * exp(clip(LN(x_t).w + b, -clip, clip)) * spt -------
template <int CPL>
__global__ void dp_head_kernel(const double* __restrict__ x, const float* __restrict__ g, const float* __restrict__ beta,
                               const float* __restrict__ w, const float* __restrict__ wb, const float* __restrict__ mask,
                               float* __restrict__ dur, const int* __restrict__ off, float eps, float clip, float spt) {
    pdl_wait(); pdl_trigger_light();
    constexpr int C = CPL * 32;
    __shared__ double part[32];
    int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int base = __ldg(off + b), N = __ldg(off + b + 1) - base;
    double acc = 0.0;
    for (int n = warp; n < N; n += nw) {
        const double* xr = x + ((size_t)base + n) * C;
        double y[CPL], s = 0;
#pragma unroll
        for (int i = 0; i < CPL; ++i) { y[i] = xr[lane + 32 * i]; s += y[i]; }
        double mean = warp_sum<double>(s) / C, v = 0;
#pragma unroll
        for (int i = 0; i < CPL; ++i) { y[i] -= mean; v += y[i] * y[i]; }
        double den = sqrt(warp_sum<double>(v) / C + (double)eps), dot = 0;
#pragma unroll
        for (int i = 0; i < CPL; ++i) { int c = lane + 32 * i; dot += (y[i] / den * (double)g[c] + (double)beta[c]) * (double)w[c]; }
        dot = warp_sum<double>(dot) + (double)wb[0];
        dot = fmin(fmax(dot, -(double)clip), (double)clip);
        acc += exp(dot) * (double)spt * (mask ? (double)mask[(size_t)base + n] : 1.0);
    }
    if (lane == 0) part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int i = 0; i < nw; ++i) t += part[i];
        dur[b] = (float)t;
    }
}

// duration /= speed; wav_len = (int64)(d*sr)   (cpp/helper.cpp:529-531, 434) — float32 IEEE, no fast-math
__global__ void dur_post_kernel(float* __restrict__ dur, int64_t* __restrict__ wav_len, int B, float speed, int sr) {
    pdl_wait(); pdl_trigger_light();
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float d = __fdiv_rn(dur[b], speed);
    dur[b] = d;
    wav_len[b] = (int64_t)__fmul_rn(d, (float)sr);
}

// ---- waveform output packing: silence insert + PCM16 quantise on the device ----------------------------------------------
// The vocoder leaves utterance b's samples at src[off[b]*cs ...) (packed frames). This writes them to dst with `gap` zeros after every
// utterance but the last — what TextToSpeech::call does on the host when it joins the chunks of a long text (cpp/helper.cpp:706-714) —
// and, for PCM16, quantised exactly like writeWavFile (cpp/helper.cpp:985-988): (int16_t)(max(-1, min(1, x)) * 32767), the cast
// truncating toward zero. One block per latent frame (cs samples, 8 per thread step).
STC_DEVINL int16_t pcm16_of(float v) {
    const float c = fmaxf(-1.0f, fminf(1.0f, v));            // std::max(-1.0f, std::min(1.0f, sample)); NaN -> -1 on both sides
    return (int16_t)__float2int_rz(__fmul_rn(c, 32767.0f));
}
template <typename TO> STC_DEVINL TO wav_out_of(float v);
template <> STC_DEVINL float wav_out_of<float>(float v) { return v; }
template <> STC_DEVINL int16_t wav_out_of<int16_t>(float v) { return pcm16_of(v); }

template <typename TO>
__global__ void __launch_bounds__(256)
wav_pack_kernel(const float* __restrict__ src, TO* __restrict__ dst, const int* __restrict__ off, int B, int cs, long long gap) {
    pdl_wait(); pdl_trigger_light();
    const int r = blockIdx.x;
    const int b = find_seq(off, B, r);
    if (b < 0) return;                                         // bucket padding frame
    const float* s = src + (size_t)r * cs;
    TO* d = dst + (size_t)r * cs + (size_t)b * gap;
    const bool vec = (reinterpret_cast<uintptr_t>(d) & 15) == 0 && cs % 8 == 0;
    if (vec) {
        for (int i = threadIdx.x * 8; i < cs; i += blockDim.x * 8) {
            const float4 v0 = *reinterpret_cast<const float4*>(s + i), v1 = *reinterpret_cast<const float4*>(s + i + 4);
            if constexpr (sizeof(TO) == 2) {
                const int16_t q[8] = {pcm16_of(v0.x), pcm16_of(v0.y), pcm16_of(v0.z), pcm16_of(v0.w), pcm16_of(v1.x), pcm16_of(v1.y), pcm16_of(v1.z), pcm16_of(v1.w)};
                *reinterpret_cast<uint4*>(d + i) = *reinterpret_cast<const uint4*>(q);
            } else {
                *reinterpret_cast<float4*>(d + i) = v0; *reinterpret_cast<float4*>(d + i + 4) = v1;
            }
        }
    } else {
        for (int i = threadIdx.x; i < cs; i += blockDim.x) d[i] = wav_out_of<TO>(s[i]);
    }
    if (gap > 0 && b + 1 < B && r + 1 == __ldg(off + b + 1)) {  // the utterance's last frame also writes the silence behind it
        TO* z = d + cs;
        for (long long i = threadIdx.x; i < gap; i += blockDim.x) z[i] = (TO)0;
    }
}

// sinusoidal time embedding: t = cur/tot; out[b] = [sin(t*f), cos(t*f)]
__global__ void time_embed_kernel(const float* __restrict__ cur, const float* __restrict__ tot,
                                  const float* __restrict__ freqs, float* __restrict__ out, int B, int half) {
    pdl_wait(); pdl_trigger_light();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * half) return;
    int b = i / half, j = i % half;
    float t = __fdiv_rn(cur[b], tot[b]);
    float a = t * freqs[j];
    out[(size_t)b * 2 * half + j] = sinf(a);
    out[(size_t)b * 2 * half + half + j] = cosf(a);
}

// copy rows [B][L*cs] out of a wider device matrix into a strided destination
__global__ void fill_kernel(float* __restrict__ p, float v, size_t n) {
    pdl_wait(); pdl_trigger_light();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace stc
