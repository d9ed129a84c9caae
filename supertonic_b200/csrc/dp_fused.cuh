// duration_predictor.onnx in fp64: one kernel per ConvNeXt block (sm_100a CUDA cores).
//
// Why fp64 at all: `wav_lengths[b] = (int64)(duration * sample_rate)` and the latent length (cpp/helper.cpp:430-438) flip on a
// one-ulp change of the float32 duration, so the graph is evaluated in double and rounded once — like the oracle — which makes the
// result independent of summation order (DESIGN.md "bit-exact durations"). The graph is tiny (C = 64, H = 256: 1.3 GFLOP for 5 800
// tokens) but sits on the critical path of every call: the latent length is data dependent, so the Euler loop cannot be enqueued
// before the durations are on the host. As nine CUDA-core GEMM launches + four conv/LayerNorm launches it took 0.61 ms.
//
// Here a block of R = 32 token rows does the whole ConvNeXt block in shared memory:
//   rows (+ K-1 halo) -> depthwise conv + bias -> LayerNorm -> a[R][C]
//   hid[R][H] = GELU(a W1 + b1)        (each thread an 8 x 4 (R/TY1 x 4) register tile, W1 read once per block through L1)
//   out[R][C] = ((hid W2 + b2) * gamma + x) * mask
// replacing Conv(group=C), LayerNormalization, MatMul+Add, Erf-GELU, MatMul+Add, Mul, Add, Mul of the graph (run by ORT at
// cpp/helper.cpp:519). Output goes to a second buffer (neighbouring blocks still read this block's rows as their halo).
#pragma once
#include "kernels.cuh"

namespace stc {

struct DpBlockParams {
    const double* x; double* out;           // [rows, C] in / out (distinct buffers)
    const float* dw_w;                       // [C][K] as in the graph
    const float* dw_b; const float* ln_g; const float* ln_b;
    const float* w1; const float* b1;       // [C][H], [H]
    const float* w2; const float* b2;       // [H][C], [C]
    const float* gamma;                      // [C]
    const float* mask;                       // per row or null
    const int* off; int B; int rows;
    int K, pad_left;
    float eps;
};

template <int C, int H>
struct DpTile {
    static constexpr int R = 32, KMAX = 7;
    static constexpr int XS = (R + KMAX - 1) * C, AS = R * (C + 1), HS = R * (H + 1);
    static constexpr size_t SMEM = (size_t)(XS + AS + HS) * sizeof(double) + 3 * R * sizeof(int);
};

template <int C, int H>
__global__ void __launch_bounds__(256)
dp_convnext_kernel(const DpBlockParams p) {
    pdl_wait(); pdl_trigger_light();
    using TL = DpTile<C, H>;
    constexpr int R = TL::R;
    extern __shared__ __align__(16) unsigned char dp_smem[];
    double* xs = reinterpret_cast<double*>(dp_smem);                 // [(R + K - 1)][C]   input rows r0 - pad_left ...
    double* as = xs + TL::XS;                                       // [R][C + 1]
    double* hs = as + TL::AS;                                       // [R][H + 1]
    int* slo = reinterpret_cast<int*>(hs + TL::HS);                  // per row: sequence bounds [lo, hi), or lo = -1 for padding rows
    int* shi = slo + R;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = blockIdx.x * R;
    const int K = p.K, halo = K - 1;

    if (tid < R) {
        const int row = r0 + tid;
        int lo = -1, hi = -1;
        if (row < p.rows) {
            const int b = find_seq(p.off, p.B, row);
            if (b >= 0) { lo = __ldg(p.off + b); hi = __ldg(p.off + b + 1); }
        }
        slo[tid] = lo; shi[tid] = hi;
    }
    for (int i = tid; i < (R + halo) * (C / 2); i += 256) {
        const int rr = i / (C / 2), c2 = i % (C / 2), row = r0 - p.pad_left + rr;
        double2 v = make_double2(0.0, 0.0);
        if (row >= 0 && row < p.rows) v = *reinterpret_cast<const double2*>(p.x + (size_t)row * C + 2 * c2);
        *reinterpret_cast<double2*>(xs + rr * C + 2 * c2) = v;
    }
    __syncthreads();

    // ---- depthwise conv + LayerNorm: one warp per row, lane owns channels lane, lane + 32, ...
    {
        constexpr int CPL = C / 32;
        for (int rl = warp; rl < R; rl += 8) {
            const int row = r0 + rl, lo = slo[rl], hi = shi[rl];
            double y[CPL];
            if (lo < 0) {
#pragma unroll
                for (int i = 0; i < CPL; ++i) as[rl * (C + 1) + lane + 32 * i] = 0.0;
                continue;
            }
#pragma unroll
            for (int i = 0; i < CPL; ++i) y[i] = (double)__ldg(p.dw_b + lane + 32 * i);
            for (int k = 0; k < K; ++k) {
                const int rk = row + k - p.pad_left;
                if (rk < lo || rk >= hi) continue;
#pragma unroll
                for (int i = 0; i < CPL; ++i) {
                    const int c = lane + 32 * i;
                    y[i] += (double)__ldg(p.dw_w + c * K + k) * xs[(rl + k) * C + c];
                }
            }
            double s = 0;
#pragma unroll
            for (int i = 0; i < CPL; ++i) s += y[i];
            const double mean = warp_sum<double>(s) / (double)C;
            double v = 0;
#pragma unroll
            for (int i = 0; i < CPL; ++i) { y[i] -= mean; v += y[i] * y[i]; }
            const double den = sqrt(warp_sum<double>(v) / (double)C + (double)p.eps);
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int c = lane + 32 * i;
                as[rl * (C + 1) + c] = y[i] / den * (double)__ldg(p.ln_g + c) + (double)__ldg(p.ln_b + c);
            }
        }
    }
    __syncthreads();

    // ---- hid = GELU(a W1 + b1): thread = RM1 rows x 4 hidden columns
    {
        constexpr int TX = H / 4, TY = 256 / TX, RM = R / TY;
        static_assert(TX * TY == 256 && RM * TY == R, "pw1 thread tile");
        const int tx = tid % TX, ty = tid / TX;
        double acc[RM][4];
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll 4
        for (int k = 0; k < C; ++k) {
            const float4 wf = __ldg(reinterpret_cast<const float4*>(p.w1 + (size_t)k * H) + tx);
            const double w[4] = {(double)wf.x, (double)wf.y, (double)wf.z, (double)wf.w};
#pragma unroll
            for (int i = 0; i < RM; ++i) {
                const double a = as[(ty * RM + i) * (C + 1) + k];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a * w[j];
            }
        }
        const float4 bf = __ldg(reinterpret_cast<const float4*>(p.b1) + tx);
        const double b[4] = {(double)bf.x, (double)bf.y, (double)bf.z, (double)bf.w};
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) hs[(ty * RM + i) * (H + 1) + tx * 4 + j] = gelu_erf<double>(acc[i][j] + b[j]);
    }
    __syncthreads();

    // ---- out = ((hid W2 + b2) * gamma + x) * mask: thread = RM2 rows x 4 channels
    {
        constexpr int TX = C / 4, TY = 256 / TX, RM = R / TY;
        static_assert(TX * TY == 256 && RM * TY == R, "pw2 thread tile");
        const int tx = tid % TX, ty = tid / TX;
        double acc[RM][4];
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll 4
        for (int k = 0; k < H; ++k) {
            const float4 wf = __ldg(reinterpret_cast<const float4*>(p.w2 + (size_t)k * C) + tx);
            const double w[4] = {(double)wf.x, (double)wf.y, (double)wf.z, (double)wf.w};
#pragma unroll
            for (int i = 0; i < RM; ++i) {
                const double h = hs[(ty * RM + i) * (H + 1) + k];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += h * w[j];
            }
        }
        const float4 bf = __ldg(reinterpret_cast<const float4*>(p.b2) + tx), gf = __ldg(reinterpret_cast<const float4*>(p.gamma) + tx);
        const double b[4] = {(double)bf.x, (double)bf.y, (double)bf.z, (double)bf.w};
        const double g[4] = {(double)gf.x, (double)gf.y, (double)gf.z, (double)gf.w};
#pragma unroll
        for (int i = 0; i < RM; ++i) {
            const int rl = ty * RM + i, row = r0 + rl;
            if (row >= p.rows) continue;
            const double mk = p.mask ? (double)__ldg(p.mask + row) : 1.0;
            double o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double v = acc[i][j] + b[j];
                v *= g[j];
                v += xs[(rl + p.pad_left) * C + tx * 4 + j];
                if (p.mask) v *= mk;
                o[j] = v;
            }
            *reinterpret_cast<double2*>(p.out + (size_t)row * C + tx * 4) = make_double2(o[0], o[1]);
            *reinterpret_cast<double2*>(p.out + (size_t)row * C + tx * 4 + 2) = make_double2(o[2], o[3]);
        }
    }
}

}  // namespace stc
