// Host-side plan of libsupertonic_cuda: weights in device buffers, workspace arena, the layer walkers
// for the four graphs, CUDA-graph cache. See model.cu for the implementation.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "../../include/supertonic_cuda.h"
#include "kernels.cuh"
#include "onnx_reader.h"

namespace stc {

struct StcError : std::runtime_error {
    int code;
    StcError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define STC_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            throw ::stc::StcError(STC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
    } while (0)

// A GEMM A-operand (or a GEMM output that feeds the next GEMM): split bf16 pair in tensor-core mode,
// plain fp32 in the CUDA-core debug mode.
struct Act {
    float* f = nullptr;
    __nv_bfloat16* hi = nullptr;
    __nv_bfloat16* lo = nullptr;
};

struct Linear {
    int K = 0, N = 0;
    float* w_kn = nullptr;            // [K,N] fp32, as stored in the ONNX MatMul initializer
    float* bias = nullptr;            // [N]
    __nv_bfloat16* w_hi = nullptr;    // [N,K] K-major split pair (tensor-core B operand)
    __nv_bfloat16* w_lo = nullptr;
    bool f16 = false;                 // w_hi holds single fp16 weights [N,K], w_lo is null (vocoder, default STC_VOC=f16 mode)
    bool has_maps = false;            // w_hi / w_lo (or w_nk) present (tensor maps are cached per box shape in the handle)
};

struct ConvNeXt {
    int C, H, K, dil, pad_left;
    bool masked;
    float *dw_w, *dw_wt, *dw_b, *ln_g, *ln_b, *gamma;     // dw_w [C][K] as in the graph, dw_wt [K][C]
    Linear pw1, pw2;
};

enum { CTX_SELF = 0, CTX_TEXT = 1, CTX_STYLE = 2 };
enum { ROPE_NONE = 0, ROPE_ABS = 1, ROPE_NORM = 2 };

struct Attention {
    int C, heads, ctx_dim, ctx_kind, rope;
    bool masked, key_masked;
    float *ln_g, *ln_b, *freqs;
    Linear q, k, v, o;
    int kv_slot = -1;                 // index into the per-call K/V cache (cross-attention only)
};

enum LayerType { L_CONVNEXT, L_ATTN, L_TIME_COND, L_PROJ_IN, L_PROJ_OUT, L_TIME_MLP, L_CONV_IN, L_HEAD };
struct Layer { LayerType type; int idx; };

struct Net {
    std::vector<Layer> layers;
    std::vector<ConvNeXt> cn;
    std::vector<Attention> at;
    std::vector<Linear> lin;          // time_cond / proj_in / proj_out / ...
    std::map<std::string, float*> vec; // misc fp32 vectors by name
    int C = 0, H = 0, heads = 0;
};

class Arena {
public:
    void reserve(size_t bytes);
    void* alloc(size_t bytes);
    void reset() { off_ = 0; }
    void rewind(size_t off) { off_ = off; }
    size_t used() const { return off_; }
    size_t capacity() const { return cap_; }
    size_t high_water = 0;
    Arena() = default;
    Arena(const Arena&) = delete;
    Arena& operator=(const Arena&) = delete;
    void swap(Arena& o) { std::swap(base_, o.base_); std::swap(cap_, o.cap_); std::swap(off_, o.off_); std::swap(high_water, o.high_water); }
    ~Arena();
private:
    char* base_ = nullptr;
    size_t cap_ = 0, off_ = 0;
};

// B variable-length sequences stored back to back (kernels.cuh "packed sequences").
struct Seq {
    const int* off = nullptr;    // device [B+1] row offsets
    int B = 0;
    int rows = 0;                // rows launched (>= off[B] when a graph bucket pads)
    int maxlen = 0;              // upper bound of any sequence length (attention grid)
    const float* len = nullptr;  // device [B] float lengths (length-aware RoPE); sum of the mask in rectangle mode
    const int* cnt = nullptr;    // device [B]: 1 + last unmasked index (lets attention skip a masked tail); null = all
    const float* mask = nullptr; // device per-row 0/1 mask; null when every row is valid (packed mode)
};

struct Handle;
}  // namespace stc

struct stc_handle {
    std::unique_ptr<stc::Handle> impl;
    std::string last_error;
};
