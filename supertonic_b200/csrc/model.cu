// libsupertonic_cuda — weights, workspace, layer walkers and the C ABI (include/supertonic_cuda.h).
//
// What it replaces in the reference: the four Ort::Session objects (cpp/helper.cpp:784-795) and their
// Run calls inside TextToSpeech::_infer (:512-523 DP, :545-556 TE, :620-647 VE step, :662-672 vocoder),
// plus the host-side latent bookkeeping between them (:424-467, :590-659).
#include <array>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <nlohmann/json.hpp>

#include "gemm_tc.cuh"
#include "attn_tc.cuh"
#include "mlp_tc.cuh"
#include "mlp_stream.cuh"
#include "gemm2_tc.cuh"
#include "gemm2_astat.cuh"
#include "model.cuh"
#include "graph_plan.h"
#include "dp_fused.cuh"
#include "dwconv_chain.cuh"
#include "text_frontend.h"

using json = nlohmann::json;

namespace stc {

// Stream captures and the "potentially unsafe" runtime calls (cudaMalloc / cudaFree / cudaMallocHost / device-wide syncs) of
// different handles must not overlap: a capture in one thread is invalidated ("operation not permitted when stream is
// capturing") by such a call in another. One process-wide lock around both; graph REPLAYS do not take it.
static std::recursive_mutex g_capture_mu;

// ------------------------------------------------------------------------------------------ arena
Arena::~Arena() { if (base_) cudaFree(base_); }
void Arena::reserve(size_t bytes) {
    if (bytes <= cap_) return;
    std::lock_guard<std::recursive_mutex> lk(g_capture_mu);
    if (base_) { cudaDeviceSynchronize(); cudaFree(base_); base_ = nullptr; }
    size_t want = bytes + (bytes >> 3) + (1u << 20);
    cudaError_t e = cudaMalloc((void**)&base_, want);
    if (e != cudaSuccess) { cap_ = 0; throw StcError(STC_ERR_CUDA, std::string("workspace cudaMalloc: ") + cudaGetErrorString(e)); }
    cap_ = want;
}
void* Arena::alloc(size_t bytes) {
    size_t a = (off_ + 1023) & ~size_t(1023);
    off_ = a + bytes;
    if (off_ > high_water) high_water = off_;
    if (!base_ || off_ > cap_) return reinterpret_cast<void*>(uintptr_t(0x1000) + a);   // measuring pass: never dereferenced
    return base_ + a;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct GraphKey {
    int stage, mode, B, T, rows, maxlen, steps; int64_t noise_ld; uintptr_t p0, p1;
    int trows = 0, tmaxlen = 0;         // packed text side: launched rows / longest-sequence bound (0: rectangle)
    // device-resident entry points: the graph bakes the caller's four input pointers (ids, mask, style_ttl, style_dp) — all of
    // them are part of the key, unhashed (a folded key let two distinct pointer sets replay each other's graph)
    std::array<uintptr_t, 4> in{};
    int64_t gap = 0;                    // silence samples between utterances baked into the output-packing kernel
    bool operator<(const GraphKey& o) const {
        return std::tie(stage, mode, B, T, rows, maxlen, steps, noise_ld, p0, p1, trows, tmaxlen, in, gap) <
               std::tie(o.stage, o.mode, o.B, o.T, o.rows, o.maxlen, o.steps, o.noise_ld, o.p0, o.p1, o.trows, o.tmaxlen, o.in, o.gap);
    }
};

// Keys / values of one attention layer, ready for the core: either fp32 [rows, C] (K already rotated; CUDA-core path) or the
// tensor-core operands — K as split-bf16 [rows, C], V transposed per (sequence, head) to [B*heads*64, ldk] split-bf16.
struct KV {
    const float* K = nullptr; const float* V = nullptr;
    __nv_bfloat16 *k_hi = nullptr, *k_lo = nullptr, *vt_hi = nullptr, *vt_lo = nullptr;
    int ldk = 0;
    bool tc = false;
};

// What follows a ConvNeXt block in the graph and can ride in its final (reduce) kernel instead of a launch of its own.
struct PostOps {
    const float* add_vec = nullptr;                      // x <- (x + add_vec) * mask   (time conditioning)
    const float* ln_g = nullptr; const float* ln_b = nullptr;   // with `out`: LayerNorm(x); without ln_g: x itself
    const Act* out = nullptr;                            // split-bf16 operand of the next layer
};

struct VeCtx {
    Seq lat, text, style;             // latent frames (packed or rectangle), text tokens [B,T], style tokens [B,S]
    std::vector<KV> kv;               // per cross-attention layer: step-invariant, hoisted out of the Euler loop
};

struct MlpPlan { int nslice; int mode; };      // mode 0: single CTAs; 1: CTA pairs + an odd last tile as single CTAs; 2: pairs, odd tile padded

struct Handle {
    int device = 0;
    int num_sms = 148;
    int precision = STC_PREC_BF16X3;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;    // the text encoder runs here, concurrently with the duration predictor on `stream`
    cudaEvent_t ev_in = nullptr, ev_te = nullptr;
    // asynchronous result copies (stc_synthesize_packed_async): a third stream carries the device->host copy of the waveform
    // while the main stream already runs the next call; two alternating device result buffers and staging halves
    cudaStream_t stream_copy = nullptr;
    // Request streams (stc_synthesize_packed_async): uploads + duration predictor of call k+1 run on stream_f
    // with their own workspace (arena_f) and the text encoder on stream2 WHILE the Euler loop / vocoder of call k still occupy the
    // main stream; the stage-1 buffers (inputs, durations, text_emb) of odd calls live in persist_alt, so that call k's stage 2 and
    // call k+1's stage 1 never share memory. The host's wait for the durations of call k+1 then ends while call k is still running.
    cudaStream_t stream_f = nullptr;
    bool overlap = true;                     // (always on; kept as a member so that a debugger can serialise the stages)
    cudaEvent_t ev_out = nullptr, copy_done[2] = {nullptr, nullptr};
    float* outbuf[2] = {nullptr, nullptr}; size_t outcap[2] = {0, 0};
    int slot = 0; bool async_pending = false;
    size_t h_stage_lim = 0;
    void wait_async();
    stc_config cfg{};
    Net dp, te, ve, voc;
    json dp_arch, te_arch, ve_arch, voc_arch;
    std::vector<void*> owned;          // weight allocations
    Arena arena;       // workspace: reset per stage
    Arena arena2;      // workspace of whatever runs on stream2 (swapped in for the duration of that stage)
    Arena persist;     // buffers that survive from stage 1 (DP/TE) into stage 2 (VE loop + vocoder)
    Arena arena_f, persist_alt;        // see stream_f
    bool dry = false;          // no launches / copies (workspace measuring pass, or re-staging before a graph replay)
    bool restage = false;      // dry, but offset arrays are still written into their pinned staging slots
    bool capturing = false;    // launches go into a stream capture; host->device copies of caller memory are deferred
    struct PreCopy { void* dst; const void* src; size_t bytes; };
    std::vector<PreCopy> pre_copies;
    void run_graphed(const GraphKey& key, const std::function<void()>& body, cudaEvent_t after_uploads = nullptr);
    uint64_t graph_replays = 0, graph_captures = 0;
    uint64_t launches = 0;
    // which kernel variant each size-dependent dispatcher chose (issued or captured launches since creation): lets a test assert
    // that the variants running at benchmark scale are the ones it compared with the oracle (stc_kernel_variants)
    std::map<std::string, uint64_t> variant_count;
    void note(const char* name) { if (!dry) ++variant_count[name]; }
    EncodeTiledFn encode = nullptr;
    std::map<std::tuple<const void*, int, int, int>, CUtensorMap> map_cache;
    bool use_graphs = true;
    bool force_simt_attn = false;     // env STC_ATTN=simt: keep the CUDA-core attention core (cross-check)
    int profile = 0;          // 0 off, 1 stage events, 2 + per-kernel events for the GEMM / dwconv+LN classes
    struct KProf { double ms = 0, flops = 0, bytes = 0; uint64_t n = 0; };
    KProf kprof[6];           // 0 = tcgen05 GEMM (split-bf16), 1 = dwconv+LayerNorm, 2 = attention core, 3 = fused ConvNeXt MLP (+ its reduce),
                              // 4 = tcgen05 GEMM, single-pass fp16 operands (vocoder), 5 = dwconv+LayerNorm launches of >= 32 MB (the vocoder's:
                              // HBM-resident; class 1 keeps the L2-resident ones)
    struct Pending { int cls; cudaEvent_t a, b; double flops, bytes; };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> ev_pool; size_t ev_next = 0;
    cudaEvent_t pool_event() {
        if (ev_next == ev_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); ev_pool.push_back(e); }
        return ev_pool[ev_next++];
    }
    void kprof_begin(int cls, double flops, double bytes) {
        if (profile < 2 || dry) return;
        Pending p{cls, pool_event(), pool_event(), flops, bytes};
        cudaEventRecord(p.a, stream); pending.push_back(p);
    }
    void kprof_end() { if (profile < 2 || dry) return; cudaEventRecord(pending.back().b, stream); }
    void kprof_resolve() {
        for (auto& p : pending) { float ms = 0; cudaEventElapsedTime(&ms, p.a, p.b); auto& k = kprof[p.cls]; k.ms += ms; k.flops += p.flops; k.bytes += p.bytes; k.n++; }
        pending.clear(); ev_next = 0;
    }
    float stage_ms[5] = {0, 0, 0, 0, 0};
    cudaEvent_t ev[7] = {};   // start, dp, te, ve, vocoder, end, duration-ready
    struct GraphEntry { cudaGraphExec_t exec; uint64_t kernels; };
    std::map<GraphKey, GraphEntry> graphs;
    std::map<GraphKey, size_t> ws_need;
    TextFrontend frontend;
    // persistent small device buffers for the fast layer
    float* d_dtvec = nullptr; int dtvec_steps = -1;
    // pinned staging
    float* h_dur = nullptr; int64_t* h_wavlen = nullptr; int h_cap = 0;

    ~Handle();
    // ---- loading
    void load(const std::string& onnx_dir);
    float* upload_f32(const float* p, size_t n);
    float* W(const OnnxFile& f, const std::string& name, size_t numel);
    Linear make_linear(const OnnxFile& f, const std::string& wname, const std::string& bname, int K, int N, int tc);
    Linear make_linear_host(const std::vector<float>& w_kn, const std::vector<float>& bias, int K, int N, int tc);
    void load_net(const OnnxFile& f, const json& arch, Net& net, int tc);
    ConvNeXt load_convnext(const OnnxFile& f, const json& l, int tc);
    Attention load_attention(const OnnxFile& f, const json& l, int tc);
    CUtensorMap encode_map(const void* ptr, int rows, int K, int box_rows);
    const CUtensorMap& tmap(const void* ptr, int rows, int K, int box_rows);
    const CUtensorMap& tmap_f32(const void* ptr, int rows, int cols);
    int pick_gemm(int M, int N, int K, bool f16) const;      // tile width 64 / 128 / 256, or 512 = the two-SM 256 x 256 form
    int force_bn = 0;                // debug / sweep override (0: heuristic)

    // ---- workspace helpers
    template <typename T> T* ws(size_t n) { return static_cast<T*>(arena.alloc(n * sizeof(T))); }
    bool tc_mode() const { return precision == STC_PREC_BF16X3; }
    Act ws_act_f16(size_t n) { Act a; a.hi = ws<__nv_bfloat16>(n); return a; }   // single fp16 operand (hi holds fp16 bits, lo stays null)
    Act ws_act_for(const Linear& w, size_t n) { return w.f16 ? ws_act_f16(n) : ws_act(n); }
    Act ws_act(size_t n) {
        Act a;
        if (tc_mode()) { a.hi = ws<__nv_bfloat16>(n); a.lo = ws<__nv_bfloat16>(n); } else a.f = ws<float>(n);
        return a;
    }
    template <typename T> T* ps(size_t n) { return static_cast<T*>(persist.alloc(n * sizeof(T))); }
    size_t mark() { return arena.used(); }
    void release(size_t m) { arena.rewind(m); }

    // ---- kernels
    void to_act(const float* x, size_t n, const Act& out);
    template <typename T> void dwconv_ln(const T* x, const ConvNeXt* cn, const float* g, const float* b, int C,
                                         const Seq& seq, float eps, T* out_plain, const Act* out_act);
    void gemm(const Act& a, int M, const Linear& w, const Epilogue& ep, float* out_f32, const Act* out_act, int ldo);
    template <typename T> void gemm_simt(const T* a, int lda, int M, const Linear& w, const Epilogue& ep, T* out, int ldo);
    template <typename T> void convnext(const ConvNeXt& c, T* x, const Seq& seq, const PostOps* post = nullptr);
    void apply_post(const PostOps& post, float* x, const Seq& seq, int C);      // the same post-ops as separate launches
    bool mlp_fused(const ConvNeXt& c) const;
    MlpPlan mlp_plan(int tiles, int rows) const;
    void fused_mlp(const Act& a, int rows, const ConvNeXt& c, float* x, const float* mask, const PostOps* post = nullptr);
    int pdl_mode = 2;                 // env STC_PDL (launch_k): 0 = plain stream-ordered launches (cross-check), 2 / 3 = release point (kernels.cuh)
    int mlp_pair = -1;                // env STC_MLP_PAIR: 0 = one-CTA stream kernel only (cross-check), default: CTA pairs where the slices allow
    int mlp_force_slices = 0;         // env STC_MLP_SLICES (tools/mlp_sweep.py): hidden slices per row tile instead of the cost model
    bool gemm_astat = true;           // env STC_ASTAT=0: the vocoder's pw1 keeps the streaming two-SM kernel (cross-check of gemm2_astat.cuh)
    bool mlp_unfused = false;         // env STC_MLP=unfused: the C = 256 / H = 1024 blocks as two tcgen05 GEMMs (cross-check)
    long long* gemm_trace = nullptr;  // stc_debug_gemm with STC_GEMM_TRACE=1
    long long* mlp_trace = nullptr;   // stc_debug_mlp with STC_MLP_TRACE=1
    bool dp_fused = true;             // env STC_DP=unfused: the duration predictor's ConvNeXt blocks as separate fp64 conv / GEMM launches (cross-check)
    bool voc_f16 = true;              // vocoder GEMMs single-pass fp16 (default; env STC_VOC=bf16x3 keeps the split-bf16 form there too)
    bool dw_chain_kernel = true;      // env STC_DW=slide: long chains keep the two-pass sliding-window kernel with the ring (cross-check)
    bool dw_slide = true;             // env STC_DW=tile: shared-memory tiled depthwise conv + LayerNorm instead of the register sliding window
    int dw_rt = 0;                    // env STC_DW_RT: rows per chain of the sliding-window kernel (0: heuristic)
    int dw_ring = -1;                 // env STC_DW_RING: 1 / 0 force the shared-memory ring prefetch on / off (-1: by chain length)
    void attention(const Attention& a, float* x, const Seq& q, const Act* ctx, const Seq& k, const KV* pre, const Act* xn_pre = nullptr);
    void attn_core(const float* Q, const float* K, const float* V, const Act& out, const Seq& q, const Seq& k, bool key_masked,
                   int heads, int dh);
    bool attn_on_tc(const Attention& a, const Seq& ks) const;
    KV make_kv(const Attention& a, const Act& ctx, const Seq& ks);      // result lives in the arena above the caller's mark
    void attn_core_tc(const Act& Q, const Attention& a, const KV& kv, const Act& out, const Seq& q, const Seq& k);
    Epilogue rope_epilogue(const Attention& a, const Seq& seq) const;
    void rope(float* x, const float* freqs, const Seq& seq, int heads, int dh, int normalise);
    // sequence descriptors (offsets staged through pinned memory)
    int* stage_ints(const std::vector<int>& v);
    Seq rect_seq(int B, int N, const float* mask, bool want_len);
    Seq packed_seq(const std::vector<int>& lens, int rows_launch, int maxlen_launch);
    Seq scaled_seq(const Seq& s, int f);
    const float* time_vectors(float cur, float tot);

    // ---- graph walkers (device pointers)
    // `seq`: the text tokens as packed rows (fast layer: only real tokens) or as the reference's [B,T] rectangle + row mask
    void run_dp(const int64_t* ids, const float* style_dp, const Seq& seq, int T, float* dur);
    void run_te(const int64_t* ids, const float* style_ttl, const Seq& seq, int T, float* text_emb_cl);
    void prepare_ve(VeCtx& vc, const float* text_emb_cl, const float* style_ttl);
    void run_ve_step(const VeCtx& vc, float* x_lat, const float* tvec, const float* dtvec);
    void run_vocoder(const float* lat_cl, const Seq& lat, float* wav);

    void check_launch(const char* what);
    void ensure_ws(const std::function<void()>& fn);
    void synth_tail(const float* d_text_emb, const Seq& text, const float* d_style_ttl, const float* d_noise, int64_t noise_ld,
                    uint64_t seed, const Seq& lat, int steps, float* d_xlat, float* d_wav, const int* d_noise_index = nullptr);
    std::map<std::pair<uint32_t, uint32_t>, float*> tvec_cache;
    int* h_stage = nullptr; size_t h_stage_cap = 0, h_stage_off = 0;   // pinned staging for offset arrays
};

// Every kernel of the library goes out through this (stream-ordered launches; inside a stream capture they become graph nodes).
// Programmatic dependent launch (default; env STC_PDL=0 for plain launches, 3 for no early release): the kernel may start while its
// predecessor still runs and waits for it itself (kernels.cuh: pdl_wait and the pre-wait rules). With every kernel releasing its
// dependents at its very top this lost twice (r1g 15.2 vs 14.3 ms/step, r2b 8.61 vs 8.50); releasing after the kernel's own wait —
// late in the tensor-core kernels — wins: end-to-end request stream 40.7 k -> 44.5 k audio-s/s, batch-1 p50 4.64 -> 4.2 ms
// (profiles/r2y_pdl_modes.txt).
template <typename... KArgs, typename... Args>
static inline void launch_k(stc::Handle* h, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = h->pdl_mode ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
    if (e != cudaSuccess) throw ::stc::StcError(STC_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
}

#define STC_LAUNCH(h, kernel, grid, block, smem, ...)                                       \
    do {                                                                                    \
        if (!(h)->dry) {                                                                    \
            launch_k((h), kernel, dim3(grid), dim3(block), (size_t)(smem), (h)->stream, __VA_ARGS__); \
            ++(h)->launches;                                                                \
        }                                                                                   \
    } while (0)

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

Handle::~Handle() {
    for (auto& g : graphs) cudaGraphExecDestroy(g.second.exec);
    for (void* p : owned) cudaFree(p);
    if (d_dtvec) cudaFree(d_dtvec);
    if (h_dur) cudaFreeHost(h_dur);
    if (h_wavlen) cudaFreeHost(h_wavlen);
    if (h_stage) cudaFreeHost(h_stage);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    for (auto& e : ev_pool) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
    if (stream2) cudaStreamDestroy(stream2);
    if (ev_in) cudaEventDestroy(ev_in);
    if (ev_out) cudaEventDestroy(ev_out);
    for (auto& e : copy_done) if (e) cudaEventDestroy(e);
    for (auto& p : outbuf) if (p) cudaFree(p);
    if (stream_copy) cudaStreamDestroy(stream_copy);
    if (stream_f) cudaStreamDestroy(stream_f);
    if (ev_te) cudaEventDestroy(ev_te);
}

void Handle::wait_async() {
    if (!async_pending) return;
    STC_CUDA(cudaStreamSynchronize(stream_copy));
    STC_CUDA(cudaStreamSynchronize(stream));
    if (stream_f) STC_CUDA(cudaStreamSynchronize(stream_f));
    if (stream2) STC_CUDA(cudaStreamSynchronize(stream2));
    async_pending = false;
    check_launch("asynchronous synthesis");
}

void Handle::check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw StcError(STC_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

// ------------------------------------------------------------------------------------------ loading
float* Handle::upload_f32(const float* p, size_t n) {
    float* d = nullptr;
    STC_CUDA(cudaMalloc((void**)&d, std::max<size_t>(n, 1) * sizeof(float)));
    owned.push_back(d);
    STC_CUDA(cudaMemcpy(d, p, n * sizeof(float), cudaMemcpyHostToDevice));
    return d;
}

static const OnnxTensor& get_tensor(const OnnxFile& f, const std::string& name, size_t numel) {
    auto it = f.initializers.find(name);
    if (it == f.initializers.end()) throw StcError(STC_ERR_IO, "initializer not found: " + name);
    if (it->second.dtype != 1) throw StcError(STC_ERR_UNSUPPORTED, "initializer " + name + " is not float32");
    if (numel && it->second.numel() != numel)
        throw StcError(STC_ERR_IO, "initializer " + name + ": expected " + std::to_string(numel) + " elements, file has " +
                                       std::to_string(it->second.numel()));
    return it->second;
}

float* Handle::W(const OnnxFile& f, const std::string& name, size_t numel) {
    const OnnxTensor& t = get_tensor(f, name, numel);
    return upload_f32(t.f32(), t.numel());
}

CUtensorMap Handle::encode_map(const void* ptr, int rows, int K, int box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)tc::BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    if (K % 8) throw StcError(STC_ERR_UNSUPPORTED, "tensor-core GEMM needs K % 8 == 0, got " + std::to_string(K));
    CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw StcError(STC_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
    return m;
}

// fp32 [rows, cols] row-major, box 32 x 32 (128-byte swizzled rows): destination of the fused MLP's TMA stores
const CUtensorMap& Handle::tmap_f32(const void* ptr, int rows, int cols) {
    auto key = std::make_tuple(ptr, rows, cols, -32);
    auto it = map_cache.find(key);
    if (it != map_cache.end()) return it->second;
    if (map_cache.size() > 16384) map_cache.clear();
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw StcError(STC_ERR_CUDA, "cuTensorMapEncodeTiled (fp32) failed: " + std::to_string((int)r));
    return map_cache.emplace(key, m).first->second;
}


const CUtensorMap& Handle::tmap(const void* ptr, int rows, int K, int box_rows) {
    auto key = std::make_tuple(ptr, rows, K, box_rows);
    auto it = map_cache.find(key);
    if (it != map_cache.end()) return it->second;
    if (map_cache.size() > 16384) map_cache.clear();
    return map_cache.emplace(key, encode_map(ptr, rows, K, box_rows)).first->second;
}

static inline uint16_t bf16_bits_rn(float v) {
    uint32_t u; memcpy(&u, &v, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static inline float bf16_to_f(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }

Linear Handle::make_linear_host(const std::vector<float>& w_kn, const std::vector<float>& bias, int K, int N, int tc) {
    Linear l; l.K = K; l.N = N;
    l.w_kn = upload_f32(w_kn.data(), w_kn.size());
    l.bias = upload_f32(bias.data(), bias.size());
    if (tc == 3) {   // single fp16 operand: [N,K] K-major, round to nearest, saturating
        std::vector<uint16_t> hf((size_t)N * K);
        for (int k = 0; k < K; ++k)
            for (int n = 0; n < N; ++n) {
                float v = std::min(65504.f, std::max(-65504.f, w_kn[(size_t)k * N + n]));
                __half hv = __float2half_rn(v);
                memcpy(&hf[(size_t)n * K + k], &hv, 2);
            }
        STC_CUDA(cudaMalloc((void**)&l.w_hi, hf.size() * 2)); owned.push_back(l.w_hi);
        STC_CUDA(cudaMemcpy(l.w_hi, hf.data(), hf.size() * 2, cudaMemcpyHostToDevice));
        if (K % 8) throw StcError(STC_ERR_UNSUPPORTED, "tensor-core GEMM needs K % 8 == 0, got " + std::to_string(K));
        if (N % 4) throw StcError(STC_ERR_UNSUPPORTED, "tensor-core GEMM needs N % 4 == 0, got " + std::to_string(N));
        l.f16 = true; l.has_maps = true;
    } else if (tc) {
        std::vector<uint16_t> hi((size_t)N * K), lo((size_t)N * K);
        for (int k = 0; k < K; ++k)
            for (int n = 0; n < N; ++n) {
                float v = w_kn[(size_t)k * N + n];
                uint16_t h = bf16_bits_rn(v);
                hi[(size_t)n * K + k] = h;
                lo[(size_t)n * K + k] = bf16_bits_rn(v - bf16_to_f(h));
            }
        STC_CUDA(cudaMalloc((void**)&l.w_hi, hi.size() * 2)); owned.push_back(l.w_hi);
        STC_CUDA(cudaMalloc((void**)&l.w_lo, lo.size() * 2)); owned.push_back(l.w_lo);
        STC_CUDA(cudaMemcpy(l.w_hi, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice));
        STC_CUDA(cudaMemcpy(l.w_lo, lo.data(), lo.size() * 2, cudaMemcpyHostToDevice));
        if (K % 8) throw StcError(STC_ERR_UNSUPPORTED, "tensor-core GEMM needs K % 8 == 0, got " + std::to_string(K));
        if (N % 4) throw StcError(STC_ERR_UNSUPPORTED, "tensor-core GEMM needs N % 4 == 0, got " + std::to_string(N));
        l.has_maps = true;
    }
    return l;
}

// Initializer name of role `key` of a layer (or of the graph): the name the node-pattern matcher found in that position
// (graph_plan.h, entry "t"), else the surrogate generator's naming scheme.
static std::string tn(const json& l, const char* key, const std::string& dflt) {
    auto it = l.find("t");
    if (it != l.end() && it->contains(key)) return it->at(key).get<std::string>();
    return dflt;
}

Linear Handle::make_linear(const OnnxFile& f, const std::string& wname, const std::string& bname, int K, int N, int tc) {
    const OnnxTensor& w = get_tensor(f, wname, (size_t)K * N);
    const OnnxTensor& b = get_tensor(f, bname, (size_t)N);
    return make_linear_host(std::vector<float>(w.f32(), w.f32() + w.numel()), std::vector<float>(b.f32(), b.f32() + b.numel()), K, N, tc);
}

ConvNeXt Handle::load_convnext(const OnnxFile& f, const json& l, int tc) {
    ConvNeXt c{};
    std::string p = l.at("name");
    c.C = l.at("C"); c.H = l.at("H"); c.K = l.at("K"); c.dil = l.at("dilation");
    bool causal = l.at("causal");
    c.masked = l.at("masked");
    int span = c.dil * (c.K - 1);
    c.pad_left = causal ? span : span / 2;
    c.dw_w = W(f, tn(l, "dw_w", p + ".dw.weight"), (size_t)c.C * c.K);
    {   // tap-major copy for the vectorised kernel
        const float* w = get_tensor(f, tn(l, "dw_w", p + ".dw.weight"), (size_t)c.C * c.K).f32();
        std::vector<float> wt((size_t)c.C * c.K);
        for (int ch = 0; ch < c.C; ++ch) for (int k = 0; k < c.K; ++k) wt[(size_t)k * c.C + ch] = w[(size_t)ch * c.K + k];
        c.dw_wt = upload_f32(wt.data(), wt.size());
    }
    c.dw_b = W(f, tn(l, "dw_b", p + ".dw.bias"), c.C);
    c.ln_g = W(f, tn(l, "ln_g", p + ".ln.weight"), c.C);
    c.ln_b = W(f, tn(l, "ln_b", p + ".ln.bias"), c.C);
    c.gamma = W(f, tn(l, "gamma", p + ".gamma"), c.C);
    c.pw1 = make_linear(f, tn(l, "w1", p + ".pw1.weight"), tn(l, "b1", p + ".pw1.bias"), c.C, c.H, tc);
    c.pw2 = make_linear(f, tn(l, "w2", p + ".pw2.weight"), tn(l, "b2", p + ".pw2.bias"), c.H, c.C, tc);
    return c;
}

Attention Handle::load_attention(const OnnxFile& f, const json& l, int tc) {
    Attention a{};
    std::string p = l.at("name");
    a.C = l.at("C"); a.heads = l.at("heads"); a.ctx_dim = l.at("ctx_dim");
    std::string ctx = l.at("ctx"), rope = l.at("rope");
    a.ctx_kind = ctx == "self" ? CTX_SELF : ctx == "text_emb" ? CTX_TEXT : ctx == "style_ttl" ? CTX_STYLE : -1;
    if (a.ctx_kind < 0) throw StcError(STC_ERR_UNSUPPORTED, "attention context " + ctx);
    a.rope = rope == "none" ? ROPE_NONE : rope == "abs" ? ROPE_ABS : ROPE_NORM;
    a.masked = l.at("masked"); a.key_masked = l.at("key_masked");
    int dh = a.C / a.heads;
    if (dh != 32 && dh != 64) throw StcError(STC_ERR_UNSUPPORTED, "attention head dim must be 32 or 64");
    a.ln_g = W(f, tn(l, "ln_g", p + ".ln.weight"), a.C);
    a.ln_b = W(f, tn(l, "ln_b", p + ".ln.bias"), a.C);
    a.freqs = a.rope != ROPE_NONE ? W(f, tn(l, "rope_freqs", p + ".rope_freqs"), dh / 2) : nullptr;
    if (a.rope != ROPE_NONE) {
        // rotary pairs (d, d + dh/2) of every head move to adjacent output columns (2i, 2i+1) of the Q and K projections
        // (kernels.cuh rope_kernel): Q.K^T does not change, and the GEMM epilogue can rotate inside one float4
        auto permuted = [&](const std::string& wname, const std::string& bname, int K) {
            const float* w = get_tensor(f, wname, (size_t)K * a.C).f32();
            const float* b = get_tensor(f, bname, (size_t)a.C).f32();
            std::vector<float> wp((size_t)K * a.C), bp(a.C);
            for (int n = 0; n < a.C; ++n) {
                const int h = n / dh, j = n % dh, src = h * dh + ((j & 1) ? dh / 2 + j / 2 : j / 2);
                bp[n] = b[src];
                for (int k = 0; k < K; ++k) wp[(size_t)k * a.C + n] = w[(size_t)k * a.C + src];
            }
            return make_linear_host(wp, bp, K, a.C, tc);
        };
        a.q = permuted(tn(l, "wq", p + ".q.weight"), tn(l, "bq", p + ".q.bias"), a.C);
        a.k = permuted(tn(l, "wk", p + ".k.weight"), tn(l, "bk", p + ".k.bias"), a.ctx_dim);
    } else {
        a.q = make_linear(f, tn(l, "wq", p + ".q.weight"), tn(l, "bq", p + ".q.bias"), a.C, a.C, tc);
        a.k = make_linear(f, tn(l, "wk", p + ".k.weight"), tn(l, "bk", p + ".k.bias"), a.ctx_dim, a.C, tc);
    }
    a.v = make_linear(f, tn(l, "wv", p + ".v.weight"), tn(l, "bv", p + ".v.bias"), a.ctx_dim, a.C, tc);
    a.o = make_linear(f, tn(l, "wo", p + ".o.weight"), tn(l, "bo", p + ".o.bias"), a.C, a.C, tc);
    return a;
}

void Handle::load_net(const OnnxFile& f, const json& arch, Net& net, int tc) {
    net.C = arch.value("C", 0); net.H = arch.value("H", 0); net.heads = arch.value("heads", 0);
    int kv = 0;
    for (const auto& l : arch.at("layers")) {
        std::string type = l.at("type");
        if (type == "convnext") { net.cn.push_back(load_convnext(f, l, tc)); net.layers.push_back({L_CONVNEXT, (int)net.cn.size() - 1}); }
        else if (type == "attention") {
            Attention a = load_attention(f, l, tc);
            if (a.ctx_kind != CTX_SELF) a.kv_slot = kv++;
            net.at.push_back(a); net.layers.push_back({L_ATTN, (int)net.at.size() - 1});
        } else if (type == "time_cond") {
            std::string p = l.at("name"); int C = l.at("C");
            net.lin.push_back(make_linear(f, tn(l, "w", p + ".weight"), tn(l, "b", p + ".bias"), C, C, false));
            net.layers.push_back({L_TIME_COND, (int)net.lin.size() - 1});
        } else if (type == "proj_in" || type == "proj_out") {
            std::string p = l.at("name"); int ci = l.at("cin"), co = l.at("cout");
            net.lin.push_back(make_linear(f, tn(l, "w", p + ".weight"), tn(l, "b", p + ".bias"), ci, co, tc));
            net.layers.push_back({type == "proj_in" ? L_PROJ_IN : L_PROJ_OUT, (int)net.lin.size() - 1});
        } else if (type == "time_mlp") {
            std::string p = l.at("name"); int td = l.at("time_dim"), C = l.at("C");
            net.vec["time.freqs"] = W(f, tn(l, "freqs", p + ".freqs"), td / 2);
            net.lin.push_back(make_linear(f, tn(l, "w1", p + ".fc1.weight"), tn(l, "b1", p + ".fc1.bias"), td, C, false));
            net.lin.push_back(make_linear(f, tn(l, "w2", p + ".fc2.weight"), tn(l, "b2", p + ".fc2.bias"), C, C, false));
            net.layers.push_back({L_TIME_MLP, (int)net.lin.size() - 2});
        } else if (type == "conv_in") {
            // Conv1d(ld -> C, K, causal) followed by eval-mode BatchNorm: fold BN into the conv (in double) and
            // express it as a [K*ld, C] linear over the im2col rows the front-end kernel writes.
            std::string p = l.at("name"), bn = l.at("bn");
            int ci = l.at("cin"), co = l.at("cout"), K = l.at("K");
            const float* w = get_tensor(f, tn(l, "w", p + ".weight"), (size_t)co * ci * K).f32();
            const float* b = get_tensor(f, tn(l, "b", p + ".bias"), co).f32();
            const float* g = get_tensor(f, tn(l, "bn_w", bn + ".weight"), co).f32();
            const float* be = get_tensor(f, tn(l, "bn_b", bn + ".bias"), co).f32();
            const float* mu = get_tensor(f, tn(l, "bn_mean", bn + ".running_mean"), co).f32();
            const float* var = get_tensor(f, tn(l, "bn_var", bn + ".running_var"), co).f32();
            std::vector<float> wk((size_t)K * ci * co), bb(co);
            for (int o = 0; o < co; ++o) {
                double s = (double)g[o] / std::sqrt((double)var[o] + 1e-5);
                bb[o] = (float)(((double)b[o] - (double)mu[o]) * s + (double)be[o]);
                for (int c = 0; c < ci; ++c)
                    for (int k = 0; k < K; ++k) wk[(size_t)(k * ci + c) * co + o] = (float)((double)w[((size_t)o * ci + c) * K + k] * s);
            }
            net.lin.push_back(make_linear_host(wk, bb, K * ci, co, tc));
            net.layers.push_back({L_CONV_IN, (int)net.lin.size() - 1});
        } else if (type == "head") {
            std::string p = l.at("name"); int ci = l.at("cin"), co = l.at("cout");
            net.vec["head.ln_g"] = W(f, tn(l, "ln_g", p + ".ln.weight"), ci);
            net.vec["head.ln_b"] = W(f, tn(l, "ln_b", p + ".ln.bias"), ci);
            net.lin.push_back(make_linear(f, tn(l, "w", p + ".proj.weight"), tn(l, "b", p + ".proj.bias"), ci, co, tc));
            net.layers.push_back({L_HEAD, (int)net.lin.size() - 1});
        } else throw StcError(STC_ERR_UNSUPPORTED, "layer type " + type);
    }
}

// Layer plan of a graph: the `stc_arch` metadata the surrogate generator writes (fast path), else derived from the NODES
// (graph_plan.h) — what a released export carries. STC_IGNORE_ARCH=1 forces the derivation (tests).
static json arch_of(const OnnxFile& f, const std::string& file, const std::string& kind) {
    auto it = f.metadata.find("stc_arch");
    const char* ig = getenv("STC_IGNORE_ARCH");
    if (it != f.metadata.end() && !(ig && ig[0] == '1')) return json::parse(it->second);
    try { return derive_arch(f, kind); }
    catch (const PlanError& e) { throw StcError(STC_ERR_UNSUPPORTED, file + ": " + e.what()); }
}

void Handle::load(const std::string& onnx_dir) {
    {   // tts.json (reference loadCfgs, cpp/helper.cpp:801-818)
        std::string p = onnx_dir + "/tts.json";
        std::ifstream file(p);
        if (!file.is_open()) throw StcError(STC_ERR_IO, "Failed to open config file: " + p);
        json j; file >> j;
        cfg.sample_rate = j["ae"]["sample_rate"]; cfg.base_chunk_size = j["ae"]["base_chunk_size"];
        cfg.chunk_compress_factor = j["ttl"]["chunk_compress_factor"]; cfg.latent_dim = j["ttl"]["latent_dim"];
        cfg.latent_channels = cfg.latent_dim * cfg.chunk_compress_factor;
        cfg.chunk_size = cfg.base_chunk_size * cfg.chunk_compress_factor;
    }
    frontend.load_indexer(onnx_dir + "/unicode_indexer.json");
    bool tc = tc_mode();
    {
        OnnxFile f = load_onnx(onnx_dir + "/duration_predictor.onnx");
        dp_arch = arch_of(f, "duration_predictor.onnx", "duration_predictor");
        load_net(f, dp_arch, dp, false);
        cfg.vocab_size = dp_arch.at("vocab");
        dp.vec["embed"] = W(f, tn(dp_arch, "embed", "dp.embed.weight"), (size_t)cfg.vocab_size * dp.C);
        int si = dp_arch.at("style_in");
        {   // style_dp[B, e1, e2]: the split comes from the graph's input signature (the reference reads it from the voice-style
            // JSON, cpp/helper.cpp:856-861, and ORT checks it against the graph)
            const OnnxValueInfo* v = f.input("style_dp");
            if (!v || v->dims.size() != 3 || v->dims[1] <= 0 || v->dims[2] <= 0 || v->dims[1] * v->dims[2] != si)
                throw StcError(STC_ERR_UNSUPPORTED, "duration_predictor.onnx: input style_dp must be [B, e1, e2] with e1*e2 = " + std::to_string(si));
            cfg.style_dp_tokens = (int)v->dims[1]; cfg.style_dp_dim = (int)v->dims[2];
        }
        dp.lin.push_back(make_linear(f, tn(dp_arch, "style_w", "dp.style.weight"), tn(dp_arch, "style_b", "dp.style.bias"), si, dp.C, false));
        dp.vec["head.ln_g"] = W(f, tn(dp_arch, "head_ln_g", "dp.head.ln.weight"), dp.C);
        dp.vec["head.ln_b"] = W(f, tn(dp_arch, "head_ln_b", "dp.head.ln.bias"), dp.C);
        dp.vec["head.w"] = W(f, tn(dp_arch, "head_w", "dp.head.proj.weight"), dp.C);
        dp.vec["head.b"] = W(f, tn(dp_arch, "head_b", "dp.head.proj.bias"), 1);
    }
    {
        OnnxFile f = load_onnx(onnx_dir + "/text_encoder.onnx");
        te_arch = arch_of(f, "text_encoder.onnx", "text_encoder");
        load_net(f, te_arch, te, tc);
        te.vec["embed"] = W(f, tn(te_arch, "embed", "te.embed.weight"), (size_t)cfg.vocab_size * te.C);
        cfg.text_emb_channels = te.C;
        cfg.style_ttl_tokens = te_arch.at("n_style"); cfg.style_ttl_dim = te_arch.at("style_dim");
        const OnnxValueInfo* v = f.input("style_ttl");
        if (v && v->dims.size() == 3 && v->dims[1] > 0 && v->dims[2] > 0 && (v->dims[1] != cfg.style_ttl_tokens || v->dims[2] != cfg.style_ttl_dim))
            throw StcError(STC_ERR_UNSUPPORTED, "text_encoder.onnx: input style_ttl dims do not match the layer plan");
    }
    {
        OnnxFile f = load_onnx(onnx_dir + "/vector_estimator.onnx");
        ve_arch = arch_of(f, "vector_estimator.onnx", "vector_estimator");
        load_net(f, ve_arch, ve, tc);
        if ((int)ve_arch.at("latent_ch") != cfg.latent_channels) throw StcError(STC_ERR_IO, "vector_estimator latent_ch != tts.json");
    }
    {
        OnnxFile f = load_onnx(onnx_dir + "/vocoder.onnx");
        voc_arch = arch_of(f, "vocoder.onnx", "vocoder");
        load_net(f, voc_arch, voc, tc ? (voc_f16 ? 3 : 1) : 0);
        voc.vec["std"] = W(f, tn(voc_arch, "latent_std", "voc.latent_std"), cfg.latent_channels);
        voc.vec["mean"] = W(f, tn(voc_arch, "latent_mean", "voc.latent_mean"), cfg.latent_channels);
    }
}

// ------------------------------------------------------------------------------------------ kernels (host side)
void Handle::to_act(const float* x, size_t n, const Act& out) {
    if (tc_mode()) STC_LAUNCH(this, (convert_kernel<OutSplit>), cdiv(n, 256), 256, 0, x, OutSplit{out.hi, out.lo}, n);
    else STC_LAUNCH(this, (convert_kernel<OutPlain<float>>), cdiv(n, 256), 256, 0, x, OutPlain<float>{out.f}, n);
}

template <typename T, typename Out>
static void launch_dwln(Handle* h, int C, const T* x, const float* w, const float* wb, const float* g, const float* b, Out out,
                        int rows, const int* off, int B, int K, int dil, int pad, float eps) {
    dim3 grid(cdiv(rows, 8)), block(256);
    if constexpr (std::is_same<T, float>::value) {
        // w is the tap-major transpose wT[K][C] for these widths (ConvNeXt::dw_wt)
        // (measured: the shared-memory tiles also beat the direct warp-per-row kernel on the L2-resident VE / TE tensors:
        //  12.97 vs 13.27 ms/step)
        if (h->dw_slide && (K == 5 || K == 7) && (C == 128 || C == 256 || C == 512) && dil >= 1) {
            // register sliding window along chains of RT rows (kernels.cuh): RT as long as every SM still gets >= ~1024 threads
            const int GT = C / 4;
            int RT = h->dw_rt;
            if (RT <= 0) {
                // long chains (shared-memory ring prefetch) when one wave of resident chains (4 blocks of 128 threads per SM) covers
                // the rows with >= 16 rows per chain, else chains of 4 rows (one iteration: latency bound, as many threads as possible)
                const int resident = h->num_sms * 4 * (512 / C);
                RT = (int)cdiv(cdiv(rows, resident), 4) * 4;
                if (RT < 16) RT = 4;
            }
            RT = std::max(4, RT / 4 * 4);
            const bool ring = h->dw_ring < 0 ? RT >= 16 : h->dw_ring > 0;
            // long chains: the one-pass kernel of dwconv_chain.cuh (K - 1 rows per iteration; chain length a multiple of that)
            if (ring && h->dw_chain_kernel && h->dw_rt <= 0) RT = (int)cdiv(RT, 4 * (K - 1)) * 4 * (K - 1);
            const bool chain_k = ring && h->dw_chain_kernel && RT % (K - 1) == 0;
            const unsigned chains = cdiv(rows, RT * dil) * dil;
            dim3 sg(cdiv(chains, 512 / C));
            if (chain_k) {
                h->note("dwconv_ln_chain");
#define STC_CHAIN(NW, KK) STC_LAUNCH(h, (dwconv_ln_chain_kernel<NW, KK, Out>), sg, 128, ChainSmem<KK>::BYTES, x, w, wb, g, b, out, rows, off, B, dil, pad, eps, RT); return
                switch (C / 128 * 10 + K) {
                    case 15: STC_CHAIN(1, 5);
                    case 17: STC_CHAIN(1, 7);
                    case 25: STC_CHAIN(2, 5);
                    case 27: STC_CHAIN(2, 7);
                    case 45: STC_CHAIN(4, 5);
                    case 47: STC_CHAIN(4, 7);
                }
#undef STC_CHAIN
            }
#define STC_SLIDE(NW, KK)                                                                                                                     \
    do {                                                                                                                                      \
        h->note(ring ? "dwconv_ln_slide_ring" : "dwconv_ln_slide");                                                                          \
        if (ring) STC_LAUNCH(h, (dwconv_ln_slide_kernel<NW, KK, true, Out>), sg, 128, 0, x, w, wb, g, b, out, rows, off, B, dil, pad, eps, RT); \
        else STC_LAUNCH(h, (dwconv_ln_slide_kernel<NW, KK, false, Out>), sg, 128, 0, x, w, wb, g, b, out, rows, off, B, dil, pad, eps, RT);     \
        return;                                                                                                                               \
    } while (0)
            switch (C / 128 * 10 + K) {
                case 15: STC_SLIDE(1, 5);
                case 17: STC_SLIDE(1, 7);
                case 25: STC_SLIDE(2, 5);
                case 27: STC_SLIDE(2, 7);
                case 45: STC_SLIDE(4, 5);
                case 47: STC_SLIDE(4, 7);
            }
#undef STC_SLIDE
        }
        if (K > 0 && (C == 128 || C == 256 || C == 512)) {
            // rows per block: as many as keep >= 2 blocks per SM in flight (shared-memory tile = (R + span) rows)
            const int span = (K - 1) * dil;
            int R = std::min(32, (int)(100 * 1024 / (C * sizeof(float))) - span) / 8 * 8;       // <= 100 KB: two blocks per SM
            while (R > 8 && (int)cdiv(rows, R) < 4 * h->num_sms) R -= 8;
            const size_t smem = (size_t)(std::max(R, 8) + span) * C * sizeof(float);
            if (R >= 8 && smem <= 200 * 1024) {
                dim3 tg(cdiv(rows, R));
                h->note("dwconv_ln_tile");
                switch (C / 32) {
                    case 4: STC_LAUNCH(h, (dwconv_ln_tile_kernel<4, Out>), tg, block, smem, x, w, wb, g, b, out, rows, off, B, K, dil, pad, eps, R); return;
                    case 8: STC_LAUNCH(h, (dwconv_ln_tile_kernel<8, Out>), tg, block, smem, x, w, wb, g, b, out, rows, off, B, K, dil, pad, eps, R); return;
                    case 16: STC_LAUNCH(h, (dwconv_ln_tile_kernel<16, Out>), tg, block, smem, x, w, wb, g, b, out, rows, off, B, K, dil, pad, eps, R); return;
                }
            }
        }
        if (C == 128 || C == 256 || C == 512) h->note("dwconv_ln_vec");
        switch (C / 32) {
            case 4: STC_LAUNCH(h, (dwconv_ln_vec_kernel<4, Out>), grid, block, 0, x, w, wb, g, b, out, rows, off, B, K, dil, pad, eps); return;
            case 8: STC_LAUNCH(h, (dwconv_ln_vec_kernel<8, Out>), grid, block, 0, x, w, wb, g, b, out, rows, off, B, K, dil, pad, eps); return;
            case 16: STC_LAUNCH(h, (dwconv_ln_vec_kernel<16, Out>), grid, block, 0, x, w, wb, g, b, out, rows, off, B, K, dil, pad, eps); return;
            default: break;
        }
    }
    h->note("dwconv_ln_generic");
    switch (C / 32) {
        case 1: STC_LAUNCH(h, (dwconv_ln_kernel<T, 1, Out>), grid, block, 0, x, w, wb, g, b, out, rows, off, B, K, dil, pad, eps); break;
        case 2: STC_LAUNCH(h, (dwconv_ln_kernel<T, 2, Out>), grid, block, 0, x, w, wb, g, b, out, rows, off, B, K, dil, pad, eps); break;
        case 4: STC_LAUNCH(h, (dwconv_ln_kernel<T, 4, Out>), grid, block, 0, x, w, wb, g, b, out, rows, off, B, K, dil, pad, eps); break;
        case 8: STC_LAUNCH(h, (dwconv_ln_kernel<T, 8, Out>), grid, block, 0, x, w, wb, g, b, out, rows, off, B, K, dil, pad, eps); break;
        case 16: STC_LAUNCH(h, (dwconv_ln_kernel<T, 16, Out>), grid, block, 0, x, w, wb, g, b, out, rows, off, B, K, dil, pad, eps); break;
        default: throw StcError(STC_ERR_UNSUPPORTED, "channel count must be 32/64/128/256/512, got " + std::to_string(C));
    }
}

template <typename T>
void Handle::dwconv_ln(const T* x, const ConvNeXt* cn, const float* g, const float* b, int C, const Seq& seq, float eps,
                       T* out_plain, const Act* out_act) {
    const bool vec = std::is_same<T, float>::value && (C == 128 || C == 256 || C == 512);
    const float* w = cn ? (vec ? cn->dw_wt : cn->dw_w) : nullptr; const float* wb = cn ? cn->dw_b : nullptr;
    int K = cn ? cn->K : 0, dil = cn ? cn->dil : 1, pad = cn ? cn->pad_left : 0, rows = seq.rows;
    if constexpr (std::is_same<T, float>::value) {
        if (out_act && out_act->hi) {
            kprof_begin(8.0 * rows * C >= 32e6 ? 5 : 1, (2.0 * K + 8.0) * rows * C, (out_act->lo || !out_act->hi ? 8.0 : 6.0) * rows * C + 4.0 * C * (K + 3));
            launch_dwln<T, OutSplit>(this, C, x, w, wb, g, b, OutSplit{out_act->hi, out_act->lo}, rows, seq.off, seq.B, K, dil, pad, eps);
            kprof_end();
            return;
        }
        if (out_act) {          // fp32 operand of the CUDA-core GEMMs (fp32_simt mode)
            kprof_begin(8.0 * rows * C >= 32e6 ? 5 : 1, (2.0 * K + 8.0) * rows * C, (out_act->lo || !out_act->hi ? 8.0 : 6.0) * rows * C + 4.0 * C * (K + 3));
            launch_dwln<T, OutPlain<T>>(this, C, x, w, wb, g, b, OutPlain<T>{out_act->f}, rows, seq.off, seq.B, K, dil, pad, eps);
            kprof_end();
            return;
        }
    }
    launch_dwln<T, OutPlain<T>>(this, C, x, w, wb, g, b, OutPlain<T>{out_plain}, rows, seq.off, seq.B, K, dil, pad, eps);
}

template <typename T>
void Handle::gemm_simt(const T* a, int lda, int M, const Linear& w, const Epilogue& ep, T* out, int ldo) {
    // small row tiles when 64-row tiles would not give every SM a block (the fp64 duration predictor: 150 blocks before)
    // (per-element accumulation order over k is the same for every tile height: the variants are bit-identical)
    // tallest tile that still gives every SM a block: 64 rows (TM = 4: 8 shared-memory loads per 16 FMAs), else 32, else 16
    if (cdiv(w.N, 64) * cdiv(M, 64) < 4u * num_sms && cdiv(w.N, 64) * cdiv(M, 64) >= (unsigned)num_sms) {
        dim3 grid(cdiv(w.N, 64), cdiv(M, 64));
        STC_LAUNCH(this, (gemm_simt_kernel<T, OutPlain<T>>), grid, 256, 0, a, lda, w.w_kn, OutPlain<T>{out}, ldo, M, w.N, w.K, ep);
        return;
    }
    if (cdiv(w.N, 64) * cdiv(M, 64) < 4u * num_sms && cdiv(w.N, 64) * cdiv(M, 32) >= (unsigned)num_sms && M > 32) {
        dim3 grid(cdiv(w.N, 64), cdiv(M, 32));
        STC_LAUNCH(this, (gemm_simt_kernel<T, OutPlain<T>, 32>), grid, 256, 0, a, lda, w.w_kn, OutPlain<T>{out}, ldo, M, w.N, w.K, ep);
        return;
    }
    if (cdiv(w.N, 64) * cdiv(M, 64) < 4u * num_sms && M > 16) {
        dim3 grid(cdiv(w.N, 64), cdiv(M, 16));
        STC_LAUNCH(this, (gemm_simt_kernel<T, OutPlain<T>, 16>), grid, 256, 0, a, lda, w.w_kn, OutPlain<T>{out}, ldo, M, w.N, w.K, ep);
        return;
    }
    dim3 grid(cdiv(w.N, 64), cdiv(M, 64));
    STC_LAUNCH(this, (gemm_simt_kernel<T, OutPlain<T>>), grid, 256, 0, a, lda, w.w_kn, OutPlain<T>{out}, ldo, M, w.N, w.K, ep);
}

void Handle::gemm(const Act& a, int M, const Linear& w, const Epilogue& ep_in, float* out_f32, const Act* out_act, int ldo) {
    Epilogue ep = ep_in;
    if (!ep.bias) ep.bias = w.bias;
    if (!tc_mode()) {
        float* o = out_f32 ? out_f32 : out_act->f;
        note("gemm_simt_f32");
        gemm_simt<float>(a.f, w.K, M, w, ep, o, ldo);
        return;
    }
    if (!w.has_maps) throw StcError(STC_ERR_INVALID, "linear has no tensor maps");
    tc::Params p{};
    p.trace = gemm_trace;
    p.M = M; p.N = w.N; p.K = w.K; p.ep = ep; p.ldo = ldo;
    const bool f16 = w.f16;                       // vocoder: single fp16 operands (a.hi, w.w_hi; no lo halves)
    if (!a.hi || (a.lo == nullptr) != f16 || (f16 && ep.rope_freqs)) throw StcError(STC_ERR_INVALID, "GEMM operand form does not match the weights");
    if (out_f32) { p.out_f32 = out_f32; p.split = 0; }
    else { p.out_hi = out_act->hi; p.out_lo = out_act->lo; p.split = 1; }
    if (dry) return;
    int bn = force_bn ? force_bn : pick_gemm(M, w.N, w.K, f16);
    if (ep.rope_freqs) bn = 64;                   // the rotary epilogue exists for the 64-wide tile only
    p.cm = p.cn = 1;
    kprof_begin(f16 ? 4 : 0, 2.0 * M * (double)w.N * w.K, 4.0 * ((double)M * w.K + (double)w.N * w.K + (double)M * w.N * (ep.resid ? 2 : 1)));
    if (bn == 512) {          // two-SM form (gemm2_tc.cuh): 256 x 256 tiles computed by CTA pairs
        if (w.N % 4) throw StcError(STC_ERR_INVALID, "two-SM GEMM: N % 4 != 0");
        p.cm = 2; p.cn = 1;
        const CUtensorMap mah = tmap(a.hi, M, w.K, tc::BM), mal = f16 ? mah : tmap(a.lo, M, w.K, tc::BM);
        const CUtensorMap mwh = tmap(w.w_hi, w.N, w.K, tc2::HALF), mwl = f16 ? mwh : tmap(w.w_lo, w.N, w.K, tc2::HALF);
        const int num_ct = cdiv(cdiv(M, tc::BM), 2) * cdiv(w.N, tc2::BN);
        const int clusters = std::max(1, std::min(num_ct, num_sms / 2));
        // K <= 512, bias + GELU -> fp16 operand (the vocoder's pw1): A rows stay in shared memory, W tiles stream (gemm2_astat.cuh)
        if (f16 && gemm_astat && w.K <= tc2a::MAX_KB * tc2a::KBLK && p.split && !p.out_lo && ep.gelu && ep.bias && !ep.scale && !ep.mask && !ep.resid &&
            w.N % tc2::BN == 0 && ldo % 16 == 0) {
            const int units = cdiv(cdiv(M, tc::BM), 2) * cdiv(w.N / tc2::BN, tc2a::NG);
            note("gemm2_f16_astat");
            launch_k(this, tc2a::gemm2_f16_astat_kernel, dim3(std::max(1, std::min(units, num_sms / 2)) * 2), dim3(tc2a::THREADS), (size_t)tc2a::SMEM_BYTES, stream, mah, mwh, p);
            ++launches;
            kprof_end();
            return;
        }
        note(f16 ? "gemm2_f16" : "gemm2_bf16x3");
        if (f16) launch_k(this, tc2::gemm2_bf16x3_kernel<true>, dim3(clusters * 2), dim3(tc2::THREADS), (size_t)tc2::SMEM_BYTES, stream, mah, mal, mwh, mwl, p);
        else launch_k(this, tc2::gemm2_bf16x3_kernel<false>, dim3(clusters * 2), dim3(tc2::THREADS), (size_t)tc2::SMEM_BYTES, stream, mah, mal, mwh, mwl, p);
        ++launches;
        kprof_end();
        return;
    }
    const CUtensorMap mah = tmap(a.hi, M, w.K, tc::BM), mal = f16 ? mah : tmap(a.lo, M, w.K, tc::BM);
    const CUtensorMap mwh = tmap(w.w_hi, w.N, w.K, bn), mwl = f16 ? mwh : tmap(w.w_lo, w.N, w.K, bn);
    const int tiles = cdiv(M, tc::BM) * cdiv(w.N, bn);
    const dim3 grid(std::max(1, std::min(tiles, num_sms))), block(tc::NUM_THREADS);          // persistent over tiles
    note(ep.rope_freqs ? "gemm_bn64_rope" : bn == 64 ? (f16 ? "gemm_f16_bn64" : "gemm_bn64") : bn == 128 ? (f16 ? "gemm_f16_bn128" : "gemm_bn128")
                                                                                              : (f16 ? "gemm_f16_bn256" : "gemm_bn256"));
    switch (ep.rope_freqs ? 1 : f16 ? 2000 + bn : bn) {
        case 2064: launch_k(this, tc::gemm_bf16x3_kernel<64, false, true>, grid, block, (size_t)tc::Tile<64>::SMEM_BYTES, stream, mah, mal, mwh, mwl, p); break;
        case 2128: launch_k(this, tc::gemm_bf16x3_kernel<128, false, true>, grid, block, (size_t)tc::Tile<128>::SMEM_BYTES, stream, mah, mal, mwh, mwl, p); break;
        case 2256: launch_k(this, tc::gemm_bf16x3_kernel<256, false, true>, grid, block, (size_t)tc::Tile<256>::SMEM_BYTES, stream, mah, mal, mwh, mwl, p); break;
        case 1: launch_k(this, tc::gemm_bf16x3_kernel<64, true>, grid, block, (size_t)tc::Tile<64>::SMEM_BYTES, stream, mah, mal, mwh, mwl, p); break;
        case 64: launch_k(this, tc::gemm_bf16x3_kernel<64>, grid, block, (size_t)tc::Tile<64>::SMEM_BYTES, stream, mah, mal, mwh, mwl, p); break;
        case 128: launch_k(this, tc::gemm_bf16x3_kernel<128>, grid, block, (size_t)tc::Tile<128>::SMEM_BYTES, stream, mah, mal, mwh, mwl, p); break;
        case 256: launch_k(this, tc::gemm_bf16x3_kernel<256>, grid, block, (size_t)tc::Tile<256>::SMEM_BYTES, stream, mah, mal, mwh, mwl, p); break;
        default: throw StcError(STC_ERR_INVALID, "bad GEMM tile width");
    }
    ++launches;
    kprof_end();
}

// Tile width per problem (tuned with stc_debug_gemm sweeps on B200, profiles/r1c_gemm_sweep.txt, r1z_gemm_epi.txt): the one-SM
// kernel is bound by L2 -> shared-memory operand bytes, so wide tiles win once every SM has many tiles (fewer operand bytes per
// flop), narrow ones when tiles are scarce; two-SM 256 x 256 tiles (gemm2_tc.cuh, returned as 512) once they fill the SMs: the
// vocoder projections (pw1 158.8 -> 142.3 us split-bf16; single-pass fp16 operands: the two-SM form wins at every K, conv_in 32 -> 22 us).
// TMA multicast inside clusters was implemented and measured as well (2x1 / 1x2 / 2x2: 5-100 % slower on every hot shape — the
// limit is bytes delivered INTO each SM, which multicast does not reduce) and is no longer reachable from the host.
int Handle::pick_gemm(int M, int N, int K, bool f16) const {
    if (N % 256 == 0 && K >= (f16 ? 128 : 512) && (int)(cdiv(M, 2 * tc::BM) * (N / 256)) >= num_sms) return 512;
    const int tiles128 = cdiv(M, tc::BM) * cdiv(N, 128);
    if (tiles128 >= 16 * num_sms && N % 256 == 0) return 256;
    if (tiles128 >= num_sms && K >= 512 && N > 64) return 128;
    return 64;
}

// The C = 256 / H = 1024 ConvNeXt blocks (vector estimator, text encoder) run the fused MLP of mlp_stream.cuh; other widths
// (vocoder C = 512 / H = 2048, tiny config) and STC_MLP=unfused run pw1 and pw2 as two GEMMs.
bool Handle::mlp_fused(const ConvNeXt& c) const {
    return tc_mode() && !mlp_unfused && c.C == mlp::C && c.H == mlp::H && c.pw1.w_hi && c.pw2.w_hi && !c.pw1.f16 && !c.pw2.f16;
}

// How a block's 128-row tiles x 1024 hidden units are dealt out to CTAs (mlp_stream.cuh): `nslice` hidden slices per row tile, as
// single CTAs or — up to 4 slices of whole 128-unit chunks — as CTA pairs of two row tiles (half of the weight bytes per SM;
// an odd last tile either runs as single CTAs inside the same launch or is padded to a pair, whichever keeps the wave count).
// The plan minimises a cost model fitted to tools/mlp_sweep.py on B200 (profiles/r2c_mlp_sweep.txt, r2t_mlp_sweep_epi16.txt):
//   waves x K(chunks of 128 units per CTA) x (1 + load x CTAs / SMs)  +  reduce kernel (3 us + 0.4 us per slice and 4 736 rows)
// with K = 9.5 / 12.5 / +4.0 us per further chunk for single CTAs (a chunk is 8 weight units of ~980 cycles, shared-memory-port
// paced) and 9.4 / 10.9 / +3.25 us for pairs (~795 cycles per unit: MMA bound), 7.1 us for a lone 64-unit slice.
// 37 tiles -> 4 slices, 18 pairs + 1 single tile = 148 CTAs; 38..49 -> 3 slices (2 + 3 + 3 chunks) in pairs; 50..74 -> 2 slices in
// pairs; >= ~100 tiles -> 1 slice in pairs; <= 9 tiles -> 16 slices of 64 units (the batch-1 latency path).
MlpPlan Handle::mlp_plan(int tiles, int rows) const {
    const double red0_us = 3.0, red_slice_us = 0.4 * std::max(rows, 1) / 4736.0;
    auto kernel_us = [&](int chunks, bool half_chunk, bool pair) {
        if (half_chunk) return 7.1;
        if (pair) return chunks == 1 ? 9.4 : 10.9 + 3.25 * (chunks - 2);
        return chunks == 1 ? 9.5 : 12.5 + 4.0 * (chunks - 2);
    };
    MlpPlan best{1, 0}; double best_t = 1e30;
    for (int s = 1; s <= 16; ++s) {
        if (mlp_force_slices && s != mlp_force_slices) continue;
        const int blk = (16 + s - 1) / s;
        const double red = red0_us + red_slice_us * s;
        for (int mode = 0; mode < 3; ++mode) {
            const bool pair_ok = tiles >= 2 && s <= 4;            // a pair's slice is whole 128-unit chunks: 8 / 4 / 2+3+3 / 2
            if (mode && (mlp_pair == 0 || !pair_ok)) continue;
            if (!mode && mlp_pair == 1 && pair_ok) continue;      // forced pairs (sweeps)
            const int chunks = mode ? (8 + s - 1) / s : (blk + 1) / 2;
            if (mode == 2 && tiles % 2 == 0) continue;
            const int ctas = mode == 0 ? tiles * s : mode == 1 ? (tiles / 2) * 2 * s + (tiles % 2) * s : (tiles + 1) / 2 * 2 * s;
            const int waves = (ctas + num_sms - 1) / num_sms;
            const double load = std::min(1.0, (double)ctas / num_sms);
            double t = waves * kernel_us(chunks, blk == 1, mode != 0) * (1.0 + (mode ? 0.23 : 0.15) * load);
            if (mode == 1 && tiles % 2) t = std::max(t, waves * kernel_us((blk + 1) / 2, false, false) * 1.08);   // the single CTAs of the odd tile
            t += red;
            if (t < best_t - 1e-9) { best_t = t; best = MlpPlan{s, mode}; }
        }
    }
    return best;
}

// pw1 -> GELU -> pw2 of one C = 256 / H = 1024 ConvNeXt block on `rows` rows: the stream kernel writes `nslice` partial outputs per
// row tile, the reduce kernel adds them in slice order and applies b2 / layer-scale / residual / mask (+ the post-ops).
void Handle::fused_mlp(const Act& a, int rows, const ConvNeXt& c, float* x, const float* mask, const PostOps* post) {
    const int tiles = cdiv(rows, mlp::BM);
    const size_t slice = (size_t)tiles * mlp::BM * mlp::C;
    const size_t mk = mark();
    const MlpPlan plan = mlp_plan(tiles, rows);
    const int nslice = plan.nslice;
    float* partial = ws<float>(slice * nslice);
    kprof_begin(3, 4.0 * rows * (double)c.C * c.H, 4.0 * (3.0 * rows * c.C + 2.0 * c.C * c.H));
    if (plan.mode) note(nslice == 1 ? "mlp_stream2_x1" : nslice == 2 ? "mlp_stream2_x2" : nslice == 3 ? "mlp_stream2_x3" : "mlp_stream2_x4");
    else note(nslice == 1 ? "mlp_stream_x1" : nslice == 2 ? "mlp_stream_x2" : nslice == 3 ? "mlp_stream_x3" : nslice == 4 ? "mlp_stream_x4" : nslice <= 8 ? "mlp_stream_x5to8" : "mlp_stream_x9to16");
    if (!dry) {
        const CUtensorMap w1h = tmap(c.pw1.w_hi, c.H, c.C, 128), w1l = tmap(c.pw1.w_lo, c.H, c.C, 128);
        const CUtensorMap w1h64 = tmap(c.pw1.w_hi, c.H, c.C, 64), w1l64 = tmap(c.pw1.w_lo, c.H, c.C, 64);
        const CUtensorMap w2h = tmap(c.pw2.w_hi, c.C, c.H, 128), w2l = tmap(c.pw2.w_lo, c.C, c.H, 128);
        const CUtensorMap mah = tmap(a.hi, rows, c.C, mlp::BM), mal = tmap(a.lo, rows, c.C, mlp::BM);
        const CUtensorMap mpart = tmap_f32(partial, tiles * mlp::BM * nslice, mlp::C);
        mlp::StreamParams sp{};
        sp.M = rows; sp.nslice = nslice; sp.b1 = c.pw1.bias; sp.trace = mlp_trace;
        if (plan.mode) {
            const CUtensorMap w2h64 = tmap(c.pw2.w_hi, c.C, c.H, 64), w2l64 = tmap(c.pw2.w_lo, c.C, c.H, 64);
            sp.npairs = plan.mode == 2 ? (tiles + 1) / 2 : tiles / 2;
            const int clusters = sp.npairs * nslice + (plan.mode == 1 && tiles % 2 ? (nslice + 1) / 2 : 0);
            launch_k(this, mlp::convnext_mlp_stream2_kernel, dim3(2 * clusters), dim3(mlp::NUM_THREADS), (size_t)mlp::ST_SMEM_BYTES, stream,
                     mah, mal, w1h, w1l, w1h64, w1l64, w2h, w2l, w2h64, w2l64, mpart, sp);
        } else
            launch_k(this, mlp::convnext_mlp_stream_kernel, dim3(tiles * nslice), dim3(mlp::NUM_THREADS), (size_t)mlp::ST_SMEM_BYTES, stream,
                     mah, mal, w1h, w1l, w1h64, w1l64, w2h, w2l, mpart, sp);
        if (post)
            launch_k(this, nslice > 4 ? mlp::mlp_reduce_post_kernel<true> : mlp::mlp_reduce_post_kernel<false>, dim3(cdiv(rows, 8)), dim3(256), (size_t)0, stream,
                     (const float*)partial, slice, c.pw2.bias, c.gamma, mask, x, rows, post->add_vec, post->ln_g, post->ln_b, 1e-6f,
                     post->out ? post->out->hi : (__nv_bfloat16*)nullptr, post->out ? post->out->lo : (__nv_bfloat16*)nullptr, nslice);
        else
            launch_k(this, nslice > 4 ? mlp::mlp_reduce_kernel<true> : mlp::mlp_reduce_kernel<false>, dim3(cdiv((size_t)rows * mlp::C / 4, 256)), dim3(256), (size_t)0, stream,
                     (const float*)partial, slice, c.pw2.bias, c.gamma, mask, x, rows, nslice);
        launches += 2;
    }
    kprof_end();
    release(mk);
}

void Handle::apply_post(const PostOps& post, float* x, const Seq& seq, int C) {
    const int rows = seq.rows;
    if (post.add_vec)
        STC_LAUNCH(this, add_rowvec_mask_kernel<float>, cdiv((size_t)rows * C, 256), 256, 0, x, post.add_vec, seq.mask, rows, C, 0, seq.off, seq.B);
    if (post.out) {
        if (post.ln_g) dwconv_ln<float>(x, nullptr, post.ln_g, post.ln_b, C, seq, 1e-6f, nullptr, post.out);
        else to_act(x, (size_t)rows * C, *post.out);
    }
}

template <typename T>
void Handle::convnext(const ConvNeXt& c, T* x, const Seq& seq, const PostOps* post) {
    size_t mk = mark();
    int rows = seq.rows;
    Epilogue e1; e1.gelu = 1;
    Epilogue e2; e2.scale = c.gamma; e2.resid = x; e2.mask = c.masked ? seq.mask : nullptr;
    if constexpr (std::is_same<T, float>::value) {
        const bool fused = mlp_fused(c);                 // the fused form's reduce kernel carries the post-ops
        Act a = ws_act_for(c.pw1, (size_t)rows * c.C);
        dwconv_ln<float>(x, &c, c.ln_g, c.ln_b, c.C, seq, 1e-6f, nullptr, &a);
        if (fused) fused_mlp(a, rows, c, x, c.masked ? seq.mask : nullptr, post);
        else {
            Act hid = ws_act_for(c.pw2, (size_t)rows * c.H);
            gemm(a, rows, c.pw1, e1, nullptr, &hid, c.H);
            gemm(hid, rows, c.pw2, e2, x, nullptr, c.C);
        }
        if (post && !fused) apply_post(*post, x, seq, c.C);
    } else {
        T* a = ws<T>((size_t)rows * c.C); T* hid = ws<T>((size_t)rows * c.H);
        dwconv_ln<T>(x, &c, c.ln_g, c.ln_b, c.C, seq, 1e-6f, a, nullptr);
        e1.bias = c.pw1.bias; e2.bias = c.pw2.bias;
        gemm_simt<T>(a, c.C, rows, c.pw1, e1, hid, c.H);
        gemm_simt<T>(hid, c.H, rows, c.pw2, e2, x, c.C);
    }
    release(mk);
}

void Handle::rope(float* x, const float* freqs, const Seq& seq, int heads, int dh, int normalise) {
    size_t n = (size_t)seq.rows * heads * (dh / 2);
    STC_LAUNCH(this, rope_kernel, cdiv(n, 256), 256, 0, x, freqs, seq.len, seq.rows, seq.off, seq.B, heads, dh, normalise);
}

void Handle::attn_core(const float* Q, const float* K, const float* V, const Act& out, const Seq& q, const Seq& k, bool key_masked,
                       int heads, int dh) {
    dim3 grid(cdiv(q.maxlen, 16), heads, q.B);
    float scale = 1.0f / std::sqrt((float)dh);
    const float* kmask = key_masked ? k.mask : nullptr;
    kprof_begin(2, 4.0 * (double)q.rows * k.maxlen * heads * dh, 4.0 * heads * dh * (2.0 * q.rows + 2.0 * k.rows));
    note("attention_simt");
    if (out.hi) {
        OutSplit o{out.hi, out.lo};
        if (dh == 64) STC_LAUNCH(this, (attention_kernel<64, OutSplit>), grid, 128, 0, Q, K, V, kmask, o, q.off, k.off, kmask ? k.cnt : nullptr, heads, scale);
        else STC_LAUNCH(this, (attention_kernel<32, OutSplit>), grid, 128, 0, Q, K, V, kmask, o, q.off, k.off, kmask ? k.cnt : nullptr, heads, scale);
    } else {
        OutPlain<float> o{out.f};
        if (dh == 64) STC_LAUNCH(this, (attention_kernel<64, OutPlain<float>>), grid, 128, 0, Q, K, V, kmask, o, q.off, k.off, kmask ? k.cnt : nullptr, heads, scale);
        else STC_LAUNCH(this, (attention_kernel<32, OutPlain<float>>), grid, 128, 0, Q, K, V, kmask, o, q.off, k.off, kmask ? k.cnt : nullptr, heads, scale);
    }
    kprof_end();
}

bool Handle::attn_on_tc(const Attention& a, const Seq& ks) const {
    return tc_mode() && a.C / a.heads == attn::DH && ks.maxlen <= attn::MAX_BLOCKS * attn::KB && !force_simt_attn;
}

// K and V projections of `ctx` for one layer (+ rotary embedding on K), in the form the chosen attention core reads.
KV Handle::make_kv(const Attention& a, const Act& ctx, const Seq& ks) {
    KV kv;
    const int dh = a.C / a.heads;
    const size_t n = (size_t)ks.rows * a.C;
    kv.tc = attn_on_tc(a, ks);
    float *k = nullptr, *v = nullptr;
    if (kv.tc) {
        kv.ldk = (ks.maxlen + attn::KB - 1) / attn::KB * attn::KB;
        const size_t nv = (size_t)ks.B * a.heads * attn::DH * kv.ldk;
        kv.k_hi = ws<__nv_bfloat16>(n); kv.k_lo = ws<__nv_bfloat16>(n);
        kv.vt_hi = ws<__nv_bfloat16>(nv); kv.vt_lo = ws<__nv_bfloat16>(nv);
    } else { k = ws<float>(n); v = ws<float>(n); }
    const size_t mk = mark();
    if (kv.tc) {
        v = ws<float>(n);
        Act kact; kact.hi = kv.k_hi; kact.lo = kv.k_lo;
        gemm(ctx, ks.rows, a.k, rope_epilogue(a, ks), nullptr, &kact, a.C);        // K: rotary + split-bf16 in the epilogue
    } else gemm(ctx, ks.rows, a.k, Epilogue{}, k, nullptr, a.C);
    gemm(ctx, ks.rows, a.v, Epilogue{}, v, nullptr, a.C);
    if (kv.tc) {
        STC_LAUNCH(this, attn::v_prep_kernel, dim3(kv.ldk / attn::KB, a.heads, ks.B), 256, 0, v, kv.vt_hi, kv.vt_lo, ks.off, a.heads, kv.ldk);
        release(mk);
    } else {
        if (a.rope != ROPE_NONE) rope(k, a.freqs, ks, a.heads, dh, a.rope == ROPE_NORM);
        kv.K = k; kv.V = v;
    }
    return kv;
}

Epilogue Handle::rope_epilogue(const Attention& a, const Seq& seq) const {
    Epilogue e;
    if (a.rope != ROPE_NONE) {
        e.rope_freqs = a.freqs; e.rope_off = seq.off; e.rope_B = seq.B; e.rope_dh = a.C / a.heads;
        e.rope_len = a.rope == ROPE_NORM ? seq.len : nullptr;
    }
    return e;
}

void Handle::attn_core_tc(const Act& Q, const Attention& a, const KV& kv, const Act& out, const Seq& q, const Seq& k) {
    const size_t mk = mark();
    __nv_bfloat16 *q_hi = Q.hi, *q_lo = Q.lo;
    attn::Params p{};
    p.qoff = q.off; p.koff = k.off; p.kcnt = (a.key_masked && k.mask) ? k.cnt : nullptr;
    p.heads = a.heads; p.scale_log2e = (1.0f / std::sqrt((float)attn::DH)) * 1.4426950408889634f;
    if (out.hi) { p.out_hi = out.hi; p.out_lo = out.lo; p.split = 1; } else { p.out_f32 = out.f; p.split = 0; }
    kprof_begin(2, 4.0 * (double)q.rows * k.maxlen * a.C, 4.0 * a.C * (2.0 * q.rows + 2.0 * k.rows));
    if (!dry) {
        const int vrows = k.B * a.heads * attn::DH;
        const CUtensorMap mqh = tmap(q_hi, q.rows, a.C, attn::BQ), mql = tmap(q_lo, q.rows, a.C, attn::BQ);
        const CUtensorMap mkh = tmap(kv.k_hi, k.rows, a.C, attn::KB), mkl = tmap(kv.k_lo, k.rows, a.C, attn::KB);
        const CUtensorMap mvh = tmap(kv.vt_hi, vrows, kv.ldk, attn::DH), mvl = tmap(kv.vt_lo, vrows, kv.ldk, attn::DH);
        dim3 grid(cdiv(q.maxlen, attn::BQ), a.heads, q.B);
        note(k.maxlen <= attn::KB ? "attention_tc_small" : "attention_tc");
        if (k.maxlen <= attn::KB)       // <= 64 keys (style attention): the two-CTAs-per-SM layout
            launch_k(this, attn::attention_tc_kernel<1>, grid, dim3(attn::NUM_THREADS), (size_t)attn::Lay<1>::SMEM_BYTES, stream, mqh, mql, mkh, mkl, mvh, mvl, p);
        else
            launch_k(this, attn::attention_tc_kernel<attn::MAX_BLOCKS>, grid, dim3(attn::NUM_THREADS), (size_t)attn::SMEM_BYTES, stream, mqh, mql, mkh, mkl, mvh, mvl, p);
        ++launches;
    }
    kprof_end();
    release(mk);
}

// Pre-LN multi-head attention with residual. Self-attention: ctx == nullptr && pre == nullptr (keys from LN(x), kseq = qseq).
void Handle::attention(const Attention& a, float* x, const Seq& qs, const Act* ctx, const Seq& ks_in, const KV* pre, const Act* xn_pre) {
    size_t mk = mark();
    int rows = qs.rows, dh = a.C / a.heads;
    Act xn;
    if (xn_pre) xn = *xn_pre;                                  // LayerNorm(x) already produced by the preceding block's reduce kernel
    else { xn = ws_act((size_t)rows * a.C); dwconv_ln<float>(x, nullptr, a.ln_g, a.ln_b, a.C, qs, 1e-6f, nullptr, &xn); }
    const Seq& ks = a.ctx_kind == CTX_SELF ? qs : ks_in;
    if (a.ctx_kind == CTX_SELF) ctx = &xn;
    KV local;
    if (!pre) { local = make_kv(a, *ctx, ks); pre = &local; }
    Act o = ws_act((size_t)rows * a.C);
    if (pre->tc) {
        Act q = ws_act((size_t)rows * a.C);
        gemm(xn, rows, a.q, rope_epilogue(a, qs), nullptr, &q, a.C);               // Q: rotary + split-bf16 in the epilogue
        attn_core_tc(q, a, *pre, o, qs, ks);
    } else {
        float* q = ws<float>((size_t)rows * a.C);
        gemm(xn, rows, a.q, Epilogue{}, q, nullptr, a.C);
        if (a.rope != ROPE_NONE) rope(q, a.freqs, qs, a.heads, dh, a.rope == ROPE_NORM);
        attn_core(q, pre->K, pre->V, o, qs, ks, a.key_masked, a.heads, dh);
    }
    Epilogue eo; eo.resid = x; eo.mask = a.masked ? qs.mask : nullptr;
    gemm(o, rows, a.o, eo, x, nullptr, a.C);
    release(mk);
}

// ------------------------------------------------------------------------------------------ sequence descriptors
int* Handle::stage_ints(const std::vector<int>& v) {
    int* d = ws<int>(v.size());
    if (dry && !restage) return d;
    if (h_stage_off + v.size() > (h_stage_lim ? h_stage_lim : h_stage_cap)) throw StcError(STC_ERR_CAPACITY, "offset staging buffer exhausted (batch too large)");
    int* hp = h_stage + h_stage_off;
    h_stage_off += v.size();
    memcpy(hp, v.data(), v.size() * sizeof(int));
    if (!restage) STC_CUDA(cudaMemcpyAsync(d, hp, v.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
    return d;
}

Seq Handle::rect_seq(int B, int N, const float* mask, bool want_len) {
    std::vector<int> off(B + 1);
    for (int b = 0; b <= B; ++b) off[b] = b * N;
    Seq s; s.off = stage_ints(off); s.B = B; s.rows = B * N; s.maxlen = N; s.mask = mask;
    if (want_len && mask) {
        float* len = ws<float>(B); int* cnt = ws<int>(B);
        STC_LAUNCH(this, mask_len_kernel, B, 32, 0, mask, len, cnt, N);
        s.len = len; s.cnt = cnt;
    }
    return s;
}

Seq Handle::packed_seq(const std::vector<int>& lens, int rows_launch, int maxlen_launch) {
    int B = (int)lens.size();
    std::vector<int> off(B + 1, 0), lf(B);
    for (int b = 0; b < B; ++b) off[b + 1] = off[b] + lens[b];
    Seq s; s.off = stage_ints(off); s.B = B; s.rows = std::max(rows_launch, off[B]); s.maxlen = maxlen_launch; s.mask = nullptr;
    // lengths as float via the same staging path (bit pattern copy)
    std::vector<int> bits(B);
    for (int b = 0; b < B; ++b) { float f = (float)lens[b]; memcpy(&bits[b], &f, 4); }
    s.len = reinterpret_cast<const float*>(stage_ints(bits));
    return s;
}

__global__ void scale_off_kernel(const int* __restrict__ in, int* __restrict__ out, int n, int f) {
    pdl_wait(); pdl_trigger_light();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] * f;
}

Seq Handle::scaled_seq(const Seq& s, int f) {
    Seq r = s;
    int* o = ws<int>(s.B + 1);
    STC_LAUNCH(this, scale_off_kernel, cdiv(s.B + 1, 128), 128, 0, s.off, o, s.B + 1, f);
    r.off = o; r.rows = s.rows * f; r.maxlen = s.maxlen * f; r.mask = nullptr; r.len = nullptr;
    return r;
}

// ------------------------------------------------------------------------------------------ graph walkers
template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, size_t n) {
    pdl_wait(); pdl_trigger_light();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (TO)in[i];
}

void Handle::run_dp(const int64_t* ids, const float* style_dp, const Seq& seq, int T, float* dur) {
    // duration_predictor.onnx evaluated in fp64 (reference call site cpp/helper.cpp:512-526)
    size_t mk = mark();
    const int C = dp.C, rows = seq.rows, B = seq.B, si = dp_arch.at("style_in");
    const float* mask = seq.mask;
    double* x = ws<double>((size_t)rows * C);
    STC_LAUNCH(this, embed_kernel<double>, cdiv(rows, 8), dim3(32, 8), 0, ids, dp.vec["embed"], mask, x, rows, C, cfg.vocab_size, seq.off, B, T);
    double* sd = ws<double>((size_t)B * si); double* s = ws<double>((size_t)B * C);
    STC_LAUNCH(this, (cast_kernel<float, double>), cdiv((size_t)B * si, 256), 256, 0, style_dp, sd, (size_t)B * si);
    const Linear& ls = dp.lin.back();
    Epilogue es; es.bias = ls.bias;
    gemm_simt<double>(sd, si, B, ls, es, s, C);
    STC_LAUNCH(this, add_rowvec_mask_kernel<double>, cdiv((size_t)rows * C, 256), 256, 0, x, s, mask, rows, C, C, seq.off, B);
    double* x2 = ws<double>((size_t)rows * C);
    for (const Layer& l : dp.layers) {
        if (l.type != L_CONVNEXT) continue;
        const ConvNeXt& c = dp.cn[l.idx];
        // (a block of 32 rows is one CTA's serial chain: below ~1 000 rows — the batch-1 latency path — the separate launches, whose
        //  GEMMs spread over 16-row tiles, finish sooner: p50 of call() 4.50 vs 4.80 ms)
        const bool fused = dp_fused && rows >= 1024 && c.dil == 1 && c.K <= 7 && ((c.C == 64 && c.H == 256) || (c.C == 32 && c.H == 64));
        if (!fused) { note("dp_convnext_unfused"); convnext<double>(c, x, seq); continue; }
        // one launch per ConvNeXt block (dp_fused.cuh): conv + LayerNorm + both projections of 32 token rows in shared memory
        DpBlockParams bp{};
        bp.x = x; bp.out = x2; bp.dw_w = c.dw_w; bp.dw_b = c.dw_b; bp.ln_g = c.ln_g; bp.ln_b = c.ln_b;
        bp.w1 = c.pw1.w_kn; bp.b1 = c.pw1.bias; bp.w2 = c.pw2.w_kn; bp.b2 = c.pw2.bias; bp.gamma = c.gamma;
        bp.mask = c.masked ? mask : nullptr; bp.off = seq.off; bp.B = B; bp.rows = rows; bp.K = c.K; bp.pad_left = c.pad_left; bp.eps = 1e-6f;
        note("dp_convnext_fused");
        if (c.C == 64) STC_LAUNCH(this, (dp_convnext_kernel<64, 256>), cdiv(rows, 32), 256, (DpTile<64, 256>::SMEM), bp);
        else STC_LAUNCH(this, (dp_convnext_kernel<32, 64>), cdiv(rows, 32), 256, (DpTile<32, 64>::SMEM), bp);
        std::swap(x, x2);
    }
    float clip = dp_arch.at("clip"), spt = dp_arch.at("sec_per_token");
    if (C == 64) STC_LAUNCH(this, dp_head_kernel<2>, B, 256, 0, x, dp.vec["head.ln_g"], dp.vec["head.ln_b"], dp.vec["head.w"], dp.vec["head.b"], mask, dur, seq.off, 1e-6f, clip, spt);
    else if (C == 32) STC_LAUNCH(this, dp_head_kernel<1>, B, 256, 0, x, dp.vec["head.ln_g"], dp.vec["head.ln_b"], dp.vec["head.w"], dp.vec["head.b"], mask, dur, seq.off, 1e-6f, clip, spt);
    else throw StcError(STC_ERR_UNSUPPORTED, "duration predictor width");
    release(mk);
}

void Handle::run_te(const int64_t* ids, const float* style_ttl, const Seq& tseq, int T, float* text_emb_cl) {
    // text_encoder.onnx (reference call site cpp/helper.cpp:545-556); output kept channels-last [rows, C]
    size_t mk = mark();
    const int C = te.C, rows = tseq.rows, B = tseq.B, S = cfg.style_ttl_tokens, Cs = cfg.style_ttl_dim;
    Seq sseq = rect_seq(B, S, nullptr, false);
    float* x = ws<float>((size_t)rows * C);
    STC_LAUNCH(this, embed_kernel<float>, cdiv(rows, 8), dim3(32, 8), 0, ids, te.vec["embed"], tseq.mask, x, rows, C, cfg.vocab_size, tseq.off, B, T);
    Act sty = ws_act((size_t)B * S * Cs);
    to_act(style_ttl, (size_t)B * S * Cs, sty);
    Act nxt = ws_act((size_t)rows * C);           // operand a ConvNeXt block's reduce kernel prepares for the layer after it
    bool nxt_ready = false;
    for (size_t li = 0; li < te.layers.size(); ++li) {
        const Layer& l = te.layers[li];
        if (l.type == L_CONVNEXT) {
            PostOps post; bool any = false;
            if (li + 1 < te.layers.size() && te.layers[li + 1].type == L_ATTN) {
                const Attention& a = te.at[te.layers[li + 1].idx];
                post.ln_g = a.ln_g; post.ln_b = a.ln_b; post.out = &nxt; any = nxt_ready = true;
            } else if (li + 1 < te.layers.size() && te.layers[li + 1].type == L_PROJ_OUT) {
                post.out = &nxt; any = nxt_ready = true;
            }
            convnext<float>(te.cn[l.idx], x, tseq, any ? &post : nullptr);
        } else if (l.type == L_ATTN) {
            const Attention& a = te.at[l.idx];
            const Act* pre_ln = nxt_ready ? &nxt : nullptr;
            nxt_ready = false;
            if (a.ctx_kind == CTX_SELF) attention(a, x, tseq, nullptr, tseq, nullptr, pre_ln);
            else attention(a, x, tseq, &sty, sseq, nullptr, pre_ln);
        } else if (l.type == L_PROJ_OUT) {
            Act xa = nxt;
            if (!nxt_ready) { xa = ws_act((size_t)rows * C); to_act(x, (size_t)rows * C, xa); }
            nxt_ready = false;
            Epilogue e; e.mask = tseq.mask;
            gemm(xa, rows, te.lin[l.idx], e, text_emb_cl, nullptr, te.lin[l.idx].N);
        }
    }
    release(mk);
}

void Handle::prepare_ve(VeCtx& vc, const float* text_emb_cl, const float* style_ttl) {
    // K/V of text_emb and style_ttl are step-invariant: hoisted out of the Euler loop (north_star).
    int Cs = cfg.style_ttl_dim, Ct = cfg.text_emb_channels;
    int nslots = 0;
    for (const Attention& a : ve.at) if (a.kv_slot >= 0) nslots = std::max(nslots, a.kv_slot + 1);
    vc.kv.assign(nslots, KV{});
    // the split-bf16 copies of the two contexts stay allocated next to the K/V they feed (arena space, a few MB)
    Act ta = ws_act((size_t)vc.text.rows * Ct), sa = ws_act((size_t)vc.style.rows * Cs);
    to_act(text_emb_cl, (size_t)vc.text.rows * Ct, ta);
    to_act(style_ttl, (size_t)vc.style.rows * Cs, sa);
    for (const Attention& a : ve.at) {
        if (a.kv_slot < 0) continue;
        const bool text = a.ctx_kind == CTX_TEXT;
        vc.kv[a.kv_slot] = make_kv(a, text ? ta : sa, text ? vc.text : vc.style);
    }
}

// Time conditioning depends only on (current_step, total_step): the sinusoid -> MLP -> per-super-block linears are
// evaluated once per distinct pair and cached for the lifetime of the handle: [n_time_cond][C] fp32.
const float* Handle::time_vectors(float cur, float tot) {
    uint32_t kc, kt; memcpy(&kc, &cur, 4); memcpy(&kt, &tot, 4);
    auto key = std::make_pair(kc, kt);
    auto it = tvec_cache.find(key);
    if (it != tvec_cache.end()) return it->second;
    int C = ve.C, td = ve_arch.at("time_dim"), ntc = 0;
    for (const Layer& l : ve.layers) if (l.type == L_TIME_COND) ++ntc;
    if (dry) return reinterpret_cast<const float*>(uintptr_t(0x1000));
    std::lock_guard<std::recursive_mutex> lk(g_capture_mu);
    float* out = nullptr;
    STC_CUDA(cudaMalloc((void**)&out, sizeof(float) * std::max(1, ntc) * C)); owned.push_back(out);
    float* tmp = nullptr;
    STC_CUDA(cudaMalloc((void**)&tmp, sizeof(float) * (td + 2 * C + 2)));
    float* d_cur = tmp; float* d_tot = tmp + 1; float* e0 = tmp + 2; float* e1 = e0 + td; float* temb = e1 + C;
    STC_LAUNCH(this, fill_kernel, 1, 32, 0, d_cur, cur, (size_t)1);
    STC_LAUNCH(this, fill_kernel, 1, 32, 0, d_tot, tot, (size_t)1);
    int i = 0;
    for (const Layer& l : ve.layers) {
        if (l.type == L_TIME_MLP) {
            STC_LAUNCH(this, time_embed_kernel, cdiv((size_t)td / 2, 128), 128, 0, d_cur, d_tot, ve.vec["time.freqs"], e0, 1, td / 2);
            Epilogue a; a.bias = ve.lin[l.idx].bias; a.gelu = 1;
            gemm_simt<float>(e0, td, 1, ve.lin[l.idx], a, e1, C);
            Epilogue b; b.bias = ve.lin[l.idx + 1].bias;
            gemm_simt<float>(e1, C, 1, ve.lin[l.idx + 1], b, temb, C);
        } else if (l.type == L_TIME_COND) {
            Epilogue e; e.bias = ve.lin[l.idx].bias;
            gemm_simt<float>(temb, C, 1, ve.lin[l.idx], e, out + (size_t)(i++) * C, C);
        }
    }
    STC_CUDA(cudaStreamSynchronize(stream));
    cudaFree(tmp);
    tvec_cache[key] = out;
    return out;
}

void Handle::run_ve_step(const VeCtx& vc, float* x_lat, const float* tvec, const float* dtvec) {
    // vector_estimator.onnx: one Euler step, update in-graph (reference cpp/helper.cpp:620-658)
    size_t mk = mark();
    const Seq& ls = vc.lat;
    int rows = ls.rows, C = ve.C, D = cfg.latent_channels, itc = 0;
    float* x = ws<float>((size_t)rows * C);
    Act nxt = ws_act((size_t)rows * C);           // operand a ConvNeXt block's reduce kernel prepares for the layer after it
    bool nxt_ready = false;
    for (size_t li = 0; li < ve.layers.size(); ++li) {
        const Layer& l = ve.layers[li];
        switch (l.type) {
            case L_PROJ_IN: {
                size_t m2 = mark();
                Act xa = ws_act((size_t)rows * D);
                to_act(x_lat, (size_t)rows * D, xa);
                Epilogue e; e.mask = ls.mask;
                gemm(xa, rows, ve.lin[l.idx], e, x, nullptr, C);
                release(m2);
                break;
            }
            case L_CONVNEXT: {
                // fold what follows into the block's last kernel: time conditioning, the pre-LayerNorm of an attention layer,
                // or the operand conversion for the output projection
                PostOps post; bool any = false;
                size_t nx = li + 1;
                if (nx < ve.layers.size() && ve.layers[nx].type == L_TIME_COND) {
                    post.add_vec = tvec + (size_t)(itc++) * C; any = true; ++li; ++nx;          // the time_cond layer is consumed here
                }
                if (nx < ve.layers.size() && ve.layers[nx].type == L_ATTN) {
                    const Attention& a = ve.at[ve.layers[nx].idx];
                    post.ln_g = a.ln_g; post.ln_b = a.ln_b; post.out = &nxt; any = true; nxt_ready = true;
                } else if (nx < ve.layers.size() && ve.layers[nx].type == L_PROJ_OUT) {
                    post.out = &nxt; any = true; nxt_ready = true;
                }
                convnext<float>(ve.cn[l.idx], x, ls, any ? &post : nullptr);
                break;
            }
            case L_TIME_COND:
                STC_LAUNCH(this, add_rowvec_mask_kernel<float>, cdiv((size_t)rows * C, 256), 256, 0, x, tvec + (size_t)(itc++) * C, ls.mask,
                           rows, C, 0, ls.off, ls.B);
                break;
            case L_ATTN: {
                const Attention& a = ve.at[l.idx];
                attention(a, x, ls, nullptr, a.ctx_kind == CTX_TEXT ? vc.text : vc.style, &vc.kv[a.kv_slot], nxt_ready ? &nxt : nullptr);
                nxt_ready = false;
                break;
            }
            case L_PROJ_OUT: {
                size_t m2 = mark();
                Act xa = nxt;
                if (!nxt_ready) { xa = ws_act((size_t)rows * C); to_act(x, (size_t)rows * C, xa); }
                nxt_ready = false;
                Epilogue e; e.scale = dtvec; e.resid = x_lat; e.mask = ls.mask;   // x <- (x + v*dt) * mask
                gemm(xa, rows, ve.lin[l.idx], e, x_lat, nullptr, D);
                release(m2);
                break;
            }
            default: break;
        }
    }
    release(mk);
}

void Handle::run_vocoder(const float* lat_cl, const Seq& lat, float* wav) {
    // vocoder.onnx (reference call site cpp/helper.cpp:662-672); no mask: every launched frame is decoded
    size_t mk = mark();
    int f = cfg.chunk_compress_factor, ld = cfg.latent_dim, C = voc.C;
    Seq s6 = scaled_seq(lat, f);
    int rows = s6.rows;
    float* x = ws<float>((size_t)rows * C);
    for (const Layer& l : voc.layers) {
        if (l.type == L_CONV_IN) {
            size_t m2 = mark();
            const Linear& w = voc.lin[l.idx];
            int K = w.K / ld;
            Act a = ws_act_for(w, (size_t)rows * w.K);
            size_t n = (size_t)rows * w.K;
            if (a.hi) STC_LAUNCH(this, voc_im2col_kernel<OutSplit>, cdiv(n, 256), 256, 0, lat_cl, voc.vec["std"], voc.vec["mean"], OutSplit{a.hi, a.lo}, rows, lat.off, lat.B, f, ld, K, w.K);
            else STC_LAUNCH(this, voc_im2col_kernel<OutPlain<float>>, cdiv(n, 256), 256, 0, lat_cl, voc.vec["std"], voc.vec["mean"], OutPlain<float>{a.f}, rows, lat.off, lat.B, f, ld, K, w.K);
            gemm(a, rows, w, Epilogue{}, x, nullptr, C);
            release(m2);
        } else if (l.type == L_CONVNEXT) convnext<float>(voc.cn[l.idx], x, s6);
        else if (l.type == L_HEAD) {
            Act hn = ws_act_for(voc.lin[l.idx], (size_t)rows * C);
            dwconv_ln<float>(x, nullptr, voc.vec["head.ln_g"], voc.vec["head.ln_b"], C, s6, 1e-6f, nullptr, &hn);
            gemm(hn, rows, voc.lin[l.idx], Epilogue{}, wav, nullptr, voc.lin[l.idx].N);
        }
    }
    release(mk);
}

// Run `fn` once in measuring mode to learn its workspace high-water marks; grow the arenas if needed.
// (`fn` must start with arena.reset()/rewind so that both passes allocate identically.)
void Handle::ensure_ws(const std::function<void()>& fn) {
    bool was = dry;
    size_t p_used = persist.used();
    dry = true; arena.high_water = 0; persist.high_water = p_used;
    try { fn(); } catch (...) { dry = was; throw; }
    dry = was;
    persist.rewind(p_used);
    if (arena.high_water > arena.capacity() || persist.high_water > persist.capacity()) {
        for (auto& g : graphs) cudaGraphExecDestroy(g.second.exec);     // captured pointers die with the old arena
        graphs.clear(); map_cache.clear();
        if (persist.high_water > persist.capacity()) {
            if (p_used) throw StcError(STC_ERR_CUDA, "persistent arena cannot grow while in use");
            persist.reserve(persist.high_water);
        }
        arena.reserve(arena.high_water);
    }
}

// Run `body` (which must allocate deterministically from the arenas) either eagerly or as a cached CUDA graph.
// Replay: `body` is re-run with launches suppressed, only to refresh the pinned staging slots (sequence offsets, seed)
// and to collect the caller-memory uploads, then the instantiated graph is launched.
void Handle::run_graphed(const GraphKey& key, const std::function<void()>& body, cudaEvent_t after_uploads) {
    if (!use_graphs || profile) {            // eager (profiling / debugging): no overlap — the event fires after the whole body
        ensure_ws(body); body();
        if (after_uploads) STC_CUDA(cudaEventRecord(after_uploads, stream));
        return;
    }
    auto it = graphs.find(key);
    if (it == graphs.end()) {
        std::lock_guard<std::recursive_mutex> lk(g_capture_mu);
        ensure_ws(body);                       // may grow the arenas (and then drops every cached graph)
        if (graphs.size() >= 48) { for (auto& g : graphs) cudaGraphExecDestroy(g.second.exec); graphs.clear(); }
        pre_copies.clear();
        uint64_t l0 = launches;
        cudaGraph_t graph = nullptr;
        STC_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        capturing = true;
        try { body(); } catch (...) { capturing = false; cudaStreamEndCapture(stream, &graph); if (graph) cudaGraphDestroy(graph); throw; }
        capturing = false;
        STC_CUDA(cudaStreamEndCapture(stream, &graph));
        cudaGraphExec_t exec = nullptr;
        cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) throw StcError(STC_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
        it = graphs.emplace(key, GraphEntry{exec, launches - l0}).first;
        launches = l0;                         // captured, not launched yet
        ++graph_captures;
    } else {
        pre_copies.clear();
        dry = true; restage = true;
        try { body(); } catch (...) { dry = false; restage = false; throw; }
        dry = false; restage = false;
        ++graph_replays;
    }
    for (const PreCopy& c : pre_copies) STC_CUDA(cudaMemcpyAsync(c.dst, c.src, c.bytes, cudaMemcpyHostToDevice, stream));
    pre_copies.clear();
    if (after_uploads) STC_CUDA(cudaEventRecord(after_uploads, stream));
    STC_CUDA(cudaGraphLaunch(it->second.exec, stream));
    launches += it->second.kernels;            // kernels executed by this replay
}

}  // namespace stc

// =============================================================================================== C ABI
using namespace stc;

static thread_local std::string g_last_error;

static int fail(stc_handle* h, int code, const std::string& msg) {
    g_last_error = msg;
    if (h) h->last_error = msg;
    return code;
}

#define STC_TRY(h, ...)                                                                      \
    try { __VA_ARGS__; return STC_OK; }                                                           \
    catch (const StcError& e) { return fail(h, e.code, e.what()); }                          \
    catch (const std::exception& e) { return fail(h, STC_ERR_INVALID, e.what()); }

extern "C" {

const char* stc_last_error(const stc_handle* h) { return h ? h->last_error.c_str() : g_last_error.c_str(); }

int stc_create(const char* onnx_dir, int device, int precision, stc_handle** out) {
    if (!out || !onnx_dir) return fail(nullptr, STC_ERR_INVALID, "stc_create: null argument");
    *out = nullptr;
    auto sh = std::make_unique<stc_handle>();
    std::lock_guard<std::recursive_mutex> lk(stc::g_capture_mu);
    try {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0)
            throw StcError(STC_ERR_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                             "); libsupertonic_cuda has no CPU fallback");
        if (device < 0 || device >= n) throw StcError(STC_ERR_INVALID, "device index out of range");
        STC_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop; STC_CUDA(cudaGetDeviceProperties(&prop, device));
        auto hd = std::make_unique<Handle>();
        hd->device = device;
        hd->num_sms = prop.multiProcessorCount;
        if (precision == STC_PREC_DEFAULT) {
            const char* env = getenv("STC_PRECISION");
            precision = (env && std::string(env) == "fp32_simt") ? STC_PREC_FP32_SIMT : STC_PREC_BF16X3;
        }
        if (precision != STC_PREC_BF16X3 && precision != STC_PREC_FP32_SIMT) throw StcError(STC_ERR_INVALID, "unknown precision");
        if (precision == STC_PREC_BF16X3 && prop.major != 10)
            throw StcError(STC_ERR_UNSUPPORTED, "the tcgen05 path needs an sm_100 device, found sm_" + std::to_string(prop.major * 10 + prop.minor));
        hd->precision = precision;
        { const char* e = getenv("STC_ATTN"); hd->force_simt_attn = e && std::string(e) == "simt"; }
        { const char* e = getenv("STC_MLP"); hd->mlp_unfused = e && !strcmp(e, "unfused"); }
        { const char* e = getenv("STC_MLP_PAIR"); if (e && *e) hd->mlp_pair = atoi(e); }
        { const char* e = getenv("STC_PDL"); if (e && *e) hd->pdl_mode = std::max(0, std::min(3, atoi(e))); }
        STC_CUDA(cudaMemcpyToSymbol(stc::c_pdl_mode, &hd->pdl_mode, sizeof(int)));
        { const char* e = getenv("STC_MLP_SLICES"); if (e && *e) hd->mlp_force_slices = std::max(0, std::min(16, atoi(e))); }
        { const char* e = getenv("STC_DP"); hd->dp_fused = !(e && !strcmp(e, "unfused")); }
        { const char* e = getenv("STC_VOC"); hd->voc_f16 = !e || !strcmp(e, "f16"); }
        { const char* e = getenv("STC_DW"); hd->dw_slide = !(e && !strcmp(e, "tile")); hd->dw_chain_kernel = !(e && !strcmp(e, "slide")); }
        { const char* e = getenv("STC_DW_RT"); hd->dw_rt = e ? atoi(e) : 0; }
        { const char* e = getenv("STC_ASTAT"); hd->gemm_astat = !(e && !strcmp(e, "0")); }
        { const char* e = getenv("STC_DW_RING"); hd->dw_ring = e ? atoi(e) : -1; }
        STC_CUDA(cudaStreamCreateWithFlags(&hd->stream, cudaStreamNonBlocking));
        STC_CUDA(cudaStreamCreateWithFlags(&hd->stream2, cudaStreamNonBlocking));
        STC_CUDA(cudaStreamCreateWithFlags(&hd->stream_copy, cudaStreamNonBlocking));
        STC_CUDA(cudaStreamCreateWithFlags(&hd->stream_f, cudaStreamNonBlocking));
        STC_CUDA(cudaEventCreateWithFlags(&hd->ev_out, cudaEventDisableTiming));
        for (auto& e : hd->copy_done) STC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        STC_CUDA(cudaEventCreateWithFlags(&hd->ev_in, cudaEventDisableTiming));
        STC_CUDA(cudaEventCreateWithFlags(&hd->ev_te, cudaEventDisableTiming));
        for (auto& ev : hd->ev) STC_CUDA(cudaEventCreate(&ev));
        if (precision == STC_PREC_BF16X3) {
            void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
            STC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
            if (!fn || qr != cudaDriverEntryPointSuccess) throw StcError(STC_ERR_CUDA, "cuTensorMapEncodeTiled not available");
            hd->encode = (EncodeTiledFn)fn;
            STC_CUDA(cudaFuncSetAttribute(tc::gemm_bf16x3_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Tile<256>::SMEM_BYTES));
            STC_CUDA(cudaFuncSetAttribute(tc::gemm_bf16x3_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Tile<128>::SMEM_BYTES));
            STC_CUDA(cudaFuncSetAttribute(tc::gemm_bf16x3_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Tile<64>::SMEM_BYTES));
            STC_CUDA(cudaFuncSetAttribute((tc::gemm_bf16x3_kernel<64, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Tile<64>::SMEM_BYTES));
            STC_CUDA(cudaFuncSetAttribute(tc2::gemm2_bf16x3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_BYTES));
            STC_CUDA(cudaFuncSetAttribute(tc2::gemm2_bf16x3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_BYTES));
            STC_CUDA(cudaFuncSetAttribute(tc2a::gemm2_f16_astat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2a::SMEM_BYTES));
            STC_CUDA(cudaFuncSetAttribute((tc::gemm_bf16x3_kernel<64, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Tile<64>::SMEM_BYTES));
            STC_CUDA(cudaFuncSetAttribute((tc::gemm_bf16x3_kernel<128, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Tile<128>::SMEM_BYTES));
            STC_CUDA(cudaFuncSetAttribute((tc::gemm_bf16x3_kernel<256, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Tile<256>::SMEM_BYTES));
            STC_CUDA(cudaFuncSetAttribute(attn::attention_tc_kernel<attn::MAX_BLOCKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::SMEM_BYTES));
            STC_CUDA(cudaFuncSetAttribute(attn::attention_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::Lay<1>::SMEM_BYTES));
            STC_CUDA(cudaFuncSetAttribute(mlp::convnext_mlp_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mlp::ST_SMEM_BYTES));
            STC_CUDA(cudaFuncSetAttribute(mlp::convnext_mlp_stream2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mlp::ST_SMEM_BYTES));
        }
#define STC_CHAIN_ATTR(NW, KK)                                                                                                                 \
    STC_CUDA(cudaFuncSetAttribute((dwconv_ln_chain_kernel<NW, KK, OutSplit>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ChainSmem<KK>::BYTES)); \
    STC_CUDA(cudaFuncSetAttribute((dwconv_ln_chain_kernel<NW, KK, OutPlain<float>>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ChainSmem<KK>::BYTES)); \
    STC_CUDA(cudaFuncSetAttribute((dwconv_ln_chain_kernel<NW, KK, OutSplit>), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
    STC_CUDA(cudaFuncSetAttribute((dwconv_ln_chain_kernel<NW, KK, OutPlain<float>>), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared))
        STC_CHAIN_ATTR(1, 5); STC_CHAIN_ATTR(1, 7); STC_CHAIN_ATTR(2, 5); STC_CHAIN_ATTR(2, 7); STC_CHAIN_ATTR(4, 5); STC_CHAIN_ATTR(4, 7);
#undef STC_CHAIN_ATTR
        STC_CUDA(cudaFuncSetAttribute((dp_convnext_kernel<64, 256>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DpTile<64, 256>::SMEM));
        STC_CUDA(cudaFuncSetAttribute((dp_convnext_kernel<32, 64>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DpTile<32, 64>::SMEM));
        {
            const int big = 200 * 1024;
            STC_CUDA(cudaFuncSetAttribute(dwconv_ln_tile_kernel<4, OutSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
            STC_CUDA(cudaFuncSetAttribute(dwconv_ln_tile_kernel<8, OutSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
            STC_CUDA(cudaFuncSetAttribute(dwconv_ln_tile_kernel<16, OutSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
            STC_CUDA(cudaFuncSetAttribute((dwconv_ln_tile_kernel<4, OutPlain<float>>), cudaFuncAttributeMaxDynamicSharedMemorySize, big));
            STC_CUDA(cudaFuncSetAttribute((dwconv_ln_tile_kernel<8, OutPlain<float>>), cudaFuncAttributeMaxDynamicSharedMemorySize, big));
            STC_CUDA(cudaFuncSetAttribute((dwconv_ln_tile_kernel<16, OutPlain<float>>), cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        }
        hd->load(onnx_dir);
        sh->impl = std::move(hd);
    } catch (const StcError& e) { return fail(nullptr, e.code, e.what()); }
    catch (const std::exception& e) { return fail(nullptr, STC_ERR_IO, e.what()); }
    *out = sh.release();
    return STC_OK;
}

void stc_destroy(stc_handle* h) {
    if (!h) return;
    std::lock_guard<std::recursive_mutex> lk(stc::g_capture_mu);
    if (h->impl) { cudaSetDevice(h->impl->device); cudaStreamSynchronize(h->impl->stream); }
    delete h;
}

int stc_get_config(const stc_handle* h, stc_config* out) {
    if (!h || !out) return STC_ERR_INVALID;
    *out = h->impl->cfg;
    return STC_OK;
}

int stc_validate_style(const stc_handle* h, int B, const int64_t ttl[3], const int64_t dp[3]) {
    if (!h || !h->impl) return fail(nullptr, STC_ERR_INVALID, "stc_validate_style: null handle");
    const stc_config& c = h->impl->cfg;
    auto bad = [&](const char* name, const int64_t* got, int d1, int d2) {
        return fail(const_cast<stc_handle*>(h), STC_ERR_INVALID,
                    std::string("Got invalid dimensions for input: ") + name + " — got [" + std::to_string(got[0]) + "," + std::to_string(got[1]) + "," +
                        std::to_string(got[2]) + "], expected [" + std::to_string(B) + "," + std::to_string(d1) + "," + std::to_string(d2) + "]");
    };
    if (ttl && (ttl[0] != B || ttl[1] != c.style_ttl_tokens || ttl[2] != c.style_ttl_dim)) return bad("style_ttl", ttl, c.style_ttl_tokens, c.style_ttl_dim);
    if (dp && (dp[0] != B || dp[1] != c.style_dp_tokens || dp[2] != c.style_dp_dim)) return bad("style_dp", dp, c.style_dp_tokens, c.style_dp_dim);
    return STC_OK;
}

uint64_t stc_launch_count(const stc_handle* h) { return h ? h->impl->launches : 0; }
int stc_kernel_variants(const stc_handle* h, char* buf, size_t cap, size_t* need) {
    if (!h || !h->impl || !need) return STC_ERR_INVALID;
    std::string s;
    for (const auto& kv : h->impl->variant_count) s += kv.first + "=" + std::to_string(kv.second) + "\n";
    s += "graph_captures=" + std::to_string(h->impl->graph_captures) + "\ngraph_replays=" + std::to_string(h->impl->graph_replays) + "\n";
    *need = s.size() + 1;
    if (!buf || cap < s.size() + 1) return STC_ERR_CAPACITY;
    memcpy(buf, s.c_str(), s.size() + 1);
    return STC_OK;
}
int stc_set_graphs(stc_handle* h, int enabled) { if (!h) return STC_ERR_INVALID; h->impl->use_graphs = enabled != 0; return STC_OK; }
void* stc_stream(stc_handle* h) { return h ? (void*)h->impl->stream : nullptr; }
int stc_set_profile(stc_handle* h, int level) { if (!h) return STC_ERR_INVALID; h->impl->profile = level; return STC_OK; }
int stc_kernel_profile(const stc_handle* h, int cls, double out[4]) {
    if (!h || !out || cls < 0 || cls > 5) return STC_ERR_INVALID;
    const auto& k = h->impl->kprof[cls];
    out[0] = k.ms; out[1] = k.flops; out[2] = k.bytes; out[3] = (double)k.n;
    return STC_OK;
}
int stc_last_stage_ms(const stc_handle* h, float out[5]) {
    if (!h || !out) return STC_ERR_INVALID;
    memcpy(out, h->impl->stage_ms, sizeof(float) * 5);
    return STC_OK;
}

}  // extern "C"

// ---- helpers shared by the entry points
namespace {
struct Scope {   // per-call: select device, reset arena
    Handle* h;
    explicit Scope(stc_handle* sh, bool keep_async = false) : h(sh ? sh->impl.get() : nullptr) {
        if (!h) throw StcError(STC_ERR_INVALID, "null handle");
        STC_CUDA(cudaSetDevice(h->device));
        if (!keep_async) h->wait_async();              // every synchronous entry point drains outstanding asynchronous calls first
        h->arena.reset(); h->persist.reset();
        if (!h->h_stage) {
            std::lock_guard<std::recursive_mutex> lk(g_capture_mu);
            h->h_stage_cap = 1 << 18; STC_CUDA(cudaMallocHost((void**)&h->h_stage, h->h_stage_cap * sizeof(int)));
        }
        h->h_stage_off = 0; h->h_stage_lim = 0;
    }
};
template <typename T> void h2d(Handle* h, T* d, const T* host, size_t n) {
    if (h->capturing || h->restage) h->pre_copies.push_back({d, host, n * sizeof(T)});     // issued before the graph launch
    else if (!h->dry) STC_CUDA(cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
}
template <typename T> T* upp(Handle* h, const T* host, size_t n) { T* d = h->ps<T>(n); h2d(h, d, host, n); return d; }   // persistent arena
template <typename T> T* up(Handle* h, const T* host, size_t n) { T* d = h->ws<T>(n); h2d(h, d, host, n); return d; }
void validate_ids(const int64_t* ids, size_t n, int V) {
    for (size_t i = 0; i < n; ++i)
        if (ids[i] < 0 || ids[i] >= V) throw StcError(STC_ERR_INVALID, "text_ids value out of the embedding table range");
}
void transpose(Handle* h, const float* in, float* out, int batch, int R, int Cc) {
    dim3 grid(cdiv(Cc, 32), cdiv(R, 32), batch);
    STC_LAUNCH(h, (transpose_kernel<float, float>), grid, dim3(32, 8), 0, in, out, R, Cc);
}
}  // namespace

extern "C" {

int stc_duration(stc_handle* sh, const int64_t* text_ids, const float* style_dp, const float* text_mask, int B, int T,
                 float* duration_out) {
    STC_TRY(sh, {
        Scope sc(sh); Handle* h = sc.h;
        if (B <= 0 || T <= 0 || !text_ids || !style_dp || !text_mask || !duration_out) throw StcError(STC_ERR_INVALID, "stc_duration: bad argument");
        validate_ids(text_ids, (size_t)B * T, h->cfg.vocab_size);
        int si = h->cfg.style_dp_tokens * h->cfg.style_dp_dim;
        auto body = [&]() {
            h->arena.reset(); h->h_stage_off = 0;
            int64_t* d_ids = up(h, text_ids, (size_t)B * T);
            float* d_sty = up(h, style_dp, (size_t)B * si);
            float* d_mask = up(h, text_mask, (size_t)B * T);
            float* d_dur = h->ws<float>(B);
            h->run_dp(d_ids, d_sty, h->rect_seq(B, T, d_mask, false), T, d_dur);
            if (!h->dry) STC_CUDA(cudaMemcpyAsync(duration_out, d_dur, B * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        };
        h->ensure_ws(body);
        body();
        STC_CUDA(cudaStreamSynchronize(h->stream));
        h->check_launch("stc_duration");
    })
}

int stc_text_encode(stc_handle* sh, const int64_t* text_ids, const float* style_ttl, const float* text_mask, int B, int T,
                    float* text_emb_out, int64_t shape_out[3]) {
    STC_TRY(sh, {
        Scope sc(sh); Handle* h = sc.h;
        if (B <= 0 || T <= 0 || !text_ids || !style_ttl || !text_mask || !text_emb_out) throw StcError(STC_ERR_INVALID, "stc_text_encode: bad argument");
        validate_ids(text_ids, (size_t)B * T, h->cfg.vocab_size);
        int C = h->cfg.text_emb_channels, S = h->cfg.style_ttl_tokens, Cs = h->cfg.style_ttl_dim;
        auto body = [&]() {
            h->arena.reset(); h->h_stage_off = 0;
            int64_t* d_ids = up(h, text_ids, (size_t)B * T);
            float* d_sty = up(h, style_ttl, (size_t)B * S * Cs);
            float* d_mask = up(h, text_mask, (size_t)B * T);
            float* d_cl = h->ws<float>((size_t)B * T * C);
            float* d_ncl = h->ws<float>((size_t)B * T * C);
            h->run_te(d_ids, d_sty, h->rect_seq(B, T, d_mask, true), T, d_cl);
            transpose(h, d_cl, d_ncl, B, T, C);                       // [B,T,C] -> [B,C,T] (reference layout)
            if (!h->dry) STC_CUDA(cudaMemcpyAsync(text_emb_out, d_ncl, (size_t)B * T * C * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        };
        h->ensure_ws(body);
        body();
        STC_CUDA(cudaStreamSynchronize(h->stream));
        h->check_launch("stc_text_encode");
        if (shape_out) { shape_out[0] = B; shape_out[1] = C; shape_out[2] = T; }
    })
}

int stc_vector_step(stc_handle* sh, const float* noisy_latent, const float* text_emb, const float* style_ttl, const float* text_mask,
                    const float* latent_mask, const float* total_step, const float* current_step, int B, int L, int T,
                    float* denoised_out) {
    STC_TRY(sh, {
        Scope sc(sh); Handle* h = sc.h;
        if (B <= 0 || L <= 0 || T <= 0 || !noisy_latent || !text_emb || !style_ttl || !text_mask || !latent_mask || !total_step ||
            !current_step || !denoised_out)
            throw StcError(STC_ERR_INVALID, "stc_vector_step: bad argument");
        for (int b = 1; b < B; ++b)
            if (total_step[b] != total_step[0] || current_step[b] != current_step[0])
                throw StcError(STC_ERR_UNSUPPORTED, "per-utterance step counters must be equal (the reference passes one value, cpp/helper.cpp:573,591)");
        int C = h->cfg.text_emb_channels, S = h->cfg.style_ttl_tokens, Cs = h->cfg.style_ttl_dim, D = h->cfg.latent_channels;
        float dt = 1.0f / total_step[0];
        const float* tvec = h->time_vectors(current_step[0], total_step[0]);
        auto body = [&]() {
            h->arena.reset(); h->h_stage_off = 0;
            float* d_x_ncl = up(h, noisy_latent, (size_t)B * D * L);
            float* d_te_ncl = up(h, text_emb, (size_t)B * C * T);
            float* d_sty = up(h, style_ttl, (size_t)B * S * Cs);
            float* d_tm = up(h, text_mask, (size_t)B * T);
            float* d_lm = up(h, latent_mask, (size_t)B * L);
            float* d_x = h->ws<float>((size_t)B * L * D);
            float* d_te = h->ws<float>((size_t)B * T * C);
            float* d_dt = h->ws<float>(D);
            transpose(h, d_x_ncl, d_x, B, D, L);
            transpose(h, d_te_ncl, d_te, B, C, T);
            STC_LAUNCH(h, fill_kernel, cdiv(D, 128), 128, 0, d_dt, dt, (size_t)D);
            VeCtx vc;
            vc.lat = h->rect_seq(B, L, d_lm, true); vc.text = h->rect_seq(B, T, d_tm, true); vc.style = h->rect_seq(B, S, nullptr, false);
            h->prepare_ve(vc, d_te, d_sty);
            h->run_ve_step(vc, d_x, tvec, d_dt);
            transpose(h, d_x, d_x_ncl, B, L, D);
            if (!h->dry) STC_CUDA(cudaMemcpyAsync(denoised_out, d_x_ncl, (size_t)B * D * L * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        };
        h->ensure_ws(body);
        body();
        STC_CUDA(cudaStreamSynchronize(h->stream));
        h->check_launch("stc_vector_step");
    })
}

int stc_vocode(stc_handle* sh, const float* latent, int B, int L, float* wav_out) {
    STC_TRY(sh, {
        Scope sc(sh); Handle* h = sc.h;
        if (B <= 0 || L <= 0 || !latent || !wav_out) throw StcError(STC_ERR_INVALID, "stc_vocode: bad argument");
        int D = h->cfg.latent_channels; size_t nw = (size_t)B * L * h->cfg.chunk_size;
        auto body = [&]() {
            h->arena.reset(); h->h_stage_off = 0;
            float* d_ncl = up(h, latent, (size_t)B * D * L);
            float* d_cl = h->ws<float>((size_t)B * D * L);
            float* d_wav = h->ws<float>(nw);
            transpose(h, d_ncl, d_cl, B, D, L);
            h->run_vocoder(d_cl, h->rect_seq(B, L, nullptr, false), d_wav);
            if (!h->dry) STC_CUDA(cudaMemcpyAsync(wav_out, d_wav, nw * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        };
        h->ensure_ws(body);
        body();
        STC_CUDA(cudaStreamSynchronize(h->stream));
        h->check_launch("stc_vocode");
    })
}

}  // extern "C"

// ---- fast layer ---------------------------------------------------------------------------------
namespace stc {

// float32 length math of sampleNoisyLatent (cpp/helper.cpp:430-438) — IEEE single precision, no contraction
static int latent_len_f32(const float* dur, int B, int sr, int cs) {
    volatile float mx = dur[0];
    for (int b = 1; b < B; ++b) if (dur[b] > mx) mx = dur[b];
    volatile float wav_len_max = mx * (float)sr;
    volatile float t = wav_len_max + (float)cs;
    t = t - 1.0f;
    volatile float q = t / (float)cs;
    return (int)q;
}

void Handle::synth_tail(const float* d_text_emb, const Seq& text, const float* d_style_ttl, const float* d_noise, int64_t noise_ld,
                        uint64_t seed, const Seq& lat, int steps, float* d_xlat, float* d_wav, const int* d_noise_index) {
    int D = cfg.latent_channels;
    std::vector<const float*> tv(steps);
    for (int s = 0; s < steps; ++s) tv[s] = time_vectors((float)s, (float)steps);
    std::vector<int> sbits = {(int)(uint32_t)(seed & 0xffffffffu), (int)(uint32_t)(seed >> 32)};
    const uint64_t* d_seed = reinterpret_cast<const uint64_t*>(stage_ints(sbits));     // per call, outside the graph's baked arguments
    STC_LAUNCH(this, init_latent_kernel, cdiv((size_t)lat.rows * D, 256), 256, 0, d_noise, noise_ld, d_seed, lat.mask, d_xlat, lat.rows, lat.off, lat.B, D, d_noise_index);
    VeCtx vc; vc.lat = lat; vc.text = text; vc.style = rect_seq(text.B, cfg.style_ttl_tokens, nullptr, false);
    prepare_ve(vc, d_text_emb, d_style_ttl);
    float* d_dt = ws<float>(D);
    STC_LAUNCH(this, fill_kernel, cdiv(D, 128), 128, 0, d_dt, 1.0f / (float)steps, (size_t)D);
    for (int s = 0; s < steps; ++s) run_ve_step(vc, d_xlat, tv[s], d_dt);
    if (profile && !dry) cudaEventRecord(ev[3], stream);
    run_vocoder(d_xlat, lat, d_wav);
}

}  // namespace stc

extern "C" {

// mode bits: 1 = host I/O (else device pointers), 2 = packed latent rows (else the reference's padded rectangle)
static int synth_impl(stc_handle* sh, int mode, const int64_t* text_ids, const float* text_mask, const float* style_ttl,
                      const float* style_dp, int B, int T, int total_step, float speed, const float* noise, int64_t noise_ld,
                      uint64_t seed, float* wav_out, int64_t wav_cap, float* duration_out, int64_t* wav_lengths_out, int64_t* L_out,
                      float* latent_out, int64_t* wav_offsets_out, const int32_t* text_lens, bool async_copy = false,
                      const stc_out_opts* oo = nullptr) {
    STC_TRY(sh, {
        const bool host_io = mode & 1, packed = mode & 2;
        Scope sc(sh, async_copy); Handle* h = sc.h;
        // output packing (packed host-I/O entry points only): PCM16 quantisation and / or silence between utterances on the device
        const bool pcm = oo && oo->pcm16;
        const int64_t gap = oo ? oo->gap_samples : 0;
        const bool pack = pcm || gap > 0;
        if (pack && !(host_io && packed)) throw StcError(STC_ERR_INVALID, "output options need the packed host entry points");
        if (gap < 0) throw StcError(STC_ERR_INVALID, "gap_samples < 0");
        const size_t esz = pcm ? sizeof(int16_t) : sizeof(float);
        if (pack) mode |= 256 | (pcm ? 512 : 0);               // graphs of the packing forms carry one more kernel
        const int64_t* nidx = oo ? oo->noise_index : nullptr;
        if (nidx && !(host_io && packed)) throw StcError(STC_ERR_INVALID, "noise_index needs the packed host entry points");
        if (nidx) mode |= 1024;
        if (async_copy && !(host_io && packed && !latent_out && !h->profile && h->use_graphs)) async_copy = false;
        int slot = 0;
        if (async_copy) {                              // this call's half of the pinned offset staging + its device result buffer
            slot = h->slot; h->slot ^= 1;
            // the call that last used this slot (two calls ago) must have run to the end of its copy before its staging half
            // is rewritten: at most two asynchronous calls are in flight, the library throttles the caller here
            STC_CUDA(cudaEventSynchronize(h->copy_done[slot]));
            h->h_stage_off = slot ? h->h_stage_cap / 2 : 0;
            h->h_stage_lim = slot ? h->h_stage_cap : h->h_stage_cap / 2;
            mode |= 16 | (slot << 5);                  // graphs bake the staging / result addresses: one set per slot
        }
        const bool ovl = async_copy && h->overlap && h->stream_f;
        // odd calls of a request stream keep their stage-1 buffers in the second persistent arena (swapped back when the call returns)
        struct PersistSwap {
            Handle* h; bool on;
            PersistSwap(Handle* hh, bool o) : h(hh), on(o) { if (on) { h->persist.swap(h->persist_alt); h->persist.reset(); } }
            ~PersistSwap() { if (on) h->persist.swap(h->persist_alt); }
        } persist_swap(h, ovl && slot == 1);
        if (ovl) mode |= 128;                          // graphs of the overlapped form bake other addresses than the serial form's
        if (B <= 0 || T <= 0 || total_step <= 0 || !(speed > 0.f) || !text_ids || !text_mask || !style_ttl || !style_dp || !wav_out)
            throw StcError(STC_ERR_INVALID, "stc_synthesize: bad argument");
        const stc_config& c = h->cfg;
        int C = c.text_emb_channels, S = c.style_ttl_tokens, Cs = c.style_ttl_dim, D = c.latent_channels;
        int si = c.style_dp_tokens * c.style_dp_dim;
        if (host_io) validate_ids(text_ids, (size_t)B * T, c.vocab_size);
        if (h->h_cap < B) {
            std::lock_guard<std::recursive_mutex> lk(g_capture_mu);
            if (h->h_dur) cudaFreeHost(h->h_dur);
            if (h->h_wavlen) cudaFreeHost(h->h_wavlen);
            STC_CUDA(cudaMallocHost((void**)&h->h_dur, sizeof(float) * B));
            STC_CUDA(cudaMallocHost((void**)&h->h_wavlen, sizeof(int64_t) * B));
            h->h_cap = B;
            for (auto& g : h->graphs) cudaGraphExecDestroy(g.second.exec);     // graphs captured the old pinned addresses
            h->graphs.clear();
        }
        // Text side: only the real tokens are computed (packed rows) when the token counts are known on the host — from the
        // mask itself (host I/O; it must be the prefix mask the reference builds, cpp/helper.cpp:740-757) or from `text_lens`
        // (device-resident inputs). Otherwise the reference's padded [B,T] rectangle + row mask.
        std::vector<int> tlens;
        if (packed) {
            tlens.resize(B);
            bool ok = true;
            for (int b = 0; b < B && ok; ++b) {
                if (host_io) {
                    int cnt = 0, last = 0;
                    for (int t = 0; t < T; ++t) if (text_mask[(size_t)b * T + t] != 0.f) { ++cnt; last = t + 1; }
                    tlens[b] = cnt; ok = cnt == last && cnt > 0;
                } else { ok = text_lens && text_lens[b] > 0 && text_lens[b] <= T; if (ok) tlens[b] = text_lens[b]; }
            }
            if (!ok) tlens.clear();
        }
        const bool tpacked = !tlens.empty();
        int trows = 0, tmaxlen = 0;
        if (tpacked) {
            int sum = 0, mx = 0;
            for (int v : tlens) { sum += v; mx = std::max(mx, v); }
            trows = h->use_graphs && !h->profile ? (sum + 127) / 128 * 128 : sum;
            tmaxlen = (mx + 15) / 16 * 16;
        }
        auto text_seq = [&](const float* d_mask, bool want_len) {
            return tpacked ? h->packed_seq(tlens, trows, tmaxlen) : h->rect_seq(B, T, d_mask, want_len);
        };
        const size_t text_rows = tpacked ? (size_t)trows : (size_t)B * T;
        cudaStream_t st = h->stream;
        for (auto& k : h->kprof) k = Handle::KProf{};
        for (int s = 0; s < total_step; ++s) h->time_vectors((float)s, (float)total_step);     // cached after the first call
        // ---- stage 1: DP (+ /speed, wav lengths) and TE; independent of L. Two graphs so that the host can start waiting
        //      for the durations while the text encoder is still running.
        const int64_t* d_ids = nullptr; const float *d_tmask = nullptr, *d_sttl = nullptr, *d_sdp = nullptr;
        float *d_dur = nullptr, *d_temb = nullptr; int64_t* d_wavlen = nullptr;
        const uintptr_t pin = 0;
        std::array<uintptr_t, 4> pins{};
        if (!host_io) pins = {(uintptr_t)text_ids, (uintptr_t)text_mask, (uintptr_t)style_ttl, (uintptr_t)style_dp};
        auto keyed = [&](GraphKey k) { k.in = pins; k.gap = gap; return k; };
        const size_t stage_base = h->h_stage_off;
        auto stage1a = [&]() {
            h->arena.reset(); h->persist.reset(); h->h_stage_off = stage_base;
            if (host_io) {
                d_ids = upp(h, text_ids, (size_t)B * T); d_tmask = upp(h, text_mask, (size_t)B * T);
                d_sttl = upp(h, style_ttl, (size_t)B * S * Cs); d_sdp = upp(h, style_dp, (size_t)B * si);
            } else { d_ids = text_ids; d_tmask = text_mask; d_sttl = style_ttl; d_sdp = style_dp; }
            d_dur = h->ps<float>(B); d_wavlen = h->ps<int64_t>(B); d_temb = h->ps<float>(text_rows * C);
            h->run_dp(d_ids, d_sdp, text_seq(d_tmask, false), T, d_dur);
            STC_LAUNCH(h, dur_post_kernel, cdiv(B, 128), 128, 0, d_dur, d_wavlen, B, speed, c.sample_rate);
            if (!h->dry) {
                // (h->stream: the main stream, or stream_f when this stage runs ahead of the previous call's stage 2)
                STC_CUDA(cudaMemcpyAsync(h->h_dur, d_dur, sizeof(float) * B, cudaMemcpyDeviceToHost, h->stream));
                STC_CUDA(cudaMemcpyAsync(h->h_wavlen, d_wavlen, sizeof(int64_t) * B, cudaMemcpyDeviceToHost, h->stream));
            }
        };
        auto stage1b = [&]() {
            h->arena.reset();
            h->run_te(d_ids, d_sttl, text_seq(d_tmask, true), T, d_temb);      // overlaps the D2H of the durations
        };
        uint32_t speed_bits; memcpy(&speed_bits, &speed, 4);
        if (h->profile) cudaEventRecord(h->ev[0], st);
        // duration predictor on the main stream (behind the input uploads), text encoder concurrently on stream2 with its own
        // workspace: the two are independent (cpp/helper.cpp:512-556 runs them back to back) and DP alone leaves the GPU idle
        if (ovl) {
            std::swap(h->stream, h->stream_f); h->arena.swap(h->arena_f);
            try { h->run_graphed(keyed(GraphKey{1, mode, B, T, 0, 0, 0, (int64_t)speed_bits, pin, 0, trows, tmaxlen}), stage1a, h->ev_in); }
            catch (...) { std::swap(h->stream, h->stream_f); h->arena.swap(h->arena_f); throw; }
            cudaEventRecord(h->ev[6], h->stream);
            std::swap(h->stream, h->stream_f); h->arena.swap(h->arena_f);
            STC_CUDA(cudaStreamWaitEvent(st, h->ev[6], 0));           // stage 2 reads the durations / wav lengths on the device
        } else {
            h->run_graphed(keyed(GraphKey{1, mode, B, T, 0, 0, 0, (int64_t)speed_bits, pin, 0, trows, tmaxlen}), stage1a, h->ev_in);
            if (h->profile) cudaEventRecord(h->ev[1], st);
            cudaEventRecord(h->ev[6], st);
        }
        {
            STC_CUDA(cudaStreamWaitEvent(h->stream2, h->ev_in, 0));
            std::swap(h->stream, h->stream2); h->arena.swap(h->arena2);
            try { h->run_graphed(keyed(GraphKey{2, mode, B, T, 0, 0, 0, 0, pin, 0, trows, tmaxlen}), stage1b); }
            catch (...) { std::swap(h->stream, h->stream2); h->arena.swap(h->arena2); throw; }
            cudaEventRecord(h->ev_te, h->stream);
            std::swap(h->stream, h->stream2); h->arena.swap(h->arena2);
            STC_CUDA(cudaStreamWaitEvent(st, h->ev_te, 0));            // everything after this point on the main stream sees text_emb
        }
        if (h->profile) cudaEventRecord(h->ev[2], st);
        const auto t_host0 = std::chrono::steady_clock::now();
        STC_CUDA(cudaEventSynchronize(h->ev[6]));                 // the one data-dependent sync: duration -> L
        const auto t_host1 = std::chrono::steady_clock::now();
        int L = latent_len_f32(h->h_dur, B, c.sample_rate, c.chunk_size);
        if (L_out) *L_out = L;
        if (duration_out && host_io) memcpy(duration_out, h->h_dur, sizeof(float) * B);
        if (wav_lengths_out) memcpy(wav_lengths_out, h->h_wavlen, sizeof(int64_t) * B);
        if (L <= 0) throw StcError(STC_ERR_INVALID, "computed latent length is 0");
        // latent frames per utterance: integer formula of getLatentMask (cpp/helper.cpp:767)
        std::vector<int> lens(B);
        int64_t R = 0; int maxlen = 0;
        for (int b = 0; b < B; ++b) {
            lens[b] = (int)((h->h_wavlen[b] + c.chunk_size - 1) / c.chunk_size);
            R += lens[b]; maxlen = std::max(maxlen, lens[b]);
        }
        // utterance b's samples start at wav_offsets_out[b]; with a gap the silence follows them, so [B] (the total) excludes a trailing gap
        if (wav_offsets_out) {
            wav_offsets_out[0] = 0;
            for (int b = 0; b < B; ++b) wav_offsets_out[b + 1] = wav_offsets_out[b] + (int64_t)lens[b] * c.chunk_size + (b + 1 < B ? gap : 0);
        }
        // graph buckets: packed rows are rounded up to a multiple of 128 (extra rows belong to no sequence), the
        // longest-utterance bound (attention grid) to a multiple of 16
        const int64_t rows = packed ? (h->use_graphs && !h->profile ? (R + 127) / 128 * 128 : R) : (int64_t)B * L;
        const int maxlen_launch = packed ? (maxlen + 15) / 16 * 16 : L;
        const int64_t wav_need = packed ? rows * c.chunk_size + (int64_t)(B - 1) * gap : (int64_t)L * c.chunk_size;   // total elements (packed) / per row (rectangle)
        const int64_t wav_real = R * c.chunk_size + (int64_t)(B - 1) * gap;                 // elements the caller receives (packed)
        if (wav_cap < wav_need) {
            if (wav_offsets_out) wav_offsets_out[B] = std::max<int64_t>(wav_offsets_out[B], wav_need);
            throw StcError(STC_ERR_CAPACITY, "wav_out too small: need " + std::to_string(wav_need));
        }
        if (noise && noise_ld < (packed ? maxlen : L)) throw StcError(STC_ERR_CAPACITY, "noise_ld smaller than the latent length");
        // ---- stage 2: everything that depends on L
        float *d_noise = nullptr, *d_lmask = nullptr, *d_xlat = nullptr, *d_wav = nullptr, *d_lat_ncl = nullptr;
        void* d_out = nullptr;                 // what is copied to the caller: d_wav itself, or the packed / quantised copy of it
        if (async_copy) {
            const size_t need = (size_t)wav_need;          // (elements; sized as floats, so a PCM16 result uses half of it)
            if (h->outcap[slot] < need) {
                std::lock_guard<std::recursive_mutex> lk(g_capture_mu);
                STC_CUDA(cudaEventSynchronize(h->copy_done[slot]));                  // the copy that last read this buffer
                if (h->outbuf[slot]) cudaFree(h->outbuf[slot]);
                h->outbuf[slot] = nullptr; h->outcap[slot] = 0;
                STC_CUDA(cudaMalloc((void**)&h->outbuf[slot], (need + need / 4) * sizeof(float)));
                h->outcap[slot] = need + need / 4;
            }
            STC_CUDA(cudaStreamWaitEvent(st, h->copy_done[slot], 0));                // device-side: do not overwrite before it is copied out
        }
        auto stage2 = [&]() {
            h->arena.reset();          // (the pinned offset staging keeps growing: stage-1 copies may still be in flight)
            d_noise = nullptr;
            if (noise) d_noise = up(h, noise, (size_t)B * D * noise_ld);
            d_xlat = h->ws<float>((size_t)rows * D);
            d_wav = (async_copy && !pack) ? h->outbuf[slot] : host_io ? h->ws<float>((size_t)rows * c.chunk_size) : wav_out;
            d_out = d_wav;
            if (pack) d_out = async_copy ? (void*)h->outbuf[slot] : (void*)h->ws<unsigned char>((size_t)wav_need * esz);
            Seq text = text_seq(d_tmask, true), lat;
            if (packed) lat = h->packed_seq(lens, (int)rows, maxlen_launch);
            else {
                d_lmask = h->ws<float>((size_t)B * L);
                STC_LAUNCH(h, latent_mask_kernel, cdiv((size_t)B * L, 256), 256, 0, d_wavlen, d_lmask, B, L, c.chunk_size);
                lat = h->rect_seq(B, L, d_lmask, true);
                if (latent_out) d_lat_ncl = h->ws<float>((size_t)B * L * D);
            }
            const int* d_nidx = nullptr;
            if (nidx) {
                std::vector<int> ni(B);
                for (int b = 0; b < B; ++b) ni[b] = (int)(nidx[b] & 0x7fffffff);
                d_nidx = h->stage_ints(ni);
            }
            h->synth_tail(d_temb, text, d_sttl, d_noise, noise_ld, seed, lat, total_step, d_xlat, d_wav, d_nidx);
            if (pack) {
                if (pcm) STC_LAUNCH(h, wav_pack_kernel<int16_t>, (unsigned)rows, 256, 0, (const float*)d_wav, (int16_t*)d_out, lat.off, B, c.chunk_size, (long long)gap);
                else STC_LAUNCH(h, wav_pack_kernel<float>, (unsigned)rows, 256, 0, (const float*)d_wav, (float*)d_out, lat.off, B, c.chunk_size, (long long)gap);
            }
        };
        h->run_graphed(keyed(GraphKey{3, mode | (latent_out ? 4 : 0), B, T, (int)rows, maxlen_launch, total_step, noise ? noise_ld : -1,
                                      pin, async_copy ? (uintptr_t)h->outbuf[slot] : host_io ? 0 : (uintptr_t)wav_out, trows, tmaxlen}), stage2);
        if (getenv("STC_TIMING")) {
            const auto t_host2 = std::chrono::steady_clock::now();
            fprintf(stderr, "[stc timing] wait for durations %.1f us, host work until stage-2 launch returned %.1f us\n",
                    std::chrono::duration<double, std::micro>(t_host1 - t_host0).count(),
                    std::chrono::duration<double, std::micro>(t_host2 - t_host1).count());
        }
        if (h->profile) cudaEventRecord(h->ev[4], st);
        if (async_copy) {
            // the copy rides its own stream; this call returns as soon as it is enqueued (stc_wait delivers it)
            STC_CUDA(cudaEventRecord(h->ev_out, st));
            STC_CUDA(cudaStreamWaitEvent(h->stream_copy, h->ev_out, 0));
            STC_CUDA(cudaMemcpyAsync(wav_out, d_out, (size_t)wav_real * esz, cudaMemcpyDeviceToHost, h->stream_copy));
            STC_CUDA(cudaEventRecord(h->copy_done[slot], h->stream_copy));
            h->async_pending = true;
            return STC_OK;
        }
        if (host_io) {
            if (packed) {
                STC_CUDA(cudaMemcpyAsync(wav_out, d_out, (size_t)wav_real * esz, cudaMemcpyDeviceToHost, st));
                if (latent_out) STC_CUDA(cudaMemcpyAsync(latent_out, d_xlat, (size_t)R * D * sizeof(float), cudaMemcpyDeviceToHost, st));
            } else {
                int64_t wav_row = (int64_t)L * c.chunk_size;
                STC_CUDA(cudaMemcpy2DAsync(wav_out, (size_t)wav_cap * sizeof(float), d_wav, (size_t)wav_row * sizeof(float),
                                           (size_t)wav_row * sizeof(float), B, cudaMemcpyDeviceToHost, st));
                if (latent_out) {
                    transpose(h, d_xlat, d_lat_ncl, B, L, D);
                    STC_CUDA(cudaMemcpyAsync(latent_out, d_lat_ncl, (size_t)B * L * D * sizeof(float), cudaMemcpyDeviceToHost, st));
                }
            }
        } else if (duration_out) {
            STC_CUDA(cudaMemcpyAsync(duration_out, d_dur, sizeof(float) * B, cudaMemcpyDeviceToDevice, st));
        }
        if (h->profile) cudaEventRecord(h->ev[5], st);
        STC_CUDA(cudaStreamSynchronize(st));
        h->check_launch("stc_synthesize");
        if (h->profile) {
            h->kprof_resolve();
            cudaEventElapsedTime(&h->stage_ms[0], h->ev[0], h->ev[1]);
            cudaEventElapsedTime(&h->stage_ms[1], h->ev[1], h->ev[2]);
            cudaEventElapsedTime(&h->stage_ms[2], h->ev[2], h->ev[3]);
            cudaEventElapsedTime(&h->stage_ms[3], h->ev[3], h->ev[4]);
            cudaEventElapsedTime(&h->stage_ms[4], h->ev[0], h->ev[5]);
        }
    })
}

int stc_synthesize(stc_handle* h, const int64_t* text_ids, const float* text_mask, const float* style_ttl, const float* style_dp,
                   int B, int T, int total_step, float speed, const float* noise, int64_t noise_ld, uint64_t seed, float* wav_out,
                   int64_t wav_ld, float* duration_out, int64_t* wav_lengths_out, int64_t* L_out, float* latent_out) {
    return synth_impl(h, 1, text_ids, text_mask, style_ttl, style_dp, B, T, total_step, speed, noise, noise_ld, seed, wav_out, wav_ld,
                      duration_out, wav_lengths_out, L_out, latent_out, nullptr, nullptr);
}

int stc_synthesize_device(stc_handle* h, const int64_t* text_ids_dev, const float* text_mask_dev, const float* style_ttl_dev,
                          const float* style_dp_dev, int B, int T, int total_step, float speed, uint64_t seed, float* wav_dev,
                          int64_t wav_ld, float* duration_dev, int64_t* L_out) {
    return synth_impl(h, 0, text_ids_dev, text_mask_dev, style_ttl_dev, style_dp_dev, B, T, total_step, speed, nullptr, 0, seed,
                      wav_dev, wav_ld, duration_dev, nullptr, L_out, nullptr, nullptr, nullptr);
}

int stc_synthesize_packed(stc_handle* h, const int64_t* text_ids, const float* text_mask, const float* style_ttl, const float* style_dp,
                          int B, int T, int total_step, float speed, const float* noise, int64_t noise_ld, uint64_t seed,
                          float* wav_out, int64_t wav_cap, int64_t* wav_offsets_out, float* duration_out, int64_t* wav_lengths_out,
                          float* latent_out) {
    return synth_impl(h, 3, text_ids, text_mask, style_ttl, style_dp, B, T, total_step, speed, noise, noise_ld, seed, wav_out, wav_cap,
                      duration_out, wav_lengths_out, nullptr, latent_out, wav_offsets_out, nullptr);
}

int stc_synthesize_packed_async(stc_handle* h, const int64_t* text_ids, const float* text_mask, const float* style_ttl,
                                const float* style_dp, int B, int T, int total_step, float speed, uint64_t seed, float* wav_out_pinned,
                                int64_t wav_cap, int64_t* wav_offsets_out, float* duration_out, int64_t* wav_lengths_out) {
    return synth_impl(h, 3, text_ids, text_mask, style_ttl, style_dp, B, T, total_step, speed, nullptr, 0, seed, wav_out_pinned, wav_cap,
                      duration_out, wav_lengths_out, nullptr, nullptr, wav_offsets_out, nullptr, true);
}

int stc_synthesize_packed_ex(stc_handle* h, const int64_t* text_ids, const float* text_mask, const float* style_ttl, const float* style_dp,
                             int B, int T, int total_step, float speed, const float* noise, int64_t noise_ld, uint64_t seed,
                             const stc_out_opts* opts, void* out, int64_t out_cap, int64_t* wav_offsets_out, float* duration_out,
                             int64_t* wav_lengths_out, int async_copy) {
    if (async_copy && noise) return fail(h, STC_ERR_INVALID, "stc_synthesize_packed_ex: the asynchronous form takes no injected noise");
    return synth_impl(h, 3, text_ids, text_mask, style_ttl, style_dp, B, T, total_step, speed, noise, noise_ld, seed, static_cast<float*>(out), out_cap,
                      duration_out, wav_lengths_out, nullptr, nullptr, wav_offsets_out, nullptr, async_copy != 0, opts);
}

int stc_wait(stc_handle* sh) {
    STC_TRY(sh, {
        if (!sh || !sh->impl) throw StcError(STC_ERR_INVALID, "null handle");
        STC_CUDA(cudaSetDevice(sh->impl->device));
        sh->impl->wait_async();
    })
}

int stc_synthesize_packed_device(stc_handle* h, const int64_t* text_ids_dev, const float* text_mask_dev, const float* style_ttl_dev,
                                 const float* style_dp_dev, const int32_t* text_lens, int B, int T, int total_step, float speed,
                                 uint64_t seed, float* wav_dev, int64_t wav_cap, int64_t* wav_offsets_out, float* duration_dev) {
    return synth_impl(h, 2, text_ids_dev, text_mask_dev, style_ttl_dev, style_dp_dev, B, T, total_step, speed, nullptr, 0, seed,
                      wav_dev, wav_cap, duration_dev, nullptr, nullptr, nullptr, wav_offsets_out, text_lens);
}

// ---- debug / tuning: one tcgen05 GEMM of a given shape and launch configuration, timed and checked -----------------
__global__ void debug_fill_kernel(float* __restrict__ p, size_t n, uint64_t seed, float scale) {
    pdl_wait(); pdl_trigger_light();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float u1 = (stc::mix32(seed + 2 * i) + 1.0f) * (1.0f / 4294967808.0f), u2 = stc::mix32(seed + 2 * i + 1) * (1.0f / 4294967296.0f);
    p[i] = scale * sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}
__global__ void debug_maxdiff_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n, float* __restrict__ out) {
    pdl_wait(); pdl_trigger_light();
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(a[i] - b[i]));
    m = stc::warp_max(m);
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));     // non-negative floats order like ints
}
__global__ void debug_join_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, float* __restrict__ out, size_t n) {
    pdl_wait(); pdl_trigger_light();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __bfloat162float(hi[i]) + __bfloat162float(lo[i]);
}
__global__ void debug_unhalf_kernel(const __half* __restrict__ h, float* __restrict__ out, size_t n) {
    pdl_wait(); pdl_trigger_light();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __half2float(h[i]);
}

int stc_debug_gemm(stc_handle* sh, int M, int N, int K, int bn, int cm, int cn, int epilogue, int iters, float* ms_per_iter,
                   float* max_abs_err) {
    STC_TRY(sh, {
        Scope sc(sh); Handle* h = sc.h;
        if (!h->tc_mode()) throw StcError(STC_ERR_UNSUPPORTED, "stc_debug_gemm needs the tcgen05 precision mode");
        if (M <= 0 || N <= 0 || K <= 0 || iters <= 0 || epilogue < 0 || epilogue > 2) throw StcError(STC_ERR_INVALID, "stc_debug_gemm: bad argument");
        if (bn && (bn != 64 && bn != 128 && bn != 256 && bn != 512)) throw StcError(STC_ERR_INVALID, "bn must be 0, 64, 128, 256 or 512 (two-SM 256 x 256)");
        if (bn != 512 && (cm != 1 || cn != 1)) throw StcError(STC_ERR_UNSUPPORTED, "cluster shapes other than 1 x 1 were measured slower and removed (DESIGN.md)");
        std::vector<float> wk((size_t)K * N), bias(N);
        uint64_t z = 12345;
        auto rnd = [&]() { z = z * 6364136223846793005ull + 1442695040888963407ull; return (float)((int64_t)(z >> 11) % 2000001 - 1000000) * 1e-6f; };
        for (auto& v : wk) v = rnd() / std::sqrt((float)K) * 1.7f;
        for (auto& v : bias) v = rnd() * 0.1f;
        const bool f16 = getenv("STC_DEBUG_F16") != nullptr;            // time / check the single-pass fp16 instantiation instead
        Linear lin = h->make_linear_host(wk, bias, K, N, f16 ? 3 : 1);
        auto body = [&]() {
            h->arena.reset(); h->h_stage_off = 0;
            float* A = h->ws<float>((size_t)M * K); float* X = h->ws<float>((size_t)M * N); float* Xr = h->ws<float>((size_t)M * N);
            float* gamma = h->ws<float>(N); float* mask = h->ws<float>(M);
            float* out = h->ws<float>((size_t)M * N); float* ref = h->ws<float>((size_t)M * N); float* err = h->ws<float>(1);
            Act a = h->ws_act_for(lin, (size_t)M * K), o = h->ws_act_for(lin, (size_t)M * N);
            if (h->dry) { h->ws<long long>(128); return; }
            debug_fill_kernel<<<cdiv((size_t)M * K, 256), 256, 0, h->stream>>>(A, (size_t)M * K, 1, 1.0f);
            debug_fill_kernel<<<cdiv((size_t)M * N, 256), 256, 0, h->stream>>>(X, (size_t)M * N, 2, 1.0f);
            debug_fill_kernel<<<cdiv(N, 256), 256, 0, h->stream>>>(gamma, N, 3, 0.1f);
            fill_kernel<<<cdiv(M, 256), 256, 0, h->stream>>>(mask, 1.0f, (size_t)M);
            fill_kernel<<<1, 32, 0, h->stream>>>(mask + M / 2, 0.0f, (size_t)std::min(M - M / 2, 3));
            fill_kernel<<<1, 32, 0, h->stream>>>(err, 0.0f, (size_t)1);
            h->to_act(A, (size_t)M * K, a);
            Epilogue ep;
            if (epilogue == 1) ep.gelu = 1;
            if (epilogue == 2) { ep.scale = gamma; ep.mask = mask; }
            // reference: CUDA-core fp32 GEMM, same epilogue
            Epilogue er = ep; er.bias = lin.bias;
            STC_CUDA(cudaMemcpyAsync(Xr, X, (size_t)M * N * 4, cudaMemcpyDeviceToDevice, h->stream));
            if (epilogue == 2) er.resid = Xr;
            h->gemm_simt<float>(A, K, M, lin, er, epilogue == 2 ? Xr : ref, N);
            const float* refp = epilogue == 2 ? Xr : ref;
            h->force_bn = bn;
            cudaEvent_t e0 = h->pool_event(), e1 = h->pool_event();
            auto one = [&](bool checked) {
                if (epilogue == 2) {
                    if (checked) STC_CUDA(cudaMemcpyAsync(out, X, (size_t)M * N * 4, cudaMemcpyDeviceToDevice, h->stream));
                    Epilogue e2 = ep; e2.resid = out;
                    h->gemm(a, M, lin, e2, out, nullptr, N);
                } else if (epilogue == 1) h->gemm(a, M, lin, ep, nullptr, &o, N);
                else h->gemm(a, M, lin, ep, out, nullptr, N);
            };
            if (getenv("STC_GEMM_TRACE")) {
                long long* tr = h->ws<long long>(128);
                STC_CUDA(cudaMemsetAsync(tr, 0, 128 * 8, h->stream));
                one(true);
                h->gemm_trace = tr; one(true); h->gemm_trace = nullptr;
                long long ht[128];
                STC_CUDA(cudaMemcpyAsync(ht, tr, 128 * 8, cudaMemcpyDeviceToHost, h->stream));
                STC_CUDA(cudaStreamSynchronize(h->stream));
                fprintf(stderr, "gemm trace (cycles; MMA warp after full-wait [0..47], producer after empty-wait [64..111], tile starts [120..]):");
                for (int i = 0; i < 128; ++i) fprintf(stderr, "%s%lld", i % 8 == 0 ? "\n  " : " ", ht[i] ? ht[i] - ht[64] : -99999);
                fprintf(stderr, "\n");
            }
            one(true);
            if (epilogue == 1 && f16) debug_unhalf_kernel<<<cdiv((size_t)M * N, 256), 256, 0, h->stream>>>(reinterpret_cast<const __half*>(o.hi), out, (size_t)M * N);
            else if (epilogue == 1) debug_join_kernel<<<cdiv((size_t)M * N, 256), 256, 0, h->stream>>>(o.hi, o.lo, out, (size_t)M * N);
            debug_maxdiff_kernel<<<592, 256, 0, h->stream>>>(out, refp, (size_t)M * N, err);
            // timed: `iters` launches replayed from a CUDA graph (as in production), so the host is not in the loop
            cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
            STC_CUDA(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            for (int it = 0; it < iters; ++it) one(false);
            STC_CUDA(cudaStreamEndCapture(h->stream, &graph));
            STC_CUDA(cudaGraphInstantiate(&exec, graph, 0));
            STC_CUDA(cudaGraphLaunch(exec, h->stream));        // warm
            cudaEventRecord(e0, h->stream);
            STC_CUDA(cudaGraphLaunch(exec, h->stream));
            cudaEventRecord(e1, h->stream);
            h->force_bn = 0;
            STC_CUDA(cudaStreamSynchronize(h->stream));
            h->check_launch("stc_debug_gemm");
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
            if (ms_per_iter) *ms_per_iter = ms / iters;
            if (max_abs_err) STC_CUDA(cudaMemcpy(max_abs_err, err, 4, cudaMemcpyDeviceToHost));
            h->ev_next = 0;
        };
        h->force_bn = 0;
        h->ensure_ws(body);
        body();
        // the temporary weights stay owned by the handle until it is destroyed (debug entry point: acceptable)
    })
}

int stc_debug_mlp(stc_handle* sh, int M, int iters, float* ms_fused, float* ms_unfused, float* max_abs_err) {
    STC_TRY(sh, {
        Scope sc(sh); Handle* h = sc.h;
        if (!h->tc_mode() || M <= 0 || iters <= 0) throw StcError(STC_ERR_INVALID, "stc_debug_mlp: bad argument");
        const int C = mlp::C, H = mlp::H;
        uint64_t z = 777;
        auto rnd = [&]() { z = z * 6364136223846793005ull + 1442695040888963407ull; return (float)((int64_t)(z >> 11) % 2000001 - 1000000) * 1e-6f; };
        std::vector<float> w1((size_t)C * H), b1(H), w2((size_t)H * C), b2(C), gm(C);
        for (auto& v : w1) v = rnd() / 16.f * 1.7f;
        for (auto& v : w2) v = rnd() / 32.f * 1.7f;
        for (auto& v : b1) v = rnd() * 0.1f;
        for (auto& v : b2) v = rnd() * 0.1f;
        for (auto& v : gm) v = 0.1f + rnd() * 0.02f;
        ConvNeXt cn{};
        cn.C = C; cn.H = H; cn.pw1 = h->make_linear_host(w1, b1, C, H, true); cn.pw2 = h->make_linear_host(w2, b2, H, C, true);
        cn.gamma = h->upload_f32(gm.data(), gm.size());
        auto body = [&]() {
            h->arena.reset(); h->h_stage_off = 0;
            float* A = h->ws<float>((size_t)M * C); float* X0 = h->ws<float>((size_t)M * C);
            float* Xa = h->ws<float>((size_t)M * C); float* Xb = h->ws<float>((size_t)M * C);
            float* mask = h->ws<float>(M); float* err = h->ws<float>(1);
            Act a = h->ws_act((size_t)M * C), hid = h->ws_act((size_t)M * H);
            if (h->dry) { h->ws<float>((size_t)cdiv(M, mlp::BM) * mlp::BM * mlp::C * 16); h->ws<long long>(64); return; }     // partial outputs (<= 16 slices)
            debug_fill_kernel<<<cdiv((size_t)M * C, 256), 256, 0, h->stream>>>(A, (size_t)M * C, 1, 1.0f);
            debug_fill_kernel<<<cdiv((size_t)M * C, 256), 256, 0, h->stream>>>(X0, (size_t)M * C, 2, 1.0f);
            fill_kernel<<<cdiv(M, 256), 256, 0, h->stream>>>(mask, 1.0f, (size_t)M);
            fill_kernel<<<1, 32, 0, h->stream>>>(mask + M / 3, 0.0f, (size_t)std::min(M - M / 3, 2));
            fill_kernel<<<1, 32, 0, h->stream>>>(err, 0.0f, (size_t)1);
            h->to_act(A, (size_t)M * C, a);
            auto fused = [&](float* x) { h->fused_mlp(a, M, cn, x, mask); };
            auto unfused = [&](float* x) {
                Epilogue e1; e1.gelu = 1;
                Epilogue e2; e2.scale = cn.gamma; e2.resid = x; e2.mask = mask;
                h->gemm(a, M, cn.pw1, e1, nullptr, &hid, H);
                h->gemm(hid, M, cn.pw2, e2, x, nullptr, C);
            };
            STC_CUDA(cudaMemcpyAsync(Xa, X0, (size_t)M * C * 4, cudaMemcpyDeviceToDevice, h->stream));
            STC_CUDA(cudaMemcpyAsync(Xb, X0, (size_t)M * C * 4, cudaMemcpyDeviceToDevice, h->stream));
            if (getenv("STC_MLP_TRACE")) {
                long long* tr = h->ws<long long>(64);
                STC_CUDA(cudaMemsetAsync(tr, 0, 64 * 8, h->stream));
                fused(Xa);                                   // warm (descriptors, L2)
                h->mlp_trace = tr; fused(Xa); h->mlp_trace = nullptr;
                long long ht[64];
                STC_CUDA(cudaMemcpyAsync(ht, tr, 64 * 8, cudaMemcpyDeviceToHost, h->stream));
                STC_CUDA(cudaStreamSynchronize(h->stream));
                fprintf(stderr, "mlp trace M=%d (cycles since prologue end; [8..23] MMA warp per weight unit, [24..27] S chunk ready, [28..31] P chunk written, [32,33] output half final, [35] stores issued, [36] store reads drained, [37] CTA joined, [34] end, [40..55] producer per unit):", M);
                for (int i = 0; i < 64; ++i) fprintf(stderr, "%s%lld", i % 8 == 0 ? "\n  " : " ", ht[i] ? ht[i] - ht[0] : -99999);
                fprintf(stderr, "\n");
                STC_CUDA(cudaMemcpyAsync(Xa, X0, (size_t)M * C * 4, cudaMemcpyDeviceToDevice, h->stream));
            }
            fused(Xa); unfused(Xb);
            debug_maxdiff_kernel<<<592, 256, 0, h->stream>>>(Xa, Xb, (size_t)M * C, err);
            float out_ms[2] = {0, 0};
            for (int which = 0; which < 2; ++which) {
                cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
                cudaEvent_t e0 = h->pool_event(), e1 = h->pool_event();
                STC_CUDA(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
                for (int it = 0; it < iters; ++it) { if (which == 0) fused(Xa); else unfused(Xb); }
                STC_CUDA(cudaStreamEndCapture(h->stream, &graph));
                STC_CUDA(cudaGraphInstantiate(&exec, graph, 0));
                STC_CUDA(cudaGraphLaunch(exec, h->stream));
                cudaEventRecord(e0, h->stream);
                STC_CUDA(cudaGraphLaunch(exec, h->stream));
                cudaEventRecord(e1, h->stream);
                STC_CUDA(cudaStreamSynchronize(h->stream));
                cudaEventElapsedTime(&out_ms[which], e0, e1);
                cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
            }
            h->check_launch("stc_debug_mlp");
            if (ms_fused) *ms_fused = out_ms[0] / iters;
            if (ms_unfused) *ms_unfused = out_ms[1] / iters;
            if (max_abs_err) STC_CUDA(cudaMemcpy(max_abs_err, err, 4, cudaMemcpyDeviceToHost));
            h->ev_next = 0;
        };
        h->ensure_ws(body);
        body();
    })
}

// Depthwise conv + LayerNorm in isolation: B ragged sequences packed into `rows` rows (the last rows/16 are bucket padding),
// sliding-window kernel (chains of `rt` rows, 0 = heuristic) against the shared-memory tiled kernel on the same input.
int stc_debug_dwconv(stc_handle* sh, int rows, int C, int K, int dil, int causal, int B, int rt, int iters, float* ms_slide,
                     float* ms_tile, float* max_abs_diff) {
    STC_TRY(sh, {
        Scope sc(sh); Handle* h = sc.h;
        if (rows <= 0 || B <= 0 || iters <= 0 || (C != 128 && C != 256 && C != 512) || (K != 5 && K != 7) || dil < 1)
            throw StcError(STC_ERR_INVALID, "stc_debug_dwconv: bad argument");
        uint64_t z = 4242;
        auto rnd = [&]() { z = z * 6364136223846793005ull + 1442695040888963407ull; return (float)((int64_t)(z >> 11) % 2000001 - 1000000) * 1e-6f; };
        std::vector<float> wt((size_t)K * C), wb(C), gm(C), bt(C);
        for (auto& v : wt) v = rnd() * 0.5f;
        for (auto& v : wb) v = rnd() * 0.1f;
        for (auto& v : gm) v = 1.0f + rnd() * 0.2f;
        for (auto& v : bt) v = rnd() * 0.1f;
        ConvNeXt cn{};
        cn.C = C; cn.K = K; cn.dil = dil; cn.pad_left = causal ? (K - 1) * dil : (K - 1) * dil / 2;
        cn.dw_wt = h->upload_f32(wt.data(), wt.size()); cn.dw_b = h->upload_f32(wb.data(), wb.size());
        const float* g = h->upload_f32(gm.data(), gm.size()); const float* bta = h->upload_f32(bt.data(), bt.size());
        // ragged lengths (one empty sequence when B > 2), real rows = rows - rows / 16
        std::vector<int> lens(B);
        { int real = rows - rows / 16, left = real;
          for (int i = 0; i < B; ++i) { int want = (i == B - 1) ? left : std::min(left, (int)((real / B) * (0.5f + (rnd() + 1.0f) * 0.5f))); if (B > 2 && i == 1) want = 0; lens[i] = want; left -= want; } }
        auto body = [&]() {
            h->arena.reset(); h->h_stage_off = 0;
            float* X = h->ws<float>((size_t)rows * C);
            float* Ya = h->ws<float>((size_t)rows * C); float* Yb = h->ws<float>((size_t)rows * C);
            float* err = h->ws<float>(1);
            // timed form: split-bf16 operands (8 B per element), or with STC_DEBUG_F16 the single fp16 operand the vocoder uses (6 B)
            Act oa = getenv("STC_DEBUG_F16") ? h->ws_act_f16((size_t)rows * C) : h->ws_act((size_t)rows * C);
            Seq seq = h->packed_seq(lens, rows, 0);
            if (h->dry) return;
            debug_fill_kernel<<<cdiv((size_t)rows * C, 256), 256, 0, h->stream>>>(X, (size_t)rows * C, 3, 1.0f);
            fill_kernel<<<1, 32, 0, h->stream>>>(err, 0.0f, (size_t)1);
            const bool keep = h->dw_slide; const int keep_rt = h->dw_rt;
            auto run = [&](bool slide, float* y, const Act* act) {
                h->dw_slide = slide; h->dw_rt = rt;
                h->dwconv_ln<float>(X, &cn, g, bta, C, seq, 1e-6f, act ? nullptr : y, act);
                h->dw_slide = keep; h->dw_rt = keep_rt;
            };
            run(true, Ya, nullptr); run(false, Yb, nullptr);
            debug_maxdiff_kernel<<<592, 256, 0, h->stream>>>(Ya, Yb, (size_t)rows * C, err);
            float out_ms[2] = {0, 0};
            for (int which = 0; which < 2; ++which) {
                cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
                cudaEvent_t e0 = h->pool_event(), e1 = h->pool_event();
                STC_CUDA(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
                for (int it = 0; it < iters; ++it) run(which == 0, nullptr, &oa);
                STC_CUDA(cudaStreamEndCapture(h->stream, &graph));
                STC_CUDA(cudaGraphInstantiate(&exec, graph, 0));
                STC_CUDA(cudaGraphLaunch(exec, h->stream));
                cudaEventRecord(e0, h->stream);
                STC_CUDA(cudaGraphLaunch(exec, h->stream));
                cudaEventRecord(e1, h->stream);
                STC_CUDA(cudaStreamSynchronize(h->stream));
                cudaEventElapsedTime(&out_ms[which], e0, e1);
                cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
            }
            h->check_launch("stc_debug_dwconv");
            if (ms_slide) *ms_slide = out_ms[0] / iters;
            if (ms_tile) *ms_tile = out_ms[1] / iters;
            if (max_abs_diff) STC_CUDA(cudaMemcpy(max_abs_diff, err, 4, cudaMemcpyDeviceToHost));
            h->ev_next = 0;
        };
        h->ensure_ws(body);
        body();
    })
}

// The device quantiser on caller-provided samples (known-answer tests against the reference's writeWavFile goldens)
int stc_debug_pcm16(stc_handle* sh, const float* samples, int64_t n, int16_t* out) {
    STC_TRY(sh, {
        Scope sc(sh); Handle* h = sc.h;
        if (n <= 0 || !samples || !out) throw StcError(STC_ERR_INVALID, "stc_debug_pcm16: bad argument");
        auto body = [&]() {
            h->arena.reset(); h->h_stage_off = 0;
            float* d_in = up(h, samples, (size_t)n);
            int16_t* d_out = h->ws<int16_t>((size_t)n);
            const int* off = h->stage_ints(std::vector<int>{0, 1});
            STC_LAUNCH(h, wav_pack_kernel<int16_t>, 1, 256, 0, (const float*)d_in, d_out, off, 1, (int)n, (long long)0);
            if (!h->dry) STC_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n * sizeof(int16_t), cudaMemcpyDeviceToHost, h->stream));
        };
        h->ensure_ws(body);
        body();
        STC_CUDA(cudaStreamSynchronize(h->stream));
        h->check_launch("stc_debug_pcm16");
    })
}

// The layer plan the node-pattern matcher (graph_plan.h) derives from one .onnx file, as JSON — host only, no GPU needed.
int stc_derive_arch(const char* onnx_path, const char* kind, char* buf, size_t cap, size_t* need) {
    STC_TRY(nullptr, {
        if (!onnx_path || !kind || !need) throw StcError(STC_ERR_INVALID, "stc_derive_arch: bad argument");
        std::string out;
        try { out = derive_arch(load_onnx(onnx_path), kind).dump(); }
        catch (const PlanError& e) { throw StcError(STC_ERR_UNSUPPORTED, e.what()); }
        catch (const std::runtime_error& e) { throw StcError(STC_ERR_IO, e.what()); }
        *need = out.size() + 1;
        if (!buf || cap < out.size() + 1) throw StcError(STC_ERR_CAPACITY, "stc_derive_arch: buffer too small");
        memcpy(buf, out.c_str(), out.size() + 1);
    })
}

int stc_pinned_alloc(size_t bytes, void** out) {
    STC_TRY(nullptr, {
        if (!out || !bytes) throw StcError(STC_ERR_INVALID, "stc_pinned_alloc: bad argument");
        *out = nullptr;
        std::lock_guard<std::recursive_mutex> lk(stc::g_capture_mu);
        STC_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    })
}
void stc_pinned_free(void* p) {
    if (!p) return;
    std::lock_guard<std::recursive_mutex> lk(stc::g_capture_mu);
    cudaFreeHost(p);
}

int stc_text_to_ids(stc_handle* sh, const char* const* texts, const char* const* langs, int n, int64_t* text_ids, float* text_mask,
                    int64_t T_cap, int64_t* T_out) {
    STC_TRY(sh, {
        if (!sh || !texts || !langs || n <= 0) throw StcError(STC_ERR_INVALID, "stc_text_to_ids: bad argument");
        sh->impl->frontend.call(texts, langs, n, text_ids, text_mask, T_cap, T_out);
    })
}

struct stc_frontend { stc::TextFrontend fe; };

int stc_frontend_open(const char* unicode_indexer_json, stc_frontend** out) {
    STC_TRY(nullptr, {
        if (!unicode_indexer_json || !out) throw StcError(STC_ERR_INVALID, "stc_frontend_open: bad argument");
        auto f = std::make_unique<stc_frontend>();
        try { f->fe.load_indexer(unicode_indexer_json); } catch (const std::exception& e) { throw StcError(STC_ERR_IO, e.what()); }
        *out = f.release();
    })
}
void stc_frontend_close(stc_frontend* fe) { delete fe; }
int stc_frontend_text_to_ids(const stc_frontend* fe, const char* const* texts, const char* const* langs, int n, int64_t* text_ids,
                             float* text_mask, int64_t T_cap, int64_t* T_out) {
    STC_TRY(nullptr, {
        if (!fe || !texts || !langs || n <= 0) throw StcError(STC_ERR_INVALID, "stc_frontend_text_to_ids: bad argument");
        fe->fe.call(texts, langs, n, text_ids, text_mask, T_cap, T_out);
    })
}

int stc_chunk_text(const char* text, int max_len, char* out_buf, size_t out_cap, size_t* out_len, int* n_out) {
    STC_TRY(nullptr, {
        if (!text || !n_out || !out_len) throw StcError(STC_ERR_INVALID, "stc_chunk_text: bad argument");
        std::vector<std::string> chunks = stc::chunk_text(text, max_len);
        size_t need = 0;
        for (auto& c : chunks) need += c.size() + 1;
        *out_len = need; *n_out = (int)chunks.size();
        if (!out_buf || out_cap < need) throw StcError(STC_ERR_CAPACITY, "stc_chunk_text: buffer too small");
        size_t o = 0;
        for (auto& c : chunks) { memcpy(out_buf + o, c.data(), c.size()); o += c.size(); out_buf[o++] = '\0'; }
    })
}

}  // extern "C"
