/*
 * Drop-in <onnxruntime_cxx_api.h> for the reference's C++ port, backed by libsupertonic_cuda.
 *
 * Put this directory first on the include path and link libsupertonic_cuda.so instead of libonnxruntime:
 * the UNMODIFIED zhoubin-me/supertonic cpp/helper.cpp + cpp/example_onnx.cpp then compile and run with
 * their four Session::Run calls (cpp/helper.cpp:519-523, 552-556, 643-647, 668-672) executed by the
 * sm_100a kernels through the parity layer of include/supertonic_cuda.h. INTEGRATION.md has the recipe.
 *
 * Only the Ort:: surface the reference touches is provided:
 *   Env, MemoryInfo::CreateCpu, SessionOptions, RunOptions{nullptr}, Session(env, path, opts), Session::Run,
 *   Value::CreateTensor<float|int64_t>, Value::GetTensorMutableData<T>, GetTensorTypeAndShapeInfo().
 *
 * Behaviour kept from ONNX Runtime:
 *   - inputs are matched by NAME, not position (other ports pass them in other orders, go/helper.go:884);
 *   - input tensors are borrowed for the duration of Run; each output Value owns its buffer;
 *   - input shapes are checked against the graph's geometry before Run ("Got invalid dimensions for input: ..."): the C ABI behind
 *     it takes bare pointers;
 *   - failures surface as Ort::Exception (derived from std::exception, what() carries the library message).
 * Graph files: the library takes its layer plan from `stc_arch` metadata or derives it from the graph's nodes; a
 * graph built from other patterns is rejected at Session construction with the unexplained nodes listed. So far it has only ever run
 * the labelled SURROGATE graphs — the released assets were never available to this project.
 * The four Sessions of one onnx directory share one GPU handle (the reference loads all four from the same
 * directory, cpp/helper.cpp:784-795); the handle dies with the last of them. Device: env STC_DEVICE (default 0).
 */
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../supertonic_cuda.h"

enum OrtLoggingLevel { ORT_LOGGING_LEVEL_VERBOSE = 0, ORT_LOGGING_LEVEL_INFO, ORT_LOGGING_LEVEL_WARNING, ORT_LOGGING_LEVEL_ERROR,
                       ORT_LOGGING_LEVEL_FATAL };
enum OrtAllocatorType { OrtInvalidAllocator = -1, OrtDeviceAllocator = 0, OrtArenaAllocator = 1 };
enum OrtMemType { OrtMemTypeCPUInput = -2, OrtMemTypeCPUOutput = -1, OrtMemTypeCPU = OrtMemTypeCPUOutput, OrtMemTypeDefault = 0 };

namespace Ort {

struct Exception : std::runtime_error {
    int code;
    Exception(const std::string& msg, int c) : std::runtime_error(msg), code(c) {}
    int GetOrtErrorCode() const { return code; }
};

struct Env { Env(OrtLoggingLevel = ORT_LOGGING_LEVEL_WARNING, const char* = "") {} };
struct MemoryInfo { static MemoryInfo CreateCpu(OrtAllocatorType, OrtMemType) { return {}; } };
struct SessionOptions {};
struct RunOptions { RunOptions(std::nullptr_t = nullptr) {} };

struct TensorTypeAndShapeInfo {
    size_t count = 0;
    std::vector<int64_t> shape;
    size_t GetElementCount() const { return count; }
    std::vector<int64_t> GetShape() const { return shape; }
    size_t GetDimensionsCount() const { return shape.size(); }
};

class Value {
public:
    Value(std::nullptr_t = nullptr) {}
    Value(Value&&) = default;
    Value& operator=(Value&&) = default;
    Value(const Value&) = delete;
    Value& operator=(const Value&) = delete;

    // Borrowed tensor over caller memory (cpp/helper.cpp:495-509, 1015-1021).
    template <typename T>
    static Value CreateTensor(const MemoryInfo&, T* data, size_t count, const int64_t* shape, size_t rank) {
        static_assert(sizeof(T) == 4 || sizeof(T) == 8, "float / int64_t tensors only");
        Value v;
        v.data_ = data; v.elem_ = sizeof(T);
        v.info_.count = count; v.info_.shape.assign(shape, shape + rank);
        return v;
    }
    template <typename T> T* GetTensorMutableData() { return static_cast<T*>(data_); }
    template <typename T> const T* GetTensorData() const { return static_cast<const T*>(data_); }
    TensorTypeAndShapeInfo GetTensorTypeAndShapeInfo() const { return info_; }
    bool IsTensor() const { return data_ != nullptr; }

    // Owning float tensor (Run outputs).
    static Value Owned(std::vector<int64_t> shape) {
        Value v;
        size_t n = 1;
        for (int64_t d : shape) n *= (size_t)d;
        v.own_ = std::make_unique<float[]>(n ? n : 1);
        v.data_ = v.own_.get(); v.elem_ = 4;
        v.info_.count = n; v.info_.shape = std::move(shape);
        return v;
    }
    size_t elem_size() const { return elem_; }

private:
    void* data_ = nullptr;
    size_t elem_ = 0;
    TensorTypeAndShapeInfo info_;
    std::unique_ptr<float[]> own_;
};

namespace shim_detail {
struct Shared {
    stc_handle* h = nullptr;
    stc_config cfg{};
    ~Shared() { if (h) stc_destroy(h); }
};
inline std::shared_ptr<Shared> acquire(const std::string& dir) {
    static std::mutex mu;
    static std::map<std::string, std::weak_ptr<Shared>> live;
    std::lock_guard<std::mutex> lk(mu);
    if (auto sp = live[dir].lock()) return sp;
    auto sp = std::make_shared<Shared>();
    const char* dev = std::getenv("STC_DEVICE");
    int rc = stc_create(dir.c_str(), dev ? std::atoi(dev) : 0, STC_PREC_DEFAULT, &sp->h);
    if (rc != STC_OK) throw Exception(std::string("libsupertonic_cuda: ") + stc_last_error(nullptr), rc);
    stc_get_config(sp->h, &sp->cfg);
    live[dir] = sp;
    return sp;
}
}  // namespace shim_detail

class Session {
public:
    Session(Env&, const char* model_path, const SessionOptions&) {
        std::string p(model_path);
        size_t slash = p.find_last_of("/\\");
        std::string dir = slash == std::string::npos ? "." : p.substr(0, slash);
        std::string file = slash == std::string::npos ? p : p.substr(slash + 1);
        if (file == "duration_predictor.onnx") kind_ = DP;
        else if (file == "text_encoder.onnx") kind_ = TE;
        else if (file == "vector_estimator.onnx") kind_ = VE;
        else if (file == "vocoder.onnx") kind_ = VOC;
        else throw Exception("libsupertonic_cuda shim: unknown graph file '" + file + "'", STC_ERR_UNSUPPORTED);
        shared_ = shim_detail::acquire(dir);
    }

    std::vector<Value> Run(const RunOptions&, const char* const* in_names, const Value* in, size_t n_in,
                           const char* const* out_names, size_t n_out) {
        if (n_out != 1) throw Exception("shim: exactly one output is supported", STC_ERR_INVALID);
        auto find = [&](const char* name, size_t elem) -> const Value& {
            for (size_t i = 0; i < n_in; ++i)
                if (std::strcmp(in_names[i], name) == 0) {
                    if (in[i].elem_size() != elem) throw Exception(std::string("shim: wrong element type for input ") + name, STC_ERR_INVALID);
                    return in[i];
                }
            throw Exception(std::string("shim: missing input ") + name, STC_ERR_INVALID);
        };
        auto fptr = [](const Value& v) { return v.GetTensorData<float>(); };
        auto dim = [](const Value& v, size_t i) -> int64_t {
            auto s = v.GetTensorTypeAndShapeInfo().shape;
            if (i >= s.size()) throw Exception("shim: input rank too small", STC_ERR_INVALID);
            return s[i];
        };
        auto want_out = [&](const char* name) {
            if (std::strcmp(out_names[0], name) != 0) throw Exception(std::string("shim: unknown output ") + out_names[0], STC_ERR_INVALID);
        };
        stc_handle* h = shared_->h;
        const stc_config& c = shared_->cfg;
        std::vector<Value> out;
        int rc = STC_OK;
        // the C ABI takes bare pointers: every borrowed input is checked against the shapes the graph expects, as ONNX Runtime
        // does before running ("Got invalid dimensions for input")
        auto want_shape = [&](const char* name, const Value& v, std::vector<int64_t> want) {
            const auto s = v.GetTensorTypeAndShapeInfo();
            bool ok = s.shape.size() == want.size();
            size_t n = 1;
            for (size_t i = 0; ok && i < want.size(); ++i) { ok = s.shape[i] == want[i]; n *= (size_t)want[i]; }
            if (!ok || s.count != n) throw Exception(std::string("Got invalid dimensions for input: ") + name, STC_ERR_INVALID);
        };
        auto style_ok = [&](int B, const Value* ttl, const Value* dp) {
            if (ttl) want_shape("style_ttl", *ttl, {B, c.style_ttl_tokens, c.style_ttl_dim});
            if (dp) want_shape("style_dp", *dp, {B, c.style_dp_tokens, c.style_dp_dim});
        };
        switch (kind_) {
            case DP: {
                want_out("duration");
                const Value& ids = find("text_ids", 8);
                int B = (int)dim(ids, 0), T = (int)dim(ids, 1);
                style_ok(B, nullptr, &find("style_dp", 4));
                want_shape("text_mask", find("text_mask", 4), {B, 1, T});
                out.push_back(Value::Owned({B}));
                rc = stc_duration(h, ids.GetTensorData<int64_t>(), fptr(find("style_dp", 4)), fptr(find("text_mask", 4)), B, T,
                                  out[0].GetTensorMutableData<float>());
                break;
            }
            case TE: {
                want_out("text_emb");
                const Value& ids = find("text_ids", 8);
                int B = (int)dim(ids, 0), T = (int)dim(ids, 1);
                style_ok(B, &find("style_ttl", 4), nullptr);
                want_shape("text_mask", find("text_mask", 4), {B, 1, T});
                out.push_back(Value::Owned({B, c.text_emb_channels, T}));
                rc = stc_text_encode(h, ids.GetTensorData<int64_t>(), fptr(find("style_ttl", 4)), fptr(find("text_mask", 4)), B, T,
                                     out[0].GetTensorMutableData<float>(), nullptr);
                break;
            }
            case VE: {
                want_out("denoised_latent");
                const Value& x = find("noisy_latent", 4);
                const Value& te = find("text_emb", 4);
                int B = (int)dim(x, 0), L = (int)dim(x, 2), T = (int)dim(te, 2);
                if (dim(x, 1) != c.latent_channels) throw Exception("shim: noisy_latent channel count != tts.json", STC_ERR_INVALID);
                style_ok(B, &find("style_ttl", 4), nullptr);
                want_shape("text_emb", te, {B, c.text_emb_channels, T});
                want_shape("text_mask", find("text_mask", 4), {B, 1, T});
                want_shape("latent_mask", find("latent_mask", 4), {B, 1, L});
                want_shape("total_step", find("total_step", 4), {B});
                want_shape("current_step", find("current_step", 4), {B});
                out.push_back(Value::Owned({B, c.latent_channels, L}));
                rc = stc_vector_step(h, fptr(x), fptr(te), fptr(find("style_ttl", 4)), fptr(find("text_mask", 4)),
                                     fptr(find("latent_mask", 4)), fptr(find("total_step", 4)), fptr(find("current_step", 4)), B, L, T,
                                     out[0].GetTensorMutableData<float>());
                break;
            }
            case VOC: {
                want_out("wav_tts");
                const Value& z = find("latent", 4);
                int B = (int)dim(z, 0), L = (int)dim(z, 2);
                want_shape("latent", z, {B, c.latent_channels, L});
                out.push_back(Value::Owned({B, (int64_t)L * c.chunk_size}));
                rc = stc_vocode(h, fptr(z), B, L, out[0].GetTensorMutableData<float>());
                break;
            }
        }
        if (rc != STC_OK) throw Exception(std::string("libsupertonic_cuda: ") + stc_last_error(h), rc);
        return out;
    }

private:
    enum Kind { DP, TE, VE, VOC } kind_;
    std::shared_ptr<shim_detail::Shared> shared_;
};

}  // namespace Ort
