// ORACLE BUILD ONLY — a FAKE <onnxruntime_cxx_api.h>: Session::Run computes small closed-form stand-ins for the four graphs and
// records every call. It lets the UNMODIFIED reference orchestration — TextToSpeech::_infer / call / batch (cpp/helper.cpp:469-734),
// compiled where it lies — run end to end on the CPU, so that oracle/host_ref.ReferenceTTS (the restatement the GPU parity tests use
// as their checker) is pinned to the reference's own control flow: which tensors each Run receives, in which order and shape, the
// speed division, the total_step loop and its scalar tensors, chunking, the silence join and the duration bookkeeping.
// The stand-ins (tests/test_oracle_pipeline.py holds the same formulas in numpy; every constant is a power of two, so float32
// arithmetic is exact on both sides):
//   duration[b]          = tokens_b / 16 + (text_ids[b][0] % 5) / 32
//   text_emb[b][c][t]    = text_mask[b][0][t] * (c + 1) / 32                         (4 channels)
//   denoised[b][d][t]    = latent_mask[b][0][t] * (1/2 + current_step[b] / 16 + total_step[b] / 128 + (d % 16) / 1024)
//   wav_tts[b][i]        = latent[b][i % D][i / chunk] / 2 + ((i % 97) - 48) / 128      (chunk = samples per latent frame, D = channels)
// The vector-estimator stand-in ignores the VALUES of noisy_latent (the reference seeds its noise from std::random_device) but records
// whether they are zero under the mask and, from the second step on, whether they are the previous step's output.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
enum OrtLoggingLevel { ORT_LOGGING_LEVEL_WARNING = 2 };
enum OrtAllocatorType { OrtArenaAllocator = 1 };
enum OrtMemType { OrtMemTypeDefault = 0 };
namespace fake_ort {
struct Call {
    std::string graph;
    std::vector<std::string> inputs;                 // names in the order passed
    std::vector<std::vector<int64_t>> shapes;
    std::vector<float> total_step, current_step;     // vector estimator only
    bool masked_zero = true, is_prev_output = true;  // vector estimator only
};
inline std::vector<Call> g_trace;
inline std::vector<float> g_prev_ve;
inline int g_chunk = 3072;                           // samples per latent frame (base_chunk_size * chunk_compress_factor), set by the driver
}  // namespace fake_ort
namespace Ort {
struct Env { Env(OrtLoggingLevel = ORT_LOGGING_LEVEL_WARNING, const char* = "") {} };
struct MemoryInfo { static MemoryInfo CreateCpu(OrtAllocatorType, OrtMemType) { return {}; } };
struct SessionOptions {};
struct RunOptions { RunOptions(std::nullptr_t = nullptr) {} };
struct TensorTypeAndShapeInfo {
    size_t n = 0; std::vector<int64_t> shape;
    size_t GetElementCount() const { return n; }
    std::vector<int64_t> GetShape() const { return shape; }
};
struct Value {
    void* p = nullptr; TensorTypeAndShapeInfo info; std::shared_ptr<std::vector<float>> own;
    Value(std::nullptr_t = nullptr) {}
    template <typename T>
    static Value CreateTensor(const MemoryInfo&, T* data, size_t count, const int64_t* shape, size_t rank) {
        Value v; v.p = data; v.info.n = count; v.info.shape.assign(shape, shape + rank); return v;
    }
    static Value Owned(std::vector<int64_t> shape) {
        Value v; size_t n = 1; for (int64_t d : shape) n *= (size_t)d;
        v.own = std::make_shared<std::vector<float>>(n, 0.0f); v.p = v.own->data(); v.info.n = n; v.info.shape = std::move(shape); return v;
    }
    template <typename T> T* GetTensorMutableData() { return static_cast<T*>(p); }
    TensorTypeAndShapeInfo GetTensorTypeAndShapeInfo() const { return info; }
};
struct Session {
    std::string graph;
    Session(Env&, const char* path, const SessionOptions&) {
        std::string p(path); size_t s = p.find_last_of("/\\"); graph = s == std::string::npos ? p : p.substr(s + 1);
    }
    std::vector<Value> Run(const RunOptions&, const char* const* in_names, const Value* in, size_t n_in, const char* const* out_names, size_t n_out) {
        using namespace fake_ort;
        if (n_out != 1) throw std::runtime_error("fake ORT: one output expected");
        Call c; c.graph = graph;
        auto find = [&](const char* name) -> const Value& {
            for (size_t i = 0; i < n_in; ++i) if (!std::strcmp(in_names[i], name)) return in[i];
            throw std::runtime_error(std::string("fake ORT: missing input ") + name);
        };
        for (size_t i = 0; i < n_in; ++i) { c.inputs.push_back(in_names[i]); c.shapes.push_back(in[i].info.shape); }
        std::vector<Value> out;
        if (graph == "duration_predictor.onnx") {
            if (std::strcmp(out_names[0], "duration")) throw std::runtime_error("fake ORT: output name");
            const Value& ids = find("text_ids"); const Value& mask = find("text_mask"); find("style_dp");
            const int64_t B = ids.info.shape[0], T = ids.info.shape[1];
            out.push_back(Value::Owned({B}));
            for (int64_t b = 0; b < B; ++b) {
                float n = 0.f; for (int64_t t = 0; t < T; ++t) n += static_cast<const float*>(mask.p)[b * T + t];
                out[0].GetTensorMutableData<float>()[b] = n * 0.0625f + (float)(static_cast<const int64_t*>(ids.p)[b * T] % 5) * 0.03125f;
            }
        } else if (graph == "text_encoder.onnx") {
            if (std::strcmp(out_names[0], "text_emb")) throw std::runtime_error("fake ORT: output name");
            const Value& ids = find("text_ids"); const Value& mask = find("text_mask"); find("style_ttl");
            const int64_t B = ids.info.shape[0], T = ids.info.shape[1], C = 4;
            out.push_back(Value::Owned({B, C, T}));
            for (int64_t b = 0; b < B; ++b) for (int64_t ch = 0; ch < C; ++ch) for (int64_t t = 0; t < T; ++t)
                out[0].GetTensorMutableData<float>()[(b * C + ch) * T + t] = static_cast<const float*>(mask.p)[b * T + t] * (float)(ch + 1) * 0.03125f;
        } else if (graph == "vector_estimator.onnx") {
            if (std::strcmp(out_names[0], "denoised_latent")) throw std::runtime_error("fake ORT: output name");
            const Value& x = find("noisy_latent"); const Value& lm = find("latent_mask"); find("text_emb"); find("style_ttl"); find("text_mask");
            const Value& tot = find("total_step"); const Value& cur = find("current_step");
            const int64_t B = x.info.shape[0], D = x.info.shape[1], L = x.info.shape[2];
            const float* xp = static_cast<const float*>(x.p); const float* mp = static_cast<const float*>(lm.p);
            c.total_step.assign(static_cast<const float*>(tot.p), static_cast<const float*>(tot.p) + tot.info.n);
            c.current_step.assign(static_cast<const float*>(cur.p), static_cast<const float*>(cur.p) + cur.info.n);
            for (int64_t b = 0; b < B; ++b) for (int64_t d = 0; d < D; ++d) for (int64_t t = 0; t < L; ++t)
                if (mp[b * L + t] == 0.f && xp[(b * D + d) * L + t] != 0.f) c.masked_zero = false;
            c.is_prev_output = c.current_step[0] == 0.f ? true : (g_prev_ve.size() == (size_t)(B * D * L) && !std::memcmp(g_prev_ve.data(), xp, sizeof(float) * B * D * L));
            out.push_back(Value::Owned({B, D, L}));
            float* o = out[0].GetTensorMutableData<float>();
            for (int64_t b = 0; b < B; ++b) for (int64_t d = 0; d < D; ++d) for (int64_t t = 0; t < L; ++t)
                o[(b * D + d) * L + t] = mp[b * L + t] * (0.5f + c.current_step[b] * 0.0625f + c.total_step[b] * 0.0078125f + (float)(d % 16) * 0.0009765625f);
            g_prev_ve.assign(o, o + B * D * L);
        } else if (graph == "vocoder.onnx") {
            if (std::strcmp(out_names[0], "wav_tts")) throw std::runtime_error("fake ORT: output name");
            const Value& x = find("latent");
            const int64_t B = x.info.shape[0], D = x.info.shape[1], L = x.info.shape[2], N = L * g_chunk;
            const float* xp = static_cast<const float*>(x.p);
            out.push_back(Value::Owned({B, N}));
            float* o = out[0].GetTensorMutableData<float>();
            for (int64_t b = 0; b < B; ++b) for (int64_t i = 0; i < N; ++i)
                o[b * N + i] = xp[(b * D + i % D) * L + i / g_chunk] * 0.5f + (float)((int)(i % 97) - 48) * 0.0078125f;
        } else throw std::runtime_error("fake ORT: unknown graph " + graph);
        g_trace.push_back(std::move(c));
        return out;
    }
};
}  // namespace Ort
