// ORACLE BUILD ONLY — drives the unmodified reference host code (cpp/helper.cpp, compiled from
// /root/reference where it lies) over a JSON list of cases and prints JSON answers. Used by
// oracle/make_golden.py to produce tests/golden/host_golden.json.
// (private members of the reference class are reached from THIS driver only: sampleNoisyLatent is private; the class layout and
//  the reference translation unit are untouched)
#include <algorithm>
#include <cstdint>
#include <fstream>
#include <iomanip>
#include <chrono>
#include <iostream>
#include <map>
#include <memory>
#include <random>
#include <regex>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>
#include <onnxruntime_cxx_api.h>
#include <nlohmann/json.hpp>
#define private public
#include "helper.h"
#undef private
#include <fstream>
#include <nlohmann/json.hpp>
using json = nlohmann::json;
std::vector<std::string> chunkText(const std::string& text, int max_len);
std::string sanitizeFilename(const std::string& text, int max_len);

int main(int argc, char** argv) {
    if (argc < 3) { std::cerr << "usage: ref_host <cases.json> <onnx_dir>\n"; return 2; }
    std::ifstream f(argv[1]); json cases; f >> cases;
    std::string onnx_dir = argv[2];
    auto proc = loadTextProcessor(onnx_dir);
    json out = json::array();
    for (auto& c : cases) {
        json r; r["case"] = c;
        std::string kind = c["kind"];
        try {
            if (kind == "text") {
                std::vector<std::vector<int64_t>> ids; std::vector<std::vector<std::vector<float>>> mask;
                proc->call(c["texts"].get<std::vector<std::string>>(), c["langs"].get<std::vector<std::string>>(), ids, mask);
                r["text_ids"] = ids; r["text_mask"] = mask;
            } else if (kind == "chunk") {
                r["chunks"] = chunkText(c["text"].get<std::string>(), c["max_len"].get<int>());
            } else if (kind == "latent_mask") {
                r["mask"] = getLatentMask(c["wav_lengths"].get<std::vector<int64_t>>(), c["base_chunk_size"], c["chunk_compress_factor"]);
            } else if (kind == "length_mask") {
                r["mask"] = lengthToMask(c["lengths"].get<std::vector<int64_t>>(), c.value("max_len", -1));
            } else if (kind == "sanitize") {
                r["name"] = sanitizeFilename(c["text"].get<std::string>(), c["max_len"].get<int>());
            } else if (kind == "wav") {
                std::string path = c["tmp"].get<std::string>();
                writeWavFile(path, c["samples"].get<std::vector<float>>(), c["sample_rate"].get<int>());
                std::ifstream w(path, std::ios::binary); std::vector<unsigned char> b((std::istreambuf_iterator<char>(w)), {});
                r["bytes"] = b;
            } else if (kind == "style") {
                Style s = loadVoiceStyle(c["paths"].get<std::vector<std::string>>());
                r["ttl_shape"] = s.getTtlShape(); r["dp_shape"] = s.getDpShape();
                double a = 0, b = 0; for (float v : s.getTtlData()) a += v; for (float v : s.getDpData()) b += v;
                r["ttl_sum"] = a; r["dp_sum"] = b;
                r["ttl_head"] = std::vector<float>(s.getTtlData().begin(), s.getTtlData().begin() + 4);
            } else if (kind == "noisy_latent") {
                // TextToSpeech::sampleNoisyLatent (cpp/helper.cpp:424-467): latent length / mask from float32 durations; the noise itself
                // is std::random_device-seeded, so only its geometry and masking are reported
                Config cfg = loadCfgs(onnx_dir);
                TextToSpeech tts(cfg, proc.get(), nullptr, nullptr, nullptr, nullptr);
                std::vector<std::vector<std::vector<float>>> lat, mask;
                tts.sampleNoisyLatent(c["duration"].get<std::vector<float>>(), lat, mask);
                r["latent_shape"] = {lat.size(), lat[0].size(), lat[0][0].size()};
                r["mask"] = mask;
                bool masked_zero = true, live_nonzero = true;
                for (size_t b = 0; b < lat.size(); ++b)
                    for (size_t t = 0; t < lat[b][0].size() && t < mask[b][0].size(); ++t) {
                        bool any = false;
                        for (size_t d = 0; d < lat[b].size(); ++d) any = any || lat[b][d][t] != 0.0f;
                        if (mask[b][0][t] == 0.0f && any) masked_zero = false;
                        if (mask[b][0][t] != 0.0f && !any) live_nonzero = false;
                    }
                r["masked_zero"] = masked_zero; r["live_nonzero"] = live_nonzero;
#ifdef STC_FAKE_ORT
            } else if (kind == "call" || kind == "batch") {
                // the reference's orchestration end to end over the closed-form stand-ins of oracle/ref_stub_fake (built as ref_pipe)
                Config cfg = loadCfgs(onnx_dir);
                fake_ort::g_chunk = cfg.ae.base_chunk_size * cfg.ttl.chunk_compress_factor;
                Ort::Env env; Ort::SessionOptions opts;
                Ort::Session dp(env, (onnx_dir + "/duration_predictor.onnx").c_str(), opts), te(env, (onnx_dir + "/text_encoder.onnx").c_str(), opts),
                    ve(env, (onnx_dir + "/vector_estimator.onnx").c_str(), opts), voc(env, (onnx_dir + "/vocoder.onnx").c_str(), opts);
                TextToSpeech tts(cfg, proc.get(), &dp, &te, &ve, &voc);
                Style style = loadVoiceStyle(c["styles"].get<std::vector<std::string>>());
                auto mem = Ort::MemoryInfo::CreateCpu(OrtArenaAllocator, OrtMemTypeDefault);
                fake_ort::g_trace.clear(); fake_ort::g_prev_ve.clear();
                TextToSpeech::SynthesisResult res = kind == "call"
                    ? tts.call(mem, c["text"].get<std::string>(), c["lang"].get<std::string>(), style, c["total_step"].get<int>(),
                               c["speed"].get<float>(), c["silence_duration"].get<float>())
                    : tts.batch(mem, c["texts"].get<std::vector<std::string>>(), c["langs"].get<std::vector<std::string>>(), style,
                                c["total_step"].get<int>(), c["speed"].get<float>());
                r["wav_len"] = res.wav.size();
                double sum = 0; for (float v : res.wav) sum += v;
                r["wav_sum"] = sum;
                std::vector<float> samp; for (size_t i = 0; i < res.wav.size(); i += 1009) samp.push_back(res.wav[i]);
                if (!res.wav.empty()) samp.push_back(res.wav.back());
                r["wav_samples"] = samp; r["duration"] = res.duration;
                json tr = json::array();
                for (auto& k : fake_ort::g_trace) {
                    json e; e["graph"] = k.graph; e["inputs"] = k.inputs; e["shapes"] = k.shapes;
                    if (k.graph == "vector_estimator.onnx") {
                        e["total_step"] = k.total_step; e["current_step"] = k.current_step; e["masked_zero"] = k.masked_zero; e["is_prev_output"] = k.is_prev_output;
                    }
                    tr.push_back(e);
                }
                r["trace"] = tr;
#endif
            } else if (kind == "cfg") {
                Config cfg = loadCfgs(onnx_dir);
                r["cfg"] = {cfg.ae.sample_rate, cfg.ae.base_chunk_size, cfg.ttl.chunk_compress_factor, cfg.ttl.latent_dim};
            }
        } catch (const std::exception& e) { r["error"] = e.what(); }
        out.push_back(r);
    }
    std::cout << out.dump() << std::endl;
    return 0;
}
