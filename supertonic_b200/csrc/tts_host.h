// C++ host side of libsupertonic_cuda: the reference's synthesis API (zhoubin-me/supertonic cpp/helper.h)
// re-implemented over the fast layer of the C ABI (include/supertonic_cuda.h).
//
// Same type / function names, argument meaning and error behaviour as the reference for this path, so a caller
// of cpp/helper.h ports by dropping the Ort::Env / Ort::MemoryInfo arguments:
//   TextToSpeech::call / batch   cpp/helper.cpp:685-734  (here: one stc_synthesize per _infer, no per-step marshalling)
//   Style, loadVoiceStyle        cpp/helper.h:57-72, cpp/helper.cpp:829-897
//   loadTextToSpeech             cpp/helper.cpp:903-937  (use_gpu=true is the ONLY mode; false throws — mirror image
//                                of the reference, which throws on true: there is no CPU path in this library)
//   loadCfgs, writeWavFile, sanitizeFilename, chunkText, timer   :801-818, :943-990, :1070-1111, :1117-1186, helper.h:213-223
// Additions (do not change reference semantics): seedable/injectable noise (the reference RNG is unseedable,
// :442-443), `many()` — independent utterances on packed latent rows — and `callBatched()` — all chunks of one
// long text in a single packed batch.
#pragma once
#include <atomic>
#include <chrono>
#include <cstdint>
#include <iomanip>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "../../include/supertonic_cuda.h"

namespace supertonic {

extern const std::vector<std::string> AVAILABLE_LANGS;

struct Config {
    struct AEConfig { int sample_rate; int base_chunk_size; } ae;
    struct TTLConfig { int chunk_compress_factor; int latent_dim; } ttl;
};

class Style {
public:
    Style(std::vector<float> ttl_data, std::vector<int64_t> ttl_shape, std::vector<float> dp_data, std::vector<int64_t> dp_shape)
        : ttl_data_(std::move(ttl_data)), dp_data_(std::move(dp_data)), ttl_shape_(std::move(ttl_shape)), dp_shape_(std::move(dp_shape)) {}
    const std::vector<float>& getTtlData() const { return ttl_data_; }
    const std::vector<float>& getDpData() const { return dp_data_; }
    const std::vector<int64_t>& getTtlShape() const { return ttl_shape_; }
    const std::vector<int64_t>& getDpShape() const { return dp_shape_; }
    // rows [i0, i0+n) of the batch dimension
    Style slice(const std::vector<int>& rows) const;
private:
    std::vector<float> ttl_data_, dp_data_;
    std::vector<int64_t> ttl_shape_, dp_shape_;
};

class TextToSpeech {
public:
    struct SynthesisResult {
        std::vector<float> wav;        // [B * L*cs] row-major, untrimmed rows (cpp/helper.cpp:679)
        std::vector<float> duration;   // [B] seconds after /speed
    };
    struct Utterance {                 // many(): one entry per input text
        std::vector<float> wav;        // trimmed to int(sr*duration) samples (what cpp/example_onnx.cpp:104-109 keeps)
        float duration;
    };

    TextToSpeech(const Config& cfgs, stc_handle* engine);     // takes ownership of `engine`
    ~TextToSpeech();
    TextToSpeech(const TextToSpeech&) = delete;
    TextToSpeech& operator=(const TextToSpeech&) = delete;

    SynthesisResult call(const std::string& text, const std::string& lang, const Style& style, int total_step,
                         float speed = 1.05f, float silence_duration = 0.3f);
    SynthesisResult batch(const std::vector<std::string>& text_list, const std::vector<std::string>& lang_list, const Style& style,
                          int total_step, float speed = 1.05f);
    // call() with every chunk of the text synthesised in ONE packed batch: the same joined waveform (untrimmed chunk rows +
    // silence), written on the device (stc_out_opts.gap_samples).
    SynthesisResult callBatched(const std::string& text, const std::string& lang, const Style& style, int total_step,
                                float speed = 1.05f, float silence_duration = 0.3f);
    // Independent utterances (style row i belongs to text i), packed latent rows, at most `max_batch` per launch group.
    std::vector<Utterance> many(const std::vector<std::string>& text_list, const std::vector<std::string>& lang_list, const Style& style,
                                int total_step, float speed = 1.05f, int max_batch = 128);

    // many() in two halves, so that MultiGpuTextToSpeech can deal the groups of one plan out over several engines:
    struct ManyPlan {
        std::vector<int64_t> ids; std::vector<float> mask; int64_t T = 0;      // front-end output for the whole request
        std::vector<int> tok;                                                   // token count per text
        std::vector<std::vector<int>> groups;                                   // text indices per launch group (equal predicted latent frames)
    };
    ManyPlan planMany(const std::vector<std::string>& text_list, const std::vector<std::string>& lang_list, int max_batch) const;
    // Synthesises the groups `group_ids` of `plan` on this engine into out[text index]. Noise streams are keyed by the text's index in
    // the request (stc_out_opts.noise_index), so the result does not depend on how the request was grouped or sharded.
    void runGroups(const ManyPlan& plan, const std::vector<int>& group_ids, const Style& style, int total_step, float speed,
                   uint64_t seed, std::vector<Utterance>& out);
    void checkStyle(const Style& style, int bsz) const;

    int getSampleRate() const { return sample_rate_; }
    // Noise: Philox keyed by (seed + call index); or inject N(0,1) as [B][D][ld] for the NEXT _infer only.
    void setNoiseSeed(uint64_t seed) { seed_ = seed; calls_ = 0; }
    void injectNoise(std::vector<float> noise, int64_t ld) { noise_ = std::move(noise); noise_ld_ = ld; }
    stc_handle* engine() const { return engine_; }

private:
    SynthesisResult _infer(const std::vector<std::string>& text_list, const std::vector<std::string>& lang_list, const Style& style,
                           int total_step, float speed);
    void textToIds(const std::vector<std::string>& texts, const std::vector<std::string>& langs, std::vector<int64_t>& ids,
                   std::vector<float>& mask, int64_t& T) const;
    Config cfgs_;
    stc_handle* engine_;
    stc_config geo_{};
    int sample_rate_;
    uint64_t seed_ = 0, calls_ = 0;
    std::atomic<double> frames_per_token_{1.0};      // latent frames per text token measured by the last runGroups (planMany's prediction)
    std::vector<float> noise_;
    int64_t noise_ld_ = 0;
};

// One process, several GPUs of one box: a full weight replica (TextToSpeech engine) + a host thread per device; a request is
// cut into launch groups of equal predicted latent frames once and the groups are dealt out longest-first (no collective, no inter-GPU traffic).
class MultiGpuTextToSpeech {
public:
    MultiGpuTextToSpeech(const std::string& onnx_dir, const std::vector<int>& devices);
    std::vector<TextToSpeech::Utterance> many(const std::vector<std::string>& text_list, const std::vector<std::string>& lang_list,
                                              const Style& style, int total_step, float speed = 1.05f, int max_batch = 128);
    int deviceCount() const { return (int)engines_.size(); }
    TextToSpeech& engine(int i) { return *engines_[i]; }
    void setNoiseSeed(uint64_t seed) { seed_ = seed; calls_ = 0; }
private:
    std::vector<std::unique_ptr<TextToSpeech>> engines_;
    uint64_t seed_ = 0, calls_ = 0;
};

Config loadCfgs(const std::string& onnx_dir);
Style loadVoiceStyle(const std::vector<std::string>& voice_style_paths, bool verbose = false);
std::unique_ptr<TextToSpeech> loadTextToSpeech(const std::string& onnx_dir, bool use_gpu = true, int device = 0);
void writeWavFile(const std::string& filename, const std::vector<float>& audio_data, int sample_rate);
std::vector<unsigned char> wavFileBytes(const std::vector<float>& audio_data, int sample_rate);
std::string sanitizeFilename(const std::string& text, int max_len);
std::vector<std::string> chunkText(const std::string& text, int max_len = 300);

template <typename Func>
auto timer(const std::string& name, Func&& func) -> decltype(func()) {
    auto t0 = std::chrono::high_resolution_clock::now();
    std::cout << name << "..." << std::endl;
    auto result = func();
    std::chrono::duration<double> dt = std::chrono::high_resolution_clock::now() - t0;
    std::cout << "  -> " << name << " completed in " << std::fixed << std::setprecision(2) << dt.count() << " sec" << std::endl;
    return result;
}

}  // namespace supertonic
