// CLI over the C++ host API (tts_host.h) — the counterpart of the reference's cpp/example_onnx.cpp with the same
// flags (--onnx-dir --total-step --speed --n-test --voice-style a,b --text "x|y" --lang en,ko --save-dir --batch) and the
// same output files (<sanitized text>_<n>.wav trimmed to int(sr*duration) samples, cpp/example_onnx.cpp:99-114), plus
// --device N, --seed S (deterministic noise), --many (packed throughput path) and --batched-chunks (long-form text with
// all chunks in one batch).
#include <filesystem>
#include <iostream>
#include <sstream>

#include "tts_host.h"

namespace fs = std::filesystem;
using namespace supertonic;

static std::vector<std::string> split(const std::string& s, char delim) {
    std::vector<std::string> parts;
    std::stringstream ss(s);
    for (std::string item; std::getline(ss, item, delim);) parts.push_back(item);
    if (s.empty() || s.back() == delim) parts.emplace_back();
    return parts;
}

int main(int argc, char** argv) {
    std::string onnx_dir = "assets/onnx", save_dir = "results";
    std::vector<std::string> voices = {"assets/voice_styles/M1.json"}, langs = {"en"};
    std::vector<std::string> texts = {"This morning, I took a walk in the park, and the sound of the birds and the breeze was so "
                                      "pleasant that I stopped for a long time just to listen."};
    int total_step = 5, n_test = 4, device = 0;
    float speed = 1.05f;
    uint64_t seed = 0;
    bool batch = false, many = false, batched_chunks = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) { std::cerr << "missing value for " << a << "\n"; exit(2); } return argv[++i]; };
        if (a == "--onnx-dir") onnx_dir = next();
        else if (a == "--total-step") total_step = std::stoi(next());
        else if (a == "--speed") speed = std::stof(next());
        else if (a == "--n-test") n_test = std::stoi(next());
        else if (a == "--voice-style") voices = split(next(), ',');
        else if (a == "--text") texts = split(next(), '|');
        else if (a == "--lang") langs = split(next(), ',');
        else if (a == "--save-dir") save_dir = next();
        else if (a == "--device") device = std::stoi(next());
        else if (a == "--seed") seed = std::stoull(next());
        else if (a == "--batch") batch = true;
        else if (a == "--many") many = true;
        else if (a == "--batched-chunks") batched_chunks = true;
        else { std::cerr << "unknown flag " << a << "\n"; return 2; }
    }
    std::cout << "=== TTS Inference with libsupertonic_cuda (C++) ===\n\n";
    if (voices.size() != texts.size()) {
        std::cerr << "Error: Number of voice styles (" << voices.size() << ") must match number of texts (" << texts.size() << ")\n";
        return 1;
    }
    const int bsz = (int)voices.size();
    try {
        auto tts = loadTextToSpeech(onnx_dir, true, device);
        tts->setNoiseSeed(seed);
        std::cout << std::endl;
        Style style = loadVoiceStyle(voices, true);
        fs::create_directories(save_dir);
        for (int n = 0; n < n_test; ++n) {
            std::cout << "\n[" << (n + 1) << "/" << n_test << "] Starting synthesis...\n";
            std::vector<std::vector<float>> wavs(bsz);
            timer("Generating speech from text", [&]() {
                if (many) {
                    auto us = tts->many(texts, langs, style, total_step, speed);
                    for (int b = 0; b < bsz; ++b) wavs[b] = std::move(us[b].wav);
                    return 0;
                }
                TextToSpeech::SynthesisResult r = batch ? tts->batch(texts, langs, style, total_step, speed)
                                                  : batched_chunks ? tts->callBatched(texts[0], langs[0], style, total_step, speed)
                                                                   : tts->call(texts[0], langs[0], style, total_step, speed);
                const size_t row = r.wav.size() / bsz;
                for (int b = 0; b < bsz; ++b) {
                    size_t keep = std::min(row, (size_t)static_cast<int>(tts->getSampleRate() * r.duration[b]));
                    wavs[b].assign(r.wav.begin() + b * row, r.wav.begin() + b * row + keep);
                }
                return 0;
            });
            for (int b = 0; b < bsz; ++b) {
                std::string path = save_dir + "/" + sanitizeFilename(texts[b], 20) + "_" + std::to_string(n + 1) + ".wav";
                writeWavFile(path, wavs[b], tts->getSampleRate());
                std::cout << "Saved: " << path << "\n";
            }
        }
    } catch (const std::exception& e) {
        std::cerr << "Error: " << e.what() << "\n";
        return 1;
    }
    std::cout << "\n=== Synthesis completed successfully! ===\n";
    return 0;
}
