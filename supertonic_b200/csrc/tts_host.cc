// See tts_host.h. Everything neural goes through the C ABI; this file is host bookkeeping only.
#include "tts_host.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <nlohmann/json.hpp>
#include <numeric>
#include <stdexcept>
#include <thread>

using json = nlohmann::json;

namespace supertonic {

const std::vector<std::string> AVAILABLE_LANGS = {"en", "ko", "es", "pt", "fr"};    // cpp/helper.cpp:15

namespace {
[[noreturn]] void raise(stc_handle* h, const char* what) {
    throw std::runtime_error(std::string(what) + ": " + stc_last_error(h));
}
std::vector<const char*> cptrs(const std::vector<std::string>& v) {
    std::vector<const char*> p(v.size());
    for (size_t i = 0; i < v.size(); ++i) p[i] = v[i].c_str();
    return p;
}
}  // namespace

Style Style::slice(const std::vector<int>& rows) const {
    size_t nt = (size_t)(ttl_shape_[1] * ttl_shape_[2]), nd = (size_t)(dp_shape_[1] * dp_shape_[2]);
    std::vector<float> t(rows.size() * nt), d(rows.size() * nd);
    for (size_t i = 0; i < rows.size(); ++i) {
        std::copy_n(ttl_data_.begin() + (size_t)rows[i] * nt, nt, t.begin() + i * nt);
        std::copy_n(dp_data_.begin() + (size_t)rows[i] * nd, nd, d.begin() + i * nd);
    }
    return Style(std::move(t), {(int64_t)rows.size(), ttl_shape_[1], ttl_shape_[2]}, std::move(d),
                 {(int64_t)rows.size(), dp_shape_[1], dp_shape_[2]});
}

// ---------------------------------------------------------------------------------------------- loading
Config loadCfgs(const std::string& onnx_dir) {
    std::string path = onnx_dir + "/tts.json";
    std::ifstream f(path);
    if (!f.is_open()) throw std::runtime_error("Failed to open config file: " + path);
    json j; f >> j;
    Config c;
    c.ae.sample_rate = j["ae"]["sample_rate"]; c.ae.base_chunk_size = j["ae"]["base_chunk_size"];
    c.ttl.chunk_compress_factor = j["ttl"]["chunk_compress_factor"]; c.ttl.latent_dim = j["ttl"]["latent_dim"];
    return c;
}

static void flatten3(const json& node, std::vector<float>& out) {
    for (const auto& plane : node) for (const auto& row : plane) for (const auto& v : row) out.push_back(v.get<float>());
}

Style loadVoiceStyle(const std::vector<std::string>& paths, bool verbose) {
    if (paths.empty()) throw std::runtime_error("Failed to open voice style file: (none given)");
    int64_t bsz = (int64_t)paths.size(), t1 = 0, t2 = 0, d1 = 0, d2 = 0;
    std::vector<float> ttl, dp;
    for (int64_t i = 0; i < bsz; ++i) {
        std::ifstream f(paths[i]);
        if (!f.is_open()) throw std::runtime_error("Failed to open voice style file: " + paths[i]);
        json j; f >> j;
        if (i == 0) {       // geometry comes from the first file only, like the reference (:839-846)
            auto td = j["style_ttl"]["dims"].get<std::vector<int64_t>>(), dd = j["style_dp"]["dims"].get<std::vector<int64_t>>();
            t1 = td.at(1); t2 = td.at(2); d1 = dd.at(1); d2 = dd.at(2);
            ttl.reserve((size_t)(bsz * t1 * t2)); dp.reserve((size_t)(bsz * d1 * d2));
        }
        size_t a = ttl.size(), b = dp.size();
        flatten3(j["style_ttl"]["data"], ttl);
        flatten3(j["style_dp"]["data"], dp);
        ttl.resize(a + (size_t)(t1 * t2)); dp.resize(b + (size_t)(d1 * d2));     // fixed slot per file, zero-filled if short
    }
    if (verbose) std::cout << "Loaded " << bsz << " voice styles" << std::endl;
    return Style(std::move(ttl), {bsz, t1, t2}, std::move(dp), {bsz, d1, d2});
}

std::unique_ptr<TextToSpeech> loadTextToSpeech(const std::string& onnx_dir, bool use_gpu, int device) {
    if (!use_gpu) throw std::runtime_error("CPU mode is not supported by libsupertonic_cuda (use the reference's ONNX Runtime path)");
    std::cout << "Using GPU " << device << " (libsupertonic_cuda, sm_100a) for inference" << std::endl;
    Config cfg = loadCfgs(onnx_dir);
    stc_handle* h = nullptr;
    if (stc_create(onnx_dir.c_str(), device, STC_PREC_DEFAULT, &h) != STC_OK) raise(nullptr, "loadTextToSpeech");
    return std::make_unique<TextToSpeech>(cfg, h);
}

// ---------------------------------------------------------------------------------------------- synthesis
TextToSpeech::TextToSpeech(const Config& cfgs, stc_handle* engine) : cfgs_(cfgs), engine_(engine), sample_rate_(cfgs.ae.sample_rate) {
    if (!engine_) throw std::runtime_error("TextToSpeech: null engine");
    stc_get_config(engine_, &geo_);
    if (geo_.sample_rate != cfgs.ae.sample_rate || geo_.latent_dim != cfgs.ttl.latent_dim)
        throw std::runtime_error("TextToSpeech: Config does not match the engine's tts.json");
}
TextToSpeech::~TextToSpeech() { stc_destroy(engine_); }

// The C ABI takes bare pointers: a voice-style file with other dims than the graphs' (or a style batch that does not match
// the texts) must be rejected here, where ONNX Runtime would raise a shape error (cpp/helper.cpp:519, 552, 643)
void TextToSpeech::checkStyle(const Style& style, int bsz) const {
    const auto& t = style.getTtlShape(); const auto& d = style.getDpShape();
    if (t.size() != 3 || d.size() != 3) throw std::runtime_error("Got invalid dimensions for input: style tensors must have rank 3");
    if ((int64_t)style.getTtlData().size() != t[0] * t[1] * t[2] || (int64_t)style.getDpData().size() != d[0] * d[1] * d[2])
        throw std::runtime_error("Got invalid dimensions for input: style data does not match its dims");
    if (stc_validate_style(engine_, bsz, t.data(), d.data()) != STC_OK) raise(engine_, "voice style");
}

void TextToSpeech::textToIds(const std::vector<std::string>& texts, const std::vector<std::string>& langs, std::vector<int64_t>& ids,
                             std::vector<float>& mask, int64_t& T) const {
    if (texts.size() != langs.size()) throw std::runtime_error("Number of texts must match number of languages");
    auto tp = cptrs(texts), lp = cptrs(langs);
    int n = (int)texts.size();
    if (stc_text_to_ids(engine_, tp.data(), lp.data(), n, nullptr, nullptr, 0, &T) != STC_OK) raise(engine_, "text front-end");
    ids.assign((size_t)n * T, 0); mask.assign((size_t)n * T, 0.f);
    if (stc_text_to_ids(engine_, tp.data(), lp.data(), n, ids.data(), mask.data(), T, &T) != STC_OK) raise(engine_, "text front-end");
}

TextToSpeech::SynthesisResult TextToSpeech::_infer(const std::vector<std::string>& text_list, const std::vector<std::string>& lang_list,
                                                   const Style& style, int total_step, float speed) {
    int bsz = (int)text_list.size();
    if (bsz != style.getTtlShape()[0]) throw std::runtime_error("Number of texts must match number of style vectors");
    checkStyle(style, bsz);
    std::vector<int64_t> ids; std::vector<float> mask; int64_t T = 0;
    textToIds(text_list, lang_list, ids, mask, T);
    const int cs = geo_.chunk_size;
    // the latent length is data dependent (duration -> L): start from a per-token guess, retry once with the exact size
    int64_t ld = std::max<int64_t>((int64_t)((double)T * 0.12 * sample_rate_ / cs) + 8, 16) * cs, L = 0;
    SynthesisResult r;
    r.duration.assign(bsz, 0.f);
    std::vector<float> noise; int64_t nld = 0;
    noise.swap(noise_); std::swap(nld, noise_ld_);                     // injected noise is consumed by this call
    ++calls_;
    std::vector<float> wav;
    for (int attempt = 0;; ++attempt) {
        wav.resize((size_t)bsz * ld);
        int rc = stc_synthesize(engine_, ids.data(), mask.data(), style.getTtlData().data(), style.getDpData().data(), bsz, (int)T,
                                total_step, speed, noise.empty() ? nullptr : noise.data(), nld, seed_ + calls_, wav.data(), ld,
                                r.duration.data(), nullptr, &L, nullptr);
        if (rc == STC_ERR_CAPACITY && attempt == 0 && L * cs > ld) { ld = L * cs; continue; }
        if (rc != STC_OK) raise(engine_, "stc_synthesize");
        break;
    }
    const int64_t row = L * cs;                                         // result.wav is dense [B][L*cs] (cpp/helper.cpp:674-679)
    if (row == ld) r.wav.swap(wav);
    else {
        r.wav.resize((size_t)bsz * row);
        for (int b = 0; b < bsz; ++b) std::copy_n(wav.begin() + (size_t)b * ld, row, r.wav.begin() + (size_t)b * row);
    }
    return r;
}

TextToSpeech::SynthesisResult TextToSpeech::call(const std::string& text, const std::string& lang, const Style& style, int total_step,
                                                 float speed, float silence_duration) {
    if (style.getTtlShape()[0] != 1) throw std::runtime_error("Single speaker text to speech only supports single style");
    SynthesisResult out;
    float total = 0.f;
    bool first = true;
    for (const std::string& chunk : chunkText(text, lang == "ko" ? 120 : 300)) {
        SynthesisResult part = _infer({chunk}, {lang}, style, total_step, speed);
        if (first) { out.wav = std::move(part.wav); total = part.duration[0]; first = false; continue; }
        out.wav.resize(out.wav.size() + (size_t)static_cast<int>(silence_duration * sample_rate_), 0.f);     // :710-712
        out.wav.insert(out.wav.end(), part.wav.begin(), part.wav.end());
        total += part.duration[0] + silence_duration;
    }
    out.duration = {total};
    return out;
}

TextToSpeech::SynthesisResult TextToSpeech::batch(const std::vector<std::string>& text_list, const std::vector<std::string>& lang_list,
                                                  const Style& style, int total_step, float speed) {
    return _infer(text_list, lang_list, style, total_step, speed);
}

// many() = plan (front-end once, launch groups) + run (packed launches of the groups on ONE engine).
// MultiGpuTextToSpeech runs the same plan with the groups dealt out over several engines.
// Groups: as few as `max_batch` utterances and 140 row tiles of PREDICTED latent frames per group allow, filled to equal predicted
// frames (longest text first into the emptiest group). Packed rows carry no padding, so a group need not hold similar lengths;
// what costs is a group whose 128-row tiles do not fill the 148 SMs, or one that spills into a second wave of the fused MLP
// (model.cu mlp_plan). The frames-per-token ratio is what the last runGroups of this object measured (1.0 until then).
TextToSpeech::ManyPlan TextToSpeech::planMany(const std::vector<std::string>& text_list, const std::vector<std::string>& lang_list,
                                              int max_batch) const {
    ManyPlan p;
    const int n = (int)text_list.size();
    textToIds(text_list, lang_list, p.ids, p.mask, p.T);
    p.tok.resize(n);
    for (int i = 0; i < n; ++i) p.tok[i] = (int)std::accumulate(p.mask.begin() + (size_t)i * p.T, p.mask.begin() + (size_t)(i + 1) * p.T, 0.f);
    if (n == 0) return p;
    max_batch = std::max(1, max_batch);
    // the groups are balanced on the integer token counts (predicted frames are proportional to them), so the same request always gives
    // the same groups — and hits the same CUDA graphs — however the measured ratio jitters; the ratio, rounded up to a multiple of
    // 1/16, only decides how many groups there are
    const double fpt = std::ceil(std::max(frames_per_token_.load(), 1e-3) * 16.0) / 16.0, max_rows = 140.0 * 128.0;
    int64_t total = 0;
    for (int i = 0; i < n; ++i) total += std::max(1, p.tok[i]);
    const int ng = std::min(n, std::max((n + max_batch - 1) / max_batch, (int)std::ceil((double)total * fpt / max_rows)));
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return p.tok[a] > p.tok[b]; });
    std::vector<int64_t> load(ng, 0);
    p.groups.assign(ng, {});
    for (int i : order) {
        int g = -1;
        for (int k = 0; k < ng; ++k)
            if ((int)p.groups[k].size() < max_batch && (g < 0 || load[k] < load[g])) g = k;
        p.groups[g].push_back(i);
        load[g] += std::max(1, p.tok[i]);
    }
    for (auto& g : p.groups) std::sort(g.begin(), g.end());
    p.groups.erase(std::remove_if(p.groups.begin(), p.groups.end(), [](const std::vector<int>& g) { return g.empty(); }), p.groups.end());
    return p;
}

void TextToSpeech::runGroups(const ManyPlan& plan, const std::vector<int>& group_ids, const Style& style, int total_step, float speed,
                             uint64_t seed, std::vector<Utterance>& out) {
    const int cs = geo_.chunk_size;
    const int64_t T = plan.T;
    // Groups go out as an asynchronous request stream (stc_synthesize_packed_ex, async): the device->host copy of group g runs
    // while group g+1 is computed. Each group gets its own page-locked result buffer; everything lands at stc_wait.
    struct Pending { const std::vector<int>* grp; float* wav = nullptr; std::vector<int64_t> off, wl, nidx; std::vector<float> dur; };
    std::vector<Pending> pend;
    auto free_all = [&]() { for (auto& p : pend) stc_pinned_free(p.wav); };
    try {
        for (int gi : group_ids) {
            Pending p;
            p.grp = &plan.groups[gi];
            const std::vector<int>& grp = *p.grp;
            const int B = (int)grp.size();
            int Tg = 1;
            for (int i : grp) Tg = std::max(Tg, plan.tok[i]);
            std::vector<int64_t> gids((size_t)B * Tg); std::vector<float> gm((size_t)B * Tg);
            int64_t toks = 0;
            for (int k = 0; k < B; ++k) {
                std::copy_n(plan.ids.begin() + (size_t)grp[k] * T, Tg, gids.begin() + (size_t)k * Tg);
                std::copy_n(plan.mask.begin() + (size_t)grp[k] * T, Tg, gm.begin() + (size_t)k * Tg);
                toks += plan.tok[grp[k]];
            }
            Style st = style.slice(grp);
            int64_t cap = (int64_t)((double)toks * 0.12 * sample_rate_) + (int64_t)(B + 8) * cs;
            p.off.assign(B + 1, 0); p.wl.assign(B, 0); p.dur.assign(B, 0.f);
            p.nidx.assign(grp.begin(), grp.end());          // noise stream = index in the REQUEST: the result does not depend on grouping / sharding
            stc_out_opts oo{}; oo.noise_index = p.nidx.data();
            for (int attempt = 0;; ++attempt) {
                void* mem = nullptr;
                if (stc_pinned_alloc((size_t)cap * sizeof(float), &mem) != STC_OK) raise(nullptr, "stc_pinned_alloc");
                p.wav = static_cast<float*>(mem);
                int rc = stc_synthesize_packed_ex(engine_, gids.data(), gm.data(), st.getTtlData().data(), st.getDpData().data(), B, Tg, total_step,
                                                  speed, nullptr, 0, seed, &oo, p.wav, cap, p.off.data(), p.dur.data(), p.wl.data(), 1);
                if (rc == STC_ERR_CAPACITY && attempt == 0 && p.off[B] > cap) { stc_pinned_free(p.wav); p.wav = nullptr; cap = p.off[B]; continue; }
                if (rc != STC_OK) { stc_pinned_free(p.wav); p.wav = nullptr; raise(engine_, "stc_synthesize_packed_ex"); }
                break;
            }
            pend.push_back(std::move(p));          // (the library keeps at most two calls in flight)
        }
        if (stc_wait(engine_) != STC_OK) raise(engine_, "stc_wait");
    } catch (...) { stc_wait(engine_); free_all(); throw; }
    int64_t frames = 0, tokens = 0;
    for (auto& p : pend)
        for (size_t k = 0; k < p.grp->size(); ++k) {
            Utterance& u = out[(*p.grp)[k]];
            u.duration = p.dur[k];
            u.wav.assign(p.wav + p.off[k], p.wav + p.off[k] + p.wl[k]);
            frames += (p.off[k + 1] - p.off[k]) / cs;
            tokens += plan.tok[(*p.grp)[k]];
        }
    if (tokens > 0) frames_per_token_.store((double)frames / (double)tokens);       // the next plan's prediction
    free_all();
}

std::vector<TextToSpeech::Utterance> TextToSpeech::many(const std::vector<std::string>& text_list, const std::vector<std::string>& lang_list,
                                                        const Style& style, int total_step, float speed, int max_batch) {
    const int n = (int)text_list.size();
    if (n != style.getTtlShape()[0]) throw std::runtime_error("Number of texts must match number of style vectors");
    checkStyle(style, n);
    ManyPlan plan = planMany(text_list, lang_list, max_batch);
    std::vector<int> all(plan.groups.size());
    std::iota(all.begin(), all.end(), 0);
    std::vector<Utterance> out(n);
    ++calls_;
    runGroups(plan, all, style, total_step, speed, seed_ + calls_, out);
    return out;
}

// ---------------------------------------------------------------------------------------------- several GPUs of one box
// north_star: "a request batch is length-bucketed and partitioned across the 8 B200s of one box. Each GPU holds a full weight replica
// and no collective runs on the hot path". The reference entry point that fans out is TextToSpeech::batch (cpp/helper.cpp:725-734).
MultiGpuTextToSpeech::MultiGpuTextToSpeech(const std::string& onnx_dir, const std::vector<int>& devices) {
    if (devices.empty()) throw std::runtime_error("MultiGpuTextToSpeech: no devices given");
    for (int d : devices) engines_.push_back(loadTextToSpeech(onnx_dir, true, d));
}

std::vector<TextToSpeech::Utterance> MultiGpuTextToSpeech::many(const std::vector<std::string>& text_list,
                                                                const std::vector<std::string>& lang_list, const Style& style,
                                                                int total_step, float speed, int max_batch) {
    const int n = (int)text_list.size(), nd = (int)engines_.size();
    if (n != style.getTtlShape()[0]) throw std::runtime_error("Number of texts must match number of style vectors");
    engines_[0]->checkStyle(style, n);
    // at least two launch groups per device: the device->host copy of one group then runs under the computation of the next
    if (nd > 1) max_batch = std::min(max_batch, std::max(16, (n + 2 * nd - 1) / (2 * nd)));
    TextToSpeech::ManyPlan plan = engines_[0]->planMany(text_list, lang_list, max_batch);
    // longest-processing-time-first over the groups; cost ~ tokens (frames are proportional) x (Euler steps + vocoder share)
    std::vector<double> cost(plan.groups.size(), 0.0);
    for (size_t g = 0; g < plan.groups.size(); ++g)
        for (int i : plan.groups[g]) cost[g] += (double)plan.tok[i] * (total_step * 1.0 + 6 * 4.0 * 0.35);
    std::vector<int> order(plan.groups.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
    std::vector<std::vector<int>> shard(nd);
    std::vector<double> load(nd, 0.0);
    for (int g : order) {
        const int r = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        shard[r].push_back(g); load[r] += cost[g];
    }
    std::vector<TextToSpeech::Utterance> out(n);
    ++calls_;
    std::vector<std::thread> th;
    std::vector<std::exception_ptr> err(nd);
    for (int r = 0; r < nd; ++r)
        th.emplace_back([&, r]() {
            try { if (!shard[r].empty()) engines_[r]->runGroups(plan, shard[r], style, total_step, speed, seed_ + calls_, out); }
            catch (...) { err[r] = std::current_exception(); }
        });
    for (auto& t : th) t.join();
    for (auto& e : err) if (e) std::rethrow_exception(e);
    return out;
}

TextToSpeech::SynthesisResult TextToSpeech::callBatched(const std::string& text, const std::string& lang, const Style& style,
                                                        int total_step, float speed, float silence_duration) {
    // `call()` with the chunks as ONE packed batch: the untrimmed chunk waveforms joined with (int)(silence * sr) zeros exactly as
    // cpp/helper.cpp:703-716 does on the host — here the library writes the joined waveform on the device (stc_out_opts.gap_samples)
    if (style.getTtlShape()[0] != 1) throw std::runtime_error("Single speaker text to speech only supports single style");
    std::vector<std::string> chunks = chunkText(text, lang == "ko" ? 120 : 300);
    const int n = (int)chunks.size();
    std::vector<std::string> langs(n, lang);
    Style st = style.slice(std::vector<int>(n, 0));
    checkStyle(st, n);
    std::vector<int64_t> ids; std::vector<float> mask; int64_t T = 0;
    textToIds(chunks, langs, ids, mask, T);
    const int cs = geo_.chunk_size;
    stc_out_opts oo{};
    oo.pcm16 = 0; oo.gap_samples = (int64_t)static_cast<int>(silence_duration * sample_rate_);
    double toks = 0;
    for (float m : mask) toks += m;
    int64_t cap = (int64_t)(toks * 0.12 * sample_rate_) + (int64_t)(n + 8) * cs + (int64_t)(n - 1) * oo.gap_samples;
    std::vector<int64_t> off(n + 1, 0), wl(n, 0);
    std::vector<float> dur(n, 0.f);
    SynthesisResult out;
    ++calls_;
    for (int attempt = 0;; ++attempt) {
        out.wav.resize((size_t)cap);
        int rc = stc_synthesize_packed_ex(engine_, ids.data(), mask.data(), st.getTtlData().data(), st.getDpData().data(), n, (int)T, total_step,
                                          speed, nullptr, 0, seed_ + calls_, &oo, out.wav.data(), cap, off.data(), dur.data(), wl.data(), 0);
        if (rc == STC_ERR_CAPACITY && attempt == 0 && off[n] > cap) { cap = off[n]; continue; }
        if (rc != STC_OK) raise(engine_, "stc_synthesize_packed_ex");
        break;
    }
    out.wav.resize((size_t)off[n]);
    float total = dur[0];
    for (int i = 1; i < n; ++i) total += dur[i] + silence_duration;
    out.duration = {total};
    return out;
}

// ---------------------------------------------------------------------------------------------- utilities
std::vector<unsigned char> wavFileBytes(const std::vector<float>& audio, int sample_rate) {
    // 44-byte canonical RIFF/WAVE header, mono PCM16; sample = (int16)(clamp(x,-1,1) * 32767), truncation (:985-988)
    const uint32_t data_bytes = (uint32_t)(audio.size() * 2);
    std::vector<unsigned char> b(44 + data_bytes);
    auto put32 = [&](size_t o, uint32_t v) { for (int i = 0; i < 4; ++i) b[o + i] = (unsigned char)(v >> (8 * i)); };
    auto put16 = [&](size_t o, uint16_t v) { b[o] = (unsigned char)v; b[o + 1] = (unsigned char)(v >> 8); };
    std::memcpy(&b[0], "RIFF", 4); put32(4, 36 + data_bytes); std::memcpy(&b[8], "WAVEfmt ", 8);
    put32(16, 16); put16(20, 1); put16(22, 1); put32(24, (uint32_t)sample_rate); put32(28, (uint32_t)sample_rate * 2); put16(32, 2); put16(34, 16);
    std::memcpy(&b[36], "data", 4); put32(40, data_bytes);
    for (size_t i = 0; i < audio.size(); ++i) {
        float c = std::fmax(-1.0f, std::fmin(1.0f, audio[i]));
        put16(44 + 2 * i, (uint16_t)static_cast<int16_t>(c * 32767));
    }
    return b;
}

void writeWavFile(const std::string& filename, const std::vector<float>& audio, int sample_rate) {
    std::ofstream f(filename, std::ios::binary);
    if (!f.is_open()) throw std::runtime_error("Failed to open file for writing: " + filename);
    std::vector<unsigned char> b = wavFileBytes(audio, sample_rate);
    f.write(reinterpret_cast<const char*>(b.data()), (std::streamsize)b.size());
}

std::string sanitizeFilename(const std::string& text, int max_len) {
    // one output unit per input "character": ASCII [A-Za-z0-9_] kept, a complete 2/3/4-byte UTF-8 sequence kept whole,
    // anything else becomes '_'
    std::string out;
    size_t i = 0;
    for (int count = 0; i < text.size() && count < max_len; ++count) {
        unsigned char c = (unsigned char)text[i];
        size_t len = 0;
        if (std::isalnum(c) || c == '_') len = 1;
        else if ((c & 0xE0) == 0xC0) len = 2;
        else if ((c & 0xF0) == 0xE0) len = 3;
        else if ((c & 0xF8) == 0xF0) len = 4;
        if (len && i + len <= text.size()) { out.append(text, i, len); i += len; }
        else { out.push_back('_'); ++i; }
    }
    return out;
}

std::vector<std::string> chunkText(const std::string& text, int max_len) {
    size_t need = 0; int n = 0;
    std::vector<char> buf(text.size() + 64);
    for (int attempt = 0; attempt < 2; ++attempt) {
        int rc = stc_chunk_text(text.c_str(), max_len, buf.data(), buf.size(), &need, &n);
        if (rc == STC_OK) break;
        if (rc != STC_ERR_CAPACITY || attempt) raise(nullptr, "chunkText");
        buf.resize(need + 16);
    }
    std::vector<std::string> chunks;
    const char* p = buf.data();
    for (int i = 0; i < n; ++i) { chunks.emplace_back(p); p += chunks.back().size() + 1; }
    if (chunks.empty()) chunks.emplace_back();
    return chunks;
}

}  // namespace supertonic

// ---- C hooks so the host utilities can be pinned against the reference's golden vectors from Python (tests/test_cabi_host.py)
extern "C" {
int stc_host_sanitize_filename(const char* text, int max_len, char* out, size_t cap) {
    std::string s = supertonic::sanitizeFilename(text, max_len);
    if (s.size() + 1 > cap) return -5;
    std::memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}
long stc_host_wav_bytes(const float* audio, size_t n, int sample_rate, unsigned char* out, size_t cap) {
    std::vector<unsigned char> b = supertonic::wavFileBytes(std::vector<float>(audio, audio + n), sample_rate);
    if (b.size() > cap) return -5;
    std::memcpy(out, b.data(), b.size());
    return (long)b.size();
}
// many() on one engine vs MultiGpuTextToSpeech::many() on `nd` devices for the same request and seed: per utterance the duration, the
// sample count and the sum of |samples| of both (tests/test_multi_gpu.py compares them).
int stc_host_many_check(const char* onnx_dir, const int* devices, int nd, const char* const* texts, const char* const* langs, int n,
                        const char* const* style_paths, int total_step, float speed, int max_batch, float* dur, int64_t* nsamp, double* asum) {
    try {
        std::vector<std::string> tx(texts, texts + n), lg(langs, langs + n);
        supertonic::Style st = supertonic::loadVoiceStyle(std::vector<std::string>(style_paths, style_paths + n));
        auto fill = [&](const std::vector<supertonic::TextToSpeech::Utterance>& u, int k) {
            for (int i = 0; i < n; ++i) {
                dur[k * n + i] = u[i].duration; nsamp[k * n + i] = (int64_t)u[i].wav.size();
                double s = 0; for (float v : u[i].wav) s += std::fabs((double)v);
                asum[k * n + i] = s;
            }
        };
        {
            auto one = supertonic::loadTextToSpeech(onnx_dir, true, devices[0]);
            one->setNoiseSeed(11);
            fill(one->many(tx, lg, st, total_step, speed, max_batch), 0);
        }
        supertonic::MultiGpuTextToSpeech multi(onnx_dir, std::vector<int>(devices, devices + nd));
        multi.setNoiseSeed(11);
        fill(multi.many(tx, lg, st, total_step, speed, max_batch), 1);
        return 0;
    } catch (const std::exception& e) { std::fprintf(stderr, "stc_host_many_check: %s\n", e.what()); return -1; }
}
int stc_host_load_voice_style(const char* const* paths, int n, int64_t shapes[6], double sums[2], float ttl_head[4]) {
    try {
        supertonic::Style s = supertonic::loadVoiceStyle(std::vector<std::string>(paths, paths + n));
        for (int i = 0; i < 3; ++i) { shapes[i] = s.getTtlShape()[i]; shapes[3 + i] = s.getDpShape()[i]; }
        sums[0] = sums[1] = 0;
        for (float v : s.getTtlData()) sums[0] += v;
        for (float v : s.getDpData()) sums[1] += v;
        for (int i = 0; i < 4; ++i) ttl_head[i] = s.getTtlData()[i];
        return 0;
    } catch (const std::exception&) { return -2; }
}
int stc_host_load_cfgs(const char* onnx_dir, int out[4]) {
    try {
        supertonic::Config c = supertonic::loadCfgs(onnx_dir);
        out[0] = c.ae.sample_rate; out[1] = c.ae.base_chunk_size; out[2] = c.ttl.chunk_compress_factor; out[3] = c.ttl.latent_dim;
        return 0;
    } catch (const std::exception&) { return -2; }
}
}
