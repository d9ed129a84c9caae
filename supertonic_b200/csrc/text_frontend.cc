#include "text_frontend.h"

#include <exception>
#include <thread>

#include <algorithm>
#include <cstring>
#include <fstream>
#include <nlohmann/json.hpp>
#include <stdexcept>

namespace stc {
namespace {

inline bool is_ws(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

std::string strip(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && is_ws((unsigned char)s[a])) ++a;
    while (b > a && is_ws((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}

// sequential find/replace; scanning resumes after the inserted text (so inserted text is never re-matched)
void substitute(std::string& s, const char* from, const char* to) {
    const size_t nf = strlen(from), nt = strlen(to);
    for (size_t pos = s.find(from); pos != std::string::npos; pos = s.find(from, pos + nt)) s.replace(pos, nf, to);
}

struct Rule { const char* from; const char* to; };

// typographic normalisation, in the order the reference applies it
const Rule kSymbols[] = {
    {"\xE2\x80\x93", "-"}, {"\xE2\x80\x91", "-"}, {"\xE2\x80\x94", "-"}, {"_", " "},
    {"\xE2\x80\x9C", "\""}, {"\xE2\x80\x9D", "\""}, {"\xE2\x80\x98", "'"}, {"\xE2\x80\x99", "'"},
    {"\xC2\xB4", "'"}, {"`", "'"}, {"[", " "}, {"]", " "}, {"|", " "}, {"/", " "}, {"#", " "},
    {"\xE2\x86\x92", " "}, {"\xE2\x86\x90", " "},
};
const char* const kDropped[] = {"\xE2\x99\xA5", "\xE2\x98\x86", "\xE2\x99\xA1", "\xC2\xA9", "\\"};
const Rule kExpansions[] = {{"@", " at "}, {"e.g.,", "for example, "}, {"i.e.,", "that is, "}};
const char kTightPunct[] = ",.!?;:'";
const char* const kClosers3[] = {"\xE2\x80\xA6", "\xE3\x80\x82", "\xE3\x80\x8D", "\xE3\x80\x8F", "\xE3\x80\x91", "\xE3\x80\x89",
                                 "\xE3\x80\x8B", "\xE2\x80\xBA", "\xE2\x80\x9C", "\xE2\x80\x9D", "\xE2\x80\x98", "\xE2\x80\x99"};
const char* const kLangs[] = {"en", "ko", "es", "pt", "fr"};

}  // namespace

std::string preprocess_text(const std::string& text, const std::string& lang) {
    std::string s = text;
    for (const Rule& r : kSymbols) substitute(s, r.from, r.to);
    {   // drop 4-byte sequences F0 9F xx xx (emoji block)
        std::string t; t.reserve(s.size());
        for (size_t i = 0; i < s.size();) {
            auto u = [&](size_t k) { return (unsigned char)s[k]; };
            if (i + 3 < s.size() && u(i) == 0xF0 && u(i + 1) == 0x9F && (u(i + 2) & 0xC0) == 0x80 && (u(i + 3) & 0xC0) == 0x80) i += 4;
            else t.push_back(s[i++]);
        }
        s.swap(t);
    }
    for (const char* d : kDropped) substitute(s, d, "");
    for (const Rule& r : kExpansions) substitute(s, r.from, r.to);
    for (const char* p = kTightPunct; *p; ++p) {      // " ," -> "," etc. (one left-to-right pass per mark, like regex_replace)
        std::string t; t.reserve(s.size());
        for (size_t i = 0; i < s.size(); ++i) {
            if (s[i] == ' ' && i + 1 < s.size() && s[i + 1] == *p) continue;
            t.push_back(s[i]);
        }
        s.swap(t);
    }
    for (const char* q : {"\"\"", "''", "``"}) {
        for (size_t pos = s.find(q); pos != std::string::npos; pos = s.find(q)) s.replace(pos, 2, std::string(1, q[0]));
    }
    {   // collapse whitespace runs, trim
        std::string t; t.reserve(s.size());
        bool in_ws = false;
        for (unsigned char c : s) {
            if (is_ws(c)) { if (!in_ws) t.push_back(' '); in_ws = true; }
            else { t.push_back((char)c); in_ws = false; }
        }
        s = strip(t);
    }
    if (!s.empty()) {
        bool closed = strchr(".!?;:,'\")]}>", s.back()) != nullptr;
        if (!closed && s.size() >= 3) {
            const char* tail = s.data() + s.size() - 3;
            for (const char* c : kClosers3) if (memcmp(tail, c, 3) == 0) { closed = true; break; }
        }
        if (!closed) s.push_back('.');
    }
    if (std::none_of(std::begin(kLangs), std::end(kLangs), [&](const char* l) { return lang == l; }))
        throw std::runtime_error("Invalid language: " + lang + ". Available: en, ko, es, pt, fr");
    return "<" + lang + ">" + s + "</" + lang + ">";
}

namespace {
// precomposed Latin letter -> (base letter, combining mark) for the es/pt/fr repertoire
struct Decomp { uint16_t cp, base, mark; };
const Decomp kLatin[] = {
    {0xC0, 'A', 0x300}, {0xC1, 'A', 0x301}, {0xC2, 'A', 0x302}, {0xC3, 'A', 0x303}, {0xC4, 'A', 0x308}, {0xC7, 'C', 0x327},
    {0xC8, 'E', 0x300}, {0xC9, 'E', 0x301}, {0xCA, 'E', 0x302}, {0xCB, 'E', 0x308}, {0xCC, 'I', 0x300}, {0xCD, 'I', 0x301},
    {0xCE, 'I', 0x302}, {0xCF, 'I', 0x308}, {0xD1, 'N', 0x303}, {0xD2, 'O', 0x300}, {0xD3, 'O', 0x301}, {0xD4, 'O', 0x302},
    {0xD5, 'O', 0x303}, {0xD6, 'O', 0x308}, {0xD9, 'U', 0x300}, {0xDA, 'U', 0x301}, {0xDB, 'U', 0x302}, {0xDC, 'U', 0x308},
    {0xE0, 'a', 0x300}, {0xE1, 'a', 0x301}, {0xE2, 'a', 0x302}, {0xE3, 'a', 0x303}, {0xE4, 'a', 0x308}, {0xE7, 'c', 0x327},
    {0xE8, 'e', 0x300}, {0xE9, 'e', 0x301}, {0xEA, 'e', 0x302}, {0xEB, 'e', 0x308}, {0xEC, 'i', 0x300}, {0xED, 'i', 0x301},
    {0xEE, 'i', 0x302}, {0xEF, 'i', 0x308}, {0xF1, 'n', 0x303}, {0xF2, 'o', 0x300}, {0xF3, 'o', 0x301}, {0xF4, 'o', 0x302},
    {0xF5, 'o', 0x303}, {0xF6, 'o', 0x308}, {0xF9, 'u', 0x300}, {0xFA, 'u', 0x301}, {0xFB, 'u', 0x302}, {0xFC, 'u', 0x308},
};

inline void emit(uint32_t cp, std::vector<uint16_t>& out) {
    if (cp >= 0xAC00 && cp < 0xAC00 + 11172) {           // Hangul syllable -> conjoining Jamo (UAX #15 arithmetic)
        uint32_t s = cp - 0xAC00, t = s % 28;
        out.push_back((uint16_t)(0x1100 + s / 588));
        out.push_back((uint16_t)(0x1161 + (s % 588) / 28));
        if (t) out.push_back((uint16_t)(0x11A7 + t));
        return;
    }
    if (cp >= 0xC0 && cp <= 0xFC) {
        const Decomp* e = std::lower_bound(std::begin(kLatin), std::end(kLatin), cp, [](const Decomp& d, uint32_t v) { return d.cp < v; });
        if (e != std::end(kLatin) && e->cp == cp) { out.push_back(e->base); out.push_back(e->mark); return; }
    }
    out.push_back((uint16_t)(cp & 0xFFFF));
}
}  // namespace

void text_to_units(const std::string& t, std::vector<uint16_t>& out) {
    const size_t n = t.size();
    for (size_t i = 0; i < n;) {
        auto u = [&](size_t k) -> uint32_t { return (unsigned char)t[k]; };
        uint32_t c = u(i), cp;
        if (c < 0x80) { cp = c; i += 1; }
        else if ((c & 0xE0) == 0xC0 && i + 1 < n) { cp = ((c & 0x1F) << 6) | (u(i + 1) & 0x3F); i += 2; }
        else if ((c & 0xF0) == 0xE0 && i + 2 < n) { cp = ((c & 0x0F) << 12) | ((u(i + 1) & 0x3F) << 6) | (u(i + 2) & 0x3F); i += 3; }
        else if ((c & 0xF8) == 0xF0 && i + 3 < n) {
            cp = ((c & 0x07) << 18) | ((u(i + 1) & 0x3F) << 12) | ((u(i + 2) & 0x3F) << 6) | (u(i + 3) & 0x3F); i += 4;
        } else { i += 1; continue; }                      // stray byte: skipped
        emit(cp, out);
    }
}

void TextFrontend::load_indexer(const std::string& path) {
    std::ifstream f(path);
    if (!f.is_open()) throw std::runtime_error("Failed to open file: " + path);
    nlohmann::json j; f >> j;
    indexer_ = j.get<std::vector<int64_t>>();
}

void TextFrontend::call(const char* const* texts, const char* const* langs, int n, int64_t* text_ids, float* text_mask,
                        int64_t T_cap, int64_t* T_out) const {
    std::vector<std::vector<uint16_t>> units((size_t)n);
    // texts are independent: large requests are normalised on a few host threads (semantics unchanged — the reference does
    // them one by one, cpp/helper.cpp:362-368); an exception in a worker (unknown language) is rethrown here
    const int workers = n >= 64 ? (int)std::min<unsigned>(8, std::max(1u, std::thread::hardware_concurrency())) : 1;
    if (workers > 1) {
        std::vector<std::thread> pool;
        std::vector<std::exception_ptr> errs((size_t)workers);
        for (int w = 0; w < workers; ++w)
            pool.emplace_back([&, w]() {
                try { for (int i = w; i < n; i += workers) text_to_units(preprocess_text(texts[i], langs[i]), units[i]); }
                catch (...) { errs[(size_t)w] = std::current_exception(); }
            });
        for (auto& t : pool) t.join();
        for (auto& e : errs) if (e) std::rethrow_exception(e);
    } else {
        for (int i = 0; i < n; ++i) text_to_units(preprocess_text(texts[i], langs[i]), units[i]);
    }
    int64_t T = 0;
    for (int i = 0; i < n; ++i) T = std::max<int64_t>(T, (int64_t)units[i].size());
    if (T_out) *T_out = T;
    if (!text_ids) return;
    if (T_cap < T) throw std::runtime_error("stc_text_to_ids: T_cap smaller than the token count");
    for (int i = 0; i < n; ++i) {
        int64_t* row = text_ids + (size_t)i * T_cap;
        float* mrow = text_mask ? text_mask + (size_t)i * T_cap : nullptr;
        std::fill(row, row + T_cap, 0);
        if (mrow) std::fill(mrow, mrow + T_cap, 0.f);
        for (size_t j = 0; j < units[i].size(); ++j) {
            if (units[i][j] < indexer_.size()) row[j] = indexer_[units[i][j]];
            if (mrow) mrow[j] = 1.f;
        }
    }
}

// ---- chunking -----------------------------------------------------------------------------------
namespace {
struct Span { size_t a, b; };
// pieces between delimiter matches; a trailing empty piece is not produced, an empty input yields one empty piece
template <typename MatchFn>
std::vector<Span> split_by(const std::string& s, MatchFn match_at) {
    std::vector<Span> out;
    size_t pos = 0, i = 0;
    while (i < s.size()) {
        size_t len = match_at(s, i);
        if (len) { out.push_back({pos, i}); i += len; pos = i; } else ++i;
    }
    if (pos < s.size() || out.empty()) out.push_back({pos, s.size()});
    return out;
}
// "[.!?]\s+" at i ?
size_t sent_delim(const std::string& s, size_t i) {
    if (s[i] != '.' && s[i] != '!' && s[i] != '?') return 0;
    size_t j = i + 1;
    while (j < s.size() && is_ws((unsigned char)s[j])) ++j;
    return j > i + 1 ? j - i : 0;
}
}  // namespace

std::vector<std::string> chunk_text(const std::string& text, int max_len) {
    std::vector<std::string> chunks;
    // paragraph delimiter = '\n' + whitespace* + '\n'+ ; the regex is greedy, so the match runs from the first newline to the
    // LAST newline of the whitespace run (trailing blanks after it stay with the next paragraph and are trimmed anyway)
    auto para_at = [](const std::string& s, size_t i) -> size_t {
        if (s[i] != '\n') return 0;
        size_t j = i + 1, last_nl = 0;
        while (j < s.size() && is_ws((unsigned char)s[j])) { if (s[j] == '\n') last_nl = j; ++j; }
        return last_nl ? last_nl + 1 - i : 0;
    };
    for (const Span& ps : split_by(text, para_at)) {
        std::string para = strip(text.substr(ps.a, ps.b - ps.a));
        if (para.empty()) continue;
        std::vector<std::string> sentences;
        for (const Span& ss : split_by(para, sent_delim)) {
            if (ss.b == ss.a) continue;
            std::string sent = para.substr(ss.a, ss.b - ss.a);
            for (size_t i = ss.a; i < para.size(); ++i) {          // re-attach the first delimiter found at/after the piece start
                size_t len = sent_delim(para, i);
                if (len) { sent += para.substr(i, len); break; }
            }
            sentences.push_back(std::move(sent));
        }
        std::string cur;
        for (const std::string& s : sentences) {
            if ((int)(cur.size() + s.size() + 1) <= max_len) { if (!cur.empty()) cur += ' '; cur += s; }
            else { if (!cur.empty()) chunks.push_back(strip(cur)); cur = s; }
        }
        if (!cur.empty()) chunks.push_back(strip(cur));
    }
    if (chunks.empty()) chunks.push_back(strip(text));
    return chunks;
}

}  // namespace stc
