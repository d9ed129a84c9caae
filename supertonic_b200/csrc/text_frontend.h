// Host text front-end of libsupertonic_cuda — same observable behaviour as the C++ reference's
// UnicodeProcessor::call (cpp/helper.cpp:52-390) and chunkText (:1117-1186), re-implemented without
// std::regex (the reference constructs >= 10 regex objects per call, SURVEY.md §7 "hard parts"):
// hand-written byte scanners, one pass per rule. Pinned by tests/golden/host_golden.json, which is
// produced by running the unmodified reference.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace stc {

std::string preprocess_text(const std::string& text, const std::string& lang);   // throws on unknown language
void text_to_units(const std::string& utf8, std::vector<uint16_t>& out);         // UTF-8 -> decomposed UTF-16 units
std::vector<std::string> chunk_text(const std::string& text, int max_len);

class TextFrontend {
public:
    void load_indexer(const std::string& json_path);
    // n texts -> text_ids[n*T] (pad 0), text_mask[n*T]; T = max token count. With text_ids == nullptr only T_out is written.
    void call(const char* const* texts, const char* const* langs, int n, int64_t* text_ids, float* text_mask, int64_t T_cap,
              int64_t* T_out) const;
    size_t vocab_entries() const { return indexer_.size(); }
private:
    std::vector<int64_t> indexer_;
};

}  // namespace stc
