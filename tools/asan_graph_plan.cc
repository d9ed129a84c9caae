// Host-only driver for fuzzing the ONNX reader + node-pattern matcher under AddressSanitizer / UBSan (tools/fuzz_graph_loader.py builds and runs it):
//   asan_graph_plan <file.onnx> <duration_predictor|text_encoder|vector_estimator|vocoder>  ->  "ok <plan bytes>" | "rejected: <why>"
#include "graph_plan.h"
#include <iostream>
int main(int argc, char** argv) {
    try { std::string out = stc::derive_arch(stc::load_onnx(argv[1]), argv[2]).dump(); std::cout << "ok " << out.size() << "\n"; }
    catch (const std::exception& e) { std::cout << "rejected: " << e.what() << "\n"; }
    return 0;
}
