// ORACLE BUILD ONLY — inert stand-in for <onnxruntime_cxx_api.h>.
// Lets the UNMODIFIED reference cpp/helper.cpp compile where it lies so that its pure host
// functions (UnicodeProcessor::call, chunkText, getLatentMask, sanitizeFilename, writeWavFile,
// loadVoiceStyle; SURVEY.md §8c "known-answer material") can be run to produce golden vectors.
// Session::Run throws: no neural arithmetic can come out of this stub.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
enum OrtLoggingLevel { ORT_LOGGING_LEVEL_WARNING = 2 };
enum OrtAllocatorType { OrtArenaAllocator = 1 };
enum OrtMemType { OrtMemTypeDefault = 0 };
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
    void* p = nullptr; TensorTypeAndShapeInfo info;
    Value(std::nullptr_t = nullptr) {}
    template <typename T>
    static Value CreateTensor(const MemoryInfo&, T* data, size_t count, const int64_t* shape, size_t rank) {
        Value v; v.p = data; v.info.n = count; v.info.shape.assign(shape, shape + rank); return v;
    }
    template <typename T> T* GetTensorMutableData() { return static_cast<T*>(p); }
    TensorTypeAndShapeInfo GetTensorTypeAndShapeInfo() const { return info; }
};
struct Session {
    Session(Env&, const char*, const SessionOptions&) {}
    std::vector<Value> Run(const RunOptions&, const char* const*, const Value*, size_t, const char* const*, size_t) {
        throw std::runtime_error("oracle stub: Session::Run is not available");
    }
};
}  // namespace Ort
