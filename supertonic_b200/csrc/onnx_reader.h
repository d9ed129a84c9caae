// Minimal ONNX (protobuf wire format) reader: initializers + metadata_props + graph I/O names.
// No protobuf / onnx dependency (neither is in the image). Field numbers from the public onnx.proto
// (SURVEY.md Appendix C). Replaces what Ort::Session(env, path, opts) does with the file at
// reference cpp/helper.cpp:781 — for this library: "load the ONNX initializers into device buffers".
#pragma once
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace stc {

struct OnnxTensor {
    std::vector<int64_t> dims;
    int dtype = 1;                    // TensorProto.DataType (1 = FLOAT, 7 = INT64, 11 = DOUBLE)
    std::vector<uint8_t> raw;         // little-endian payload
    size_t numel() const { size_t n = 1; for (auto d : dims) n *= (size_t)d; return n; }
    const float* f32() const { return reinterpret_cast<const float*>(raw.data()); }
};

struct OnnxFile {
    std::map<std::string, OnnxTensor> initializers;
    std::map<std::string, std::string> metadata;
    std::vector<std::string> inputs, outputs;
    size_t n_nodes = 0;
};

namespace pb {
struct Reader {
    const uint8_t* p; const uint8_t* end;
    bool ok() const { return p < end; }
    uint64_t varint() {
        uint64_t r = 0; int s = 0;
        while (p < end) { uint8_t b = *p++; r |= uint64_t(b & 0x7F) << s; if (!(b & 0x80)) return r; s += 7; }
        throw std::runtime_error("onnx: truncated varint");
    }
    // returns field number; sets wire type; for wt==2 sets sub-reader
    int next(int& wt, uint64_t& val, Reader& sub) {
        uint64_t k = varint(); wt = int(k & 7); int f = int(k >> 3);
        if (wt == 0) val = varint();
        else if (wt == 1) { if (end - p < 8) throw std::runtime_error("onnx: truncated"); memcpy(&val, p, 8); p += 8; }
        else if (wt == 5) { if (end - p < 4) throw std::runtime_error("onnx: truncated"); uint32_t v; memcpy(&v, p, 4); val = v; p += 4; }
        else if (wt == 2) {
            uint64_t n = varint(); if (uint64_t(end - p) < n) throw std::runtime_error("onnx: truncated bytes");
            sub.p = p; sub.end = p + n; p += n;
        } else throw std::runtime_error("onnx: unsupported wire type");
        return f;
    }
    std::string str() const { return std::string(reinterpret_cast<const char*>(p), size_t(end - p)); }
};
}  // namespace pb

inline void parse_tensor(pb::Reader r, std::string& name, OnnxTensor& t) {
    std::vector<float> floats; std::vector<int64_t> i64; bool has_raw = false;
    while (r.ok()) {
        int wt; uint64_t v; pb::Reader s{nullptr, nullptr};
        int f = r.next(wt, v, s);
        if (f == 1) { if (wt == 2) { while (s.ok()) t.dims.push_back((int64_t)s.varint()); } else t.dims.push_back((int64_t)v); }
        else if (f == 2) t.dtype = (int)v;
        else if (f == 8) name = s.str();
        else if (f == 9) { t.raw.assign(s.p, s.end); has_raw = true; }
        else if (f == 4) { if (wt == 2) { size_t n = (s.end - s.p) / 4; size_t o = floats.size(); floats.resize(o + n); memcpy(floats.data() + o, s.p, n * 4); }
                           else { float x; uint32_t u = (uint32_t)v; memcpy(&x, &u, 4); floats.push_back(x); } }
        else if (f == 7) { if (wt == 2) { while (s.ok()) i64.push_back((int64_t)s.varint()); } else i64.push_back((int64_t)v); }
        else if (f == 14 && v == 1) throw std::runtime_error("onnx: external tensor data is not supported");
    }
    if (!has_raw) {
        if (!floats.empty()) { t.raw.resize(floats.size() * 4); memcpy(t.raw.data(), floats.data(), t.raw.size()); }
        else if (!i64.empty()) { t.raw.resize(i64.size() * 8); memcpy(t.raw.data(), i64.data(), t.raw.size()); }
    }
}

inline std::string parse_value_info_name(pb::Reader r) {
    while (r.ok()) { int wt; uint64_t v; pb::Reader s{nullptr, nullptr}; if (r.next(wt, v, s) == 1) return s.str(); }
    return "";
}

inline OnnxFile load_onnx(const std::string& path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f.is_open()) throw std::runtime_error("Failed to open file: " + path);
    std::vector<uint8_t> buf((size_t)f.tellg());
    f.seekg(0); f.read(reinterpret_cast<char*>(buf.data()), (std::streamsize)buf.size());
    OnnxFile out;
    pb::Reader r{buf.data(), buf.data() + buf.size()};
    while (r.ok()) {
        int wt; uint64_t v; pb::Reader s{nullptr, nullptr};
        int fno = r.next(wt, v, s);
        if (fno == 7 && wt == 2) {                       // GraphProto
            pb::Reader g = s;
            while (g.ok()) {
                pb::Reader s2{nullptr, nullptr};
                int f2 = g.next(wt, v, s2);
                if (f2 == 1) out.n_nodes++;
                else if (f2 == 5) { std::string name; OnnxTensor t; parse_tensor(s2, name, t); out.initializers.emplace(std::move(name), std::move(t)); }
                else if (f2 == 11) out.inputs.push_back(parse_value_info_name(s2));
                else if (f2 == 12) out.outputs.push_back(parse_value_info_name(s2));
            }
        } else if (fno == 14 && wt == 2) {               // metadata_props
            std::string k, val; pb::Reader m = s;
            while (m.ok()) { pb::Reader s2{nullptr, nullptr}; int f2 = m.next(wt, v, s2); if (f2 == 1) k = s2.str(); else if (f2 == 2) val = s2.str(); }
            out.metadata[k] = val;
        }
    }
    return out;
}

}  // namespace stc
