// Minimal ONNX (protobuf wire format) reader: nodes, initializers, metadata_props, graph I/O signatures.
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

struct OnnxAttr {                     // AttributeProto: the scalar / list kinds the four graphs use
    int64_t i = 0; float f = 0.f; std::string s;
    std::vector<int64_t> ints;
};
// Input / output names of a node with CHECKED indexing: the pattern matcher (graph_plan.h) reads `n.in[1]`, `n.out[0]` of nodes it has
// identified by op type only — a damaged or unusual file (a MatMul with one input, a node without outputs) must end in an error, not
// in an out-of-bounds read.
struct OnnxNames : std::vector<std::string> {
    using std::vector<std::string>::vector;
    const std::string& operator[](size_t i) const {
        if (i >= size()) throw std::runtime_error("ONNX node has no input / output #" + std::to_string(i));
        return std::vector<std::string>::operator[](i);
    }
    std::string& operator[](size_t i) {
        if (i >= size()) throw std::runtime_error("ONNX node has no input / output #" + std::to_string(i));
        return std::vector<std::string>::operator[](i);
    }
};
struct OnnxNode {                     // NodeProto
    std::string op, name;
    OnnxNames in, out;
    std::map<std::string, OnnxAttr> attr;
    int64_t attr_i(const std::string& k, int64_t dflt) const { auto it = attr.find(k); return it == attr.end() ? dflt : it->second.i; }
};
struct OnnxValueInfo {                // ValueInfoProto of a graph input / output: dims < 0 are symbolic (dim_param)
    std::string name; int elem_type = 0;
    std::vector<int64_t> dims;
};

struct OnnxFile {
    std::map<std::string, OnnxTensor> initializers;
    std::map<std::string, std::string> metadata;
    std::vector<std::string> inputs, outputs;
    std::vector<OnnxValueInfo> input_info, output_info;
    std::vector<OnnxNode> nodes;      // in file order (ONNX requires a topological order)
    size_t n_nodes = 0;
    const OnnxValueInfo* input(const std::string& name) const {
        for (const auto& v : input_info) if (v.name == name) return &v;
        return nullptr;
    }
};

namespace pb {
struct Reader {
    const uint8_t* p; const uint8_t* end;
    bool ok() const { return p < end; }
    uint64_t varint() {
        uint64_t r = 0; int s = 0;
        while (p < end) {
            uint8_t b = *p++;
            if (s >= 64) throw std::runtime_error("onnx: varint longer than 10 bytes");
            r |= uint64_t(b & 0x7F) << s; if (!(b & 0x80)) return r; s += 7;
        }
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
        int wt; uint64_t v = 0; pb::Reader s{nullptr, nullptr};
        int f = r.next(wt, v, s);
        if (f == 1) { if (wt == 2) { while (s.ok()) t.dims.push_back((int64_t)s.varint()); } else t.dims.push_back((int64_t)v); }
        else if (f == 2) t.dtype = (int)v;
        else if (f == 8) { if (wt != 2) throw std::runtime_error("onnx: tensor name with wire type " + std::to_string(wt)); name = s.str(); }
        else if (f == 9) { if (wt != 2) throw std::runtime_error("onnx: raw_data with wire type " + std::to_string(wt)); t.raw.assign(s.p, s.end); has_raw = true; }
        else if (f == 4) { if (wt == 2) { size_t n = (s.end - s.p) / 4; size_t o = floats.size(); floats.resize(o + n); memcpy(floats.data() + o, s.p, n * 4); }
                           else { float x; uint32_t u = (uint32_t)v; memcpy(&x, &u, 4); floats.push_back(x); } }
        else if (f == 7) { if (wt == 2) { while (s.ok()) i64.push_back((int64_t)s.varint()); } else i64.push_back((int64_t)v); }
        else if (f == 14 && v == 1) throw std::runtime_error("onnx: external tensor data is not supported");
    }
    if (!has_raw) {
        if (!floats.empty()) { t.raw.resize(floats.size() * 4); memcpy(t.raw.data(), floats.data(), t.raw.size()); }
        else if (!i64.empty()) { t.raw.resize(i64.size() * 8); memcpy(t.raw.data(), i64.data(), t.raw.size()); }
    }
    // the payload must hold exactly numel elements of the declared type (a truncated file would otherwise be read past its end)
    const size_t esz = t.dtype == 1 ? 4 : t.dtype == 7 ? 8 : t.dtype == 11 ? 8 : t.dtype == 6 ? 4 : t.dtype == 10 || t.dtype == 16 ? 2 : t.dtype == 2 || t.dtype == 3 || t.dtype == 9 ? 1 : 0;
    for (auto d : t.dims) if (d < 0) throw std::runtime_error("onnx: initializer " + name + " has a negative dimension");
    if (esz && t.raw.size() != t.numel() * esz)
        throw std::runtime_error("onnx: initializer " + name + " holds " + std::to_string(t.raw.size()) + " bytes, its dims need " +
                                 std::to_string(t.numel() * esz));
}

// ValueInfoProto: name = 1, type = 2 -> TypeProto.tensor_type = 1 -> {elem_type = 1, shape = 2 -> dim = 1 -> {dim_value = 1, dim_param = 2}}
inline OnnxValueInfo parse_value_info(pb::Reader r) {
    OnnxValueInfo out;
    while (r.ok()) {
        int wt; uint64_t v = 0; pb::Reader s{nullptr, nullptr};
        const int f = r.next(wt, v, s);
        if (f == 1 && wt == 2) out.name = s.str();
        else if (f == 2 && wt == 2) {
            pb::Reader ty = s;
            while (ty.ok()) {
                pb::Reader tt{nullptr, nullptr};
                if (ty.next(wt, v, tt) != 1 || wt != 2) continue;
                while (tt.ok()) {
                    pb::Reader sh{nullptr, nullptr};
                    const int f3 = tt.next(wt, v, sh);
                    if (f3 == 1 && wt == 0) out.elem_type = (int)v;
                    else if (f3 == 2 && wt == 2) {
                        while (sh.ok()) {
                            pb::Reader dm{nullptr, nullptr};
                            if (sh.next(wt, v, dm) != 1 || wt != 2) continue;
                            int64_t dv = -1;
                            while (dm.ok()) { pb::Reader q{nullptr, nullptr}; uint64_t vv; int w2; if (dm.next(w2, vv, q) == 1 && w2 == 0) dv = (int64_t)vv; }
                            out.dims.push_back(dv);
                        }
                    }
                }
            }
        }
    }
    return out;
}

inline void parse_attr(pb::Reader r, std::string& name, OnnxAttr& a) {
    while (r.ok()) {
        int wt; uint64_t v = 0; pb::Reader s{nullptr, nullptr};
        const int f = r.next(wt, v, s);
        if (f == 1 && wt == 2) name = s.str();
        else if (f == 2 && wt == 5) { uint32_t u = (uint32_t)v; memcpy(&a.f, &u, 4); }
        else if (f == 3 && wt == 0) a.i = (int64_t)v;
        else if (f == 4 && wt == 2) a.s = s.str();
        else if (f == 8) { if (wt == 2) { while (s.ok()) a.ints.push_back((int64_t)s.varint()); } else a.ints.push_back((int64_t)v); }
    }
}

inline OnnxNode parse_node(pb::Reader r) {
    OnnxNode n;
    while (r.ok()) {
        int wt; uint64_t v = 0; pb::Reader s{nullptr, nullptr};
        const int f = r.next(wt, v, s);
        if (wt != 2) continue;
        if (f == 1) n.in.push_back(s.str());
        else if (f == 2) n.out.push_back(s.str());
        else if (f == 3) n.name = s.str();
        else if (f == 4) n.op = s.str();
        else if (f == 5) { std::string k; OnnxAttr a; parse_attr(s, k, a); n.attr.emplace(std::move(k), std::move(a)); }
    }
    return n;
}

inline OnnxFile load_onnx(const std::string& path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f.is_open()) throw std::runtime_error("Failed to open file: " + path);
    const std::streamoff fsize = f.tellg();
    if (fsize < 0) throw std::runtime_error("Failed to size file: " + path);
    std::vector<uint8_t> buf((size_t)fsize);
    f.seekg(0); f.read(reinterpret_cast<char*>(buf.data()), (std::streamsize)buf.size());
    OnnxFile out;
    pb::Reader r{buf.data(), buf.data() + buf.size()};
    while (r.ok()) {
        int wt; uint64_t v = 0; pb::Reader s{nullptr, nullptr};
        int fno = r.next(wt, v, s);
        if (fno == 7 && wt == 2) {                       // GraphProto
            pb::Reader g = s;
            while (g.ok()) {
                pb::Reader s2{nullptr, nullptr};
                int f2 = g.next(wt, v, s2);
                if (f2 != 1 && f2 != 5 && f2 != 11 && f2 != 12) continue;
                if (wt != 2) throw std::runtime_error("onnx: graph field " + std::to_string(f2) + " with wire type " + std::to_string(wt));
                if (f2 == 1) { out.n_nodes++; out.nodes.push_back(parse_node(s2)); }
                else if (f2 == 5) { std::string name; OnnxTensor t; parse_tensor(s2, name, t); out.initializers.emplace(std::move(name), std::move(t)); }
                else if (f2 == 11) { out.input_info.push_back(parse_value_info(s2)); out.inputs.push_back(out.input_info.back().name); }
                else if (f2 == 12) { out.output_info.push_back(parse_value_info(s2)); out.outputs.push_back(out.output_info.back().name); }
            }
        } else if (fno == 14 && wt == 2) {               // metadata_props
            std::string k, val; pb::Reader m = s;
            while (m.ok()) { pb::Reader s2{nullptr, nullptr}; int f2 = m.next(wt, v, s2); if (wt != 2) continue; if (f2 == 1) k = s2.str(); else if (f2 == 2) val = s2.str(); }
            out.metadata[k] = val;
        }
    }
    return out;
}

}  // namespace stc
