// Layer plan of a Supertonic graph derived from its ONNX NODES (no private metadata needed).
//
// The reference builds its sessions from whatever .onnx files are on disk (loadOnnx / loadOnnxAll, cpp/helper.cpp:776-795) and lets
// ONNX Runtime interpret the nodes. This library runs fused kernels instead, so it needs to know which nodes form a ConvNeXt block, an
// attention layer, a time-conditioning add, ...: derive_arch() walks the dataflow from the graph inputs and recognises
//   ConvNeXt     Conv(group = C) -> Transpose -> LayerNormalization -> MatMul + Add -> Div/Erf/Add/Mul/Mul (exact GELU) -> MatMul + Add
//                -> Mul(layer scale) -> Transpose -> Add(residual) [-> Mul(mask)]
//   attention    Transpose -> LayerNormalization -> MatMul + Add (q) | (k) | (v) -> Reshape/Transpose head split [-> rotary chain] ->
//                MatMul(q, k^T) -> Mul(scale) [-> Add(key-mask bias)] -> Softmax -> MatMul(., v) -> Transpose/Reshape -> MatMul + Add (o)
//                -> Transpose -> Add(residual) [-> Mul(mask)];  self / text_emb / style_ttl context by where k and v come from
//   time_cond    Add(x, Unsqueeze(MatMul + Add(time embedding))) -> Mul(mask)
//   proj_in/out, time MLP (Div -> Unsqueeze -> Mul(freqs) -> Sin | Cos -> Concat -> MatMul + Add -> GELU -> MatMul + Add),
//   conv_in (Conv(group = 1) -> BatchNormalization), head (Transpose -> LayerNormalization -> MatMul + Add -> Reshape), the
//   duration head (-> Clip -> Exp -> Mul -> Transpose -> Mul(mask) -> ReduceSum)
// and records, for every layer, the NAMES of the initializers it found in those positions — the loader never assumes a naming
// scheme. Every node that lies on a path from the inputs to the output must belong to a recognised layer: otherwise the graph is
// rejected with the list of unexplained nodes (so a differently exported model fails with a to-do list, not with wrong audio).
// The plan has the schema of the `stc_arch` metadata the surrogate generator writes (surrogate.py), which remains the fast path.
#pragma once
#include <nlohmann/json.hpp>

#include <algorithm>
#include <set>
#include <sstream>
#include <unordered_map>

#include "onnx_reader.h"

namespace stc {

struct PlanError : std::runtime_error { using std::runtime_error::runtime_error; };

class GraphPlanner {
public:
    using json = nlohmann::json;
    explicit GraphPlanner(const OnnxFile& f) : f_(f), used_(f.nodes.size(), 0) {
        for (size_t i = 0; i < f.nodes.size(); ++i) {
            for (const auto& v : f.nodes[i].in) cons_[v].push_back((int)i);
            for (const auto& v : f.nodes[i].out) prod_[v] = (int)i;
        }
        for (const auto& v : f.inputs) inputs_.insert(v);
    }

    json derive(const std::string& kind) {
        json arch;
        // the one mask every masked stage of this graph multiplies by (the engine applies ITS mask wherever the plan says "masked")
        mask_name_ = kind == "vector_estimator" ? "latent_mask" : "text_mask";
        if (kind == "duration_predictor") arch = duration_predictor();
        else if (kind == "text_encoder") arch = text_encoder();
        else if (kind == "vector_estimator") arch = vector_estimator();
        else if (kind == "vocoder") arch = vocoder();
        else throw PlanError("unknown graph kind " + kind);
        check_all_explained(kind);
        arch["kind"] = kind;
        arch["derived_from"] = "nodes";
        return arch;
    }

private:
    const OnnxFile& f_;
    std::unordered_map<std::string, std::vector<int>> cons_;
    std::unordered_map<std::string, int> prod_;
    std::set<std::string> inputs_;
    std::vector<char> used_;
    std::string mask_name_;

    // ---- graph access
    const OnnxNode& N(int i) const { return f_.nodes[i]; }
    bool is_init(const std::string& v) const { return f_.initializers.count(v) != 0; }
    bool is_input(const std::string& v) const { return inputs_.count(v) != 0; }
    const OnnxTensor& init(const std::string& v) const {
        auto it = f_.initializers.find(v);
        if (it == f_.initializers.end()) throw PlanError("expected an initializer, found value '" + v + "'");
        return it->second;
    }
    int producer(const std::string& v) const { auto it = prod_.find(v); return it == prod_.end() ? -1 : it->second; }
    [[noreturn]] void fail(const std::string& what, const std::string& at) const {
        throw PlanError("graph pattern not recognised: " + what + " (at value '" + at + "')");
    }
    // the consumer of `v` with op type `op` (optionally: whose other input satisfies pred); -1 if none
    template <typename Pred> int consumer(const std::string& v, const char* op, Pred pred) const {
        auto it = cons_.find(v);
        if (it == cons_.end()) return -1;
        for (int i : it->second) if (N(i).op == op && pred(N(i))) return i;
        return -1;
    }
    int consumer(const std::string& v, const char* op) const { return consumer(v, op, [](const OnnxNode&) { return true; }); }
    int need(const std::string& v, const char* op, const char* what) const {
        int i = consumer(v, op);
        if (i < 0) fail(std::string(what) + ": no " + op + " consumes it", v);
        return i;
    }
    int need_prod(const std::string& v, const char* op, const char* what) const {
        int i = producer(v);
        if (i < 0 || N(i).op != op) fail(std::string(what) + ": expected to be produced by " + op + (i >= 0 ? ", found " + N(i).op : ""), v);
        return i;
    }
    static const std::string& other(const OnnxNode& n, const std::string& v) { return n.in[0] == v ? n.in[1] : n.in[0]; }
    float scalar(const std::string& v) const {
        const OnnxTensor& t = init(v);
        if (t.numel() != 1 || t.dtype != 1) throw PlanError("expected a float scalar constant at '" + v + "'");
        return t.f32()[0];
    }
    std::vector<int64_t> ints(const std::string& v) const {
        const OnnxTensor& t = init(v);
        if (t.dtype != 7) throw PlanError("expected an int64 constant at '" + v + "'");
        std::vector<int64_t> r(t.numel());
        if (t.raw.size() != r.size() * 8) throw PlanError("int64 constant '" + v + "' has a payload of the wrong size");
        if (!r.empty()) memcpy(r.data(), t.raw.data(), r.size() * 8);
        return r;
    }
    bool ints_are(const std::string& v, std::initializer_list<int64_t> want) const {
        if (!is_init(v)) return false;
        auto g = ints(v);
        return g.size() == want.size() && std::equal(want.begin(), want.end(), g.begin());
    }
    // axes of an Unsqueeze / ReduceSum: second input (opset >= 13) or the `axes` attribute
    bool axes_are(const OnnxNode& n, std::initializer_list<int64_t> want) const {
        if (n.in.size() >= 2 && !n.in[1].empty()) return ints_are(n.in[1], want);
        auto it = n.attr.find("axes");
        return it != n.attr.end() && it->second.ints.size() == want.size() && std::equal(want.begin(), want.end(), it->second.ints.begin());
    }
    // every node between `y` and the values in `stops` (exclusive) is explained by the layer that ends in y
    void mark_cone(const std::string& y, const std::set<std::string>& stops) {
        std::vector<std::string> st{y};
        while (!st.empty()) {
            std::string v = st.back(); st.pop_back();
            if (stops.count(v) || is_init(v) || is_input(v)) continue;
            int i = producer(v);
            if (i < 0 || used_[i]) continue;
            used_[i] = 1;
            for (const auto& in : N(i).in) st.push_back(in);
        }
    }
    void check_all_explained(const std::string& kind) const {
        // nodes on a path to a graph output
        std::vector<char> needed(f_.nodes.size(), 0);
        std::vector<std::string> st(f_.outputs.begin(), f_.outputs.end());
        while (!st.empty()) {
            std::string v = st.back(); st.pop_back();
            int i = producer(v);
            if (i < 0 || needed[i]) continue;
            needed[i] = 1;
            for (const auto& in : N(i).in) st.push_back(in);
        }
        std::ostringstream os; int n = 0;
        for (size_t i = 0; i < f_.nodes.size(); ++i)
            if (needed[i] && !used_[i]) { if (n++ < 40) os << (n > 1 ? ", " : "") << N(i).op << ":" << (N(i).name.empty() ? N(i).out[0] : N(i).name); }
        if (n) throw PlanError(kind + ": " + std::to_string(n) + " node(s) on the path to the output belong to no recognised layer: " + os.str());
    }

    // ---- small patterns
    struct Lin { std::string w, b, out; int K = 0, N = 0; };
    Lin linear(const std::string& v, const char* what) const {
        int m = consumer(v, "MatMul", [&](const OnnxNode& n) { return n.in[0] == v && is_init(n.in[1]); });
        if (m < 0) fail(std::string(what) + ": no MatMul with an initializer weight", v);
        return linear_at(m, what);
    }
    Lin linear_at(int m, const char* what) const {
        Lin l; l.w = N(m).in[1];
        const OnnxTensor& w = init(l.w);
        if (w.dims.size() != 2) fail(std::string(what) + ": weight is not a matrix", l.w);
        l.K = (int)w.dims[0]; l.N = (int)w.dims[1];
        int a = consumer(N(m).out[0], "Add", [&](const OnnxNode& n) { return is_init(other(n, N(m).out[0])); });
        if (a < 0) fail(std::string(what) + ": MatMul without a bias Add", N(m).out[0]);
        l.b = other(N(a), N(m).out[0]); l.out = N(a).out[0];
        if ((int)init(l.b).numel() != l.N) fail(std::string(what) + ": bias length", l.b);
        return l;
    }
    static float attr_f(const OnnxNode& n, const char* k, float dflt) { auto it = n.attr.find(k); return it == n.attr.end() ? dflt : it->second.f; }
    // the `perm` of a Transpose the kernels' fixed layouts stand for: {0,2,1} (NCL <-> NLC), {0,2,1,3} (head split / merge),
    // {0,1,3,2} (K^T), {0,2,3,1} (the vocoder's un-compress); anything else computes something else
    void check_perm(int t, std::initializer_list<int64_t> want, const char* what) const {
        auto it = N(t).attr.find("perm");
        if (it == N(t).attr.end() || it->second.ints.size() != want.size() || !std::equal(want.begin(), want.end(), it->second.ints.begin()))
            fail(std::string(what) + ": unexpected Transpose perm", N(t).out[0]);
    }
    // a Transpose consuming v with the given perm (-1 if none) / required
    int needT(const std::string& v, std::initializer_list<int64_t> want, const char* what) const {
        int t = need(v, "Transpose", what);
        check_perm(t, want, what);
        return t;
    }
    // exact GELU as exported for opset < 20: Mul(Mul(x, Add(Erf(Div(x, sqrt2)), 1)), 0.5)
    std::string gelu(const std::string& v) const {
        int d = consumer(v, "Div", [&](const OnnxNode& n) { return n.in[0] == v; });           // x / sqrt2, not sqrt2 / x
        if (d < 0) fail("GELU: no Div(x, sqrt2) consumes it", v);
        if (std::fabs(scalar(N(d).in[1]) - 1.41421356f) > 1e-5f) fail("GELU: divisor is not sqrt(2)", v);
        int e = need(N(d).out[0], "Erf", "GELU");
        int a = need(N(e).out[0], "Add", "GELU");
        int m1 = consumer(v, "Mul", [&](const OnnxNode& n) { return other(n, v) == N(a).out[0]; });
        if (m1 < 0) fail("GELU: no Mul(x, erf + 1)", v);
        int m2 = need(N(m1).out[0], "Mul", "GELU");
        if (std::fabs(scalar(other(N(m2), N(m1).out[0])) - 0.5f) > 1e-6f || std::fabs(scalar(other(N(a), N(e).out[0])) - 1.0f) > 1e-6f)
            fail("GELU constants", v);
        return N(m2).out[0];
    }
    struct LN { std::string g, b, out; };
    LN layernorm(const std::string& v, const char* what) const {
        int n = need(v, "LayerNormalization", what);
        // the kernels normalise over the last (channel) axis with epsilon 1e-6; the ONNX default is 1e-5
        if (std::fabs(attr_f(N(n), "epsilon", 1e-5f) - 1e-6f) > 1e-9f) fail(std::string(what) + ": LayerNormalization epsilon is not 1e-6", v);
        if (N(n).attr_i("axis", -1) != -1 && N(n).attr_i("axis", -1) != 2) fail(std::string(what) + ": LayerNormalization axis is not the last one", v);
        LN l; l.g = N(n).in[1]; l.b = N(n).in[2]; l.out = N(n).out[0];
        init(l.g); init(l.b);
        return l;
    }
    // trailing Mul by a graph-input mask
    bool masked_tail(std::string& y) const {
        int m = consumer(y, "Mul", [&](const OnnxNode& n) { return is_input(other(n, y)); });
        if (m < 0) return false;
        if (other(N(m), y) != mask_name_) fail("masked by '" + other(N(m), y) + "', expected " + mask_name_, y);
        y = N(m).out[0];
        return true;
    }

    // ---- ConvNeXt block at x; returns the layer entry, advances x
    bool at_convnext(const std::string& x) const {
        return consumer(x, "Conv", [&](const OnnxNode& n) { return n.in[0] == x && n.attr_i("group", 1) > 1; }) >= 0;
    }
    json convnext(std::string& x) {
        const std::string x0 = x;
        int cv = consumer(x, "Conv", [&](const OnnxNode& n) { return n.in[0] == x && n.attr_i("group", 1) > 1; });
        const OnnxNode& c = N(cv);
        const OnnxTensor& w = init(c.in[1]);
        const int C = (int)c.attr_i("group", 1);
        if (w.dims.size() != 3 || w.dims[0] != C || w.dims[1] != 1) fail("depthwise Conv weight must be [C,1,K]", c.in[1]);
        const int K = (int)w.dims[2];
        auto attr_ints = [&](const char* k) { auto it = c.attr.find(k); return it == c.attr.end() ? std::vector<int64_t>{} : it->second.ints; };
        auto dil = attr_ints("dilations"), pads = attr_ints("pads"), str = attr_ints("strides");
        const int d = dil.empty() ? 1 : (int)dil[0], span = d * (K - 1);
        if (!str.empty() && str[0] != 1) fail("strided depthwise Conv", x);
        if (pads.size() != 2 || pads[0] + pads[1] != span) fail("depthwise Conv must be length-preserving", x);
        bool causal;
        if (pads[1] == 0 && pads[0] == span) causal = true;
        else if (pads[0] == span / 2) causal = false;
        else fail("depthwise Conv padding is neither causal nor centred", x);
        if (c.in.size() < 3) fail("depthwise Conv without bias", x);
        int t1 = needT(c.out[0], {0, 2, 1}, "ConvNeXt");
        LN ln = layernorm(N(t1).out[0], "ConvNeXt LayerNorm");
        Lin l1 = linear(ln.out, "ConvNeXt pw1");
        std::string g = gelu(l1.out);
        Lin l2 = linear(g, "ConvNeXt pw2");
        int mg = consumer(l2.out, "Mul", [&](const OnnxNode& n) { return is_init(other(n, l2.out)); });
        if (mg < 0) fail("ConvNeXt layer scale", l2.out);
        const std::string gamma = other(N(mg), l2.out);
        int t2 = needT(N(mg).out[0], {0, 2, 1}, "ConvNeXt");
        int ad = consumer(N(t2).out[0], "Add", [&](const OnnxNode& n) { return other(n, N(t2).out[0]) == x0; });
        if (ad < 0) fail("ConvNeXt residual Add", N(t2).out[0]);
        std::string y = N(ad).out[0];
        const bool masked = masked_tail(y);
        if (l1.K != C || l2.N != C || l1.N != l2.K) fail("ConvNeXt projection shapes", x);
        mark_cone(y, {x0});
        x = y;
        return json{{"type", "convnext"}, {"name", c.in[1]}, {"C", C}, {"H", l1.N}, {"K", K}, {"dilation", d}, {"causal", causal}, {"masked", masked},
                    {"t", {{"dw_w", c.in[1]}, {"dw_b", c.in[2]}, {"ln_g", ln.g}, {"ln_b", ln.b}, {"w1", l1.w}, {"b1", l1.b}, {"w2", l2.w}, {"b2", l2.b},
                           {"gamma", gamma}}}};
    }

    // ---- attention layer at x
    bool at_attention(const std::string& x) const {
        int t = consumer(x, "Transpose");
        if (t < 0 || consumer(N(t).out[0], "LayerNormalization") < 0) return false;
        return consumer(x, "Add") >= 0;          // the residual: a head (Transpose -> LayerNorm -> linear) has no second consumer
    }
    // value -> (after the [B,N,C] -> [B,h,N,dh] split): Reshape -> Transpose
    std::string split_heads(const std::string& v, int& heads, int& dh) const {
        int r = need(v, "Reshape", "attention head split");
        auto shp = ints(N(r).in[1]);
        if (shp.size() != 4) fail("attention head split shape", v);
        heads = (int)shp[2]; dh = (int)shp[3];
        return N(needT(N(r).out[0], {0, 2, 1, 3}, "attention head split")).out[0];
    }
    struct Rope { std::string kind = "none", freqs; };
    // forward through an optional rotary chain starting at the split value t; returns the value that enters Q.K^T
    std::string through_rope(const std::string& t, Rope& rp, int dh) const {
        int s1 = consumer(t, "Slice");
        if (s1 < 0) return t;
        // Concat(r1, r2) reached through either half
        int m = need(N(s1).out[0], "Mul", "rotary");
        int nx = consumer(N(m).out[0], "Sub"); if (nx < 0) nx = need(N(m).out[0], "Add", "rotary");
        int cc = need(N(nx).out[0], "Concat", "rotary");
        rp = verify_rope(cc, t, dh);
        return N(cc).out[0];
    }
    // The whole rotate-half pattern behind a Concat, operand by operand (other rotary variants — interleaved pairs, a different sign
    // convention — are refused, not approximated):
    //   t1 = Slice(t, 0, dh/2, axis 3), t2 = Slice(t, dh/2, dh, axis 3); c = Cos(a), s = Sin(a) of ONE angle a = pos * freqs
    //   Concat(axis 3)( Sub(Mul(t1, c), Mul(t2, s)),  Add(Mul(t1, s), Mul(t2, c)) )
    Rope verify_rope(int cc, const std::string& t, int dh) const {
        const OnnxNode& c = N(cc);
        const int64_t cax = c.attr_i("axis", 0);
        if (c.in.size() != 2 || (cax != 3 && cax != -1)) fail("rotary: Concat of two halves along the head dimension expected", c.out[0]);
        struct Term { std::string slice, trig, angle; int mul; };
        auto term = [&](const std::string& v) {
            Term r; r.mul = need_prod(v, "Mul", "rotary");
            for (int k = 0; k < 2; ++k) {
                int ps = producer(N(r.mul).in[k]), pt = producer(N(r.mul).in[1 - k]);
                if (ps >= 0 && N(ps).op == "Slice" && pt >= 0 && (N(pt).op == "Cos" || N(pt).op == "Sin")) {
                    r.slice = N(r.mul).in[k]; r.trig = N(pt).op; r.angle = N(pt).in[0];
                    return r;
                }
            }
            fail("rotary: expected Mul(Slice, Cos | Sin)", v);
        };
        int sb = need_prod(c.in[0], "Sub", "rotary"), ad = need_prod(c.in[1], "Add", "rotary");
        Term a0 = term(N(sb).in[0]), a1 = term(N(sb).in[1]), b0 = term(N(ad).in[0]), b1 = term(N(ad).in[1]);
        if (b0.trig == "Cos") std::swap(b0, b1);                                 // Add commutes: b0 = t1 * sin, b1 = t2 * cos
        if (a0.trig != "Cos" || a1.trig != "Sin" || b0.trig != "Sin" || b1.trig != "Cos" || a0.slice != b0.slice || a1.slice != b1.slice ||
            a0.slice == a1.slice || a0.angle != a1.angle || a0.angle != b0.angle || a0.angle != b1.angle)
            fail("rotary: expected Sub(t1 cos, t2 sin) and Add(t1 sin, t2 cos) of one angle", c.out[0]);
        auto slice_ok = [&](const std::string& v, int64_t lo, int64_t hi_min) {
            const OnnxNode& sl = N(producer(v));
            if (sl.in.size() < 4 || sl.in[0] != t) return false;
            auto st = ints(sl.in[1]), en = ints(sl.in[2]), ax = ints(sl.in[3]);
            if (sl.in.size() > 4 && !sl.in[4].empty()) { auto sp = ints(sl.in[4]); if (sp.size() != 1 || sp[0] != 1) return false; }
            return st.size() == 1 && en.size() == 1 && ax.size() == 1 && st[0] == lo && en[0] >= hi_min && (lo == 0 ? en[0] == hi_min : true) &&
                   (ax[0] == 3 || ax[0] == -1);
        };
        if (dh % 2 || !slice_ok(a0.slice, 0, dh / 2) || !slice_ok(a1.slice, dh / 2, dh)) fail("rotary: halves are not t[..., :dh/2] and t[..., dh/2:]", t);
        return rope_of(a0.mul, a0.slice);
    }
    // the rotary parameters behind Mul(slice, cos|sin): cos / sin <- Mul(pos, freqs), pos <- Unsqueeze(cumsum(mask) - 1 [/ sum(mask)])
    Rope rope_of(int m, const std::string& slice_value) const {
        Rope rp;
        const std::string& trig = other(N(m), slice_value);
        int tg = producer(trig);
        if (tg < 0 || (N(tg).op != "Cos" && N(tg).op != "Sin")) fail("rotary: expected Cos / Sin", trig);
        int ang = need_prod(N(tg).in[0], "Mul", "rotary angle");
        const std::string& a0 = N(ang).in[0]; const std::string& a1 = N(ang).in[1];
        rp.freqs = is_init(a0) ? a0 : a1;
        init(rp.freqs);
        const std::string& pos = is_init(a0) ? a1 : a0;
        int un = need_prod(pos, "Unsqueeze", "rotary positions");
        if (!axes_are(N(un), {3})) fail("rotary positions: Unsqueeze axes != [3]", pos);
        int pp = producer(N(un).in[0]);
        auto is_cumsum_minus_1 = [&](int sb) {                                  // Sub(CumSum(mask), 1), in this operand order
            if (sb < 0 || N(sb).op != "Sub" || !is_init(N(sb).in[1]) || std::fabs(scalar(N(sb).in[1]) - 1.0f) > 1e-6f) return false;
            int cs = producer(N(sb).in[0]);
            return cs >= 0 && N(cs).op == "CumSum" && is_input(N(cs).in[0]) && (ints_are(N(cs).in[1], {2}) || ints_are(N(cs).in[1], {-1})) &&
                   N(cs).attr_i("exclusive", 0) == 0 && N(cs).attr_i("reverse", 0) == 0;                 // along the sequence axis of mask [B,1,N]
        };
        if (pp >= 0 && N(pp).op == "Div") {                                     // positions / sequence length (length-aware RoPE)
            int len = producer(N(pp).in[1]);
            if (!is_cumsum_minus_1(producer(N(pp).in[0])) || len < 0 || N(len).op != "ReduceSum" || !is_input(N(len).in[0]) ||
                !(axes_are(N(len), {2}) || axes_are(N(len), {-1})) || N(len).attr_i("keepdims", 1) != 1)
                fail("rotary positions: expected (cumsum(mask) - 1) / sum(mask)", pos);
            rp.kind = "norm";
        } else if (is_cumsum_minus_1(pp)) rp.kind = "abs";                      // cumsum(mask) - 1
        else fail("rotary positions are neither cumsum(mask) - 1 nor that divided by the length", pos);
        return rp;
    }
    // backward from the value entering a MatMul to the projection (MatMul + Add) that made it: undo rotary, head split
    int back_to_projection(std::string v, const char* what, int heads, int dh, Rope* rope = nullptr) const {
        int p = producer(v);
        if (p >= 0 && N(p).op == "Concat") {                                   // rotary: Concat(Sub(Mul(Slice(t),..),..), ..)
            int sb = need_prod(N(p).in[0], "Sub", what);
            int ml = need_prod(N(sb).in[0], "Mul", what);
            int sl = producer(N(ml).in[0]);
            if (sl < 0 || N(sl).op != "Slice") sl = need_prod(N(ml).in[1], "Slice", what);
            Rope r = verify_rope(p, N(sl).in[0], dh);
            if (rope) *rope = r;
            v = N(sl).in[0];
            p = producer(v);
        }
        if (p < 0 || N(p).op != "Transpose") fail(std::string(what) + ": expected the head-split Transpose", v);
        check_perm(p, {0, 2, 1, 3}, what);
        int r = need_prod(N(p).in[0], "Reshape", what);
        if (!ints_are(N(r).in[1], {0, -1, heads, dh})) fail(std::string(what) + ": head split is not [0, -1, heads, dh]", v);
        int a = need_prod(N(r).in[0], "Add", what);
        const std::string& mmv = is_init(N(a).in[0]) ? N(a).in[1] : N(a).in[0];
        return need_prod(mmv, "MatMul", what);
    }
    json attention(std::string& x) {
        const std::string x0 = x;
        int t0 = consumer(x, "Transpose");
        check_perm(t0, {0, 2, 1}, "attention input");
        LN ln = layernorm(N(t0).out[0], "attention pre-LayerNorm");
        // q: the projection of LN(x) whose split (and rotated) output is the FIRST operand of a MatMul with a transposed second operand
        int score = -1, heads = 0, dh = 0; Lin q; Rope rq;
        for (int m : cons_.at(ln.out)) {
            if (N(m).op != "MatMul" || N(m).in[0] != ln.out || !is_init(N(m).in[1])) continue;
            Lin cand = linear_at(m, "attention projection");
            int h = 0, d = 0; Rope r;
            const std::string sp = split_heads(cand.out, h, d);
            std::string qs = through_rope(sp, r, d);
            int mm = consumer(qs, "MatMul", [&](const OnnxNode& n) { int pt = producer(n.in[1]); return n.in[0] == qs && pt >= 0 && N(pt).op == "Transpose"; });
            if (mm >= 0) { score = mm; q = cand; heads = h; dh = d; rq = r; break; }
        }
        if (score < 0) fail("attention: no Q.K^T product found", ln.out);
        const int C = q.N;
        int kt = producer(N(score).in[1]);
        check_perm(kt, {0, 1, 3, 2}, "attention K^T");
        Rope rk;
        int mk = back_to_projection(N(kt).in[0], "attention K", heads, dh, &rk);
        if (rk.kind != rq.kind || rk.freqs != rq.freqs) fail("attention: Q and K carry different rotary embeddings", N(kt).in[0]);
        Lin k = linear_at(mk, "attention K");
        const std::string ctx_src = N(mk).in[0];
        // scale, optional key mask, softmax, P.V
        int sc = need(N(score).out[0], "Mul", "attention scale");
        const float scale = scalar(other(N(sc), N(score).out[0]));
        if (std::fabs(scale - 1.0f / std::sqrt((float)dh)) > 1e-6f) fail("attention scale is not 1/sqrt(dh)", N(score).out[0]);
        std::string s = N(sc).out[0];
        bool key_masked = false;
        int mb = consumer(s, "Add");
        if (mb >= 0) {
            // the additive key mask: Unsqueeze(Mul(Sub(mask, 1), big)) — masked keys get -big, in this operand order
            int un = need_prod(other(N(mb), s), "Unsqueeze", "attention key mask");
            if (!axes_are(N(un), {1})) fail("attention key mask: Unsqueeze axes != [1]", s);
            int ml = need_prod(N(un).in[0], "Mul", "attention key mask");
            const std::string& big = is_init(N(ml).in[0]) ? N(ml).in[0] : N(ml).in[1];
            int sb = need_prod(other(N(ml), big), "Sub", "attention key mask");
            if (scalar(big) < 1e4f || N(sb).in[0] != "text_mask" || !is_init(N(sb).in[1]) || std::fabs(scalar(N(sb).in[1]) - 1.0f) > 1e-6f)
                fail("attention key mask is not (mask - 1) * big", N(mb).out[0]);
            key_masked = true; s = N(mb).out[0];
        }
        int sm = need(s, "Softmax", "attention");
        int pv = need(N(sm).out[0], "MatMul", "attention P.V");
        int mv = back_to_projection(N(pv).in[1], "attention V", heads, dh);
        Lin v = linear_at(mv, "attention V");
        if (N(mv).in[0] != ctx_src) fail("attention K and V read different contexts", N(mv).in[0]);
        int t1 = needT(N(pv).out[0], {0, 2, 1, 3}, "attention merge");
        int r1 = need(N(t1).out[0], "Reshape", "attention merge");
        if (!ints_are(N(r1).in[1], {0, -1, (int64_t)heads * dh})) fail("attention merge is not [0, -1, heads * dh]", N(t1).out[0]);
        Lin o = linear(N(r1).out[0], "attention output projection");
        int t2 = needT(o.out, {0, 2, 1}, "attention");
        int ad = consumer(N(t2).out[0], "Add", [&](const OnnxNode& n) { return other(n, N(t2).out[0]) == x0; });
        if (ad < 0) fail("attention residual Add", N(t2).out[0]);
        std::string y = N(ad).out[0];
        const bool masked = masked_tail(y);
        // context kind
        std::string ctx;
        std::set<std::string> stops{x0};
        if (ctx_src == ln.out) ctx = "self";
        else if (is_input(ctx_src)) { ctx = ctx_src; }
        else {
            int tp = producer(ctx_src);
            if (tp >= 0 && N(tp).op == "Transpose" && is_input(N(tp).in[0])) { check_perm(tp, {0, 2, 1}, "attention context"); ctx = N(tp).in[0]; }
            else fail("attention context is neither LN(x), a graph input nor its transpose", ctx_src);
        }
        if (ctx != "self" && ctx != "text_emb" && ctx != "style_ttl") fail("attention context input '" + ctx + "' is not text_emb / style_ttl", ctx_src);
        if (heads * dh != C || q.K != C || o.K != C || o.N != C || k.N != C || v.N != C || k.K != v.K) fail("attention projection shapes", x0);
        mark_cone(y, stops);
        x = y;
        json t = {{"ln_g", ln.g}, {"ln_b", ln.b}, {"wq", q.w}, {"bq", q.b}, {"wk", k.w}, {"bk", k.b}, {"wv", v.w}, {"bv", v.b}, {"wo", o.w}, {"bo", o.b}};
        if (rq.kind != "none") t["rope_freqs"] = rq.freqs;
        return json{{"type", "attention"}, {"name", q.w}, {"C", C}, {"heads", heads}, {"ctx", ctx}, {"ctx_dim", k.K}, {"rope", rq.kind},
                    {"masked", masked}, {"key_masked", key_masked}, {"t", t}};
    }

    // ---- embedding front (DP / TE): Gather(emb, text_ids) -> Transpose -> Mul(mask)
    std::string embed_front(std::string& emb_name, int& V, int& C) {
        int g = -1;
        for (size_t i = 0; i < f_.nodes.size(); ++i) if (N((int)i).op == "Gather" && N((int)i).in.size() == 2 && N((int)i).in[1] == "text_ids") g = (int)i;
        if (g < 0) throw PlanError("no Gather(embedding, text_ids) found");
        emb_name = N(g).in[0];
        const OnnxTensor& e = init(emb_name);
        if (e.dims.size() != 2) fail("embedding table rank", emb_name);
        V = (int)e.dims[0]; C = (int)e.dims[1];
        int t = needT(N(g).out[0], {0, 2, 1}, "embedding");
        std::string x = N(t).out[0];
        if (!masked_tail(x)) fail("embedding is not masked", x);
        mark_cone(x, {});
        return x;
    }
    // the common trunk: ConvNeXt / attention / time_cond layers while they match
    void trunk(std::string& x, json& layers, const std::string& time_emb = "") {
        for (;;) {
            if (at_convnext(x)) layers.push_back(convnext(x));
            else if (at_attention(x)) layers.push_back(attention(x));
            else if (!time_emb.empty() && at_time_cond(x, time_emb)) layers.push_back(time_cond(x, time_emb));
            else return;
        }
    }
    bool at_time_cond(const std::string& x, const std::string& te) const {
        return consumer(x, "Add", [&](const OnnxNode& n) {
                   int u = producer(other(n, x));
                   if (u < 0 || N(u).op != "Unsqueeze") return false;
                   int a = producer(N(u).in[0]);
                   if (a < 0 || N(a).op != "Add") return false;
                   const std::string& mmv = is_init(N(a).in[0]) ? N(a).in[1] : N(a).in[0];
                   int m = producer(mmv);
                   return m >= 0 && N(m).op == "MatMul" && N(m).in[0] == te;
               }) >= 0;
    }
    json time_cond(std::string& x, const std::string& te) {
        const std::string x0 = x;
        int ad = consumer(x, "Add", [&](const OnnxNode& n) { int u = producer(other(n, x)); return u >= 0 && N(u).op == "Unsqueeze"; });
        int u = producer(other(N(ad), x));
        int a = producer(N(u).in[0]);
        const std::string& mmv = is_init(N(a).in[0]) ? N(a).in[1] : N(a).in[0];
        Lin l = linear_at(producer(mmv), "time conditioning");
        std::string y = N(ad).out[0];
        if (!masked_tail(y)) fail("time conditioning without the mask multiply", y);
        mark_cone(y, {x0, te});
        x = y;
        return json{{"type", "time_cond"}, {"name", l.w}, {"C", l.N}, {"t", {{"w", l.w}, {"b", l.b}}}};
    }

    // ---- the four graphs
    json duration_predictor() {
        json arch, layers = json::array();
        std::string emb; int V = 0, C = 0;
        std::string x = embed_front(emb, V, C);
        // style: Reshape(style_dp) -> MatMul + Add -> Unsqueeze; x = Mul(Add(x, .), mask)
        int ad = consumer(x, "Add", [&](const OnnxNode& n) { int u = producer(other(n, x)); return u >= 0 && N(u).op == "Unsqueeze"; });
        if (ad < 0) fail("duration predictor: style add", x);
        int u = producer(other(N(ad), x));
        int a = need_prod(N(u).in[0], "Add", "style projection");
        const std::string& mmv = is_init(N(a).in[0]) ? N(a).in[1] : N(a).in[0];
        int mm = need_prod(mmv, "MatMul", "style projection");
        Lin st = linear_at(mm, "style projection");
        int rs = need_prod(N(mm).in[0], "Reshape", "style projection input");
        if (N(rs).in[0] != "style_dp") fail("style projection does not read style_dp", N(rs).in[0]);
        std::string y = N(ad).out[0];
        if (!masked_tail(y)) fail("style add without the mask multiply", y);
        mark_cone(y, {x});
        x = y;
        trunk(x, layers);
        // head: Transpose -> LN -> MatMul + Add -> Clip -> Exp -> Mul(sec per token) -> Transpose -> Mul(mask) -> ReduceSum
        int t = needT(x, {0, 2, 1}, "duration head");
        LN ln = layernorm(N(t).out[0], "duration head");
        Lin pj = linear(ln.out, "duration head projection");
        if (pj.N != 1) fail("duration head projection must map to one value per token", pj.w);
        int cl = need(pj.out, "Clip", "duration head");
        const float lo = scalar(N(cl).in[1]), hi = scalar(N(cl).in[2]);
        if (std::fabs(lo + hi) > 1e-6f) fail("duration head Clip is not symmetric", pj.out);
        int ex = need(N(cl).out[0], "Exp", "duration head");
        int ms = need(N(ex).out[0], "Mul", "duration head");
        const float spt = scalar(other(N(ms), N(ex).out[0]));
        int t2 = needT(N(ms).out[0], {0, 2, 1}, "duration head");
        std::string d = N(t2).out[0];
        if (!masked_tail(d)) fail("duration head without the mask multiply", d);
        int rsum = need(d, "ReduceSum", "duration head");
        if (!axes_are(N(rsum), {1, 2})) fail("duration head: ReduceSum over axes other than [1, 2]", d);
        std::string out = N(rsum).out[0];
        int idn = consumer(out, "Identity"); if (idn >= 0) out = N(idn).out[0];
        mark_cone(out, {x});
        arch["C"] = C; arch["vocab"] = V; arch["style_in"] = st.K; arch["clip"] = hi; arch["sec_per_token"] = spt;
        arch["H"] = layers.empty() ? 0 : layers[0].value("H", 0); arch["K"] = layers.empty() ? 0 : layers[0].value("K", 0);
        arch["layers"] = layers;
        arch["t"] = {{"embed", emb}, {"style_w", st.w}, {"style_b", st.b}, {"head_ln_g", ln.g}, {"head_ln_b", ln.b}, {"head_w", pj.w}, {"head_b", pj.b}};
        return arch;
    }

    json text_encoder() {
        json arch, layers = json::array();
        std::string emb; int V = 0, C = 0;
        std::string x = embed_front(emb, V, C);
        trunk(x, layers);
        // proj_out: Transpose -> MatMul + Add -> Transpose -> Mul(mask)
        const std::string x0 = x;
        int t = needT(x, {0, 2, 1}, "text encoder output projection");
        Lin pj = linear(N(t).out[0], "text encoder output projection");
        int t2 = needT(pj.out, {0, 2, 1}, "text encoder output projection");
        std::string y = N(t2).out[0];
        if (!masked_tail(y)) fail("text_emb is not masked", y);
        int idn = consumer(y, "Identity"); if (idn >= 0) y = N(idn).out[0];
        mark_cone(y, {x0});
        layers.push_back(json{{"type", "proj_out"}, {"name", pj.w}, {"cin", pj.K}, {"cout", pj.N}, {"t", {{"w", pj.w}, {"b", pj.b}}}});
        int heads = 0, H = 0, K = 0, S = 0, Cs = 0;
        for (const auto& l : layers) {
            if (l["type"] == "attention") { heads = l["heads"]; if (l["ctx"] == "style_ttl") Cs = l["ctx_dim"]; }
            if (l["type"] == "convnext") { H = l["H"]; K = l["K"]; }
        }
        if (const OnnxValueInfo* vi = f_.input("style_ttl")) { if (vi->dims.size() == 3) { S = (int)vi->dims[1]; if (!Cs) Cs = (int)vi->dims[2]; } }
        arch["C"] = C; arch["H"] = H; arch["K"] = K; arch["heads"] = heads; arch["vocab"] = V; arch["n_style"] = S; arch["style_dim"] = Cs;
        arch["layers"] = layers;
        arch["t"] = {{"embed", emb}};
        return arch;
    }

    json vector_estimator() {
        json arch, layers = json::array();
        // time embedding: Div(current_step, total_step) -> Unsqueeze -> Mul(freqs) -> Sin | Cos -> Concat -> fc1 -> GELU -> fc2
        int dv = consumer("current_step", "Div", [&](const OnnxNode& n) { return n.in[0] == "current_step" && n.in[1] == "total_step"; });
        if (dv < 0) throw PlanError("vector_estimator: no Div(current_step, total_step)");
        int un = need(N(dv).out[0], "Unsqueeze", "time embedding");
        int mf = need(N(un).out[0], "Mul", "time embedding");
        const std::string freqs = other(N(mf), N(un).out[0]);
        const int half = (int)init(freqs).numel();
        int sn = need(N(mf).out[0], "Sin", "time embedding");
        need(N(mf).out[0], "Cos", "time embedding");
        int cc = need(N(sn).out[0], "Concat", "time embedding");
        if (N(cc).in[0] != N(sn).out[0]) fail("time embedding must be Concat(sin, cos)", N(cc).out[0]);
        Lin f1 = linear(N(cc).out[0], "time MLP fc1");
        Lin f2 = linear(gelu(f1.out), "time MLP fc2");
        const std::string te = f2.out;
        mark_cone(te, {});
        layers.push_back(json{{"type", "time_mlp"}, {"name", f1.w}, {"time_dim", 2 * half}, {"C", f2.N},
                              {"t", {{"freqs", freqs}, {"w1", f1.w}, {"b1", f1.b}, {"w2", f2.w}, {"b2", f2.b}}}});
        // proj_in: Transpose(noisy_latent) -> MatMul + Add -> Transpose -> Mul(latent_mask)
        int t = consumer("noisy_latent", "Transpose");
        if (t < 0) throw PlanError("vector_estimator: noisy_latent is not transposed into the input projection");
        check_perm(t, {0, 2, 1}, "input projection");
        Lin pi = linear(N(t).out[0], "input projection");
        int t2 = needT(pi.out, {0, 2, 1}, "input projection");
        std::string x = N(t2).out[0];
        if (!masked_tail(x)) fail("input projection is not masked", x);
        mark_cone(x, {});
        layers.push_back(json{{"type", "proj_in"}, {"name", pi.w}, {"cin", pi.K}, {"cout", pi.N}, {"t", {{"w", pi.w}, {"b", pi.b}}}});
        trunk(x, layers, te);
        // proj_out + in-graph Euler update: y = (noisy_latent + Transpose(linear(Transpose(x))) * (1 / total_step)) * latent_mask
        const std::string x0 = x;
        int t3 = needT(x, {0, 2, 1}, "output projection");
        Lin po = linear(N(t3).out[0], "output projection");
        int t4 = needT(po.out, {0, 2, 1}, "output projection");
        int ms = need(N(t4).out[0], "Mul", "Euler update");
        {   // dt = Unsqueeze(Div(1, total_step))
            int ud = need_prod(other(N(ms), N(t4).out[0]), "Unsqueeze", "Euler step size");
            int dd = need_prod(N(ud).in[0], "Div", "Euler step size");
            if (N(dd).in[1] != "total_step" || std::fabs(scalar(N(dd).in[0]) - 1.0f) > 1e-6f) fail("Euler step size is not 1 / total_step", N(dd).out[0]);
        }
        int ad = consumer(N(ms).out[0], "Add", [&](const OnnxNode& n) { return other(n, N(ms).out[0]) == "noisy_latent"; });
        if (ad < 0) fail("Euler update: no Add(noisy_latent, v * dt)", N(ms).out[0]);
        std::string y = N(ad).out[0];
        if (!masked_tail(y)) fail("Euler update is not masked", y);
        int idn = consumer(y, "Identity"); if (idn >= 0) y = N(idn).out[0];
        mark_cone(y, {x0});
        layers.push_back(json{{"type", "proj_out"}, {"name", po.w}, {"cin", po.K}, {"cout", po.N}, {"t", {{"w", po.w}, {"b", po.b}}}});
        int heads = 0, H = 0, K = 0, S = 0, Cs = 0, Ct = 0;
        for (const auto& l : layers) {
            if (l["type"] == "attention") { heads = l["heads"]; if (l["ctx"] == "style_ttl") Cs = l["ctx_dim"]; if (l["ctx"] == "text_emb") Ct = l["ctx_dim"]; }
            if (l["type"] == "convnext") { H = l["H"]; K = l["K"]; }
        }
        if (const OnnxValueInfo* vi = f_.input("style_ttl")) if (vi->dims.size() == 3) S = (int)vi->dims[1];
        arch["C"] = pi.N; arch["H"] = H; arch["K"] = K; arch["heads"] = heads; arch["latent_ch"] = pi.K; arch["time_dim"] = 2 * half;
        arch["text_dim"] = Ct; arch["n_style"] = S; arch["style_dim"] = Cs;
        arch["layers"] = layers;
        return arch;
    }

    json vocoder() {
        json arch, layers = json::array();
        // de-normalise + un-compress: Mul(latent, std) -> Add(mean) -> Reshape[0,f,ld,-1] -> Transpose(0,2,3,1) -> Reshape[0,ld,-1]
        int ms = consumer("latent", "Mul", [&](const OnnxNode& n) { return is_init(other(n, "latent")); });
        if (ms < 0) throw PlanError("vocoder: latent is not de-normalised by Mul(std)");
        const std::string stdn = other(N(ms), "latent");
        int ad = consumer(N(ms).out[0], "Add", [&](const OnnxNode& n) { return is_init(other(n, N(ms).out[0])); });
        if (ad < 0) fail("vocoder: no Add(mean)", N(ms).out[0]);
        const std::string mean = other(N(ad), N(ms).out[0]);
        int r1 = need(N(ad).out[0], "Reshape", "vocoder un-compress");
        auto s1 = ints(N(r1).in[1]);
        if (s1.size() != 4) fail("vocoder un-compress shape", N(ad).out[0]);
        const int f = (int)s1[1], ld = (int)s1[2];
        int tp = needT(N(r1).out[0], {0, 2, 3, 1}, "vocoder un-compress");
        int r2 = need(N(tp).out[0], "Reshape", "vocoder un-compress");
        // conv_in: Conv(group = 1, causal) -> BatchNormalization
        int cv = need(N(r2).out[0], "Conv", "vocoder input convolution");
        const OnnxNode& c = N(cv);
        const OnnxTensor& w = init(c.in[1]);
        if (c.attr_i("group", 1) != 1 || w.dims.size() != 3 || w.dims[1] != ld) fail("vocoder input convolution shape", c.in[1]);
        const int C = (int)w.dims[0], K = (int)w.dims[2];
        auto it = c.attr.find("pads");
        if (it == c.attr.end() || it->second.ints.size() != 2 || it->second.ints[0] != K - 1 || it->second.ints[1] != 0) fail("vocoder input convolution must be causal", c.in[1]);
        for (const char* k : {"dilations", "strides"}) { auto a = c.attr.find(k); if (a != c.attr.end()) for (int64_t d : a->second.ints) if (d != 1) fail(std::string("vocoder input convolution: ") + k + " != 1", c.in[1]); }
        int bn = need(c.out[0], "BatchNormalization", "vocoder input normalisation");
        if (std::fabs(attr_f(N(bn), "epsilon", 1e-5f) - 1e-5f) > 1e-9f) fail("vocoder input normalisation: BatchNormalization epsilon is not 1e-5", c.out[0]);
        std::string x = N(bn).out[0];
        mark_cone(x, {});
        layers.push_back(json{{"type", "conv_in"}, {"name", c.in[1]}, {"cin", ld}, {"cout", C}, {"K", K}, {"causal", true}, {"bn", N(bn).in[1]},
                              {"t", {{"w", c.in[1]}, {"b", c.in[2]}, {"bn_w", N(bn).in[1]}, {"bn_b", N(bn).in[2]}, {"bn_mean", N(bn).in[3]}, {"bn_var", N(bn).in[4]}}}});
        trunk(x, layers);
        const std::string x0 = x;
        int t = needT(x, {0, 2, 1}, "vocoder head");
        LN ln = layernorm(N(t).out[0], "vocoder head");
        Lin pj = linear(ln.out, "vocoder head projection");
        int rs = need(pj.out, "Reshape", "vocoder head");
        std::string y = N(rs).out[0];
        int idn = consumer(y, "Identity"); if (idn >= 0) y = N(idn).out[0];
        mark_cone(y, {x0});
        layers.push_back(json{{"type", "head"}, {"name", pj.w}, {"cin", pj.K}, {"cout", pj.N}, {"t", {{"ln_g", ln.g}, {"ln_b", ln.b}, {"w", pj.w}, {"b", pj.b}}}});
        int H = 0, Kc = 0;
        for (const auto& l : layers) if (l["type"] == "convnext") { H = l["H"]; Kc = l["K"]; }
        arch["C"] = C; arch["H"] = H; arch["K"] = Kc; arch["latent_ch"] = f * ld; arch["latent_dim"] = ld; arch["compress"] = f; arch["hop"] = pj.N;
        arch["layers"] = layers;
        arch["t"] = {{"latent_std", stdn}, {"latent_mean", mean}};
        return arch;
    }
};

inline nlohmann::json derive_arch(const OnnxFile& f, const std::string& kind) { return GraphPlanner(f).derive(kind); }

}  // namespace stc
