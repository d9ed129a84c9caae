"""The node-pattern matcher (supertonic_b200/csrc/graph_plan.h): the layer plan the library derives from a graph's ONNX nodes — what it
uses for graphs without `stc_arch` metadata, i.e. any export the reference's loader would accept (loadOnnx, cpp/helper.cpp:776-795).

CPU part (stc_derive_arch, no GPU): for every surrogate graph the derived plan equals the plan the generator wrote down when it built
the graph; initializer names come from the node inputs (a renamed export loads); a graph with an unknown sub-graph is rejected with
the list of unexplained nodes. GPU part: an engine built from node-derived plans computes bit-identical results."""
import copy
import json
import os

import numpy as np
import pytest

from tests import _util as U

KINDS = ("duration_predictor", "text_encoder", "vector_estimator", "vocoder")
FIELDS = ("type", "C", "H", "K", "dilation", "causal", "masked", "heads", "ctx", "ctx_dim", "rope", "key_masked", "cin", "cout", "time_dim")


def _assets(size):
    from supertonic_b200 import surrogate
    return surrogate.ensure_assets(size)


@pytest.mark.parametrize("size", ["tiny", "full"])
@pytest.mark.parametrize("kind", KINDS)
def test_derived_plan_equals_the_generators_plan(size, kind):
    from supertonic_b200 import capi, onnx_lite as ol
    path = os.path.join(_assets(size), "onnx", kind + ".onnx")
    got = capi.derive_arch(path, kind)
    model = ol.load_model(path)
    want = json.loads(model.metadata["stc_arch"])
    assert got["derived_from"] == "nodes"
    assert len(got["layers"]) == len(want["layers"])
    for a, b in zip(got["layers"], want["layers"]):
        for k in FIELDS:
            if k in b:
                assert a.get(k) == b[k], (b.get("name"), k, a.get(k), b[k])
        for role, name in a["t"].items():                      # every recorded tensor is an initializer of the file
            assert name in model.graph.initializers, (role, name)
    for k, v in want.items():
        if k in ("layers", "surrogate_version", "kind"):
            continue
        if isinstance(v, float):
            assert np.float32(got[k]) == np.float32(v), k
        else:
            assert got[k] == v, k


def test_renamed_initializers_and_stripped_metadata_still_load(tmp_path):
    """A released export names its initializers arbitrarily (onnx::MatMul_123 ...) and carries no private metadata: rename every
    initializer, drop the metadata, and the derived plan must point at the renamed tensors in the right roles."""
    from supertonic_b200 import capi, onnx_lite as ol
    src = os.path.join(_assets("tiny"), "onnx", "vector_estimator.onnx")
    ref = capi.derive_arch(src, "vector_estimator")
    m = ol.load_model(src)
    ren = {name: f"onnx::T_{i}" for i, name in enumerate(m.graph.initializers)}
    m.graph.initializers = {ren[k]: v for k, v in m.graph.initializers.items()}
    for n in m.graph.nodes:
        n.inputs = [ren.get(x, x) for x in n.inputs]
    m.metadata = {}
    out = str(tmp_path / "ve_renamed.onnx")
    ol.save_model(m, out)
    got = capi.derive_arch(out, "vector_estimator")
    assert [l["type"] for l in got["layers"]] == [l["type"] for l in ref["layers"]]
    for a, b in zip(got["layers"], ref["layers"]):
        assert {k: ren[v] for k, v in b["t"].items()} == a["t"]


def test_unknown_subgraph_is_rejected_with_the_unexplained_nodes(tmp_path):
    from supertonic_b200 import capi, onnx_lite as ol
    src = os.path.join(_assets("tiny"), "onnx", "text_encoder.onnx")
    # (1) a GELU that is not the exact-erf form
    m = ol.load_model(src)
    erf = next(n for n in m.graph.nodes if n.op_type == "Erf")
    erf.op_type = "Tanh"
    p1 = str(tmp_path / "te_tanh.onnx"); ol.save_model(m, p1)
    with pytest.raises(capi.StcError) as e:
        capi.derive_arch(p1, "text_encoder")
    assert e.value.code == -4 and "GELU" in str(e.value)
    # (2) an extra operator on the main path that belongs to no layer
    m = ol.load_model(src)
    last = next(n for n in m.graph.nodes if n.outputs == ["text_emb"])             # Identity -> text_emb
    extra = copy.deepcopy(last)
    extra.op_type, extra.name = "Relu", "mystery_relu"
    extra.inputs, extra.outputs = [last.inputs[0]], ["/mystery"]
    last.inputs = ["/mystery"]
    m.graph.nodes.insert(m.graph.nodes.index(last), extra)
    p2 = str(tmp_path / "te_relu.onnx"); ol.save_model(m, p2)
    with pytest.raises(capi.StcError) as e:
        capi.derive_arch(p2, "text_encoder")
    assert e.value.code == -4 and "mystery_relu" in str(e.value)
    # (3) wrong graph for the requested kind
    with pytest.raises(capi.StcError):
        capi.derive_arch(src, "vocoder")


@pytest.mark.gpu
@pytest.mark.parametrize("size", ["tiny", "full"])
def test_engine_from_node_derived_plans_is_bit_identical(size):
    from supertonic_b200 import capi
    root = _assets(size)
    texts, langs = U.make_batch(5, 4, 20, 200)
    a = capi.Engine(root + "/onnx")
    os.environ["STC_IGNORE_ARCH"] = "1"
    try:
        b = capi.Engine(root + "/onnx")
    finally:
        del os.environ["STC_IGNORE_ARCH"]
    try:
        ids, mask = a.text_to_ids(texts, langs)
        ttl, dp = U.styles(root, ["M1", "F1", "M2", "F2"])
        np.testing.assert_array_equal(a.duration(ids, dp, mask), b.duration(ids, dp, mask))
        np.testing.assert_array_equal(a.text_encode(ids, ttl, mask), b.text_encode(ids, ttl, mask))
        ra = a.synthesize_packed(ids, mask, ttl, dp, 3, 1.05, seed=3, want_latent=True)
        rb = b.synthesize_packed(ids, mask, ttl, dp, 3, 1.05, seed=3, want_latent=True)
        for k in range(4):
            np.testing.assert_array_equal(ra["latent"][k], rb["latent"][k])
            np.testing.assert_array_equal(ra["wavs"][k], rb["wavs"][k])
    finally:
        a.close(); b.close()


def test_corrupted_graph_files_are_rejected_not_crashed_on(tmp_path):
    """The ONNX reader and the node-pattern matcher on damaged files (truncated, bytes overwritten anywhere or in the first 4 KB of
    headers / varints / lengths, spans deleted) of all four graphs: every file either yields a plan (the damage hit a weight payload) or is
    refused with an StcError whose message survives non-UTF-8 bytes from the file. Runs in a child process: a crash of the native reader
    would be a failed test, not a dead test session."""
    import subprocess, sys, textwrap
    code = textwrap.dedent('''
        import os, sys
        import numpy as np
        sys.path.insert(0, %r)
        from supertonic_b200 import capi, surrogate
        root, td = surrogate.ensure_assets("tiny"), %r
        rng = np.random.default_rng(7)
        ok = rejected = 0
        for kind in ("duration_predictor", "text_encoder", "vector_estimator", "vocoder"):
            raw = open(os.path.join(root, "onnx", kind + ".onnx"), "rb").read()
            for i in range(60):
                b = bytearray(raw)
                if i %% 4 == 0:
                    b = b[:int(rng.integers(0, len(b)))]
                elif i %% 4 == 1:
                    for _ in range(int(rng.integers(1, 8))):
                        b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
                elif i %% 4 == 2:
                    for _ in range(int(rng.integers(1, 6))):
                        b[int(rng.integers(0, min(len(b), 4096)))] = int(rng.integers(0, 256))
                else:
                    p = int(rng.integers(0, len(b)))
                    del b[p:min(len(b), p + int(rng.integers(1, 64)))]
                path = os.path.join(td, "g.onnx")
                open(path, "wb").write(bytes(b))
                try:
                    capi.derive_arch(path, kind); ok += 1
                except capi.StcError as e:
                    assert str(e); rejected += 1
        print("ok", ok, "rejected", rejected)
    ''') % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    _, ok, _, rejected = r.stdout.split()
    assert int(ok) + int(rejected) == 240 and int(rejected) >= 100, r.stdout


@pytest.mark.parametrize("kind", KINDS)
def test_reversed_operands_of_non_commutative_nodes_are_rejected(tmp_path, kind):
    """x / sqrt2 is not sqrt2 / x: every Sub, Div and MatMul of every graph with its two operands swapped (GELU's Div, the rotary
    Sub(t1 cos, t2 sin), cumsum(mask) - 1 and its division by the length on the Q and the K side, the additive key mask (mask - 1) * big,
    1 / total_step, current_step / total_step, every projection) must make the matcher refuse the graph — it computes something else."""
    import copy
    from supertonic_b200 import capi, onnx_lite as ol
    base = ol.load_model(os.path.join(_assets("tiny"), "onnx", kind + ".onnx"))
    base.metadata = {}
    n = 0
    for k, nd in enumerate(base.graph.nodes):
        if nd.op_type not in ("Sub", "Div", "MatMul") or len(nd.inputs) != 2:
            continue
        m = copy.deepcopy(base)
        m.graph.nodes[k].inputs = list(reversed(nd.inputs))
        path = str(tmp_path / "rev.onnx")
        ol.save_model(m, path)
        with pytest.raises(capi.StcError):
            capi.derive_arch(path, kind)
        n += 1
    assert n >= 7


@pytest.mark.parametrize("kind", KINDS)
def test_every_node_matters(tmp_path, kind):
    """Remove any single node — a rotary Slice, one of the four rotary products, the additive key mask, a Transpose — and the matcher
    must refuse the graph: no node of a recognised layer is taken on trust. (The one removal that changes nothing is the Identity that
    gives the graph output its name.)"""
    import copy
    from supertonic_b200 import capi, onnx_lite as ol
    base = ol.load_model(os.path.join(_assets("tiny"), "onnx", kind + ".onnx"))
    base.metadata = {}
    accepted = []
    for k, nd in enumerate(base.graph.nodes):
        m = copy.deepcopy(base)
        del m.graph.nodes[k]
        path = str(tmp_path / "del.onnx")
        ol.save_model(m, path)
        try:
            capi.derive_arch(path, kind)
            accepted.append(nd.op_type)
        except capi.StcError:
            pass
    assert accepted in ([], ["Identity"]), accepted


def test_attribute_values_the_kernels_hard_code_are_checked(tmp_path):
    """LayerNormalization epsilon (kernels: 1e-6; the ONNX default is 1e-5), BatchNormalization epsilon (1e-5, folded at load), Transpose
    perms (fixed layouts), dilation of the vocoder's dense input convolution: a graph that differs is refused."""
    import copy
    from supertonic_b200 import capi, onnx_lite as ol

    def mutated(kind, pick, change):
        base = ol.load_model(os.path.join(_assets("tiny"), "onnx", kind + ".onnx"))
        base.metadata = {}
        hits = [n for n in base.graph.nodes if pick(n)]
        assert hits
        for i in range(len(hits)):
            m = copy.deepcopy(base)
            change([n for n in m.graph.nodes if pick(n)][i])
            path = str(tmp_path / "attr.onnx")
            ol.save_model(m, path)
            with pytest.raises(capi.StcError):
                capi.derive_arch(path, kind)
    for kind in KINDS:
        mutated(kind, lambda n: n.op_type == "LayerNormalization", lambda n: n.attrs.update(epsilon=1e-5))
        mutated(kind, lambda n: n.op_type == "LayerNormalization", lambda n: n.attrs.pop("epsilon"))
        mutated(kind, lambda n: n.op_type == "Transpose", lambda n: n.attrs.update(perm=list(range(len(n.attrs["perm"])))))
    mutated("vocoder", lambda n: n.op_type == "BatchNormalization", lambda n: n.attrs.update(epsilon=1e-3))
    mutated("vocoder", lambda n: n.op_type == "Conv" and n.attrs.get("group", 1) == 1, lambda n: n.attrs.update(dilations=[2]))


def test_a_stage_masked_by_the_wrong_mask_is_rejected(tmp_path):
    """The plan only says "masked"; the engine then multiplies by the latent mask in the vector estimator and by the text mask on the
    text side. A graph that masks a latent-side stage with text_mask (or the attention keys with latent_mask) is not that graph."""
    import copy
    from supertonic_b200 import capi, onnx_lite as ol
    base = ol.load_model(os.path.join(_assets("tiny"), "onnx", "vector_estimator.onnx"))
    base.metadata = {}
    swap = {"latent_mask": "text_mask", "text_mask": "latent_mask"}
    hits = [k for k, n in enumerate(base.graph.nodes) if n.op_type in ("Mul", "Sub") and any(i in swap for i in n.inputs)]
    assert len(hits) >= 6
    for k in hits:
        m = copy.deepcopy(base)
        m.graph.nodes[k].inputs = [swap.get(i, i) for i in m.graph.nodes[k].inputs]
        path = str(tmp_path / "mask.onnx")
        ol.save_model(m, path)
        with pytest.raises(capi.StcError):
            capi.derive_arch(path, "vector_estimator")


@pytest.mark.parametrize("kind", KINDS)
def test_small_constants_are_verified_or_carried_into_the_plan(tmp_path, kind):
    """Every small constant of a graph (scalars, axes, shapes, slice bounds: <= 8 elements) changed by x 1.5 / + 1: the matcher must either
    refuse the graph, or return a DIFFERENT plan (the value is carried: clip, seconds per token, heads, ...). What may pass unchanged is
    only what cannot change the result: the end bound of the upper rotary Slice (clamped to the dimension), the size of the key-mask
    constant (any large value masks), single-element weights (read by name at load), and broadcasting plumbing (Unsqueeze axes / Reshape
    shapes whose other values make an invalid graph)."""
    import copy
    from supertonic_b200 import capi, onnx_lite as ol
    base = ol.load_model(os.path.join(_assets("tiny"), "onnx", kind + ".onnx"))
    base.metadata = {}
    p0, p1 = str(tmp_path / "c0.onnx"), str(tmp_path / "c1.onnx")
    ol.save_model(base, p0)
    ref = capi.derive_arch(p0, kind)
    used = {i for n in base.graph.nodes for i in n.inputs}
    checked = 0
    for name, a in base.graph.initializers.items():
        if a.size > 8 or name not in used:
            continue
        m = copy.deepcopy(base)
        m.graph.initializers[name] = ((a * 1.5 + (0.25 if np.all(a == 0) else 0)) if np.issubdtype(a.dtype, np.floating) else a + 1).astype(a.dtype)
        ol.save_model(m, p1)
        try:
            same = capi.derive_arch(p1, kind) == ref
        except capi.StcError:
            same = False
        checked += 1
        if same:
            ops = {n.op_type for n in base.graph.nodes if name in n.inputs}
            assert ops <= {"Slice", "Mul", "Add", "Unsqueeze", "Reshape"}, (name, ops)
            if "Mul" in ops:
                assert a.size == 1 and float(a.reshape(-1)[0]) >= 1e4, name           # the key-mask constant
            if "Slice" in ops:
                assert all(n.inputs.index(name) == 2 for n in base.graph.nodes if name in n.inputs), name   # `ends` only
    assert checked >= 9
