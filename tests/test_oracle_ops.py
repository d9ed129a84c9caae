"""Known-answer tests for the ONNX interpreter oracle (oracle/onnx_interp.py).

The reference holds no golden vectors for the Session::Run boundary (parity unpinned),
so every operator the surrogate graphs use is pinned here against hand-computed answers written out in
plain Python/numpy loops (the ONNX operator specification, opset 17), and the whole interpreter against
its own float64 shadow run.
"""
import math

import numpy as np
import pytest
import torch

from oracle import onnx_interp as oi
from supertonic_b200 import onnx_lite as ol


def T(a, dt=np.float32):
    return torch.from_numpy(np.asarray(a, dt))


def test_elementwise_and_broadcast():
    a, b = T([[1, 2, 3], [4, 5, 6]]), T([10, 20, 30])
    assert oi.OPS["Add"](a, b).tolist() == [[11, 22, 33], [14, 25, 36]]
    assert oi.OPS["Sub"](a, b).tolist() == [[-9, -18, -27], [-6, -15, -24]]
    assert oi.OPS["Mul"](a, b).tolist() == [[10, 40, 90], [40, 100, 180]]
    np.testing.assert_allclose(oi.OPS["Div"](a, b).numpy(), np.asarray([[.1, .1, .1], [.4, .25, .2]], np.float32))
    assert oi.OPS["Div"](T([7, -7], np.int64), T([2, 2], np.int64)).tolist() == [3, -3]      # integer Div truncates
    assert oi.OPS["Clip"](T([-5, 0.5, 5]), T(-3.0), T(3.0)).tolist() == [-3, 0.5, 3]
    for name, fn in (("Erf", math.erf), ("Exp", math.exp), ("Sin", math.sin), ("Cos", math.cos)):
        x = [-1.5, -0.1, 0.0, 0.7, 2.0]
        np.testing.assert_allclose(oi.OPS[name](T(x, np.float64)).numpy(), [fn(v) for v in x], rtol=1e-14, atol=1e-15)


def test_matmul_and_gemm_loops():
    rng = np.random.default_rng(0)
    a, b, c = rng.standard_normal((2, 3, 4)), rng.standard_normal((4, 5)), rng.standard_normal(5)
    want = np.zeros((2, 3, 5))
    for i in range(2):
        for m in range(3):
            for n in range(5):
                want[i, m, n] = sum(a[i, m, k] * b[k, n] for k in range(4))
    np.testing.assert_allclose(oi.OPS["MatMul"](T(a, np.float64), T(b, np.float64)).numpy(), want, rtol=1e-13)
    g = oi.OPS["Gemm"](T(a[0], np.float64), T(b.T.copy(), np.float64), T(c, np.float64), alpha=0.5, beta=2.0, transB=1)
    np.testing.assert_allclose(g.numpy(), 0.5 * want[0] + 2.0 * c, rtol=1e-13)


@pytest.mark.parametrize("K,dil,pads", [(5, 1, (2, 2)), (5, 4, (8, 8)), (7, 2, (12, 0)), (5, 8, (16, 16))])
def test_depthwise_conv1d_loops(K, dil, pads):
    """Conv(group=C) with explicit begin/end pads and dilation — same-pad and causal variants."""
    rng = np.random.default_rng(K * 10 + dil)
    C, N = 3, 21
    x, w, b = rng.standard_normal((2, C, N)), rng.standard_normal((C, 1, K)), rng.standard_normal(C)
    want = np.zeros((2, C, N + pads[0] + pads[1] - dil * (K - 1)))
    for bi in range(2):
        for c in range(C):
            for n in range(want.shape[2]):
                acc = b[c]
                for k in range(K):
                    src = n + k * dil - pads[0]
                    if 0 <= src < N:
                        acc += w[c, 0, k] * x[bi, c, src]
                want[bi, c, n] = acc
    got = oi.OPS["Conv"](T(x, np.float64), T(w, np.float64), T(b, np.float64), group=C, dilations=[dil], pads=list(pads),
                         strides=[1], kernel_shape=[K])
    assert got.shape[2] == N           # every conv in the graphs preserves the frame count
    np.testing.assert_allclose(got.numpy(), want, rtol=1e-12, atol=1e-13)


def test_dense_conv1d_loops():
    rng = np.random.default_rng(3)
    x, w, b = rng.standard_normal((1, 4, 9)), rng.standard_normal((6, 4, 3)), rng.standard_normal(6)
    want = np.zeros((1, 6, 9))
    for o in range(6):
        for n in range(9):
            want[0, o, n] = b[o] + sum(w[o, c, k] * x[0, c, n + k - 2] for c in range(4) for k in range(3) if n + k - 2 >= 0)
    got = oi.OPS["Conv"](T(x, np.float64), T(w, np.float64), T(b, np.float64), group=1, dilations=[1], pads=[2, 0], strides=[1])
    np.testing.assert_allclose(got.numpy(), want, rtol=1e-12)


def test_layernorm_batchnorm_softmax_loops():
    rng = np.random.default_rng(4)
    x, g, b = rng.standard_normal((2, 5, 8)), rng.standard_normal(8), rng.standard_normal(8)
    want = np.zeros_like(x)
    for i in range(2):
        for j in range(5):
            mu = x[i, j].sum() / 8
            var = ((x[i, j] - mu) ** 2).sum() / 8           # biased variance, as the spec says
            want[i, j] = (x[i, j] - mu) / math.sqrt(var + 1e-6) * g + b
    got = oi.OPS["LayerNormalization"](T(x, np.float64), T(g, np.float64), T(b, np.float64), axis=-1, epsilon=1e-6)
    np.testing.assert_allclose(got.numpy(), want, rtol=1e-12)
    xc = rng.standard_normal((2, 3, 7))
    sc, bi, mu, var = rng.standard_normal(3), rng.standard_normal(3), rng.standard_normal(3), rng.random(3) + 0.5
    wantb = (xc - mu[None, :, None]) / np.sqrt(var[None, :, None] + 1e-5) * sc[None, :, None] + bi[None, :, None]
    gotb = oi.OPS["BatchNormalization"](*(T(v, np.float64) for v in (xc, sc, bi, mu, var)), epsilon=1e-5)
    np.testing.assert_allclose(gotb.numpy(), wantb, rtol=1e-12)
    s = rng.standard_normal((2, 4))
    e = np.exp(s - s.max(1, keepdims=True))
    np.testing.assert_allclose(oi.OPS["Softmax"](T(s, np.float64), axis=-1).numpy(), e / e.sum(1, keepdims=True), rtol=1e-13)


def test_shape_ops():
    x = T(np.arange(24).reshape(2, 3, 4))
    assert oi.OPS["Transpose"](x, perm=[0, 2, 1]).shape == (2, 4, 3)
    assert oi.OPS["Reshape"](x, T([0, -1], np.int64)).shape == (2, 12)              # 0 copies the input dim
    assert oi.OPS["Unsqueeze"](x, T([1, 4], np.int64)).shape == (2, 1, 3, 4, 1)
    assert oi.OPS["Slice"](x, T([1], np.int64), T([3], np.int64), T([2], np.int64)).tolist() == x[:, :, 1:3].tolist()
    assert oi.OPS["Concat"](x, x, axis=1).shape == (2, 6, 4)
    emb = T(np.arange(10).reshape(5, 2))
    assert oi.OPS["Gather"](emb, T([[4, 0], [1, 1]], np.int64), axis=0).tolist() == [[[8, 9], [0, 1]], [[2, 3], [2, 3]]]
    m = T([[[1, 1, 1, 0, 0]]])
    assert oi.OPS["CumSum"](m, T(2, np.int64)).tolist() == [[[1, 2, 3, 3, 3]]]
    assert oi.OPS["ReduceSum"](m, T([2], np.int64), keepdims=1).tolist() == [[[3]]]
    assert oi.OPS["ReduceSum"](x, T([1, 2], np.int64), keepdims=0).tolist() == [66, 210]


def test_interpreter_runs_a_graph_and_float64_shadow_bounds_it():
    """A hand-built two-node graph through the wire-format codec, and the fp32-vs-fp64 gap of a full surrogate
    vector_estimator step (the oracle's own error bar, far below the 1e-3 parity bound)."""
    g = ol.Graph(name="t")
    g.inputs.append(ol.ValueInfo("x", ol.FLOAT, [1, 3]))
    g.initializers["w"] = np.asarray([[1, 2], [3, 4], [5, 6]], np.float32)
    g.nodes.append(ol.Node("MatMul", ["x", "w"], ["y0"], {}, name="mm"))
    g.nodes.append(ol.Node("Relu", ["y0"], ["y"], {}, name="r"))
    g.outputs.append(ol.ValueInfo("y", ol.FLOAT, [1, 2]))
    blob = ol.encode_model(ol.Model(g, metadata={"k": "v"}))
    it = oi.Interpreter(ol.decode_model(blob))
    assert it.run({"x": np.asarray([[1, -1, 0.5]], np.float32)})[0].tolist() == [[0.5, 1.0]]

    from supertonic_b200 import surrogate
    root = surrogate.ensure_assets("tiny")
    p = root + "/onnx/vector_estimator.onnx"
    rng = np.random.default_rng(1)
    B, L, Tn = 2, 19, 11
    feeds = dict(noisy_latent=rng.standard_normal((B, 144, L)).astype(np.float32),
                 text_emb=rng.standard_normal((B, 64, Tn)).astype(np.float32),
                 style_ttl=rng.standard_normal((B, 50, 64)).astype(np.float32),
                 text_mask=np.ones((B, 1, Tn), np.float32), latent_mask=np.ones((B, 1, L), np.float32),
                 total_step=np.full(B, 5, np.float32), current_step=np.full(B, 2, np.float32))
    y32 = oi.Interpreter(p, torch.float32).run(feeds)[0]
    y64 = oi.Interpreter(p, torch.float64).run(feeds)[0]
    assert y32.shape == (B, 144, L)
    assert np.abs(y32 - y64).max() < 2e-5
