"""Surrogate Supertonic asset set (LABELLED SURROGATE — not the released weights).

The released graphs live in an external Hugging Face repo that is not mounted
(SURVEY.md §0 facts 1+4), so everything measured in this repository runs on
random-init graphs of the hypothesised architecture:
real `.onnx` files written with `onnx_lite`, carrying stock ONNX nodes (so the
oracle interprets them op by op, exactly as ONNX Runtime would) plus a JSON
description of the layer sequence under metadata key ``stc_arch`` (what the
CUDA library instantiates its kernels from).

I/O names, dtypes and ranks follow the reference call sites exactly
(reference cpp/helper.cpp:512-523, 545-556, 620-647, 662-672).

Directory layout written by :func:`write_assets` (== reference `assets/`,
README.md:97-105): ``onnx/{duration_predictor,text_encoder,vector_estimator,
vocoder}.onnx``, ``onnx/tts.json``, ``onnx/unicode_indexer.json``,
``voice_styles/{M1,M2,F1,F2}.json``.
"""
from __future__ import annotations

import hashlib
import json
import math
import os
from typing import Any, Dict, List, Optional

import numpy as np

from . import onnx_lite as ol

SURROGATE_VERSION = 3

# --------------------------------------------------------------------------- configs
FULL = dict(
    name="full",
    sample_rate=44100, base_chunk_size=512, chunk_compress_factor=6, latent_dim=24,
    n_style=50, style_dim=256, dp_style=(8, 16),
    ve=dict(C=256, H=1024, K=5, heads=4, time_dim=64, superblocks=4, dil=(1, 2, 4, 8), mid=2, tail=4),
    te=dict(C=256, H=1024, K=5, heads=4, convnext=6, self_attn=4, style_attn=2),
    dp=dict(C=64, H=256, K=5, convnext=4),
    voc=dict(C=512, H=2048, K=7, dil=(1, 2, 4, 1, 2, 4, 1, 1, 1, 1)),
)

TINY = dict(
    name="tiny",
    sample_rate=44100, base_chunk_size=512, chunk_compress_factor=6, latent_dim=24,
    n_style=50, style_dim=64, dp_style=(8, 16),
    ve=dict(C=64, H=128, K=5, heads=2, time_dim=64, superblocks=1, dil=(1, 2), mid=1, tail=1),
    te=dict(C=64, H=128, K=5, heads=2, convnext=1, self_attn=1, style_attn=1),
    dp=dict(C=32, H=64, K=5, convnext=2),
    voc=dict(C=64, H=128, K=7, dil=(1, 2)),
)

CONFIGS = {"full": FULL, "tiny": TINY}


# --------------------------------------------------------------------------- unicode indexer
def build_indexer() -> List[int]:
    """65536-entry code-point → token-id table (format: reference cpp/helper.cpp:48-50, 383-385).
    id 0 = pad/unknown. Covers what the C++ front-end can emit for en/ko/es/pt/fr."""
    cps: List[int] = []
    cps += list(range(0x20, 0x7F))          # printable ASCII (includes < > / for language tags)
    cps += list(range(0xA1, 0x100))         # Latin-1 (non-decomposed leftovers)
    cps += list(range(0x300, 0x370))        # combining marks produced by the Latin table
    cps += list(range(0x1100, 0x1200))      # Hangul Jamo produced by syllable decomposition
    cps += [0x2026, 0x3002, 0x20AC]         # … 。 €
    table = [0] * 65536
    for i, cp in enumerate(cps):
        table[cp] = i + 1
    return table


def vocab_size() -> int:
    return max(build_indexer()) + 1


# --------------------------------------------------------------------------- graph builder
class GB:
    """Tiny ONNX graph builder: every method appends stock ONNX nodes and returns value names."""

    def __init__(self, name: str, rng: np.random.Generator):
        self.g = ol.Graph(name=name)
        self.rng = rng
        self._n = 0
        self.layers: List[Dict[str, Any]] = []

    # -- plumbing
    def tmp(self, hint: str = "t") -> str:
        self._n += 1
        return f"/{hint}_{self._n}"

    def node(self, op: str, inputs: List[str], attrs: Optional[dict] = None, hint: Optional[str] = None,
             n_out: int = 1):
        outs = [self.tmp(hint or op.lower()) for _ in range(n_out)]
        self.g.nodes.append(ol.Node(op, list(inputs), outs, dict(attrs or {}), name=outs[0][1:]))
        return outs[0] if n_out == 1 else outs

    def init(self, name: str, arr) -> str:
        assert name not in self.g.initializers, name
        self.g.initializers[name] = np.ascontiguousarray(arr)
        return name

    def const(self, arr, hint="c") -> str:
        return self.init(self.tmp(hint)[1:], np.asarray(arr))

    def randn(self, name: str, shape, std: float, mean: float = 0.0) -> str:
        w = self.rng.standard_normal(shape, dtype=np.float32) * np.float32(std) + np.float32(mean)
        return self.init(name, w.astype(np.float32))

    def input(self, name, dt, shape):
        self.g.inputs.append(ol.ValueInfo(name, dt, shape))
        return name

    def output(self, value: str, name: str, dt, shape):
        self.g.nodes.append(ol.Node("Identity", [value], [name], {}, name="out_" + name))
        self.g.outputs.append(ol.ValueInfo(name, dt, shape))

    # -- arithmetic sugar
    def add(self, a, b): return self.node("Add", [a, b])
    def sub(self, a, b): return self.node("Sub", [a, b])
    def mul(self, a, b): return self.node("Mul", [a, b])
    def div(self, a, b): return self.node("Div", [a, b])
    def matmul(self, a, b): return self.node("MatMul", [a, b])
    def transpose(self, a, perm): return self.node("Transpose", [a], {"perm": list(perm)})
    def reshape(self, a, shape): return self.node("Reshape", [a, self.const(np.asarray(shape, np.int64), "shape")])
    def unsqueeze(self, a, axes): return self.node("Unsqueeze", [a, self.const(np.asarray(axes, np.int64), "axes")])

    def gelu(self, x):
        """Exact (erf) GELU as PyTorch exports it for opset < 20: 0.5·x·(1+erf(x/√2))."""
        e = self.node("Erf", [self.div(x, self.const(np.float32(math.sqrt(2.0))))])
        return self.mul(self.mul(x, self.add(e, self.const(np.float32(1.0)))), self.const(np.float32(0.5)))

    def linear(self, x, prefix: str, cin: int, cout: int, std: Optional[float] = None):
        w = self.randn(prefix + ".weight", (cin, cout), std if std is not None else 1.0 / math.sqrt(cin))
        b = self.randn(prefix + ".bias", (cout,), 0.02)
        return self.add(self.matmul(x, w), b)

    def layernorm(self, x, prefix: str, c: int, eps: float = 1e-6):
        g = self.randn(prefix + ".weight", (c,), 0.1, 1.0)
        b = self.randn(prefix + ".bias", (c,), 0.1)
        return self.node("LayerNormalization", [x, g, b], {"axis": -1, "epsilon": float(eps)})

    # -- blocks
    def convnext(self, x, prefix: str, C: int, H: int, K: int, dil: int, causal: bool, mask: Optional[str]):
        """ConvNeXt-1D block on NCL input: depthwise conv → LN(C) → C→H → GELU →
        H→C → layer-scale → residual [→ ×mask]."""
        span = dil * (K - 1)
        pads = [span, 0] if causal else [span // 2, span - span // 2]
        dw_w = self.randn(prefix + ".dw.weight", (C, 1, K), 1.0 / math.sqrt(K))
        dw_b = self.randn(prefix + ".dw.bias", (C,), 0.02)
        h = self.node("Conv", [x, dw_w, dw_b], {"group": C, "kernel_shape": [K], "dilations": [dil],
                                                 "pads": pads, "strides": [1]})
        h = self.transpose(h, (0, 2, 1))
        h = self.layernorm(h, prefix + ".ln", C)
        h = self.linear(h, prefix + ".pw1", C, H)
        h = self.gelu(h)
        h = self.linear(h, prefix + ".pw2", H, C)
        h = self.mul(h, self.randn(prefix + ".gamma", (C,), 0.02, 0.1))
        h = self.transpose(h, (0, 2, 1))
        y = self.add(x, h)
        if mask is not None:
            y = self.mul(y, mask)
        self.layers.append(dict(type="convnext", name=prefix, C=C, H=H, K=K, dilation=dil,
                                causal=bool(causal), masked=mask is not None))
        return y

    def _positions(self, mask: str, normalise: bool):
        """mask [B,1,N] → rotary position [B,1,N,1]: cumsum(mask)−1, optionally ÷ length (LARoPE)."""
        pos = self.sub(self.node("CumSum", [mask, self.const(np.asarray(2, np.int64), "axis")]),
                       self.const(np.float32(1.0)))
        if normalise:
            ln = self.node("ReduceSum", [mask, self.const(np.asarray([2], np.int64), "axes")], {"keepdims": 1})
            pos = self.div(pos, ln)
        return self.unsqueeze(pos, [3])

    def _rope(self, t, pos, freqs, dh: int):
        """t [B,h,N,dh], pos [B,1,N,1], freqs [dh/2] → rotate-half rotary."""
        ang = self.mul(pos, freqs)                       # [B,1,N,dh/2]
        c, s = self.node("Cos", [ang]), self.node("Sin", [ang])
        half = dh // 2
        ax = self.const(np.asarray([3], np.int64), "axes")
        t1 = self.node("Slice", [t, self.const(np.asarray([0], np.int64)), self.const(np.asarray([half], np.int64)), ax])
        t2 = self.node("Slice", [t, self.const(np.asarray([half], np.int64)), self.const(np.asarray([dh], np.int64)), ax])
        r1 = self.sub(self.mul(t1, c), self.mul(t2, s))
        r2 = self.add(self.mul(t1, s), self.mul(t2, c))
        return self.node("Concat", [r1, r2], {"axis": 3})

    def attention(self, x, prefix: str, C: int, heads: int, ctx: Optional[str], ctx_dim: int, ctx_layout: str,
                  q_mask: Optional[str], k_mask: Optional[str], rope: str):
        """Pre-LN multi-head attention with residual on NCL `x`.
        ctx: None → self-attention (keys/values from LN(x)); else `ctx` is [B,Cctx,M] ("ncl") or
        [B,M,Cctx] ("nlc"). rope ∈ {"none","abs","norm"}; "norm" = length-aware RoPE (positions divided
        by the sequence length taken from the masks, arXiv 2509.11084 as cited in SURVEY.md App. B)."""
        dh = C // heads
        xt = self.transpose(x, (0, 2, 1))                                   # [B,N,C]
        xn = self.layernorm(xt, prefix + ".ln", C)
        if ctx is None:
            kv_src = xn
        else:
            kv_src = self.transpose(ctx, (0, 2, 1)) if ctx_layout == "ncl" else ctx
        q = self.linear(xn, prefix + ".q", C, C)
        k = self.linear(kv_src, prefix + ".k", ctx_dim, C)
        v = self.linear(kv_src, prefix + ".v", ctx_dim, C)

        def split(t):
            return self.transpose(self.reshape(t, [0, -1, heads, dh]), (0, 2, 1, 3))  # [B,h,N,dh]
        q, k, v = split(q), split(k), split(v)
        if rope != "none":
            base = 10000.0
            gamma = 100.0 if rope == "norm" else 1.0
            fr = (gamma * base ** (-np.arange(dh // 2, dtype=np.float64) / (dh // 2))).astype(np.float32)
            freqs = self.init(prefix + ".rope_freqs", fr)
            q = self._rope(q, self._positions(q_mask, rope == "norm"), freqs, dh)
            k = self._rope(k, self._positions(k_mask if ctx is not None else q_mask, rope == "norm"), freqs, dh)
        s = self.matmul(q, self.transpose(k, (0, 1, 3, 2)))
        s = self.mul(s, self.const(np.float32(1.0 / math.sqrt(dh))))
        km = k_mask if ctx is not None else q_mask
        if km is not None:
            bias = self.mul(self.sub(km, self.const(np.float32(1.0))), self.const(np.float32(1e9)))  # [B,1,M]
            s = self.add(s, self.unsqueeze(bias, [1]))
        p = self.node("Softmax", [s], {"axis": -1})
        o = self.matmul(p, v)                                               # [B,h,N,dh]
        o = self.reshape(self.transpose(o, (0, 2, 1, 3)), [0, -1, C])
        o = self.linear(o, prefix + ".o", C, C, std=0.3 / math.sqrt(C))
        y = self.add(x, self.transpose(o, (0, 2, 1)))
        if q_mask is not None:
            y = self.mul(y, q_mask)
        self.layers.append(dict(type="attention", name=prefix, C=C, heads=heads,
                                ctx=("self" if ctx is None else ctx), ctx_dim=ctx_dim, rope=rope,
                                masked=q_mask is not None, key_masked=km is not None))
        return y


def _model(gb: GB, arch: Dict[str, Any]) -> ol.Model:
    arch = dict(arch, layers=gb.layers, surrogate_version=SURROGATE_VERSION)
    return ol.Model(gb.g, metadata={"stc_arch": json.dumps(arch), "stc_surrogate": "true"})


# --------------------------------------------------------------------------- the four graphs
def build_duration_predictor(cfg, seed) -> ol.Model:
    """text_ids[B,T] i64, style_dp[B,e1,e2], text_mask[B,1,T] → duration[B] seconds
    (reference cpp/helper.cpp:512-526)."""
    c = cfg["dp"]
    C, V = c["C"], vocab_size()
    e1, e2 = cfg["dp_style"]
    gb = GB("duration_predictor", np.random.default_rng(seed))
    ids = gb.input("text_ids", ol.INT64, ["B", "T"])
    sty = gb.input("style_dp", ol.FLOAT, ["B", e1, e2])
    mask = gb.input("text_mask", ol.FLOAT, ["B", 1, "T"])
    emb = gb.randn("dp.embed.weight", (V, C), 1.0)
    x = gb.node("Gather", [emb, ids], {"axis": 0})                         # [B,T,C]
    x = gb.mul(gb.transpose(x, (0, 2, 1)), mask)                           # [B,C,T]
    s = gb.linear(gb.reshape(sty, [0, e1 * e2]), "dp.style", e1 * e2, C)   # [B,C]
    x = gb.mul(gb.add(x, gb.unsqueeze(s, [2])), mask)
    for i in range(c["convnext"]):
        x = gb.convnext(x, f"dp.cn{i}", C, c["H"], c["K"], 1, False, mask)
    h = gb.layernorm(gb.transpose(x, (0, 2, 1)), "dp.head.ln", C)          # [B,T,C]
    logd = gb.linear(h, "dp.head.proj", C, 1, std=0.3 / math.sqrt(C))      # [B,T,1]
    logd = gb.node("Clip", [logd, gb.const(np.float32(-3.0)), gb.const(np.float32(3.0))])
    dtok = gb.mul(gb.node("Exp", [logd]), gb.const(np.float32(0.065)))     # seconds per token
    dtok = gb.mul(gb.transpose(dtok, (0, 2, 1)), mask)                     # [B,1,T]
    dur = gb.node("ReduceSum", [dtok, gb.const(np.asarray([1, 2], np.int64), "axes")], {"keepdims": 0})
    gb.output(dur, "duration", ol.FLOAT, ["B"])
    return _model(gb, dict(kind="duration_predictor", C=C, H=c["H"], K=c["K"], vocab=V,
                           style_in=e1 * e2, sec_per_token=0.065, clip=3.0))


def build_text_encoder(cfg, seed) -> ol.Model:
    """text_ids, style_ttl[B,S,Cs], text_mask → text_emb[B,C,T] (reference cpp/helper.cpp:545-556)."""
    c = cfg["te"]
    C, V = c["C"], vocab_size()
    S, Cs = cfg["n_style"], cfg["style_dim"]
    gb = GB("text_encoder", np.random.default_rng(seed))
    ids = gb.input("text_ids", ol.INT64, ["B", "T"])
    sty = gb.input("style_ttl", ol.FLOAT, ["B", S, Cs])
    mask = gb.input("text_mask", ol.FLOAT, ["B", 1, "T"])
    emb = gb.randn("te.embed.weight", (V, C), 1.0)
    x = gb.node("Gather", [emb, ids], {"axis": 0})
    x = gb.mul(gb.transpose(x, (0, 2, 1)), mask)
    for i in range(c["convnext"]):
        x = gb.convnext(x, f"te.cn{i}", C, c["H"], c["K"], 1, False, mask)
    for i in range(c["self_attn"]):
        x = gb.attention(x, f"te.sa{i}.attn", C, c["heads"], None, C, "nlc", mask, None, "abs")
        x = gb.convnext(x, f"te.sa{i}.cn", C, c["H"], c["K"], 1, False, mask)
    for i in range(c["style_attn"]):
        x = gb.attention(x, f"te.st{i}.attn", C, c["heads"], sty, Cs, "nlc", mask, None, "none")
        x = gb.convnext(x, f"te.st{i}.cn", C, c["H"], c["K"], 1, False, mask)
    y = gb.linear(gb.transpose(x, (0, 2, 1)), "te.proj_out", C, C)
    y = gb.mul(gb.transpose(y, (0, 2, 1)), mask)
    gb.layers.append(dict(type="proj_out", name="te.proj_out", cin=C, cout=C))
    gb.output(y, "text_emb", ol.FLOAT, ["B", C, "T"])
    return _model(gb, dict(kind="text_encoder", C=C, H=c["H"], K=c["K"], heads=c["heads"], vocab=V,
                           n_style=S, style_dim=Cs))


def build_vector_estimator(cfg, seed) -> ol.Model:
    """One Euler step of the flow-matching ODE with the update in-graph
    (reference cpp/helper.cpp:620-658: the loop only reassigns the output)."""
    c = cfg["ve"]
    C, H, K, heads = c["C"], c["H"], c["K"], c["heads"]
    D = cfg["latent_dim"] * cfg["chunk_compress_factor"]
    S, Cs, Ct = cfg["n_style"], cfg["style_dim"], cfg["te"]["C"]
    gb = GB("vector_estimator", np.random.default_rng(seed))
    xin = gb.input("noisy_latent", ol.FLOAT, ["B", D, "L"])
    temb = gb.input("text_emb", ol.FLOAT, ["B", Ct, "T"])
    sty = gb.input("style_ttl", ol.FLOAT, ["B", S, Cs])
    tmask = gb.input("text_mask", ol.FLOAT, ["B", 1, "T"])
    lmask = gb.input("latent_mask", ol.FLOAT, ["B", 1, "L"])
    tot = gb.input("total_step", ol.FLOAT, ["B"])
    cur = gb.input("current_step", ol.FLOAT, ["B"])

    # time embedding: t = current/total → sinusoid(time_dim) → MLP
    td = c["time_dim"]
    t = gb.div(cur, tot)                                                    # [B]
    fr = (1000.0 * np.exp(-math.log(10000.0) * np.arange(td // 2, dtype=np.float64) / (td // 2))).astype(np.float32)
    arg = gb.mul(gb.unsqueeze(t, [1]), gb.init("ve.time.freqs", fr))       # [B,td/2]
    te = gb.node("Concat", [gb.node("Sin", [arg]), gb.node("Cos", [arg])], {"axis": 1})
    te = gb.linear(gb.gelu(gb.linear(te, "ve.time.fc1", td, C)), "ve.time.fc2", C, C)   # [B,C]
    gb.layers.append(dict(type="time_mlp", name="ve.time", time_dim=td, C=C))

    x = gb.linear(gb.transpose(xin, (0, 2, 1)), "ve.proj_in", D, C)
    x = gb.mul(gb.transpose(x, (0, 2, 1)), lmask)                           # [B,C,L]
    gb.layers.append(dict(type="proj_in", name="ve.proj_in", cin=D, cout=C))
    for sb in range(c["superblocks"]):
        p = f"ve.sb{sb}"
        for j, d in enumerate(c["dil"]):
            x = gb.convnext(x, f"{p}.dil{j}", C, H, K, d, False, lmask)
        tc = gb.linear(te, f"{p}.time", C, C)
        x = gb.mul(gb.add(x, gb.unsqueeze(tc, [2])), lmask)
        gb.layers.append(dict(type="time_cond", name=f"{p}.time", C=C))
        for j in range(c["mid"]):
            x = gb.convnext(x, f"{p}.mid{j}", C, H, K, 1, False, lmask)
        x = gb.attention(x, f"{p}.text_attn", C, heads, temb, Ct, "ncl", lmask, tmask, "norm")
        x = gb.convnext(x, f"{p}.post", C, H, K, 1, False, lmask)
        x = gb.attention(x, f"{p}.style_attn", C, heads, sty, Cs, "nlc", lmask, None, "none")
    for j in range(c["tail"]):
        x = gb.convnext(x, f"ve.tail{j}", C, H, K, 1, False, lmask)
    v = gb.linear(gb.transpose(x, (0, 2, 1)), "ve.proj_out", C, D, std=0.5 / math.sqrt(C))
    v = gb.transpose(v, (0, 2, 1))                                          # [B,D,L]
    gb.layers.append(dict(type="proj_out", name="ve.proj_out", cin=C, cout=D))
    dt = gb.unsqueeze(gb.div(gb.const(np.float32(1.0)), tot), [1, 2])      # [B,1,1]
    y = gb.mul(gb.add(xin, gb.mul(v, dt)), lmask)
    gb.output(y, "denoised_latent", ol.FLOAT, ["B", D, "L"])
    return _model(gb, dict(kind="vector_estimator", C=C, H=H, K=K, heads=heads, latent_ch=D, time_dim=td,
                           text_dim=Ct, n_style=S, style_dim=Cs))


def build_vocoder(cfg, seed) -> ol.Model:
    """latent[B,D,L] → wav_tts[B, L·cs]; no mask input (reference cpp/helper.cpp:662-672)."""
    c = cfg["voc"]
    C, H, K = c["C"], c["H"], c["K"]
    ld, f, hop = cfg["latent_dim"], cfg["chunk_compress_factor"], cfg["base_chunk_size"]
    D = ld * f
    gb = GB("vocoder", np.random.default_rng(seed))
    z = gb.input("latent", ol.FLOAT, ["B", D, "L"])
    std = gb.randn("voc.latent_std", (1, D, 1), 0.1, 1.0)
    mean = gb.randn("voc.latent_mean", (1, D, 1), 0.1)
    z = gb.add(gb.mul(z, std), mean)
    # un-compress: channel j·ld+c at frame l → channel c at frame f·l+j
    z = gb.reshape(z, [0, f, ld, -1])
    z = gb.reshape(gb.transpose(z, (0, 2, 3, 1)), [0, ld, -1])             # [B,ld,f·L]
    w = gb.randn("voc.conv_in.weight", (C, ld, K), 1.0 / math.sqrt(ld * K))
    b = gb.randn("voc.conv_in.bias", (C,), 0.02)
    h = gb.node("Conv", [z, w, b], {"group": 1, "kernel_shape": [K], "dilations": [1], "pads": [K - 1, 0],
                                    "strides": [1]})
    bn = [gb.randn("voc.bn.weight", (C,), 0.1, 1.0), gb.randn("voc.bn.bias", (C,), 0.1),
          gb.randn("voc.bn.running_mean", (C,), 0.1),
          gb.init("voc.bn.running_var", (1.0 + 0.2 * gb.rng.random(C)).astype(np.float32))]
    h = gb.node("BatchNormalization", [h] + bn, {"epsilon": 1e-5})
    gb.layers.append(dict(type="conv_in", name="voc.conv_in", cin=ld, cout=C, K=K, causal=True, bn="voc.bn"))
    for i, d in enumerate(c["dil"]):
        h = gb.convnext(h, f"voc.cn{i}", C, H, K, d, True, None)
    h = gb.layernorm(gb.transpose(h, (0, 2, 1)), "voc.head.ln", C)
    h = gb.linear(h, "voc.head.proj", C, hop, std=0.1 / math.sqrt(C))       # [B,fL,hop]
    gb.layers.append(dict(type="head", name="voc.head", cin=C, cout=hop))
    wav = gb.reshape(h, [0, -1])
    gb.output(wav, "wav_tts", ol.FLOAT, ["B", "T_wav"])
    return _model(gb, dict(kind="vocoder", C=C, H=H, K=K, latent_ch=D, latent_dim=ld, compress=f, hop=hop))


# --------------------------------------------------------------------------- styles + config files
def make_style(cfg, seed: int) -> Dict[str, Any]:
    """Voice-style JSON in the reference's schema (cpp/helper.cpp:829-897)."""
    rng = np.random.default_rng(seed)
    S, Cs = cfg["n_style"], cfg["style_dim"]
    e1, e2 = cfg["dp_style"]
    ttl = rng.standard_normal((1, S, Cs)).astype(np.float32)
    dp = rng.standard_normal((1, e1, e2)).astype(np.float32)
    return {"style_ttl": {"data": ttl.tolist(), "dims": [1, S, Cs], "type": "float32"},
            "style_dp": {"data": dp.tolist(), "dims": [1, e1, e2], "type": "float32"}}


STYLE_SEEDS = {"M1": 7, "M2": 8, "F1": 9, "F2": 10}


def tts_json(cfg) -> Dict[str, Any]:
    """Keys consumed by loadCfgs (cpp/helper.cpp:801-818) + the extra ones go/helper.go:25-84 names."""
    return {"surrogate": True, "surrogate_config": cfg["name"],
            "ae": {"sample_rate": cfg["sample_rate"], "base_chunk_size": cfg["base_chunk_size"]},
            "ttl": {"chunk_compress_factor": cfg["chunk_compress_factor"], "latent_dim": cfg["latent_dim"],
                    "style_encoder": {"style_token_layer": {"n_style": cfg["n_style"],
                                                            "style_value_dim": cfg["style_dim"]}}},
            "dp": {"latent_dim": cfg["latent_dim"], "chunk_compress_factor": cfg["chunk_compress_factor"],
                   "style_encoder": {"style_token_layer": {"n_style": cfg["dp_style"][0],
                                                           "style_value_dim": cfg["dp_style"][1]}}}}


def _stamp(cfg, seed) -> str:
    blob = json.dumps([SURROGATE_VERSION, cfg, seed], sort_keys=True, default=list)
    return hashlib.sha256(blob.encode()).hexdigest()[:16]


def write_assets(root: str, config: str = "full", seed: int = 0) -> str:
    cfg = CONFIGS[config]
    onnx_dir = os.path.join(root, "onnx")
    os.makedirs(onnx_dir, exist_ok=True)
    os.makedirs(os.path.join(root, "voice_styles"), exist_ok=True)
    builders = {"duration_predictor": build_duration_predictor, "text_encoder": build_text_encoder,
                "vector_estimator": build_vector_estimator, "vocoder": build_vocoder}
    for i, (name, fn) in enumerate(builders.items()):
        ol.save_model(fn(cfg, seed * 100 + i), os.path.join(onnx_dir, name + ".onnx"))
    with open(os.path.join(onnx_dir, "tts.json"), "w") as f:
        json.dump(tts_json(cfg), f, indent=1)
    with open(os.path.join(onnx_dir, "unicode_indexer.json"), "w") as f:
        json.dump(build_indexer(), f)
    for name, s in STYLE_SEEDS.items():
        with open(os.path.join(root, "voice_styles", name + ".json"), "w") as f:
            json.dump(make_style(cfg, s), f)
    with open(os.path.join(root, "STAMP"), "w") as f:
        f.write(_stamp(cfg, seed))
    return root


def default_root(config: str = "full") -> str:
    base = os.environ.get("SUPERTONIC_SURROGATE_DIR") or os.path.join(
        os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets_surrogate")
    return os.path.join(base, config)


def ensure_assets(config: str = "full", seed: int = 0, root: Optional[str] = None) -> str:
    """Return a directory holding the surrogate asset set, generating it on first use
    (seeded numpy ⇒ byte-identical here and on the GPU box)."""
    root = root or default_root(config)
    stamp = os.path.join(root, "STAMP")
    want = _stamp(CONFIGS[config], seed)
    if os.path.exists(stamp) and open(stamp).read().strip() == want:
        return root
    return write_assets(root, config, seed)


if __name__ == "__main__":
    import sys
    cfgname = sys.argv[1] if len(sys.argv) > 1 else "full"
    r = ensure_assets(cfgname)
    for n in ("duration_predictor", "text_encoder", "vector_estimator", "vocoder"):
        print(ol.describe(ol.load_model(os.path.join(r, "onnx", n + ".onnx"))))
