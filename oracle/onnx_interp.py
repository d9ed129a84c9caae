"""ORACLE — test infrastructure, never the product path.

CPU restatement of what ONNX Runtime's CPU execution provider does at the four
`Ort::Session::Run` call sites of the reference (cpp/helper.cpp:519, 552, 643, 668;
py/helper.py:190, 194, 202, 214): evaluate an ONNX graph node by node, in
topological (file) order, following the published ONNX operator specification
(opset 17) for every op type present in the graphs.

The arithmetic lives in a third-party dependency that is absent from
/root/reference and from this image (onnxruntime, pinned 1.23.1 in
py/requirements.txt:1; C++ unpinned, cpp/README.md:31) and the reference holds no
golden vectors for this path ⇒ **PARITY UNPINNED**: this
interpreter is pinned only against (a) hand-computed known-answer tests of each
operator (tests/test_oracle_ops.py) and (b) its own float64 shadow run.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.

dtype policy: `dtype=torch.float32` reproduces ORT's fp32 graph; `torch.float64`
upcasts every float tensor (the shadow run used to bound the oracle's own rounding,
and the canonical mode for duration_predictor — DESIGN.md §"bit-exact durations").
"""
from __future__ import annotations

import math
import os
import sys
from typing import Any, Callable, Dict, List, Optional

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from supertonic_b200 import onnx_lite as ol  # noqa: E402  (wire-format codec only; no product compute)

OPS: Dict[str, Callable[..., Any]] = {}


def op(name):
    def deco(fn):
        OPS[name] = fn
        return fn
    return deco


def _ints(t) -> List[int]:
    return [int(v) for v in (t.reshape(-1).tolist() if isinstance(t, torch.Tensor) else t)]


# ---- elementwise
@op("Add")
def _add(a, b, **_): return a + b
@op("Sub")
def _sub(a, b, **_): return a - b
@op("Mul")
def _mul(a, b, **_): return a * b
@op("Div")
def _div(a, b, **_):
    if not a.is_floating_point() and not b.is_floating_point():
        return torch.div(a, b, rounding_mode="trunc")
    return a / b
@op("Neg")
def _neg(a, **_): return -a
@op("Abs")
def _abs(a, **_): return a.abs()
@op("Sqrt")
def _sqrt(a, **_): return a.sqrt()
@op("Exp")
def _exp(a, **_): return a.exp()
@op("Log")
def _log(a, **_): return a.log()
@op("Erf")
def _erf(a, **_): return torch.erf(a)
@op("Sin")
def _sin(a, **_): return a.sin()
@op("Cos")
def _cos(a, **_): return a.cos()
@op("Tanh")
def _tanh(a, **_): return a.tanh()
@op("Sigmoid")
def _sigmoid(a, **_): return torch.sigmoid(a)
@op("Relu")
def _relu(a, **_): return torch.relu(a)
@op("Softplus")
def _softplus(a, **_): return torch.nn.functional.softplus(a)
@op("LeakyRelu")
def _leaky(a, alpha=0.01, **_): return torch.nn.functional.leaky_relu(a, alpha)
@op("Gelu")
def _gelu(a, approximate="none", **_): return torch.nn.functional.gelu(a, approximate=approximate)
@op("Pow")
def _pow(a, b, **_): return torch.pow(a, b.to(a.dtype) if b.is_floating_point() else b)
@op("Identity")
def _identity(a, **_): return a
@op("Max")
def _max(*xs, **_):
    r = xs[0]
    for x in xs[1:]:
        r = torch.maximum(r, x)
    return r
@op("Min")
def _min(*xs, **_):
    r = xs[0]
    for x in xs[1:]:
        r = torch.minimum(r, x)
    return r
@op("Clip")
def _clip(a, lo=None, hi=None, **at):
    lo = lo if lo is not None else at.get("min")
    hi = hi if hi is not None else at.get("max")
    return torch.clamp(a, min=None if lo is None else float(lo), max=None if hi is None else float(hi))
@op("Where")
def _where(c, a, b, **_): return torch.where(c.bool(), a, b)
@op("Equal")
def _equal(a, b, **_): return a == b
@op("Less")
def _less(a, b, **_): return a < b
@op("Greater")
def _greater(a, b, **_): return a > b
@op("Not")
def _not(a, **_): return ~a.bool()


# ---- linear algebra
@op("MatMul")
def _matmul(a, b, **_): return torch.matmul(a, b)


@op("Gemm")
def _gemm(a, b, c=None, alpha=1.0, beta=1.0, transA=0, transB=0, **_):
    a = a.t() if transA else a
    b = b.t() if transB else b
    y = alpha * (a @ b)
    return y if c is None else y + beta * c


@op("Conv")
def _conv(x, w, b=None, group=1, dilations=None, pads=None, strides=None, kernel_shape=None,
          auto_pad="NOTSET", **_):
    nd = x.dim() - 2
    dilations = dilations or [1] * nd
    strides = strides or [1] * nd
    pads = pads or [0] * (2 * nd)
    if auto_pad not in ("NOTSET", b"NOTSET"):
        raise NotImplementedError("Conv auto_pad")
    if nd != 1:
        raise NotImplementedError("only Conv1d appears in these graphs")
    x = torch.nn.functional.pad(x, (pads[0], pads[1]))       # explicit begin/end pads (ONNX order)
    return torch.nn.functional.conv1d(x, w, b, stride=strides[0], dilation=dilations[0], groups=group)


@op("LayerNormalization")
def _layernorm(x, scale, bias=None, axis=-1, epsilon=1e-5, **_):
    axis = axis % x.dim()
    dims = list(range(axis, x.dim()))
    mean = x.mean(dim=dims, keepdim=True)
    d = x - mean
    var = (d * d).mean(dim=dims, keepdim=True)
    y = d / torch.sqrt(var + epsilon) * scale
    return y if bias is None else y + bias


@op("BatchNormalization")
def _batchnorm(x, scale, bias, mean, var, epsilon=1e-5, **_):
    shp = [1, -1] + [1] * (x.dim() - 2)
    return (x - mean.reshape(shp)) / torch.sqrt(var.reshape(shp) + epsilon) * scale.reshape(shp) + bias.reshape(shp)


@op("Softmax")
def _softmax(x, axis=-1, **_): return torch.softmax(x, dim=axis)


# ---- reductions
def _reduce(fn):
    def run(x, axes=None, keepdims=1, noop_with_empty_axes=0, **at):
        ax = _ints(axes) if axes is not None else at.get("axes_attr")
        if ax is None or len(ax) == 0:
            if noop_with_empty_axes:
                return x
            ax = list(range(x.dim()))
        return fn(x, dim=ax, keepdim=bool(keepdims))
    return run


OPS["ReduceSum"] = _reduce(torch.sum)
OPS["ReduceMean"] = _reduce(torch.mean)
OPS["ReduceMax"] = _reduce(torch.amax)


@op("CumSum")
def _cumsum(x, axis, exclusive=0, reverse=0, **_):
    if exclusive or reverse:
        raise NotImplementedError
    return torch.cumsum(x, dim=int(axis))


# ---- shape / data movement
@op("Transpose")
def _transpose(x, perm=None, **_): return x.permute(perm if perm is not None else list(range(x.dim()))[::-1])


@op("Reshape")
def _reshape(x, shape, allowzero=0, **_):
    s = _ints(shape)
    if not allowzero:
        s = [x.shape[i] if v == 0 else v for i, v in enumerate(s)]
    return x.reshape(s)


@op("Flatten")
def _flatten(x, axis=1, **_): return x.reshape(int(np.prod(x.shape[:axis])), -1)


@op("Unsqueeze")
def _unsqueeze(x, axes=None, **at):
    ax = _ints(axes) if axes is not None else at["axes_attr"]
    rank = x.dim() + len(ax)
    for a in sorted(a % rank for a in ax):
        x = x.unsqueeze(a)
    return x


@op("Squeeze")
def _squeeze(x, axes=None, **at):
    ax = _ints(axes) if axes is not None else at.get("axes_attr")
    if ax is None:
        return x.squeeze()
    for a in sorted((a % x.dim() for a in ax), reverse=True):
        x = x.squeeze(a)
    return x


@op("Concat")
def _concat(*xs, axis=0, **_): return torch.cat(xs, dim=axis)


@op("Slice")
def _slice(x, starts, ends, axes=None, steps=None, **_):
    starts, ends = _ints(starts), _ints(ends)
    axes = _ints(axes) if axes is not None else list(range(len(starts)))
    steps = _ints(steps) if steps is not None else [1] * len(starts)
    idx = [slice(None)] * x.dim()
    for s, e, a, st in zip(starts, ends, axes, steps):
        if st < 1:
            raise NotImplementedError("negative Slice step")
        n = x.shape[a]
        s = max(0, min(n, s + n if s < 0 else s))
        e = max(0, min(n, e + n if e < 0 else e))
        idx[a] = slice(s, e, st)
    return x[tuple(idx)]


@op("Split")
def _split(x, split=None, axis=0, num_outputs=None, **_):
    if split is not None:
        return list(torch.split(x, _ints(split), dim=axis))
    return list(torch.chunk(x, num_outputs, dim=axis))


@op("Gather")
def _gather(x, idx, axis=0, **_):
    idx = idx.long()
    idx = torch.where(idx < 0, idx + x.shape[axis], idx)
    return torch.index_select(x, axis, idx.reshape(-1)).reshape(
        list(x.shape[:axis]) + list(idx.shape) + list(x.shape[axis + 1:]))


@op("Expand")
def _expand(x, shape, **_):
    s = _ints(shape)
    return x * torch.ones(s, dtype=x.dtype) if x.is_floating_point() else x.expand(torch.broadcast_shapes(x.shape, s))


@op("Tile")
def _tile(x, reps, **_): return x.repeat(_ints(reps))


@op("Shape")
def _shape(x, **_): return torch.tensor(list(x.shape), dtype=torch.int64)


@op("Range")
def _range(s, l, d, **_): return torch.arange(s.item(), l.item(), d.item(), dtype=s.dtype)


@op("ConstantOfShape")
def _cos_(shape, value=None, **_):
    v = value if value is not None else torch.zeros(1)
    return torch.full(_ints(shape), v.reshape(-1)[0].item(), dtype=v.dtype)


@op("Pad")
def _pad(x, pads, value=None, axes=None, mode="constant", **_):
    if mode not in ("constant", b"constant"):
        raise NotImplementedError
    p = _ints(pads)
    n = x.dim()
    tp = []
    for a in reversed(range(n)):
        tp += [p[a], p[a + n]]
    return torch.nn.functional.pad(x, tp, value=0.0 if value is None else float(value))


_CAST = {ol.FLOAT: torch.float32, ol.DOUBLE: torch.float64, ol.INT64: torch.int64, ol.INT32: torch.int32,
         ol.BOOL: torch.bool, ol.FLOAT16: torch.float16}


class Interpreter:
    """Runs one ONNX graph. `run(feeds)` mirrors `Session::Run(names, values) → outputs`."""

    def __init__(self, path_or_model, dtype=torch.float32, gemm_operand_round: Optional[str] = None):
        self.model = ol.load_model(path_or_model) if isinstance(path_or_model, str) else path_or_model
        self.dtype = dtype
        self.round = gemm_operand_round     # None | "tf32" | "bf16": emulate tensor-core operand rounding
        g = self.model.graph
        self.consts = {k: self._lift(torch.from_numpy(v.copy())) for k, v in g.initializers.items()}
        self.input_names = [v.name for v in g.inputs]
        self.output_names = [v.name for v in g.outputs]

    def _lift(self, t: torch.Tensor) -> torch.Tensor:
        return t.to(self.dtype) if t.is_floating_point() else t

    def _rnd(self, t):
        if self.round == "bf16":
            return t.to(torch.bfloat16).to(t.dtype)
        if self.round == "tf32":          # round-to-nearest-even to 10 mantissa bits
            i = t.contiguous().view(torch.int32)
            i = (i + 0xFFF + ((i >> 13) & 1)) & ~0x1FFF
            return i.view(torch.float32)
        return t

    def run(self, feeds: Dict[str, Any], outputs: Optional[List[str]] = None, keep: bool = False):
        env: Dict[str, Any] = dict(self.consts)
        for k, v in feeds.items():
            t = torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v
            env[k] = self._lift(t)
        with torch.no_grad():
            for n in self.model.graph.nodes:
                ins = [env[i] if i else None for i in n.inputs]
                at = dict(n.attrs)
                if n.op_type == "Constant":
                    env[n.outputs[0]] = self._lift(torch.from_numpy(np.asarray(at["value"]).copy()))
                    continue
                if n.op_type == "Cast":
                    to = _CAST[at["to"]]
                    env[n.outputs[0]] = ins[0].to(self.dtype if to.is_floating_point else to)
                    continue
                if "axes" in at and n.op_type in ("Unsqueeze", "Squeeze", "ReduceSum", "ReduceMean", "ReduceMax"):
                    at["axes_attr"] = at.pop("axes")      # opset < 13 carries axes as an attribute
                fn = OPS.get(n.op_type)
                if fn is None:
                    raise NotImplementedError(f"ONNX op {n.op_type} (node {n.name}) not in the oracle")
                if self.round and n.op_type == "MatMul" and ins[0].dim() == 3 and ins[1].dim() == 2:
                    ins = [self._rnd(ins[0]), self._rnd(ins[1])]
                res = fn(*ins, **at)
                if isinstance(res, (list, tuple)):
                    for o, r in zip(n.outputs, res):
                        env[o] = r
                else:
                    env[n.outputs[0]] = res
        names = outputs or self.output_names
        out = [env[o] for o in names]
        if keep:
            self.env = env
        return [o.numpy() if isinstance(o, torch.Tensor) else o for o in out]
