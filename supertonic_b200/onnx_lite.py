"""Dependency-free ONNX (protobuf wire format) reader / writer.

Neither `onnx` nor `protoc` exists in this image, so the
subset of onnx.proto needed to carry a graph + initializers is coded by hand
(field numbers: SURVEY.md Appendix C, from the public onnx.proto).

The same bytes are consumed by three parties:
  * the surrogate asset generator (`supertonic_b200/surrogate.py`) writes them,
  * the CPU oracle (`oracle/onnx_interp.py`) interprets them node by node,
  * the CUDA library (`supertonic_b200/csrc/onnx_reader.cc`) loads the
    initializers into device buffers (north_star: "load the ONNX initializers
    into device buffers").
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

# TensorProto.DataType
FLOAT, UINT8, INT8, INT32, INT64, BOOL, FLOAT16, DOUBLE, BFLOAT16 = 1, 2, 3, 6, 7, 9, 10, 11, 16
_NP_OF = {FLOAT: np.float32, UINT8: np.uint8, INT8: np.int8, INT32: np.int32, INT64: np.int64,
          BOOL: np.bool_, FLOAT16: np.float16, DOUBLE: np.float64}
_DT_OF = {np.dtype(v): k for k, v in _NP_OF.items()}

# AttributeProto.AttributeType
A_FLOAT, A_INT, A_STRING, A_TENSOR, A_FLOATS, A_INTS, A_STRINGS = 1, 2, 3, 4, 6, 7, 8


# --------------------------------------------------------------------------- wire helpers
def _varint(n: int) -> bytes:
    if n < 0:
        n += 1 << 64
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(fieldno: int, wt: int) -> bytes:
    return _varint((fieldno << 3) | wt)


def _f_varint(fieldno: int, v: int) -> bytes:
    return _key(fieldno, 0) + _varint(int(v))


def _f_bytes(fieldno: int, b: bytes) -> bytes:
    return _key(fieldno, 2) + _varint(len(b)) + b


def _f_str(fieldno: int, s: str) -> bytes:
    return _f_bytes(fieldno, s.encode("utf-8"))


def _f_float(fieldno: int, v: float) -> bytes:
    return _key(fieldno, 5) + struct.pack("<f", v)


def _read_varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not (b & 0x80):
            return result, pos
        shift += 7


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _fields(buf: memoryview):
    """Yield (fieldno, wiretype, value) — value is int for varint/fixed, memoryview for bytes."""
    pos, n = 0, len(buf)
    while pos < n:
        k, pos = _read_varint(buf, pos)
        fno, wt = k >> 3, k & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
            yield fno, wt, v
        elif wt == 1:
            yield fno, wt, bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            yield fno, wt, buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            yield fno, wt, bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")


def _packed_varints(v) -> List[int]:
    out, pos, n = [], 0, len(v)
    while pos < n:
        x, pos = _read_varint(v, pos)
        out.append(_signed(x))
    return out


# --------------------------------------------------------------------------- data model
@dataclass
class Node:
    op_type: str
    inputs: List[str]
    outputs: List[str]
    attrs: Dict[str, Any] = field(default_factory=dict)
    name: str = ""


@dataclass
class ValueInfo:
    name: str
    elem_type: int
    shape: List[Any]  # ints or str (dim_param)


@dataclass
class Graph:
    name: str = "graph"
    nodes: List[Node] = field(default_factory=list)
    initializers: Dict[str, np.ndarray] = field(default_factory=dict)
    inputs: List[ValueInfo] = field(default_factory=list)
    outputs: List[ValueInfo] = field(default_factory=list)


@dataclass
class Model:
    graph: Graph
    ir_version: int = 8
    opset: int = 17
    producer_name: str = "supertonic_b200.surrogate"
    metadata: Dict[str, str] = field(default_factory=dict)


# --------------------------------------------------------------------------- encode
def _enc_tensor(name: str, arr: np.ndarray) -> bytes:
    arr = np.ascontiguousarray(arr)
    dt = _DT_OF[arr.dtype]
    out = bytearray()
    for d in arr.shape:
        out += _f_varint(1, d)
    out += _f_varint(2, dt)
    out += _f_str(8, name)
    out += _f_bytes(9, arr.tobytes())
    return bytes(out)


def _enc_attr(name: str, v: Any) -> bytes:
    out = bytearray(_f_str(1, name))
    if isinstance(v, bool):
        v = int(v)
    if isinstance(v, float):
        out += _f_float(2, v) + _f_varint(20, A_FLOAT)
    elif isinstance(v, (int, np.integer)):
        out += _f_varint(3, int(v)) + _f_varint(20, A_INT)
    elif isinstance(v, str):
        out += _f_bytes(4, v.encode()) + _f_varint(20, A_STRING)
    elif isinstance(v, np.ndarray):
        out += _f_bytes(5, _enc_tensor("", v)) + _f_varint(20, A_TENSOR)
    elif isinstance(v, (list, tuple)) and all(isinstance(x, (int, np.integer)) for x in v):
        out += _f_bytes(8, b"".join(_varint(int(x)) for x in v)) + _f_varint(20, A_INTS)
    elif isinstance(v, (list, tuple)) and all(isinstance(x, float) for x in v):
        out += _f_bytes(7, b"".join(struct.pack("<f", x) for x in v)) + _f_varint(20, A_FLOATS)
    else:
        raise TypeError(f"attribute {name}: unsupported value {v!r}")
    return bytes(out)


def _enc_node(n: Node) -> bytes:
    out = bytearray()
    for i in n.inputs:
        out += _f_str(1, i)
    for o in n.outputs:
        out += _f_str(2, o)
    if n.name:
        out += _f_str(3, n.name)
    out += _f_str(4, n.op_type)
    for k, v in n.attrs.items():
        out += _f_bytes(5, _enc_attr(k, v))
    return bytes(out)


def _enc_value_info(v: ValueInfo) -> bytes:
    dims = bytearray()
    for d in v.shape:
        if isinstance(d, str):
            dims += _f_bytes(1, _f_str(2, d))
        else:
            dims += _f_bytes(1, _f_varint(1, d))
    tensor_type = _f_varint(1, v.elem_type) + _f_bytes(2, bytes(dims))
    type_proto = _f_bytes(1, tensor_type)
    return _f_str(1, v.name) + _f_bytes(2, type_proto)


def encode_model(m: Model) -> bytes:
    g = m.graph
    gb = bytearray()
    for n in g.nodes:
        gb += _f_bytes(1, _enc_node(n))
    gb += _f_str(2, g.name)
    for name, arr in g.initializers.items():
        gb += _f_bytes(5, _enc_tensor(name, arr))
    for v in g.inputs:
        gb += _f_bytes(11, _enc_value_info(v))
    for v in g.outputs:
        gb += _f_bytes(12, _enc_value_info(v))
    out = bytearray()
    out += _f_varint(1, m.ir_version)
    out += _f_str(2, m.producer_name)
    out += _f_bytes(7, bytes(gb))
    out += _f_bytes(8, _f_str(1, "") + _f_varint(2, m.opset))
    for k, v in m.metadata.items():
        out += _f_bytes(14, _f_str(1, k) + _f_str(2, v))
    return bytes(out)


def save_model(m: Model, path: str) -> None:
    with open(path, "wb") as f:
        f.write(encode_model(m))


# --------------------------------------------------------------------------- decode
def _dec_tensor(buf: memoryview) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dt = FLOAT
    name = ""
    raw: Optional[bytes] = None
    floats: List[float] = []
    i32: List[int] = []
    i64: List[int] = []
    f64: List[float] = []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            dims += _packed_varints(v) if wt == 2 else [_signed(v)]
        elif fno == 2:
            dt = v
        elif fno == 8:
            name = bytes(v).decode()
        elif fno == 9:
            raw = bytes(v)
        elif fno == 4:
            floats += list(np.frombuffer(bytes(v), "<f4")) if wt == 2 else [struct.unpack("<f", v)[0]]
        elif fno == 5:
            i32 += _packed_varints(v) if wt == 2 else [_signed(v)]
        elif fno == 7:
            i64 += _packed_varints(v) if wt == 2 else [_signed(v)]
        elif fno == 10:
            f64 += list(np.frombuffer(bytes(v), "<f8")) if wt == 2 else [struct.unpack("<d", v)[0]]
        elif fno in (13, 14):
            if fno == 14 and v == 1:
                raise ValueError(f"tensor {name}: external data is not supported")
    npdt = _NP_OF[dt]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=npdt).copy()
    elif floats:
        arr = np.asarray(floats, np.float32)
    elif i64:
        arr = np.asarray(i64, np.int64)
    elif i32:
        arr = np.asarray(i32).astype(npdt)
    elif f64:
        arr = np.asarray(f64, np.float64)
    else:
        arr = np.zeros(0, npdt)
    return name, arr.reshape(dims)


def _dec_attr(buf: memoryview) -> Tuple[str, Any]:
    name, typ = "", 0
    f = i = s = t = None
    floats: List[float] = []
    ints: List[int] = []
    strings: List[str] = []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = bytes(v).decode()
        elif fno == 2:
            f = struct.unpack("<f", v)[0]
        elif fno == 3:
            i = _signed(v)
        elif fno == 4:
            s = bytes(v).decode("utf-8", "replace")
        elif fno == 5:
            t = _dec_tensor(v)[1]
        elif fno == 7:
            floats += list(np.frombuffer(bytes(v), "<f4").astype(float)) if wt == 2 else [struct.unpack("<f", v)[0]]
        elif fno == 8:
            ints += _packed_varints(v) if wt == 2 else [_signed(v)]
        elif fno == 9:
            strings.append(bytes(v).decode())
        elif fno == 20:
            typ = v
    if typ == A_FLOAT or (typ == 0 and f is not None):
        return name, float(f)
    if typ == A_INT or (typ == 0 and i is not None):
        return name, int(i)
    if typ == A_STRING or (typ == 0 and s is not None):
        return name, s
    if typ == A_TENSOR or (typ == 0 and t is not None):
        return name, t
    if typ == A_FLOATS:
        return name, floats
    if typ == A_INTS:
        return name, ints
    if typ == A_STRINGS:
        return name, strings
    return name, ints or floats or strings or None


def _dec_node(buf: memoryview) -> Node:
    n = Node("", [], [])
    for fno, wt, v in _fields(buf):
        if fno == 1:
            n.inputs.append(bytes(v).decode())
        elif fno == 2:
            n.outputs.append(bytes(v).decode())
        elif fno == 3:
            n.name = bytes(v).decode()
        elif fno == 4:
            n.op_type = bytes(v).decode()
        elif fno == 5:
            k, val = _dec_attr(v)
            n.attrs[k] = val
    return n


def _dec_value_info(buf: memoryview) -> ValueInfo:
    vi = ValueInfo("", FLOAT, [])
    for fno, wt, v in _fields(buf):
        if fno == 1:
            vi.name = bytes(v).decode()
        elif fno == 2:
            for f2, _, v2 in _fields(v):
                if f2 != 1:
                    continue
                for f3, _, v3 in _fields(v2):
                    if f3 == 1:
                        vi.elem_type = v3
                    elif f3 == 2:
                        for f4, _, v4 in _fields(v3):
                            if f4 != 1:
                                continue
                            dim: Any = "?"
                            for f5, _, v5 in _fields(v4):
                                if f5 == 1:
                                    dim = _signed(v5)
                                elif f5 == 2:
                                    dim = bytes(v5).decode()
                            vi.shape.append(dim)
    return vi


def decode_model(data: bytes) -> Model:
    buf = memoryview(data)
    m = Model(Graph())
    m.metadata = {}
    for fno, wt, v in _fields(buf):
        if fno == 1:
            m.ir_version = v
        elif fno == 2:
            m.producer_name = bytes(v).decode()
        elif fno == 7:
            g = m.graph
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    g.nodes.append(_dec_node(v2))
                elif f2 == 2:
                    g.name = bytes(v2).decode()
                elif f2 == 5:
                    name, arr = _dec_tensor(v2)
                    g.initializers[name] = arr
                elif f2 == 11:
                    g.inputs.append(_dec_value_info(v2))
                elif f2 == 12:
                    g.outputs.append(_dec_value_info(v2))
        elif fno == 8:
            for f2, _, v2 in _fields(v):
                if f2 == 2:
                    m.opset = v2
        elif fno == 14:
            k = val = ""
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    k = bytes(v2).decode()
                elif f2 == 2:
                    val = bytes(v2).decode()
            m.metadata[k] = val
    # graph inputs that are initializers are not runtime inputs
    m.graph.inputs = [i for i in m.graph.inputs if i.name not in m.graph.initializers]
    return m


def load_model(path: str) -> Model:
    with open(path, "rb") as f:
        return decode_model(f.read())


def describe(m: Model) -> str:
    """Op histogram + I/O signature + initializer table."""
    from collections import Counter
    g = m.graph
    lines = [f"graph {g.name}: {len(g.nodes)} nodes, {len(g.initializers)} initializers, opset {m.opset}"]
    for v in g.inputs:
        lines.append(f"  in  {v.name}: dtype {v.elem_type} shape {v.shape}")
    for v in g.outputs:
        lines.append(f"  out {v.name}: dtype {v.elem_type} shape {v.shape}")
    hist = Counter(n.op_type for n in g.nodes)
    lines.append("  ops: " + ", ".join(f"{k}×{c}" for k, c in sorted(hist.items())))
    nparam = sum(int(a.size) for a in g.initializers.values())
    lines.append(f"  parameters: {nparam:,}")
    return "\n".join(lines)
