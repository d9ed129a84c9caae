"""ORACLE — test infrastructure, never the product path.

Wires the ONNX interpreter (oracle/onnx_interp.py) into the restated `_infer`
(oracle/host_ref.py): the CPU stand-in for "reference ONNX Runtime CPU path on identical
inputs, with the same injected Gaussian noise tensor and the same style vectors" (north_star).
duration_predictor is evaluated in float64 and rounded once to float32 (DESIGN.md
§bit-exact durations); the other three graphs in float32 (float64 available as a shadow).
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import numpy as np
import torch

from . import host_ref
from .onnx_interp import Interpreter


def load_style(paths):
    """cpp/helper.cpp:829-897."""
    ttl, dp = [], []
    for p in paths:
        j = json.load(open(p))
        ttl.append(np.asarray(j["style_ttl"]["data"], np.float32).reshape(j["style_ttl"]["dims"]))
        dp.append(np.asarray(j["style_dp"]["data"], np.float32).reshape(j["style_dp"]["dims"]))
    return np.concatenate(ttl, 0), np.concatenate(dp, 0)


def make_noise(seed: int):
    """Injected noise: default_rng(seed).standard_normal((B,D,L)) as float32."""
    def fn(B, D, L):
        return np.random.default_rng(seed).standard_normal((B, D, L)).astype(np.float32)
    return fn


class OraclePipeline(host_ref.ReferenceTTS):
    def __init__(self, asset_root: str, dtype=torch.float32, gemm_operand_round: Optional[str] = None,
                 dp_dtype=torch.float64):
        onnx_dir = os.path.join(asset_root, "onnx")
        cfg = json.load(open(os.path.join(onnx_dir, "tts.json")))
        indexer = json.load(open(os.path.join(onnx_dir, "unicode_indexer.json")))
        self.sessions: Dict[str, Interpreter] = {
            "dp": Interpreter(os.path.join(onnx_dir, "duration_predictor.onnx"), dp_dtype),
            "te": Interpreter(os.path.join(onnx_dir, "text_encoder.onnx"), dtype, gemm_operand_round),
            "ve": Interpreter(os.path.join(onnx_dir, "vector_estimator.onnx"), dtype, gemm_operand_round),
            "voc": Interpreter(os.path.join(onnx_dir, "vocoder.onnx"), dtype, gemm_operand_round),
        }
        def runner(key):
            sess = self.sessions[key]
            return lambda feeds: sess.run(feeds)[0].astype(np.float32)
        super().__init__(cfg, indexer, runner("dp"), runner("te"), runner("ve"), runner("voc"))
        self.asset_root = asset_root

    def style(self, names):
        return load_style([os.path.join(self.asset_root, "voice_styles", n + ".json") for n in names])


def ort_version():
    """onnxruntime's version if it can be imported here, else None. (It never could in this environment: no wheel, no network —
    DESIGN.md §2. The probe exists so that the day it can, the TRUE reference runtime becomes the oracle and the CPU baseline.)"""
    try:
        import onnxruntime as ort           # noqa: F401
        return ort.__version__
    except Exception:                       # noqa: BLE001
        return None


class OrtPipeline(host_ref.ReferenceTTS):
    """The reference's own runtime: the four graphs through ONNX Runtime's CPU execution provider, orchestrated by the restated host
    arithmetic (the same call sites as py/helper.py:190-214 / cpp/helper.cpp:519, 552, 643, 668). Used as the primary oracle and CPU
    baseline whenever `import onnxruntime` succeeds (SURVEY.md §8c "closing the gap")."""

    def __init__(self, asset_root: str, threads: Optional[int] = None):
        import onnxruntime as ort
        onnx_dir = os.path.join(asset_root, "onnx")
        cfg = json.load(open(os.path.join(onnx_dir, "tts.json")))
        indexer = json.load(open(os.path.join(onnx_dir, "unicode_indexer.json")))
        so = ort.SessionOptions()
        if threads:
            so.intra_op_num_threads = int(threads)
        self.sessions = {k: ort.InferenceSession(os.path.join(onnx_dir, f + ".onnx"), so, providers=["CPUExecutionProvider"])
                         for k, f in (("dp", "duration_predictor"), ("te", "text_encoder"), ("ve", "vector_estimator"), ("voc", "vocoder"))}

        def runner(key):
            sess = self.sessions[key]
            names = {i.name for i in sess.get_inputs()}
            return lambda feeds: np.asarray(sess.run(None, {k: v for k, v in feeds.items() if k in names})[0], np.float32)
        super().__init__(cfg, indexer, runner("dp"), runner("te"), runner("ve"), runner("voc"))
        self.asset_root = asset_root

    def style(self, names):
        return load_style([os.path.join(self.asset_root, "voice_styles", n + ".json") for n in names])


def best_oracle(asset_root: str):
    """(pipeline, kind): ONNX Runtime when it is importable ("ort"), else the torch-CPU interpreter ("port")."""
    if ort_version():
        try:
            return OrtPipeline(asset_root), "ort"
        except Exception:                   # noqa: BLE001  (e.g. an ORT build that rejects the surrogate graphs' opset)
            pass
    return OraclePipeline(asset_root), "port"
