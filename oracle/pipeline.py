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
