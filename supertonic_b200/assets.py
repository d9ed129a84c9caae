"""Where the model files are.

The released Supertonic assets (the four .onnx graphs, tts.json, unicode_indexer.json, voice_styles/*.json) live in an external
Hugging Face repository that the reference clones into `assets/` (reference README.md:97-105); they are not part of the reference
tree and were never mounted in this environment. `find_released_assets()` looks where they would appear; everything falls back to
the labelled SURROGATE set (`surrogate.ensure_assets`) otherwise."""
from __future__ import annotations

import os
from typing import Optional

_FILES = ("duration_predictor.onnx", "text_encoder.onnx", "vector_estimator.onnx", "vocoder.onnx", "tts.json", "unicode_indexer.json")
_HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_released_assets() -> Optional[str]:
    """Root directory (the one holding `onnx/` and `voice_styles/`) of a released asset set, or None.
    Search order: $SUPERTONIC_ASSETS, <repo>/assets, <repo>/baseline/_ref/assets, /root/reference/assets."""
    cands = [os.environ.get("SUPERTONIC_ASSETS"), os.path.join(_HERE, "assets"), os.path.join(_HERE, "baseline", "_ref", "assets"),
             "/root/reference/assets"]
    for c in cands:
        if c and all(os.path.exists(os.path.join(c, "onnx", f)) for f in _FILES):
            return c
    return None


def asset_root(config: str = "full") -> tuple:
    """(root, kind): the released assets if present ("released"), else the surrogate set of that size ("surrogate")."""
    rel = find_released_assets()
    if rel:
        return rel, "released"
    from . import surrogate
    return surrogate.ensure_assets(config), "surrogate"
