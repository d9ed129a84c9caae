"""Fused ConvNeXt MLP vs two GEMMs over row counts (run on the B200 box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from supertonic_b200 import capi, surrogate
for form in ("fused", "split"):
    os.environ["STC_MLP"] = form
    eng = capi.Engine(os.path.join(surrogate.ensure_assets("tiny"), "onnx"))
    for M in (128, 1024, 1536, 2048, 4096, 4224, 4352, 4736, 4864, 8448, 9600):
        f, u, e = eng.debug_mlp(M, 20)
        print(f"{form:6s} M={M:6d} tiles={-(-M // 128):4d}  fused {f:7.2f} us   two GEMMs {u:7.2f} us   max|diff| {e:.2e}", flush=True)
    eng.close()
