"""Fused ConvNeXt MLP vs two GEMMs over row counts (run on the B200 box). Forms: STC_MLP = fused (cluster) | split | ts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from supertonic_b200 import capi, surrogate
forms = sys.argv[1].split(",") if len(sys.argv) > 1 else ["fused", "split", "thin", "thin64", "ts", "ts16"]
rows = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [128, 1024, 1536, 2048, 4096, 4224, 4352, 4736, 4864, 8448, 9600]
for form in forms:
    os.environ["STC_MLP"] = "ts" if form.startswith("ts") else form
    os.environ["STC_MLP_EPI"] = "16" if form == "ts16" else "8"
    eng = capi.Engine(os.path.join(surrogate.ensure_assets("tiny"), "onnx"))
    for M in rows:
        f, u, e = eng.debug_mlp(M, 20)
        print(f"{form:6s} M={M:6d} tiles={-(-M // 128):4d}  fused {f:7.2f} us   two GEMMs {u:7.2f} us   max|diff| {e:.2e}", flush=True)
    eng.close()
