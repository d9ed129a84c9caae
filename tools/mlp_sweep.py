"""Fused ConvNeXt MLP (mlp_stream.cuh + reduce kernel) vs pw1 / pw2 as two tcgen05 GEMMs over row counts (run on the B200 box).
usage: mlp_sweep.py [rows,rows,...] [--slices 1,2,4,8] [--pair 0,1]
  --slices: force the hidden-slice count (STC_MLP_SLICES) instead of the cost model; --pair: one-CTA kernel (0) / CTA pairs (1)
  (STC_MLP_TRACE=1 prints CTA 0's pipeline stamps for each row count)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from supertonic_b200 import capi, surrogate

def opt(name, default):
    if name in sys.argv:
        return [int(v) for v in sys.argv[sys.argv.index(name) + 1].split(",")]
    return default

args = [a for a in sys.argv[1:] if a[0].isdigit() and (sys.argv[sys.argv.index(a) - 1] not in ("--slices", "--pair"))]
rows = [int(x) for x in args[0].split(",")] if args else [128, 1152, 2432, 3712, 4736, 4864, 5632, 6272, 6400, 9472, 9600, 12800, 18560, 37120]
onnx = os.path.join(surrogate.ensure_assets("tiny"), "onnx")
for pair in opt("--pair", [1]):
    for ns in opt("--slices", [0]):
        os.environ["STC_MLP_PAIR"] = str(pair)
        os.environ["STC_MLP_SLICES"] = str(ns)
        eng = capi.Engine(onnx)
        for M in rows:
            f, u, e = eng.debug_mlp(M, 20)
            print(f"stream pair={pair} slices={ns if ns else 'auto':>4} M={M:6d} tiles={-(-M // 128):4d}  fused {f:7.2f} us   two GEMMs {u:7.2f} us   max|diff| {e:.2e}", flush=True)
        eng.close()
