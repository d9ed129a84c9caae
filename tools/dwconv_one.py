"""One depthwise conv + LayerNorm shape, few iterations (for ncu): dwconv_one.py rows C K dil [rt] [iters]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from supertonic_b200 import capi, surrogate
rows, C, K, dil = (int(x) for x in sys.argv[1:5])
rt = int(sys.argv[5]) if len(sys.argv) > 5 else 0
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 3
eng = capi.Engine(os.path.join(surrogate.ensure_assets("tiny"), "onnx"))
s, t, e = eng.debug_dwconv(rows, C, K, dil, True, 32, rt, iters)
print(f"rows={rows} C={C} K={K} dil={dil} rt={rt}: {s:.2f} us ({8.0 * rows * C / 1e3 / s:.0f} GB/s at 8 B/elem), tile {t:.2f} us, max|diff| {e:.2e}")
eng.close()
