"""Depthwise conv + LayerNorm: register sliding-window kernel vs shared-memory tiled kernel (run on the B200 box).
Algorithmic bytes = 8*rows*C (fp32 in, split-bf16 out); GB/s is quoted against them.
usage: dwconv_sweep.py [rt,rt,...] [ring modes, e.g. -1,0,1]   (rt 0 = the library's heuristic; ring -1 = by chain length)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from supertonic_b200 import capi, surrogate
shapes = [(27726, 512, 7, 1, 32), (27726, 512, 7, 2, 32), (27726, 512, 7, 4, 32),
          (4736, 256, 5, 1, 32), (4736, 256, 5, 2, 32), (4736, 256, 5, 4, 32), (4736, 256, 5, 8, 32),
          (1280, 256, 5, 1, 32), (160, 256, 5, 1, 1), (896, 512, 7, 1, 1), (4736, 128, 5, 1, 32)]
rts = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 4, 8, 16, 32, 48, 64]
rings = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [-1]
for ring in rings:
    os.environ["STC_DW_RING"] = str(ring)
    eng = capi.Engine(os.path.join(surrogate.ensure_assets("tiny"), "onnx"))
    for rows, C, K, dil, B in shapes:
        for rt in rts:
            s, t, e = eng.debug_dwconv(rows, C, K, dil, False, B, rt, 20)
            gb = 8.0 * rows * C / 1e3
            print(f"ring={ring:2d} rows={rows:6d} C={C:3d} K={K} dil={dil} rt={rt:3d}  slide {s:7.2f} us ({gb / s:7.1f} GB/s)   "
                  f"tile {t:7.2f} us ({gb / t:7.1f} GB/s)   max|diff| {e:.2e}", flush=True)
    eng.close()
