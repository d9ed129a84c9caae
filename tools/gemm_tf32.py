"""kind::tf32 vs split-bf16 (3 MMAs) instantiations of the tcgen05 GEMM on the vocoder shapes (run on the B200 box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from supertonic_b200 import capi, surrogate
SHAPES = [("voc.pw1", 27726, 2048, 512, 1), ("voc.pw2", 27726, 512, 2048, 2), ("voc.in", 27726, 512, 168, 0), ("voc.head", 27726, 512, 512, 0)]
for mode in ("bf16x3", "tf32"):
    if mode == "tf32":
        os.environ["STC_DEBUG_TF32"] = "1"
    eng = capi.Engine(os.path.join(surrogate.ensure_assets("tiny"), "onnx"))
    for name, M, N, K, ep in SHAPES:
        for bn in (64, 128, 256):
            if N % bn:
                continue
            us, err = eng.debug_gemm(M, N, K, bn, 1, 1, ep, iters=10)
            tf = 2.0 * M * N * K / us / 1e6
            print(f"{mode:6s} {name:9s} M={M:6d} N={N:5d} K={K:5d} ep={ep} bn={bn:3d}  {us:8.2f} us  {tf:7.1f} TF/s alg  err={err:.2e}", flush=True)
    eng.close()
