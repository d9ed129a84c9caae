"""Single-pass fp16 vs split-bf16 (3 MMAs) instantiations of the tcgen05 GEMMs on the vocoder shapes (run on the B200 box).
bn = 512 is the two-SM (cta_group::2) 256 x 256 form. Odd shapes check the K tails (K % 128 in 8..120) and ragged M."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from supertonic_b200 import capi, surrogate
SHAPES = [("voc.pw1", 27726, 2048, 512, 1), ("voc.pw2", 27726, 512, 2048, 2), ("voc.in", 27726, 512, 168, 0), ("voc.head", 27726, 512, 512, 0),
          ("odd.k72", 333, 256, 72, 1), ("odd.k200", 1000, 512, 200, 2), ("odd.k64", 130, 64, 64, 0)]
for mode in ("bf16x3", "f16"):
    if mode == "f16":
        os.environ["STC_DEBUG_F16"] = "1"
    eng = capi.Engine(os.path.join(surrogate.ensure_assets("tiny"), "onnx"))
    for name, M, N, K, ep in SHAPES:
        for bn in (64, 128, 256, 512):
            if N % min(bn, 256) or (bn == 512 and K < 64):
                continue
            us, err = eng.debug_gemm(M, N, K, bn, 2 if bn == 512 else 1, 1, ep, iters=10)
            tf = 2.0 * M * N * K / us / 1e6
            print(f"{mode:6s} {name:9s} M={M:6d} N={N:5d} K={K:5d} ep={ep} bn={bn:3d}  {us:8.2f} us  {tf:7.1f} TF/s alg  err={err:.2e}", flush=True)
    eng.close()
