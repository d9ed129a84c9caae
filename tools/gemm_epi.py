"""Epilogue cost of the tcgen05 GEMMs on the vocoder shapes: epilogue 0 (bias, fp32 store), 1 (bias + GELU, split-bf16 store),
2 (bias, layer-scale, residual, mask, fp32 store); one-SM tiles (bn 128 / 256) against the two-SM 256 x 256 form (bn 512)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from supertonic_b200 import capi, surrogate
eng = capi.Engine(os.path.join(surrogate.ensure_assets("tiny"), "onnx"))
for name, M, N, K in (("voc.pw1", 27726, 2048, 512), ("voc.pw2", 27726, 512, 2048)):
    for ep in (0, 1, 2):
        for bn in (128, 256, 512):
            us, err = eng.debug_gemm(M, N, K, bn, 1, 1, ep, iters=10)
            print(f"{name} ep={ep} bn={bn:3d}  {us:8.2f} us  {2.0 * M * N * K / us / 1e6:7.1f} TF/s alg  err={err:.1e}", flush=True)
eng.close()
