"""One forced-configuration tcgen05 GEMM (for ncu): python tools/gemm_one.py M N K bn epilogue [f16|bf16x3] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
M, N, K, bn, ep = (int(v) for v in sys.argv[1:6])
mode = sys.argv[6] if len(sys.argv) > 6 else "bf16x3"
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 10
if mode == "f16":
    os.environ["STC_DEBUG_F16"] = "1"
from supertonic_b200 import capi, surrogate
eng = capi.Engine(os.path.join(surrogate.ensure_assets("tiny"), "onnx"))
us, err = eng.debug_gemm(M, N, K, bn, 2 if bn == 512 else 1, 1, ep, iters=iters)
print(f"{mode} M={M} N={N} K={K} ep={ep} bn={bn}: {us:.2f} us, {2.0 * M * N * K / us / 1e6:.1f} TF/s alg, err={err:.2e}")
eng.close()
