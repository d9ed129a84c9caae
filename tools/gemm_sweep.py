"""GPU tuning sweep for the tcgen05 GEMM (run on the B200 box): every tile width (64 / 128 / 256 one-SM, 512 = two-SM 256 x 256) on the hot shapes; cluster shapes other than 1 x 1 were measured
slower in round 1 and are no longer reachable.
usage: python tools/gemm_sweep.py [quick|full] > gpurun_out/gemm_sweep.txt"""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from supertonic_b200 import capi, surrogate  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
eng = capi.Engine(os.path.join(surrogate.ensure_assets("tiny"), "onnx"))
SHAPES = [  # (name, M, N, K, epilogue)
    ("ve.pw1", 4736, 1024, 256, 1), ("ve.pw2", 4736, 256, 1024, 2), ("ve.qo", 4736, 256, 256, 0), ("ve.out", 4736, 144, 256, 2),
    ("voc.pw1", 27726, 2048, 512, 1), ("voc.pw2", 27726, 512, 2048, 2), ("voc.head", 27726, 512, 512, 0),
    ("te.pw1", 9600, 1024, 256, 1), ("te.pw2", 9600, 256, 1024, 2), ("b1.pw1", 140, 1024, 256, 1), ("b1.pw2", 140, 256, 1024, 2),
]
CL = [(1, 1)]
if mode == "quick":
    SHAPES = SHAPES[:2]
if mode == "one":       # a single configuration, for ncu
    name, M, N, K, ep = [x for x in SHAPES if x[0] == sys.argv[2]][0]
    bn, cm, cn = (int(v) for v in sys.argv[3:6])
    print(eng.debug_gemm(M, N, K, bn, cm, cn, ep, iters=2))
    eng.close()
    sys.exit(0)
for name, M, N, K, ep in SHAPES:
    best = None
    for bn, (cm, cn) in itertools.product((64, 128, 256, 512), CL):
        if (128 // cn) % 8 or (bn // cm) % 8:
            continue
        try:
            us, err = eng.debug_gemm(M, N, K, bn, cm, cn, ep, iters=10 if M > 20000 else 30)
        except capi.StcError as e:
            print(f"{name:9s} M={M} N={N} K={K} bn={bn} c={cm}x{cn}: ERROR {e}", flush=True)
            continue
        tf = 2.0 * M * N * K / us / 1e6
        print(f"{name:9s} M={M:6d} N={N:5d} K={K:5d} ep={ep} bn={bn:3d} c={cm}x{cn}  {us:8.2f} us  {tf:7.1f} TF/s alg ({3 * tf:7.1f} executed)  err={err:.2e}", flush=True)
        if err < 1e-3 and (best is None or us < best[0]):
            best = (us, bn, cm, cn)
    print(f"# best {name}: {best}", flush=True)
eng.close()
