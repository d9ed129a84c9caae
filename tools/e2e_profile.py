"""Where does the end-to-end (public API) time go? Run on the B200 box."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from supertonic_b200 import surrogate, tts as T
root = surrogate.ensure_assets("full")
tt = T.load_text_to_speech(os.path.join(root, "onnx"))
texts, langs, voices = bench.workload(32)
style = T.load_voice_style([os.path.join(root, "voice_styles", v + ".json") for v in voices])
eng = tt.engine
for _ in range(3):
    tt.synthesize_many(texts, langs, style, 5, 1.05)
def t(fn, n=10):
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    return (time.perf_counter() - t0) / n * 1000, r
ms, (ids, mask) = t(lambda: eng.text_to_ids(texts, langs))
print(f"text_to_ids            {ms:7.3f} ms")
ms, _ = t(lambda: eng.synthesize_packed(ids, mask, style.ttl, style.dp, 5, 1.05, pinned=True))
print(f"synthesize_packed pin  {ms:7.3f} ms")
ms, _ = t(lambda: eng.synthesize_packed(ids, mask, style.ttl, style.dp, 5, 1.05, pinned=False))
print(f"synthesize_packed page {ms:7.3f} ms")
ms, _ = t(lambda: tt.synthesize_many(texts, langs, style, 5, 1.05))
print(f"synthesize_many        {ms:7.3f} ms")
eng.set_profile(1)
eng.synthesize_packed(ids, mask, style.ttl, style.dp, 5, 1.05, pinned=True)
print("stage ms", eng.stage_ms())
eng.set_profile(0)
