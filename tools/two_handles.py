"""Does splitting a batch across several handles (streams) on ONE GPU beat one handle? Run on the B200 box."""
import os, sys, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from supertonic_b200 import surrogate, tts as T
root = surrogate.ensure_assets("full")
texts, langs, voices = bench.workload(32)
style = T.load_voice_style([os.path.join(root, "voice_styles", v + ".json") for v in voices])
for nh in (1, 2, 4):
    tts = [T.load_text_to_speech(os.path.join(root, "onnx")) for _ in range(nh)]
    order = np.argsort([len(t) for t in texts])
    parts = [sorted(order[i::nh]) for i in range(nh)]         # interleaved by length: equal work per handle
    subs = [([texts[i] for i in p], [langs[i] for i in p], T.Style(style.ttl[p], style.dp[p])) for p in parts]
    def run(k, out):
        t, l, s = subs[k]
        out[k] = tts[k].synthesize_many(t, l, s, 5, 1.05)
    def step():
        out = [None] * nh
        th = [threading.Thread(target=run, args=(k, out)) for k in range(nh)]
        for t in th: t.start()
        for t in th: t.join()
        return out
    for _ in range(4): step()
    t0 = time.perf_counter()
    for _ in range(10): res = step()
    dt = (time.perf_counter() - t0) / 10
    audio = sum(r[1] for o in res for r in o)
    print(f"handles={nh}: {dt*1000:.2f} ms per 32-utterance batch, {audio/dt:.0f} audio-s/s end to end", flush=True)
    for t in tts: t.engine.close()
