"""Does splitting a batch across several handles (streams) on ONE GPU beat one handle? Run on the B200 box.
Splits are balanced by latent frames (known from a first pass) so that the fused-MLP grids of the handles add up to <= 148 CTAs
(two handles: 18 + 19 row tiles) and can be resident together."""
import os, sys, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from supertonic_b200 import surrogate, tts as T
root = surrogate.ensure_assets("full")
texts, langs, voices = bench.workload(32)
style = T.load_voice_style([os.path.join(root, "voice_styles", v + ".json") for v in voices])
t0 = T.load_text_to_speech(os.path.join(root, "onnx"))
cs = t0.cfg.chunk_size
frames = [int(-(-int(np.float32(d) * np.float32(t0.sample_rate)) // cs)) for _, d in t0.synthesize_many(texts, langs, style, 5, 1.05)]
t0.engine.close()
print("frames total", sum(frames), "tiles", -(-sum(frames) // 128), flush=True)
for nh in (1, 2, 3):
    tts = [T.load_text_to_speech(os.path.join(root, "onnx")) for _ in range(nh)]
    # greedy balance by frames (largest first)
    parts, load = [[] for _ in range(nh)], [0] * nh
    for i in sorted(range(len(texts)), key=lambda i: -frames[i]):
        k = int(np.argmin(load)); parts[k].append(i); load[k] += frames[i]
    parts = [sorted(p) for p in parts]
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
    t1 = time.perf_counter()
    for _ in range(10): res = step()
    dt = (time.perf_counter() - t1) / 10
    audio = sum(r[1] for o in res for r in o)
    print(f"handles={nh}: frames per handle {load} (tiles {[-(-l // 128) for l in load]}): {dt*1000:.2f} ms per 32-utterance batch, {audio/dt:.0f} audio-s/s end to end", flush=True)
    for t in tts: t.engine.close()
