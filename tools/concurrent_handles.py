"""Does a second batch in flight on the same GPU buy throughput? configs[1] (32 utterances, total_step 5), device-resident inputs,
H handles (own streams, arenas and CUDA graphs each) driven by H host threads, K calls each; wall clock over all calls against
one handle doing them back to back. usage: concurrent_handles.py [handles=2] [calls=20] [batch=32]"""
import os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from supertonic_b200 import capi, surrogate, tts as T

H = int(sys.argv[1]) if len(sys.argv) > 1 else 2
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
root = surrogate.ensure_assets("full")
texts, langs, voices = bench.workload(B, 1234)
style = T.load_voice_style([os.path.join(root, "voice_styles", v + ".json") for v in voices])
engs = [capi.Engine(os.path.join(root, "onnx")) for _ in range(H)]
ids, mask = engs[0].text_to_ids(texts, langs)
lens = mask.reshape(B, -1).sum(1).astype(np.int32)
cs = engs[0].cfg.base_chunk_size * engs[0].cfg.chunk_compress_factor
cap = int(lens.sum() * 0.12 * engs[0].cfg.sample_rate) + (B + 8) * cs
sets = []
for h in range(H):
    sets.append(dict(ids=torch.from_numpy(ids).cuda(), mask=torch.from_numpy(mask).cuda(), ttl=torch.from_numpy(np.ascontiguousarray(style.ttl)).cuda(),
                     dp=torch.from_numpy(np.ascontiguousarray(style.dp)).cuda(), wav=torch.empty(cap, dtype=torch.float32, device="cuda"),
                     dur=torch.empty(B, dtype=torch.float32, device="cuda")))

def run(h, n, seed0):
    s, e = sets[h], engs[h]
    for k in range(n):
        e.synthesize_packed_device(s["ids"].data_ptr(), s["mask"].data_ptr(), s["ttl"].data_ptr(), s["dp"].data_ptr(), B, ids.shape[1], 5, 1.05,
                                   seed0 + k, s["wav"].data_ptr(), cap, s["dur"].data_ptr(), text_lens=lens)
    e.wait() if hasattr(e, "wait") else None

for h in range(H):
    run(h, 3, 0)
torch.cuda.synchronize()
audio = float(sets[0]["dur"].sum().item())
for rep in range(3):
    t0 = time.perf_counter(); run(0, H * K, 100); torch.cuda.synchronize(); t1 = time.perf_counter() - t0
    th = [threading.Thread(target=run, args=(h, K, 100)) for h in range(H)]
    t0 = time.perf_counter(); [t.start() for t in th]; [t.join() for t in th]; torch.cuda.synchronize(); t2 = time.perf_counter() - t0
    print(f"B={B} handles={H}: one handle {1000 * t1 / (H * K):.3f} ms/call = {audio * H * K / t1:.0f} audio-s/s;  {H} in flight {1000 * t2 / (H * K):.3f} ms/call = {audio * H * K / t2:.0f} audio-s/s  ({t1 / t2:.3f} x)", flush=True)
for e in engs:
    e.close()
