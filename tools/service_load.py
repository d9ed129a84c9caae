"""Closed-loop load on the HTTP serving wrapper (in-process TestClient, run on the B200 box): `clients` threads each POST
`per_client` single-utterance /tts requests back to back; prints audio-s/s, requests/s, p50/p90 request latency and what the
dynamic batcher coalesced. usage: service_load.py [clients=32] [per_client=20] [max_wait_ms=2] [full|tiny]"""
import json, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from starlette.testclient import TestClient
from supertonic_b200 import service, surrogate, tts as T
from bench import workload

clients = int(sys.argv[1]) if len(sys.argv) > 1 else 32
per_client = int(sys.argv[2]) if len(sys.argv) > 2 else 20
max_wait = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
root = surrogate.ensure_assets(sys.argv[4] if len(sys.argv) > 4 else "full")
tt = T.load_text_to_speech(root + "/onnx")
app = service.create_app(tt, max_batch=64, max_wait_ms=max_wait)
texts, langs, voices = workload(clients * per_client, 99)
lat, audio = [], [0.0]
lock = threading.Lock()
with TestClient(app) as c:
    def run(k, warm):
        for i in range(2 if warm else per_client):
            j = k * per_client + i
            t0 = time.perf_counter()
            r = c.post("/tts", json={"text": texts[j], "lang": langs[j], "voice_style": os.path.join(root, "voice_styles", voices[j] + ".json")})
            dt = time.perf_counter() - t0
            assert r.status_code == 200, r.text
            if not warm:
                with lock:
                    lat.append(dt); audio[0] += (len(r.content) - 44) / 2 / tt.sample_rate
    for warm in (True, False):
        th = [threading.Thread(target=run, args=(k, warm)) for k in range(clients)]
        t0 = time.perf_counter()
        [t.start() for t in th]; [t.join() for t in th]
        wall = time.perf_counter() - t0
    st = c.get("/stats").json()
app.state.batcher.close()
print(json.dumps({"clients": clients, "requests": len(lat), "wall_s": wall, "audio_s_per_s": audio[0] / wall, "requests_per_s": len(lat) / wall,
                  "p50_ms": 1000 * float(np.median(lat)), "p90_ms": 1000 * float(np.quantile(lat, 0.9)), "max_wait_ms": max_wait,
                  "batcher": st, "note": "in-process starlette TestClient (no sockets); WAV encoding and JSON parsing on the request threads"}))
