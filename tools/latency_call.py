"""Batch-1 latency path: TextToSpeech.call() on the reference's default sentence, n times (for ncu launch lists)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from supertonic_b200 import surrogate, tts as T
root = surrogate.ensure_assets("full")
tt = T.load_text_to_speech(os.path.join(root, "onnx"))
one = T.load_voice_style([os.path.join(root, "voice_styles", "M1.json")])
s = ("This morning, I took a walk in the park, and the sound of the birds and the breeze was so pleasant that "
     "I stopped for a long time just to listen.")
ts = []
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    t0 = time.perf_counter(); r = tt.call(s, "en", one, 5, 1.05); ts.append(time.perf_counter() - t0)
print("call ms:", [round(1000 * t, 2) for t in ts], "audio s", float(r.duration[0]), "launches", tt.engine.launches)
