"""configs[4] through the IN-PROCESS multi-GPU engine (tts.MultiGpuTextToSpeech: one engine + host thread per device, one plan,
LPT over launch groups): 1 024 synthetic utterances, end to end (host text front-end, H2D, synthesis, D2H into page-locked
buffers), wall clock, for 1 / 2 / 4 / 8 devices of the box. usage: multi_gpu_many.py [n_utt] [reps] [max_batch] [pcm16]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                    # noqa: E402
from supertonic_b200 import surrogate, tts as T                # noqa: E402


def main():
    import torch
    n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    max_batch = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    pcm16 = len(sys.argv) > 4 and sys.argv[4] == "pcm16"
    ndev_max = torch.cuda.device_count()
    root = surrogate.ensure_assets("full")
    texts, langs, voices = bench.workload(n_utt, 1234)
    style = T.load_voice_style([os.path.join(root, "voice_styles", v + ".json") for v in voices])
    rows = []
    base = None
    for nd in (1, 2, 4, 8):
        if nd > ndev_max:
            break
        eng = T.MultiGpuTextToSpeech(os.path.join(root, "onnx"), list(range(nd)))
        for w in range(2):
            res = eng.synthesize_many(texts, langs, style, 5, 1.05, max_batch=max_batch, seed=w, pcm16=pcm16)
        audio = float(sum(d for _, d in res))
        ts = []
        for r in range(reps):
            t0 = time.perf_counter()
            res = eng.synthesize_many(texts, langs, style, 5, 1.05, max_batch=max_batch, seed=10 + r, pcm16=pcm16)
            ts.append(time.perf_counter() - t0)
        t = float(np.median(ts))
        if base is None:
            base = audio / t
        rows.append({"devices": nd, "max_batch": max_batch, "pcm16": pcm16, "ms_per_request": 1000 * t, "audio_s_per_s": audio / t, "speedup_vs_1": (audio / t) / base,
                     "efficiency": (audio / t) / base / nd})
        print(json.dumps(rows[-1]), flush=True)
        eng.close()
    print(json.dumps({"workload": f"configs[4]: {n_utt} utterances (seed 1234), total_step 5, one process, MultiGpuTextToSpeech.synthesize_many "
                                  "(front-end + H2D + synthesis + D2H inside the timed region)", "rows": rows}))


if __name__ == "__main__":
    main()
