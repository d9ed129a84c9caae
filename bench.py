#!/usr/bin/env python
"""bench.py — audio-seconds per wall-second of the Supertonic synthesis forward pass (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[1] — a batch of 32 mixed-length English utterances
(character lengths uniform{20..300}, numpy default_rng(1234), README lengths 59/152/266 included),
total_step=5, speed=1.05, voices cycling M1/F1/M2/F2, surrogate full-size graphs (the released weights are not
mounted: SURVEY.md §0). One "step" = one pass of the hot path over that batch; with N GPUs every rank
synthesises that same configs[1] batch with its own noise seeds (weak scaling: the per-GPU work is FIXED as N grows; replicas
only, no collective on the data path). --vary-batches gives every rank its own draw of 32 utterances instead (the total
latent frames then differ by +-15 % between ranks, and a rank whose frames need 38+ row tiles takes two waves of the fused
MLP kernel instead of one — DESIGN.md §6); `per_rank` in the JSON line lists each rank's frames and time either way.

  value : device-resident leg — text_ids/masks/styles already in HBM, stc_synthesize_device per length bucket,
          CUDA events on the library's stream, L2 flushed (untimed) between steps.
  e2e   : same work through the public API (TextToSpeech.synthesize_many): host text front-end, H2D of the
          inputs and D2H of every waveform inside the timed region; steps are issued as a request stream (the copy of step k
          overlaps the computation of step k+1; everything has landed when the timed region ends).
  --impl reference : the reference's CPU path restated (oracle/, torch-CPU ONNX interpreter on all host cores;
          ONNX Runtime itself is not installable here — DESIGN.md) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

README_LENGTHS = (59, 152, 266)          # reference README.md:192
WORDS = ("the quick brown fox jumps over a lazy dog while seven silver ships sail south toward quiet harbours and "
         "nobody knows why morning light feels warmer after rain or how distant thunder rolls across open fields "
         "yesterday we walked along the river talking about music science and old friends from school").split()


def workload(n=32, seed=1234):
    rng = np.random.default_rng(seed)
    lens = [int(x) for x in rng.integers(20, 301, size=n)]
    for i, v in enumerate(README_LENGTHS):
        if i < n:
            lens[i] = v
    texts = []
    for ln in lens:
        out = []
        while sum(len(w) + 1 for w in out) < ln:
            out.append(WORDS[rng.integers(len(WORDS))])
        s = " ".join(out)[:ln].strip()
        texts.append(s[0].upper() + s[1:])
    return texts, ["en"] * n, [("M1", "F1", "M2", "F2")[i % 4] for i in range(n)]


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md 'clocks' line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.p = [], None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=j["hbm_gbs"], bf16=j["bf16_tflops"], bf16_sustained=j["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


def oracle_run(texts, langs, voices, total_step, reps):
    """CPU arm: the oracle port of the reference `_infer` (one padded batch, like TextToSpeech::batch)."""
    import torch
    from oracle.pipeline import best_oracle, make_noise, ort_version
    from supertonic_b200 import assets
    root, _ = assets.asset_root("full")
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ora, kind = best_oracle(root)           # ONNX Runtime CPU if it can be imported here, else the torch-CPU interpreter port
    ttl, dp = ora.style(voices)
    times, audio = [], 0.0
    for r in range(reps):
        t0 = time.perf_counter()
        wav, dur = ora.batch(texts, langs, ttl, dp, total_step, 1.05, make_noise(r))
        times.append(time.perf_counter() - t0)
        audio = float(dur.sum())
    runtime = (f"onnxruntime {ort_version()} CPU execution provider" if kind == "ort"
               else "oracle/ torch-CPU ONNX interpreter (not ONNX Runtime: it cannot be installed here)")
    return audio, times, cores, kind, runtime


def parity_check(eng, root, ids, mask, style, total_step, which):
    """Outside the timed region: the timed batch once more through stc_synthesize_packed with INJECTED noise (same kernels, same
    row-tile count), and the utterances `which` of it against the oracle (oracle/: the CPU restatement of the reference's `_infer`)
    run on each of them alone with the same noise rows."""
    from oracle.pipeline import best_oracle
    ora, okind = best_oracle(root)
    B = ids.shape[0]
    nz = np.random.default_rng(4321).standard_normal((B, eng.cfg.latent_channels, 420)).astype(np.float32)
    out = eng.synthesize_packed(ids, mask, style.ttl, style.dp, total_step, 1.05, noise=nz, want_latent=True)
    worst_err, worst_snr, exact = 0.0, 1e9, True
    for b in which:
        t = int(mask[b].sum())
        tr = {}
        wav_ref, dur_ref = ora.infer_ids(ids[b:b + 1, :t], mask[b:b + 1, :, :t], style.ttl[b:b + 1], style.dp[b:b + 1], total_step,
                                         np.float32(1.05), lambda B_, D_, L_, b=b: nz[b:b + 1, :, :L_], tr)
        n = int(tr["wav_lengths"][0])
        exact = exact and bool(out["duration"][b] == dur_ref[0]) and int(out["wav_lengths"][b]) == n and int(out["frames"][b]) == tr["latent_len"]
        worst_err = max(worst_err, float(np.abs(out["latent"][b].T - tr["xs"][-1][0]).max()))
        ref = wav_ref[:n].astype(np.float64)
        err = ((out["wavs"][b].astype(np.float64) - ref) ** 2).sum()
        worst_snr = min(worst_snr, 200.0 if err == 0 else float(10 * np.log10((ref ** 2).sum() / err)))
    return {"utterances": [int(b) for b in which], "durations_and_frame_counts_bit_exact": exact, "latent_max_abs": worst_err,
            "snr_db": worst_snr, "against": ("ONNX Runtime CPU (the reference's own runtime)" if okind == "ort" else "oracle/ (CPU restatement of the reference _infer)") + ", same injected noise",
            "tolerance": {"latent_max_abs": 2e-4, "snr_db": 60.0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--total-step", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the configs[4] strong-scaling sub-object")
    ap.add_argument("--no-vary", action="store_true", help="skip the varied-draw sub-object")
    ap.add_argument("--lanes", type=int, default=2, help="handles per GPU for the end-to-end request stream (TextToSpeech lanes); 1 = one handle")
    ap.add_argument("--vary-batches", action="store_true", help="N > 1: every rank draws its own 32 utterances (seed 1234 + 1000 rank)")
    ap.add_argument("--workload", default="batch", choices=["batch", "sweep1024"],
                    help="batch: configs[1], every GPU its own --batch utterances (weak scaling, the default and the headline). "
                         "sweep1024: configs[4], 1024 utterances sharded over the GPUs in groups of 128 (strong scaling)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lrank = int(os.environ.get("LOCAL_RANK", 0))
    cfg = {"workload": f"configs[1]: batch {a.batch} mixed-length English utterances (chars uniform 20..300, seed 1234), "
                       f"total_step={a.total_step}, speed=1.05, per GPU", "weights": "surrogate full-size graphs (random init, seed 0)",
           "l2": "flushed (512 MiB write) between timed steps", "parallelism": f"replicas x{world}, utterance-sharded, no collective",
           "layout": "latent frames and text tokens both as packed rows (no padded frames or tokens are computed)"}

    if a.impl == "reference":
        if rank != 0:
            return
        texts, langs, voices = workload(a.batch)
        k = min(a.cpu_sample, a.batch)
        idx = list(np.linspace(0, a.batch - 1, k).astype(int))
        st, sl, sv = [texts[i] for i in idx], [langs[i] for i in idx], [voices[i] for i in idx]
        audio, times, cores, kind, runtime = oracle_run(st, sl, sv, a.total_step, a.warmup + a.steps)
        t = times[a.warmup:]
        ms = 1000 * float(np.mean(t))
        val = audio / (ms / 1000)
        sample = f"{k} of the {a.batch} utterances (evenly spaced by index) as one padded batch per step"
        cfg = dict(cfg, workload=cfg["workload"] + f" — CPU arm: a bounded SAMPLE of it, {sample}", reference_runtime=runtime)
        print(json.dumps({"impl": "reference", "metric": "audio-sec/sec", "value": val, "unit": "audio-s/s", "n_gpus": a.gpus,
                          "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                          "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": kind, "sample": sample, "runtime": runtime},
                          "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    from supertonic_b200 import capi, surrogate, tts as T
    from supertonic_b200.scheduler import shard_for_rank
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    torch.cuda.set_device(lrank)
    bound = None
    if world > 1 and os.environ.get("STC_BIND", "1") != "0":
        from supertonic_b200 import affinity
        bound = affinity.bind_to_gpu(lrank)          # before any page-locked buffer exists: they land in the GPU's NUMA node
    cfg["host_binding"] = f"rank 0 bound to {len(bound)} CPUs local to its GPU" if bound else "none"
    from supertonic_b200 import assets
    if lrank == 0:
        root, asset_kind = assets.asset_root("full")
    if world > 1:
        dist.barrier()
    root, asset_kind = assets.asset_root("full")
    cfg["weights"] = ("released assets at " + root) if asset_kind == "released" else cfg["weights"]
    tt = T.load_text_to_speech(os.path.join(root, "onnx"), use_gpu=True, device=lrank, lanes=a.lanes)
    eng = tt.engine
    ext = torch.cuda.ExternalStream(eng.stream)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    cs = eng.cfg.chunk_size

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(texts, langs, voices, group, steps, warmup, clocks=False):
        """Device-resident leg + end-to-end leg of one workload on this rank's utterances; reductions over ranks inside."""
        n_utt = len(texts)
        style = T.load_voice_style([os.path.join(root, "voice_styles", v + ".json") for v in voices])
        # ---- device-resident leg ---------------------------------------------------------------------
        # the launch groups are the product's own (tts.plan_many: at most `group` utterances and 140 row tiles of predicted latent
        # frames per group, equal predicted frames); one untimed request first, so that the plan uses the measured frames per token
        tt.synthesize_many(texts, langs, style, a.total_step, 1.05, max_batch=group)
        plan = T.plan_many(eng, texts, langs, group)
        ids, mask, lens = plan.ids, plan.mask, plan.lens
        buckets = []
        for grp in plan.groups:                                   # inside a group everything is packed rows
            g = np.asarray(grp); Tg = int(lens[g].max())
            cap = int(lens[g].sum() * 0.12 * eng.cfg.sample_rate) + (len(g) + 8) * cs
            buckets.append(dict(B=len(g), T=Tg, cap=cap,
                                ids=torch.from_numpy(np.ascontiguousarray(ids[g, :Tg])).cuda(),
                                mask=torch.from_numpy(np.ascontiguousarray(mask[g, :, :Tg])).cuda(),
                                lens=np.ascontiguousarray(lens[g], dtype=np.int32),
                                ttl=torch.from_numpy(np.ascontiguousarray(style.ttl[g])).cuda(),
                                dp=torch.from_numpy(np.ascontiguousarray(style.dp[g])).cuda(),
                                wav=torch.empty(cap, dtype=torch.float32, device="cuda"),
                                dur=torch.empty(len(g), dtype=torch.float32, device="cuda")))

        def device_step(seed):
            for b in buckets:
                for attempt in range(2):
                    try:
                        off = eng.synthesize_packed_device(b["ids"].data_ptr(), b["mask"].data_ptr(), b["ttl"].data_ptr(), b["dp"].data_ptr(),
                                                           b["B"], b["T"], a.total_step, 1.05, seed, b["wav"].data_ptr(), b["cap"],
                                                           b["dur"].data_ptr(), text_lens=b["lens"])
                        break
                    except capi.StcError as e:          # result buffer too small for the (128-row bucketed) frame count: grow once
                        if e.code != capi.ERR_CAPACITY or attempt:
                            raise
                        b["cap"] = int(e.need)
                        b["wav"] = torch.empty(b["cap"], dtype=torch.float32, device="cuda")
                b["L"] = int(off[-1] // cs)

        seed0 = 100 + 1000 * rank
        for w in range(warmup):
            device_step(w)
        audio = float(sum(b["dur"].sum().item() for b in buckets))
        barrier()
        sampler = ClockSampler(lrank) if (rank == 0 and clocks) else None
        l0 = eng.launches
        step_ms = []
        for k in range(steps):
            flush.fill_(k & 0xFF)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(ext):
                e0.record()
                device_step(seed0 + k)
                e1.record()
            e1.synchronize()
            step_ms.append(e0.elapsed_time(e1))
        barrier()
        launches = eng.launches - l0
        clk = sampler.stop() if sampler else None
        t = torch.tensor([float(np.sum(step_ms)), audio], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        if world > 1:
            dist.all_reduce(tmax[0:1], op=dist.ReduceOp.MAX)
            dist.all_reduce(t[1:2], op=dist.ReduceOp.SUM)
        total_ms, audio_all = float(tmax[0].item()), float(t[1].item())
        mine = torch.tensor([float(np.sum(step_ms)) / steps, audio, float(sum(b.get("L", 0) for b in buckets))], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        if world > 1:
            dist.all_gather(allr, mine)
        else:
            allr = [mine]
        per_rank = [{"rank": i, "ms_per_step": float(v[0]), "audio_s_per_step": float(v[1]), "latent_frames": int(v[2]),
                     "row_tiles_of_128": int(-(-int(v[2]) // 128))} for i, v in enumerate(allr)]
        ms_per_step = total_ms / steps
        # ---- end-to-end leg: public API, host buffers, front-end + H2D + D2H inside the timed region ---------
        for w in range(max(2 * a.lanes, warmup - 1)):          # both alternating pinned result sets (and CUDA graphs) of every lane exist before the timed region
            tt.synthesize_many(texts, langs, style, a.total_step, 1.05, max_batch=group, wait=False)
        tt.wait()
        # Three windows of exactly K steps each, the MEDIAN window is reported (all three are in `windows_ms_per_step`): this leg is
        # wall-clock on the host (front-end threads, pinned copies) and a single hiccup of a shared box moved a 100 ms window by 20 %.
        wins = []
        for rep in range(3):
            barrier()
            t0 = time.perf_counter()
            for k in range(steps):
                # request stream: step k+1 is issued before step k's waveform copy has landed (two alternating pinned result sets)
                res = tt.synthesize_many(texts, langs, style, a.total_step, 1.05, max_batch=group, seed=k, wait=False)
            tt.wait()
            torch.cuda.synchronize()
            wins.append(time.perf_counter() - t0)
        tw = torch.tensor(wins, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)          # a window ends when its slowest rank ends
        wins = [float(v) for v in tw.tolist()]
        e2e_s = sorted(wins)[1]
        e2e_audio = float(sum(r[1] for r in res))
        d2h = int(sum(b.get("L", 0) * cs * 4 + b["B"] * 12 for b in buckets))
        h2d = int(sum(b["ids"].numel() * 8 + b["mask"].numel() * 4 + b["ttl"].numel() * 4 + b["dp"].numel() * 4 for b in buckets))
        te = torch.tensor([e2e_audio], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.SUM)
        e2e = {"value": float(te[0].item()) / (e2e_s / steps), "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": 1000 * e2e_s / steps, "windows_ms_per_step": [1000 * w / steps for w in wins],
               "window": "median of three windows of exactly `steps` steps each (max over ranks per window)",
               "request_lanes": a.lanes,
               "how": "TextToSpeech.synthesize_many(wait=False) request stream: host text front-end, H2D, synthesis, D2H into page-locked buffers; "
                      f"{a.lanes} handle(s) per GPU take the launch groups alternately, so consecutive requests overlap on the device"}
        # the same request stream with 16-bit PCM results (quantised on the device like writeWavFile, cpp/helper.cpp:985-988): what a
        # server that writes WAV needs, at half the device->host bytes — at N = 8 the float32 waveforms of all ranks (1.7 GB per pass of
        # configs[4]) are what the end-to-end leg waits for on this box (~50 GB/s aggregate D2H)
        for w in range(2 * a.lanes):
            tt.synthesize_many(texts, langs, style, a.total_step, 1.05, max_batch=group, pcm16=True, wait=False)
        tt.wait()
        barrier()
        t0 = time.perf_counter()
        for k in range(steps):
            res16 = tt.synthesize_many(texts, langs, style, a.total_step, 1.05, max_batch=group, seed=k, wait=False, pcm16=True)
        tt.wait()
        torch.cuda.synchronize()
        t16 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t16, op=dist.ReduceOp.MAX)
        e2e["pcm16"] = {"value": float(te[0].item()) / (float(t16[0].item()) / steps), "unit": "audio-s/s", "d2h_bytes_per_step": d2h // 2,
                        "ms_per_step": 1000 * float(t16[0].item()) / steps, "window": "one window of `steps` steps"}
        return dict(value=audio_all / (ms_per_step / 1000), ms_per_step=ms_per_step, audio=audio, audio_all=audio_all, launches=int(launches),
                    clocks=clk, per_rank=per_rank, e2e=e2e, p50_step_ms=float(np.median(step_ms)), step_ms=[float(v) for v in step_ms], buckets=buckets, style=style,
                    ids=ids, mask=mask, lens=lens)

    if a.workload == "sweep1024":
        texts, langs, voices = workload(1024, 1234)
        mine = shard_for_rank([len(t) + 9 for t in texts], a.total_step, rank, world)      # same plan on every rank, no communication
        texts, langs, voices = [texts[i] for i in mine], [langs[i] for i in mine], [voices[i] for i in mine]
        group = 128
        cfg["workload"] = (f"configs[4]: 1024 synthetic utterances (chars uniform 20..300, seed 1234) sharded over {world} GPU(s) by LPT, "
                           f"launch groups of <= {group} utterances and equal predicted latent frames (tts.plan_many), total_step={a.total_step}, speed=1.05")
    else:
        texts, langs, voices = workload(a.batch, 1234 + (1000 * rank if a.vary_batches else 0))
        group = a.batch
        if world > 1:
            cfg["workload"] += (" (every rank its own draw, seed 1234 + 1000 rank)" if a.vary_batches
                                else " (the same batch on every rank, rank-specific noise seeds)")
    cfg["working_set"] = "per step ~0.3 GB of weights + ~0.6 GB of activations per GPU (>> the 126 MB L2), and the L2 is flushed between steps"
    main = measure(texts, langs, voices, group, a.steps, a.warmup, clocks=True)
    value, ms_per_step, audio, launches, clocks, per_rank = main["value"], main["ms_per_step"], main["audio"], main["launches"], main["clocks"], main["per_rank"]
    buckets, style, ids, mask, lens, n_utt = main["buckets"], main["style"], main["ids"], main["mask"], main["lens"], len(texts)

    # ---- strong-scaling sub-object: configs[4] (1024 utterances sharded over the ranks by LPT, groups of 128) at the same N ----
    strong = None
    if a.workload == "batch" and not a.no_strong:
        t4, l4, v4 = workload(1024, 1234)
        mine4 = shard_for_rank([len(t) + 9 for t in t4], a.total_step, rank, world)
        steps4 = max(2, a.steps // 2)           # (two passes were too few at N = 8: one slow pass on one rank moved the line by 10 %)
        m4 = measure([t4[i] for i in mine4], [l4[i] for i in mine4], [v4[i] for i in mine4], 128, steps4, 2)
        strong = {"workload": f"configs[4]: 1024 synthetic utterances (seed 1234) sharded over {world} GPU(s) by LPT (scheduler.shard_for_rank), "
                              f"packed launch groups of <= 128 utterances and equal predicted latent frames (tts.plan_many), total_step={a.total_step}", "scaling": "strong", "value": m4["value"], "unit": "audio-s/s",
                  "ms_per_pass": m4["ms_per_step"], "e2e": m4["e2e"], "per_rank": m4["per_rank"], "steps": steps4,
                  "groups": [[b["B"], b["T"], b.get("L", 0)] for b in m4["buckets"]]}

    def device_ms(tv, lv, vv, total_step, reps=3):
        """Median device time (CUDA events on the library's stream, L2 flushed) of one device-resident synthesis of the given utterances."""
        sv = T.load_voice_style([os.path.join(root, "voice_styles", v + ".json") for v in vv])
        iv, mv = eng.text_to_ids(tv, lv)
        dv = {k: torch.from_numpy(np.ascontiguousarray(x)).cuda() for k, x in dict(ids=iv, mask=mv, ttl=sv.ttl, dp=sv.dp).items()}
        lv32 = mv.reshape(len(tv), -1).sum(1).astype(np.int32)
        st = dict(cap=int(lv32.sum() * 0.12 * eng.cfg.sample_rate) + (len(tv) + 8) * cs)
        st["wav"] = torch.empty(st["cap"], dtype=torch.float32, device="cuda")
        duv = torch.empty(len(tv), dtype=torch.float32, device="cuda")

        def one(seed):
            for attempt in range(2):
                try:
                    return eng.synthesize_packed_device(dv["ids"].data_ptr(), dv["mask"].data_ptr(), dv["ttl"].data_ptr(), dv["dp"].data_ptr(), len(tv),
                                                        iv.shape[1], total_step, 1.05, seed, st["wav"].data_ptr(), st["cap"], duv.data_ptr(), text_lens=lv32)
                except capi.StcError as e:
                    if e.code != capi.ERR_CAPACITY or attempt:
                        raise
                    st["cap"] = int(e.need); st["wav"] = torch.empty(st["cap"], dtype=torch.float32, device="cuda")
        for w in range(2):
            off = one(w)
        ts = []
        for k in range(reps):
            flush.fill_(k); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(ext):
                e0.record(); off = one(10 + k); e1.record()
            e1.synchronize(); ts.append(e0.elapsed_time(e1))
        return float(np.median(ts)), int(off[-1] // cs), float(duv.sum().item())

    # ---- varied draws (rank 0 only at N = 1; per rank at N > 1): every seed another 32-utterance batch, i.e. another row-tile count ----
    vary = None
    if a.workload == "batch" and not a.no_vary and rank == 0:
        rows_v = []
        for sd in range(8):
            tv, lv, vv = workload(a.batch, 2234 + 1000 * sd)
            ms, fr, au = device_ms(tv, lv, vv, a.total_step)
            rows_v.append({"seed": 2234 + 1000 * sd, "latent_frames": fr, "row_tiles_of_128": -(-fr // 128), "ms_per_step": ms, "audio_s": au})
        msv = [r["ms_per_step"] for r in rows_v]
        vary = {"what": "eight other draws of the 32-utterance batch (seeds 2234 + 1000 k), device-resident leg, median of 3 steps each",
                "draws": rows_v, "mean_ms_per_step": float(np.mean(msv)), "worst_ms_per_step": float(np.max(msv)),
                "mean_audio_s_per_s": float(np.mean([r["audio_s"] / (r["ms_per_step"] / 1000) for r in rows_v])),
                "worst_over_headline": float(np.max(msv) / ms_per_step)}

    # ---- total_step sweep (rank 0): configs[2] (en/ko/es/pt/fr in one batch, one voice style per utterance, total_step 2/5/10/20) and the
    #      same sweep on the headline batch; the slope over total_step is the device time of ONE Euler step of the vector estimator,
    #      the intercept what surrounds the loop (duration predictor || text encoder, vocoder) ----
    sweep = None
    if a.workload == "batch" and not a.no_vary and rank == 0:
        def sweep_of(tv, lv, vv):
            rows_s = []
            for ts_ in (2, 5, 10, 20):
                ms, fr, au = device_ms(tv, lv, vv, ts_)
                rows_s.append({"total_step": ts_, "ms": ms, "audio_s_per_s": au / (ms / 1000)})
            slope, icpt = np.polyfit([r["total_step"] for r in rows_s], [r["ms"] for r in rows_s], 1)
            return {"utterances": len(tv), "latent_frames": fr, "row_tiles_of_128": -(-fr // 128), "audio_s": au, "by_total_step": rows_s,
                    "ms_per_euler_step": float(slope), "ms_outside_the_loop": float(icpt)}
        base = {"en": "This morning, I took a walk in the park, and the sound of the birds and the breeze was so pleasant.",
                "ko": "오늘 아침에 공원을 산책했는데, 새소리와 바람 소리가 너무 기분 좋았어요.", "es": "El niño comió piñas en la montaña, ¿verdad? ¡Sí, señor!",
                "pt": "A ação e o coração não são fáceis de explicar, mas vovô tentou.", "fr": "Où est l'hôtel? Ça coûte très cher, naïve Zoë préfère le café déjà."}
        tm, lm, vm = [], [], []
        for i in range(30):                                      # 6 utterances per language, 1..3 repetitions of its sentence (<= 300 characters,
            lg = ("en", "ko", "es", "pt", "fr")[i % 5]           # the reference's chunk size, cpp/helper.cpp:698), voices cycling
            tm.append(" ".join([base[lg]] * (1 + (i // 5) % 3))); lm.append(lg); vm.append(("M1", "F1", "M2", "F2")[i % 4])
        sweep = {"what": "device-resident synthesis at total_step 2/5/10/20 (median of 3, L2 flushed): ms = ms_outside_the_loop + total_step x ms_per_euler_step (least squares)",
                 "multilingual": dict(sweep_of(tm, lm, vm), workload="configs[2]: 30 utterances en/ko/es/pt/fr (6 each, 1-3 sentences, <= 300 characters), voices M1/F1/M2/F2 per utterance, speed 1.05"),
                 "headline_batch": dict(sweep_of(texts[:group], langs[:group], voices[:group]), workload="configs[1] batch")}

    # ---- latency leg (rank 0): p50 wall time of TextToSpeech.call() on the reference's default sentence (configs[0]) and on a
    #      ~2 000-character text (configs[3]: chunkText -> sequential chunks like the reference, and all chunks as one packed batch) ----
    lat = None
    if rank == 0:
        sentence = ("This morning, I took a walk in the park, and the sound of the birds and the breeze was so pleasant that "
                    "I stopped for a long time just to listen.")              # reference cpp/example_onnx.cpp:17
        one = T.load_voice_style([os.path.join(root, "voice_styles", "M1.json")])

        def p50(fn, n=18, skip=3):
            ts, r = [], None
            for i in range(n):
                t0 = time.perf_counter()
                r = fn()
                ts.append(time.perf_counter() - t0)
            ts = ts[skip:]
            return 1000 * float(np.median(ts)), 1000 * float(np.quantile(ts, 0.9)), r
        a50, a90, r1 = p50(lambda: tt.call(sentence, "en", one, a.total_step, 1.05))
        rng = np.random.default_rng(7)
        sents = []
        while sum(len(x) + 1 for x in sents) < 2000:
            k = int(rng.integers(8, 22))
            w = [WORDS[rng.integers(len(WORDS))] for _ in range(k)]
            sents.append((" ".join(w) + ".").capitalize())
        long_text = " ".join(sents)
        n_chunks = len(T.chunk_text(long_text, 300))
        s50, s90, rs = p50(lambda: tt.call(long_text, "en", one, a.total_step, 1.05), n=8, skip=2)
        b50, b90, rb = p50(lambda: tt.call_batched(long_text, "en", one, a.total_step, 1.05), n=8, skip=2)
        q50, q90, rq = p50(lambda: tt.call_batched(long_text, "en", one, a.total_step, 1.05, pcm16=True), n=8, skip=2)
        lat = {"p50_ms": a50, "p90_ms": a90, "audio_s": float(r1.duration[0]),
               "what": "configs[0]: TextToSpeech.call(default sentence, M1, total_step, speed 1.05): text front-end + H2D + synthesis + D2H, batch 1",
               "long_form": {"what": f"configs[3]: {len(long_text)}-character English text, chunkText(300) -> {n_chunks} chunks, speed 1.05, silence 0.3 s",
                             "audio_s": float(rs.duration[0]),
                             "sequential_call_p50_ms": s50, "sequential_call_p90_ms": s90,
                             "call_batched_p50_ms": b50, "call_batched_p90_ms": b90,
                             "call_batched_pcm16_p50_ms": q50, "d2h_bytes": {"float32": int(rb.wav.nbytes), "pcm16": int(rq.wav.nbytes)}}}

    # ---- roofline leg (rank 0): per-launch CUDA events around the dominant kernel class -----------------
    roof = None
    stage = None
    if rank == 0:
        eng.set_profile(2)
        prof = {k: dict(ms=0, flops=0, bytes=0, launches=0) for k in ("gemm_tc", "dwconv_ln", "attention", "fused_mlp", "gemm_f16", "dwconv_ln_hbm")}
        stage = dict(dp=0.0, te=0.0, ve=0.0, vocoder=0.0, whole=0.0)
        for b in buckets:
            eng.synthesize_packed_device(b["ids"].data_ptr(), b["mask"].data_ptr(), b["ttl"].data_ptr(), b["dp"].data_ptr(), b["B"], b["T"],
                                         a.total_step, 1.05, 7, b["wav"].data_ptr(), b["cap"], b["dur"].data_ptr(), text_lens=b["lens"])
            for k, v in eng.kernel_profile().items():
                for kk in v:
                    prof[k][kk] += v[kk]
            for k, v in eng.stage_ms().items():
                stage[k] += v
        eng.set_profile(0)
        pk = peaks()
        tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        traffic = json.load(open(tp)).get("dram_bytes_per_launch") if os.path.exists(tp) else None
        names = {"gemm_tc": "tc::gemm_bf16x3_kernel<64|128|256> (TMA -> tcgen05.mma kind::f16 -> TMEM, 3 MMAs per K-slice)",
                 "gemm_f16": "tc2a::gemm2_f16_astat_kernel (pw1: A rows resident) + tc2::gemm2_bf16x3_kernel<true> (pw2, conv_in, head) — vocoder projections: two-SM cta_group::2 tcgen05.mma kind::f16, single-pass fp16 operands",
                 "fused_mlp": "mlp::convnext_mlp_stream2_kernel (CTA pairs, cta_group::2) / convnext_mlp_stream_kernel + mlp_reduce[_post]_kernel (pw1 -> GELU -> pw2 fused: S in TMEM, P written back into TMEM, "
                              "O accumulated from the TMEM operand; tcgen05, 3 MMAs per K-slice)"}

        def tensor_line(key, passes=3):
            g = prof[key]
            ach = g["flops"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] else 0.0
            return {"kernel": names[key], "bound": "tensor", "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                    "frac": ach / pk["bf16_sustained"], "frac_executed_mma": passes * ach / pk["bf16_sustained"],
                    "frac_of_burst_peak": ach / pk["bf16"], "frac_executed_mma_of_burst_peak": passes * ach / pk["bf16"], "burst_peak": pk["bf16"],
                    "peak_source": f"{pk['src']} bf16 sustained (kernel timed inside a long step; the step runs un-capped at ~1965 MHz, so the "
                                   "burst figure is quoted beside it)",
                    "timing": "CUDA events around each launch of the class, eager launches (profile level 2): includes the launch gaps, "
                              "i.e. pessimistic against the in-graph kernel time (ncu launch list under profiles/)",
                    "launches": g["launches"], "avg_launch_us": 1000 * g["ms"] / max(g["launches"], 1),
                    "algorithmic_flops_per_step": g["flops"], "executed_mma_flops_per_step": passes * g["flops"],
                    "share_of_step": g["ms"] / max(stage["whole"], 1e-9)}
        # In-graph timing of the two headline kernels on the shapes of THIS batch: CUDA events on the library's stream around a CUDA
        # graph of 20 back-to-back launches (stc_debug_mlp / stc_debug_dwconv: same kernels, same dispatch, seeded random operands).
        # The per-class numbers above come from eager launches with an event pair around each, i.e. they carry ~8 us of launch gap.
        frames = int(sum(b.get("L", 0) for b in buckets)); tokens = int(lens.sum())
        ig = {}
        try:
            f_lat, u_lat, _ = eng.debug_mlp(frames, 20)
            f_txt, u_txt, _ = eng.debug_mlp(tokens, 20)
            n_ve = 160 if a.total_step == 5 else 32 * a.total_step
            us_total = n_ve * f_lat + 12 * f_txt
            fl_total = 4.0 * 256 * 1024 * (n_ve * frames + 12 * tokens)
            ig["fused_mlp"] = {"us_per_block_latent_rows": f_lat, "us_per_block_text_rows": f_txt, "rows": [frames, tokens],
                               "two_gemm_form_us": [u_lat, u_txt], "achieved": fl_total / us_total / 1e6,
                               "large_group": None,
                               "how": "stc_debug_mlp: 20 launches of (stream kernel + reduce kernel) replayed from one CUDA graph, CUDA events on the library's stream"}
            # the same kernel pair on a launch group of configs[4] (tts.plan_many fills groups to ~136 row tiles: one CTA pair per two tiles,
            # the whole hidden layer per pair — the wave-limited 37-tile headline shape is the small end of the kernel's range)
            g_rows = 136 * 128
            f_big, u_big, _ = eng.debug_mlp(g_rows, 20)
            a_big = 4.0 * 256 * 1024 * g_rows / f_big / 1e6
            ig["fused_mlp"]["large_group"] = {"rows": g_rows, "us_per_block": f_big, "two_gemm_form_us": u_big, "achieved": a_big,
                                              "frac": a_big / pk["bf16_sustained"], "frac_executed_mma": 3 * a_big / pk["bf16_sustained"],
                                              "frac_executed_mma_of_burst_peak": 3 * a_big / pk["bf16"]}
            os.environ["STC_DEBUG_F16"] = "1"
            vrows = frames * eng.cfg.chunk_compress_factor
            s_us, t_us, _ = eng.debug_dwconv(vrows, 512, 7, 1, True, len(texts), 0, 20)
            del os.environ["STC_DEBUG_F16"]
            ig["dwconv_ln_hbm"] = {"us": s_us, "rows": vrows, "bytes": 6.0 * vrows * 512, "achieved": 6.0 * vrows * 512 / s_us / 1e3,
                                   "how": "stc_debug_dwconv (fp32 in, fp16 operand out, causal K = 7, C = 512): 20 launches replayed from one CUDA graph; the "
                                          "57 MB input and 28 MB output of consecutive launches partly stay in the 126 MB L2, as they do between the vocoder's kernels"}
        except Exception as e:          # noqa: BLE001
            ig["error"] = str(e)
        dom = max(("gemm_tc", "fused_mlp"), key=lambda k: prof[k]["ms"])
        other = "fused_mlp" if dom == "gemm_tc" else "gemm_tc"
        roof = tensor_line(dom)
        tpm = os.path.join(ROOT, "profiles", "mlp_traffic.json")
        roof["traffic"] = traffic if dom == "gemm_tc" else (json.load(open(tpm)).get("dram_bytes_per_launch") if os.path.exists(tpm) else None)
        roof["traffic_source"] = "profiles/gemm_traffic.json" if dom == "gemm_tc" else "profiles/mlp_traffic.json"
        if dom == "fused_mlp" and "fused_mlp" in ig:
            g = ig["fused_mlp"]
            roof["eager"] = {k: roof[k] for k in ("achieved", "frac", "frac_executed_mma", "frac_of_burst_peak", "frac_executed_mma_of_burst_peak", "avg_launch_us")}
            roof.update({"achieved": g["achieved"], "frac": g["achieved"] / pk["bf16_sustained"], "frac_executed_mma": 3 * g["achieved"] / pk["bf16_sustained"],
                         "frac_of_burst_peak": g["achieved"] / pk["bf16"], "frac_executed_mma_of_burst_peak": 3 * g["achieved"] / pk["bf16"],
                         "avg_launch_us": g["us_per_block_latent_rows"], "timing": g["how"] + " (`eager`: the per-launch event timing of the step itself)",
                         "in_graph": g})
        roof["note"] = ("split-bf16 arithmetic executes 3 MMAs per algorithmic multiply-add, so `frac` (algorithmic) cannot exceed 1/3; "
                        "`frac_executed_mma` is the tensor-pipe load")
        roof[other] = tensor_line(other)
        roof["gemm_f16"] = tensor_line("gemm_f16", passes=1)
        tpf = os.path.join(ROOT, "profiles", "gemm_f16_traffic.json")
        roof["gemm_f16"]["traffic"] = json.load(open(tpf)).get("dram_bytes_per_launch") if os.path.exists(tpf) else None
        roof["gemm_f16"]["traffic_source"] = "profiles/gemm_f16_traffic.json (pw1, one launch under ncu --set full)"
        roof["dwconv_ln"] = {"bound": "hbm", "achieved_gbs": prof["dwconv_ln"]["bytes"] / max(prof["dwconv_ln"]["ms"], 1e-9) / 1e6,
                             "peak_gbs": pk["hbm"], "launches": prof["dwconv_ln"]["launches"],
                             "share_of_step": prof["dwconv_ln"]["ms"] / max(stage["whole"], 1e-9)}
        dh = prof["dwconv_ln_hbm"]
        roof["dwconv_ln_hbm"] = {"kernel": "stc::dwconv_ln_chain_kernel<4, 7> (vocoder: 27.7k x 512 rows, fp32 in, fp16 operand out)", "bound": "hbm",
                                 "achieved": dh["bytes"] / max(dh["ms"], 1e-9) / 1e6, "peak": pk["hbm"], "unit": "GB/s",
                                 "frac": dh["bytes"] / max(dh["ms"], 1e-9) / 1e6 / pk["hbm"], "launches": dh["launches"],
                                 "avg_launch_us": 1000 * dh["ms"] / max(dh["launches"], 1), "share_of_step": dh["ms"] / max(stage["whole"], 1e-9)}
        if "dwconv_ln_hbm" in ig:
            g = ig["dwconv_ln_hbm"]
            roof["dwconv_ln_hbm"]["eager"] = {k: roof["dwconv_ln_hbm"][k] for k in ("achieved", "frac", "avg_launch_us")}
            roof["dwconv_ln_hbm"].update({"achieved": g["achieved"], "frac": g["achieved"] / pk["hbm"], "avg_launch_us": g["us"], "timing": g["how"]})
        roof["attention"] = {"achieved_tflops": prof["attention"]["flops"] / max(prof["attention"]["ms"], 1e-9) / 1e9,
                             "launches": prof["attention"]["launches"],
                             "share_of_step": prof["attention"]["ms"] / max(stage["whole"], 1e-9)}

    pcheck = None
    if rank == 0 and not a.no_parity_check:
        # the shortest and the longest utterance of the timed batch, checked against the oracle outside the timed region
        pcheck = parity_check(eng, root, ids[:group] if n_utt > group else ids, mask[:group] if n_utt > group else mask,
                              T.Style(style.ttl[:group], style.dp[:group]) if n_utt > group else style, a.total_step,
                              [int(np.argmin(lens[:group])), int(np.argmax(lens[:group]))])
        if not (pcheck["durations_and_frame_counts_bit_exact"] and pcheck["latent_max_abs"] <= 2e-4 and pcheck["snr_db"] >= 60.0):
            print("bench.py: PARITY CHECK FAILED " + json.dumps(pcheck), file=sys.stderr)
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        k = min(a.cpu_sample, a.batch)
        idx = list(np.linspace(0, a.batch - 1, k).astype(int))
        audio_c, times, cores, kind, runtime = oracle_run([texts[i] for i in idx], [langs[i] for i in idx], [voices[i] for i in idx], a.total_step, 3)
        cpu = {"value": audio_c / float(np.mean(times[1:])), "unit": "audio-s/s", "cores": cores, "kind": kind,
               "sample": f"{k} of the {a.batch} utterances as one padded batch, 2 timed repetitions after 1 warm-up", "runtime": runtime}
    out = {"metric": "audio-sec/sec", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if a.workload == "sweep1024" else "weak", "vs_baseline": None, "dtype": "bf16x3->f32 (Euler loop, text side); f16->f32 (vocoder GEMMs)",
           "data": "synthetic", "config": dict(cfg, buckets=[[b["B"], b["T"], b.get("L")] for b in buckets],
                                               audio_s_per_step_per_gpu=audio),
           "clocks": clocks, "e2e": main["e2e"],
           "gpu_launches": int(launches), "parity_check": pcheck, "per_rank": per_rank, "latency": lat, "strong": strong, "vary_batches": vary, "step_sweep": sweep, "roofline": roof, "cpu_baseline": cpu, "stage_ms": stage,
           "p50_step_ms": main["p50_step_ms"], "step_ms": main["step_ms"]}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
