"""Host-side mirror of the reference's C++ synthesis API (cpp/helper.h / helper.cpp) over libsupertonic_cuda.

Same names, argument meaning and error behaviour as the reference for this path:
`TextToSpeech.call / batch` (cpp/helper.cpp:685-734), `Style` + `load_voice_style` (:829-897),
`load_text_to_speech` (:903-937 — here `use_gpu=True` is the only mode and `False` raises, the mirror image
of the reference's throw), `chunk_text` (:1117-1186), `write_wav_file` (:943-990), `sanitize_filename`
(:1070-1111), `timer` (cpp/helper.h:213-223).

Additions that do not change the reference semantics: a seedable / injectable noise source (the reference
RNG is unseedable, :442-443) and `synthesize_many`, the length-bucketed throughput path (north_star).
"""
from __future__ import annotations

import json
import os
import struct
import time
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import affinity, capi


@dataclass
class Style:
    """Voice style tensors stacked on the batch dim (cpp/helper.h:57-72)."""
    ttl: np.ndarray     # [B, d1, d2] float32
    dp: np.ndarray      # [B, e1, e2] float32


def load_voice_style(paths: Sequence[str], verbose: bool = False) -> Style:
    """cpp/helper.cpp:829-897: dims come from the first file's dims[1:3]."""
    ttl, dp = [], []
    dims_t = dims_d = None
    for p in paths:
        try:
            with open(p) as f:
                j = json.load(f)
        except OSError:
            raise RuntimeError(f"Failed to open voice style file: {p}")
        if dims_t is None:
            dims_t, dims_d = j["style_ttl"]["dims"][1:3], j["style_dp"]["dims"][1:3]
        ttl.append(np.asarray(j["style_ttl"]["data"], np.float32).reshape(1, *dims_t))
        dp.append(np.asarray(j["style_dp"]["data"], np.float32).reshape(1, *dims_d))
    if verbose:
        print(f"Loaded {len(paths)} voice styles")
    return Style(np.concatenate(ttl, 0), np.concatenate(dp, 0))


def chunk_text(text, max_len: int = 300) -> List[bytes]:
    return capi.chunk_text(text, max_len)


def sanitize_filename(text, max_len: int) -> str:
    """cpp/helper.cpp:1070-1111 (ASCII alnum/_ kept, multi-byte UTF-8 sequences kept whole, rest → '_')."""
    t = text if isinstance(text, bytes) else text.encode("utf-8", "surrogateescape")
    out, i, cnt = bytearray(), 0, 0
    while i < len(t) and cnt < max_len:
        c = t[i]
        if chr(c).isascii() and (chr(c).isalnum() or c == 95):
            n = 1
        elif c & 0xE0 == 0xC0 and i + 1 < len(t):
            n = 2
        elif c & 0xF0 == 0xE0 and i + 2 < len(t):
            n = 3
        elif c & 0xF8 == 0xF0 and i + 3 < len(t):
            n = 4
        else:
            out += b"_"; i += 1; cnt += 1
            continue
        out += t[i:i + n]; i += n; cnt += 1
    return out.decode("utf-8", "surrogateescape")


def wav_file_bytes(audio: np.ndarray, sample_rate: int) -> bytes:
    """16-bit PCM mono WAV; quantisation = clamp·32767 truncated toward zero (cpp/helper.cpp:985-988)."""
    x = np.clip(np.asarray(audio, np.float32), np.float32(-1), np.float32(1)) * np.float32(32767)
    data = x.astype(np.int16).astype("<i2").tobytes()
    return (b"RIFF" + struct.pack("<i", 36 + len(data)) + b"WAVEfmt " +
            struct.pack("<ihhiihh", 16, 1, 1, sample_rate, sample_rate * 2, 2, 16) + b"data" + struct.pack("<i", len(data)) + data)


def wav_file_bytes_pcm16(pcm: np.ndarray, sample_rate: int) -> bytes:
    """The same file from samples already quantised on the device (stc_out_opts.pcm16)."""
    data = np.ascontiguousarray(pcm, dtype="<i2").tobytes()
    return (b"RIFF" + struct.pack("<i", 36 + len(data)) + b"WAVEfmt " +
            struct.pack("<ihhiihh", 16, 1, 1, sample_rate, sample_rate * 2, 2, 16) + b"data" + struct.pack("<i", len(data)) + data)


def write_wav_file(filename: str, audio: np.ndarray, sample_rate: int) -> None:
    try:
        with open(filename, "wb") as f:
            f.write(wav_file_bytes(audio, sample_rate))
    except OSError:
        raise RuntimeError(f"Failed to open file for writing: {filename}")


def timer(name: str, fn):
    """cpp/helper.h:213-223."""
    t0 = time.perf_counter()
    print(f"{name}...")
    r = fn()
    print(f"  -> {name} completed in {time.perf_counter() - t0:.2f} sec")
    return r


@dataclass
class SynthesisResult:
    wav: np.ndarray         # flat float32, [B * L*cs] row-major like the reference's result.wav (:679)
    duration: np.ndarray    # float32 [B] seconds


class TextToSpeech:
    def __init__(self, engine: capi.Engine, lanes: Sequence[capi.Engine] = ()):
        """`lanes`: further engines (handles) on the SAME GPU — request lanes. Inside one handle the Euler loop / vocoder of
        consecutive requests run back to back on one stream; a second handle has its own streams, workspaces and CUDA graphs, so
        the launch groups `synthesize_many` deals out alternately overlap on the device: a stream of 32-utterance requests gains
        ~10 % with two lanes (tools/concurrent_handles.py), groups of 128 utterances fill the GPU on their own and gain nothing."""
        self.engine = engine
        self.lanes = [engine, *lanes]
        self._lane = 0
        self.cfg = engine.cfg
        self.sample_rate = engine.cfg.sample_rate
        self.noise_seed = 0            # advanced per _infer call unless noise is injected
        self._calls = 0

    def get_sample_rate(self) -> int:
        return self.sample_rate

    # -- cpp/helper.cpp:469-683
    def _infer(self, text_list, lang_list, style: Style, total_step: int, speed: float = 1.05,
               noise: Optional[np.ndarray] = None) -> SynthesisResult:
        if len(text_list) != style.ttl.shape[0]:
            raise RuntimeError("Number of texts must match number of style vectors")
        ids, mask = self.engine.text_to_ids(text_list, lang_list)
        self._calls += 1
        r = self.engine.synthesize(ids, mask, style.ttl, style.dp, total_step, speed, noise=noise,
                                   seed=self.noise_seed + self._calls)
        return SynthesisResult(r["wav"].reshape(-1), r["duration"])

    # -- cpp/helper.cpp:685-723: sequential chunks, 0.3 s silence, untrimmed chunk wavs concatenated
    def call(self, text, lang: str, style: Style, total_step: int, speed: float = 1.05,
             silence_duration: float = 0.3) -> SynthesisResult:
        if style.ttl.shape[0] != 1:
            raise RuntimeError("Single speaker text to speech only supports single style")
        wav_cat: Optional[np.ndarray] = None
        dur_cat = np.float32(0)
        for chunk in chunk_text(text, 120 if lang == "ko" else 300):
            r = self._infer([chunk], [lang], style, total_step, speed)
            if wav_cat is None:
                wav_cat, dur_cat = r.wav, r.duration[0]
            else:
                sil = np.zeros(int(np.float32(silence_duration) * np.float32(self.sample_rate)), np.float32)
                wav_cat = np.concatenate([wav_cat, sil, r.wav])
                dur_cat = np.float32(dur_cat + np.float32(r.duration[0] + np.float32(silence_duration)))
        return SynthesisResult(wav_cat, np.asarray([dur_cat], np.float32))

    def call_batched(self, text, lang: str, style: Style, total_step: int, speed: float = 1.05, silence_duration: float = 0.3,
                     pcm16: bool = False, seed: Optional[int] = None) -> SynthesisResult:
        """`call()` with the chunks of the text as ONE packed batch: same result layout — the untrimmed chunk waveforms joined with
        (int)(silence_duration * sample_rate) zeros (cpp/helper.cpp:703-716) — but the join (and, with pcm16=True, writeWavFile's
        quantisation, cpp/helper.cpp:985-988) happens on the device and one launch serves the whole text.
        `wav` is float32, or int16 PCM with pcm16=True (half the device->host bytes)."""
        if style.ttl.shape[0] != 1:
            raise RuntimeError("Single speaker text to speech only supports single style")
        chunks = chunk_text(text, 120 if lang == "ko" else 300)
        n = len(chunks)
        ids, mask = self.engine.text_to_ids(chunks, [lang] * n)
        self._calls += 1
        gap = int(np.float32(silence_duration) * np.float32(self.sample_rate))
        r = self.engine.synthesize_joined(ids, mask, np.repeat(style.ttl, n, 0), np.repeat(style.dp, n, 0), total_step, speed,
                                          seed=self.noise_seed + self._calls if seed is None else seed, pcm16=pcm16, gap_samples=gap)
        dur = np.float32(r["duration"][0])
        for d in r["duration"][1:]:
            dur = np.float32(dur + np.float32(d + np.float32(silence_duration)))
        return SynthesisResult(r["out"], np.asarray([dur], np.float32))

    # -- cpp/helper.cpp:725-734
    def batch(self, text_list, lang_list, style: Style, total_step: int, speed: float = 1.05) -> SynthesisResult:
        return self._infer(text_list, lang_list, style, total_step, speed)

    # -- throughput path (north_star: "a request batch is length-bucketed")
    def synthesize_many(self, texts, langs, style: Style, total_step: int, speed: float = 1.05,
                        max_batch: int = 128, seed: int = 0, noise: Optional[np.ndarray] = None, copy: bool = False,
                        wait: bool = True, pcm16: bool = False):
        """Independent utterances -> list of (trimmed wav, duration) in input order. Both sides run on packed rows (no padded
        frames or tokens are computed); the request is cut into launch groups of at most `max_batch` utterances and 140 row
        tiles of predicted latent frames, filled to equal predicted frames (plan_many). Results do not depend on the grouping
        (tests: batch-composition invariance; device noise streams are keyed by the utterance's index in the REQUEST).
        With copy=False the waveforms are views into the engine's page-locked result buffers (one per group, two alternating
        sets): they stay valid until the SECOND-next synthesize_many on this object — copy them (copy=True) to keep them longer.
        Groups are issued asynchronously: the device->host copy of one group overlaps the computation of the next. wait=False
        returns before the last copies have landed — call `wait()` before reading the waveforms; issuing the next
        synthesize_many first overlaps its computation with these copies (request streams). pcm16=True: int16 samples quantised
        on the device (writeWavFile's rule), half the device->host bytes."""
        lanes = getattr(self, "lanes", None) or [self.engine]
        plan = plan_many(self.engine, texts, langs, max_batch)
        out: List[Optional[Tuple[np.ndarray, float]]] = [None] * len(texts)
        self._parity = 1 - getattr(self, "_parity", 1)
        if len(lanes) == 1 or noise is not None:
            run_groups(self.engine, plan, range(len(plan.groups)), style, total_step, speed, seed, out, tag=f"p{self._parity}", noise=noise, pcm16=pcm16)
        else:
            # request lanes: consecutive launch groups (of this and of the previous calls) go to alternating handles; the tag keeps
            # the page-locked result buffers of the last two calls apart per handle
            for gi in range(len(plan.groups)):
                eng = lanes[self._lane % len(lanes)]
                self._lane += 1
                run_groups(eng, plan, [gi], style, total_step, speed, seed, out, tag=f"p{self._parity}", pcm16=pcm16)
            self.engine.frames_per_token = lanes[(self._lane - 1) % len(lanes)].frames_per_token
        if wait or copy:
            self.wait()
            if copy:
                out = [(w.copy(), d) for w, d in out]
        return out

    def wait(self) -> None:
        """Blocks until every asynchronous result of synthesize_many(wait=False) has landed (all request lanes)."""
        for e in getattr(self, "lanes", None) or [self.engine]:
            e.wait()


@dataclass
class ManyPlan:
    ids: np.ndarray             # [n, T] token ids of the whole request (text front-end run once)
    mask: np.ndarray            # [n, 1, T]
    lens: np.ndarray            # [n] token counts
    groups: List[List[int]]     # text indices per launch group (equal predicted latent frames, scheduler.frame_balanced_groups)


def plan_many(engine: capi.Engine, texts, langs, max_batch: int) -> ManyPlan:
    """Front-end once, then launch groups of equal predicted latent frames. The frames-per-token ratio is what this engine's
    last synthesize_many measured (`engine.frames_per_token`; 1.0 — about one 70 ms latent frame per character — until then)."""
    from .scheduler import frame_balanced_groups
    ids, mask = engine.text_to_ids(texts, langs)
    lens = mask.reshape(len(texts), -1).sum(1).astype(np.int64)
    return ManyPlan(ids, mask, lens, frame_balanced_groups(lens, max_batch, getattr(engine, "frames_per_token", 1.0)))


def run_groups(engine: capi.Engine, plan: ManyPlan, group_ids, style: Style, total_step: int, speed: float, seed: int, out: list,
               tag: str = "", noise: Optional[np.ndarray] = None, pcm16: bool = False) -> None:
    """The launch groups `group_ids` of `plan` on ONE engine, as an asynchronous request stream; fills out[text index] with
    (trimmed waveform view, duration). The caller waits (`engine.wait()`) before reading the waveforms."""
    frames = tokens = 0
    for gi in group_ids:
        g = np.asarray(plan.groups[gi])
        T = int(plan.lens[g].max())
        r = engine.synthesize_joined(plan.ids[g, :T], plan.mask[g, :, :T], style.ttl[g], style.dp[g], total_step, speed, seed=seed,
                                     noise=None if noise is None else noise[g], noise_index=g, pcm16=pcm16,
                                     pinned=f"many{gi}_{tag}", wait=noise is not None)
        for k, i in enumerate(g):
            o = int(r["offsets"][k])
            out[int(i)] = (r["out"][o:o + int(r["wav_lengths"][k])], float(r["duration"][k]))
        frames += int(np.asarray(r["frames"]).sum())
        tokens += int(plan.lens[g].sum())
    if tokens:
        engine.frames_per_token = frames / tokens      # the next plan's prediction (plan_many)


class MultiGpuTextToSpeech:
    """One process, several GPUs of one box (north_star: "a request batch is length-bucketed and partitioned across the 8 B200s of
    one box; each GPU holds a full weight replica and no collective runs on the hot path"): one engine (handle, streams, CUDA-graph
    cache, page-locked result buffers) and one host thread per device. The reference entry point that fans out is
    TextToSpeech::batch (cpp/helper.cpp:725-734). A request is bucketed into launch groups ONCE (front-end on the host), the groups
    are dealt out longest-processing-time-first by predicted cost, and results come back in input order. Noise streams are keyed by
    the utterance's index in the request, so the audio does not depend on the number of devices."""

    def __init__(self, onnx_dir: Optional[str], devices: Sequence[int], precision: int = capi.PREC_DEFAULT, engines=None):
        if not len(devices):
            raise RuntimeError("MultiGpuTextToSpeech: no devices given")
        self.devices = [int(d) for d in devices]
        # (`engines`: already-built capi.Engine objects, one per entry of `devices` — used by tests)
        self.engines = list(engines) if engines is not None else [capi.Engine(onnx_dir, d, precision) for d in self.devices]
        self.engine = self.engines[0]
        self.cfg = self.engine.cfg
        self.sample_rate = self.cfg.sample_rate
        self._parity = 1
        self._first = TextToSpeech(self.engine)        # single-utterance entry points (latency path) run on the first device

    def get_sample_rate(self) -> int:
        return self.sample_rate

    def call(self, *a, **k):
        return self._first.call(*a, **k)

    def call_batched(self, *a, **k):
        return self._first.call_batched(*a, **k)

    def batch(self, *a, **k):
        return self._first.batch(*a, **k)

    def close(self):
        for e in self.engines:
            e.close()

    def synthesize_many(self, texts, langs, style: Style, total_step: int, speed: float = 1.05, max_batch: int = 128, seed: int = 0,
                        copy: bool = False, pcm16: bool = False):
        import threading
        from .scheduler import shard_lpt, synth_cost
        if len(texts) != style.ttl.shape[0]:
            raise RuntimeError("Number of texts must match number of style vectors")
        # at least two launch groups per device, so that the device->host copy of one group runs under the computation of the next
        # (measured on 8 GPUs, 1 024 utterances: 173 k audio-s/s with one group of 128 per device, 212 k with two of 64)
        nd = len(self.engines)
        if nd > 1:
            max_batch = min(max_batch, max(16, -(-len(texts) // (2 * nd))))
        plan = plan_many(self.engine, texts, langs, max_batch)
        costs = [sum(synth_cost(int(plan.lens[i]), total_step) for i in g) for g in plan.groups]
        shards = shard_lpt(costs, len(self.engines))
        out: List[Optional[Tuple[np.ndarray, float]]] = [None] * len(texts)
        self._parity = 1 - self._parity
        errs: List[BaseException] = []

        def work(r):
            try:
                if len(self.engines) > 1 and os.environ.get("STC_BIND", "1") != "0":
                    affinity.bind_to_gpu(self.devices[r])      # this thread (and the pinned buffers it allocates) next to its GPU
                run_groups(self.engines[r], plan, shards[r], style, total_step, speed, seed, out, tag=f"p{self._parity}", pcm16=pcm16)
                self.engines[r].wait()
            except BaseException as e:      # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=work, args=(r,)) for r in range(len(self.engines)) if shards[r]]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]
        return [(w.copy(), d) for w, d in out] if copy else out


def load_text_to_speech(onnx_dir: str, use_gpu: bool = True, device: int = 0, precision: int = capi.PREC_DEFAULT, lanes: int = 1) -> TextToSpeech:
    """Mirror of loadTextToSpeech (cpp/helper.cpp:903-937). The reference throws on use_gpu=True; this library is
    GPU-only, so it throws on use_gpu=False instead — there is no CPU path to fall back to. `lanes` > 1: that many handles on
    the device (TextToSpeech `lanes`)."""
    if not use_gpu:
        raise RuntimeError("CPU mode is not supported by supertonic_b200 (use the reference's ONNX Runtime path)")
    return TextToSpeech(capi.Engine(onnx_dir, device, precision), [capi.Engine(onnx_dir, device, precision) for _ in range(max(1, int(lanes)) - 1)])


def parse_devices(spec: str) -> List[int]:
    """'0-3' / '0,2,5' / '0-1,4' -> [device indices] (service env TTS_DEVICES, CLI --devices)."""
    out: List[int] = []
    for part in str(spec).split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-", 1)
            out += list(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    if not out:
        raise ValueError(f"no devices in {spec!r}")
    return out
