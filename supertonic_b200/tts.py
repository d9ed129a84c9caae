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

from . import capi


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
    def __init__(self, engine: capi.Engine):
        self.engine = engine
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

    # -- cpp/helper.cpp:725-734
    def batch(self, text_list, lang_list, style: Style, total_step: int, speed: float = 1.05) -> SynthesisResult:
        return self._infer(text_list, lang_list, style, total_step, speed)

    # -- throughput path (north_star: "a request batch is length-bucketed")
    def synthesize_many(self, texts, langs, style: Style, total_step: int, speed: float = 1.05,
                        max_batch: int = 128, seed: int = 0, noise: Optional[np.ndarray] = None, copy: bool = False,
                        wait: bool = True):
        """Independent utterances -> list of (trimmed wav, duration) in input order. The latent side runs on packed
        rows (no padded frames); the text side is one [B, T_max] rectangle per group of at most `max_batch`
        utterances, grouped by token count so text padding stays small. Results do not depend on the grouping
        (tests: batch-composition invariance). With copy=False the waveforms are views into the engine's page-locked result
        buffers (one per group, two alternating sets), valid until the call after the next one. Groups are issued
        asynchronously: the device->host copy of one group overlaps the computation of the next. wait=False returns before the
        last copies have landed — call `engine.wait()` before reading the waveforms; issuing the next synthesize_many first
        overlaps its computation with these copies (request streams)."""
        from .scheduler import length_buckets
        n = len(texts)
        ids, mask = self.engine.text_to_ids(texts, langs)
        lens = mask.reshape(n, -1).sum(1).astype(np.int64)
        out: List[Optional[Tuple[np.ndarray, float]]] = [None] * n
        self._parity = 1 - getattr(self, "_parity", 1)
        for gi, grp in enumerate(length_buckets(lens, max_batch, 1e9)):
            g = np.asarray(grp)
            T = int(lens[g].max())
            r = self.engine.synthesize_packed(ids[g, :T], mask[g, :, :T], style.ttl[g], style.dp[g], total_step, speed,
                                              seed=seed, noise=None if noise is None else noise[g],
                                              pinned=f"wav_packed{gi}_{self._parity}", wait=False)
            for k, i in enumerate(grp):
                out[i] = (r["wavs"][k], float(r["duration"][k]))
        if wait or copy:
            self.engine.wait()
            if copy:
                out = [(w.copy(), d) for w, d in out]
        return out


def load_text_to_speech(onnx_dir: str, use_gpu: bool = True, device: int = 0, precision: int = capi.PREC_DEFAULT) -> TextToSpeech:
    """Mirror of loadTextToSpeech (cpp/helper.cpp:903-937). The reference throws on use_gpu=True; this library is
    GPU-only, so it throws on use_gpu=False instead — there is no CPU path to fall back to."""
    if not use_gpu:
        raise RuntimeError("CPU mode is not supported by supertonic_b200 (use the reference's ONNX Runtime path)")
    return TextToSpeech(capi.Engine(onnx_dir, device, precision))
