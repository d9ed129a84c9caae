"""HTTP serving wrapper over libsupertonic_cuda with dynamic batching.

Same contract as the reference's `py/service.py` (zhoubin-me/supertonic):
  * `GET /health` -> `{"status": "ok"}`                                                  (py/service.py:79-81)
  * `POST /tts`   JSON `{text, lang="en", voice_style=".../M1.json", total_step=5 (1..50), speed=1.05 (>0), batch=false,
    silence_duration=0.3 (>=0)}`; `text` / `lang` / `voice_style` are strings or lists of strings   (py/service.py:28-39)
      - non-batch mode takes exactly one text / lang / style, else 400                   (py/service.py:91-96)
      - batch mode needs equally long lists, else 400                                    (py/service.py:44-49, 88-89)
      - languages outside en/ko/es/pt/fr -> 400 "Invalid language(s): ..."               (py/service.py:52-58)
      - one result  -> `audio/wav`, `Content-Disposition: attachment; filename="<sanitize_filename(text, 40) or tts>.wav"`
      - several     -> `application/zip` (deflate) of `<sanitize_filename(text_i, 40) or tts_<i+1>>.wav`, `tts_outputs.zip`
      - every waveform is cut to `int(sample_rate * duration)` samples                   (py/service.py:61-69)
    WAV payload: 16-bit PCM mono, quantised as the reference's C++ `writeWavFile` does (clamp * 32767, truncation,
    cpp/helper.cpp:985-988); the reference service goes through libsndfile, which rounds to nearest — at most 1 LSB apart.

What is new is underneath: the reference runs every request as its own ONNX Runtime call on the request thread. Here one worker
thread owns the GPU engine (the C ABI handle is single-caller) and **coalesces** the requests that arrive while the previous
launch is running — or within `max_wait_ms` of the first one when the engine is idle — into one length-bucketed
`synthesize_many` call of up to `max_batch` utterances (packed rows: a coalesced batch costs what its frames cost, no padding to
the longest request). Utterance results do not depend on what they are batched with (tests/test_parity_gpu.py: batch-composition
invariance), so coalescing is invisible to the client. Requests with different `total_step` / `speed` go into separate launches
of the same drain; long texts (more than one chunk of 300 bytes, 120 for `ko`) keep the reference's `call()` semantics
(chunks + `silence_duration` of silence) and run as their own job.

Run: `TTS_ONNX_DIR=... uvicorn supertonic_b200.service:app` (the module-level `app` is built on first access, so importing this
module needs no GPU).
"""
import io
import os
import queue
import threading
import time
import zipfile
from concurrent.futures import Future
from typing import Callable, List, Optional, Sequence, Tuple, Union

import numpy as np
from pydantic import BaseModel, Field

AVAILABLE_LANGS = ["en", "ko", "es", "pt", "fr"]          # reference py/helper.py:13


class TTSRequest(BaseModel):
    """Request schema of the reference service (py/service.py:28-39), field for field."""
    text: Union[str, List[str]] = Field(..., description="Text to synthesize.")
    lang: Union[str, List[str]] = Field("en", description="Language(s) for text.")
    voice_style: Union[str, List[str]] = Field("assets/voice_styles/M1.json", description="Voice style path(s).")
    total_step: int = Field(5, ge=1, le=50)
    speed: float = Field(1.05, gt=0.0)
    batch: bool = False
    silence_duration: float = Field(0.3, ge=0.0, description="Silence between chunks for non-batch mode.")


class _Job:
    __slots__ = ("kind", "texts", "langs", "style", "total_step", "speed", "silence", "future")

    def __init__(self, kind, texts, langs, style, total_step, speed, silence):
        self.kind, self.texts, self.langs, self.style = kind, texts, langs, style
        self.total_step, self.speed, self.silence = total_step, speed, silence
        self.future: Future = Future()


class DynamicBatcher:
    """One worker thread in front of a `TextToSpeech`: jobs queued by the request threads are drained together.

    kind "utt"  : independent utterances (a batch-mode request, or a single short text) -> coalesced across requests
    kind "call" : one long text through `TextToSpeech.call` (chunks + silence), on its own
    Result of a job: list of (waveform trimmed to int(sr * duration), duration) in the order of its texts."""

    def __init__(self, tts, max_batch: int = 32, max_wait_ms: float = 2.0):
        self.tts, self.max_batch, self.max_wait = tts, int(max_batch), float(max_wait_ms) / 1000.0
        self.q: "queue.Queue" = queue.Queue()
        self.stats = {"launches": 0, "utterances": 0, "max_coalesced": 0, "jobs": 0}
        self._seed = 0
        self._thread = threading.Thread(target=self._run, name="stc-batcher", daemon=True)
        self._thread.start()

    def submit(self, job: _Job) -> Future:
        self.q.put(job)
        return job.future

    def close(self):
        self.q.put(None)
        self._thread.join(timeout=30)

    # -- worker
    def _drain(self, first: _Job) -> Tuple[List[_Job], bool]:
        jobs, n = [first], len(first.texts)
        deadline = time.perf_counter() + self.max_wait
        while n < self.max_batch:
            left = deadline - time.perf_counter()
            try:
                j = self.q.get(timeout=left) if left > 0 else self.q.get_nowait()
            except queue.Empty:
                break
            if j is None:
                return jobs, True
            jobs.append(j)
            n += len(j.texts)
        return jobs, False

    def _run(self):
        stop = False
        while not stop:
            first = self.q.get()
            if first is None:
                return
            jobs, stop = self._drain(first)
            self.stats["jobs"] += len(jobs)
            groups = {}
            for j in jobs:
                if j.kind == "call":
                    self._run_call(j)
                else:
                    groups.setdefault((j.total_step, j.speed), []).append(j)
            for (total_step, speed), js in groups.items():
                self._run_utts(js, total_step, speed)

    def _run_call(self, j: _Job):
        try:
            r = self.tts.call(j.texts[0], j.langs[0], j.style, j.total_step, j.speed, j.silence)
            d = float(r.duration[0])
            self.stats["launches"] += 1
            self.stats["utterances"] += 1
            j.future.set_result([(np.array(r.wav[: int(self.tts.sample_rate * d)], np.float32), d)])
        except Exception as e:              # noqa: BLE001 - handed to the request thread
            j.future.set_exception(e)

    def _run_utts(self, js: Sequence[_Job], total_step: int, speed: float):
        from .tts import Style
        try:
            texts = [t for j in js for t in j.texts]
            langs = [l for j in js for l in j.langs]
            style = Style(np.concatenate([j.style.ttl for j in js], 0), np.concatenate([j.style.dp for j in js], 0))
            self._seed += 1
            res = self.tts.synthesize_many(texts, langs, style, total_step, speed, max_batch=self.max_batch, seed=self._seed, copy=True)
            self.stats["launches"] += 1
            self.stats["utterances"] += len(texts)
            self.stats["max_coalesced"] = max(self.stats["max_coalesced"], len(js))
            o = 0
            for j in js:
                j.future.set_result(res[o:o + len(j.texts)])
                o += len(j.texts)
        except Exception as e:              # noqa: BLE001
            for j in js:
                if not j.future.done():
                    j.future.set_exception(e)


def create_app(tts, style_loader: Optional[Callable] = None, max_batch: int = 32, max_wait_ms: float = 2.0):
    """FastAPI app over a `TextToSpeech` (supertonic_b200.tts) — or anything with its `call` / `synthesize_many` /
    `sample_rate` (the CPU tests pass a stand-in)."""
    from fastapi import FastAPI, HTTPException
    from fastapi.responses import JSONResponse, Response
    from . import tts as T

    if style_loader is None:
        # voice-style JSON (50 x 256 + 8 x 16 floats as text) takes longer to parse than a short utterance takes to synthesise:
        # keep the parsed tensors per (path, mtime, size)
        cache, cache_mu = {}, threading.Lock()

        def style_loader(paths):
            ttl, dp = [], []
            for p in paths:
                try:
                    st = os.stat(p)
                except OSError:
                    raise RuntimeError(f"Failed to open voice style file: {p}")
                key = (p, st.st_mtime_ns, st.st_size)
                with cache_mu:
                    one = cache.get(key)
                if one is None:
                    one = T.load_voice_style([p])
                    with cache_mu:
                        if len(cache) >= 256:
                            cache.clear()
                        cache[key] = one
                ttl.append(one.ttl); dp.append(one.dp)
            if any(t.shape[1:] != ttl[0].shape[1:] for t in ttl) or any(d.shape[1:] != dp[0].shape[1:] for d in dp):
                raise RuntimeError("voice styles of different dimensions in one request")
            return T.Style(np.concatenate(ttl, 0), np.concatenate(dp, 0))
    batcher = DynamicBatcher(tts, max_batch, max_wait_ms)
    app = FastAPI(title="Supertonic TTS Service (libsupertonic_cuda)")
    app.state.batcher = batcher

    def ensure_list(v):
        return v if isinstance(v, list) else [v]

    @app.get("/health")
    def health():
        return JSONResponse({"status": "ok"})

    @app.get("/stats")
    def stats():
        return JSONResponse(dict(batcher.stats))

    @app.post("/tts")
    def synthesize(req: TTSRequest):
        texts, langs, styles = ensure_list(req.text), ensure_list(req.lang), ensure_list(req.voice_style)
        if req.batch:
            if not (len(texts) == len(langs) == len(styles)):
                raise HTTPException(status_code=400, detail="text, lang, and voice_style must have the same length.")
        elif len(texts) != 1 or len(langs) != 1 or len(styles) != 1:
            raise HTTPException(status_code=400, detail="Non-batch mode requires single text, lang, and voice_style.")
        invalid = sorted({l for l in langs if l not in AVAILABLE_LANGS})
        if invalid:
            raise HTTPException(status_code=400, detail=f"Invalid language(s): {', '.join(invalid)}")
        try:
            style = style_loader(styles)
        except RuntimeError as e:
            raise HTTPException(status_code=400, detail=str(e))
        kind = "utt"
        if not req.batch and len(T.chunk_text(texts[0], 120 if langs[0] == "ko" else 300)) > 1:
            kind = "call"
        job = _Job(kind, texts, langs, style, req.total_step, float(req.speed), float(req.silence_duration))
        try:
            results = batcher.submit(job).result()
        except RuntimeError as e:
            raise HTTPException(status_code=500, detail=str(e))
        sr = tts.sample_rate
        chunks = [np.asarray(w[: int(sr * d)], np.float32) for w, d in results]
        if len(chunks) == 1:
            name = T.sanitize_filename(texts[0], 40) or "tts"
            return Response(T.wav_file_bytes(chunks[0], sr), media_type="audio/wav",
                            headers={"Content-Disposition": f'attachment; filename="{_ascii(name)}.wav"'})
        zbuf = io.BytesIO()
        with zipfile.ZipFile(zbuf, "w", compression=zipfile.ZIP_DEFLATED) as zf:
            for i, c in enumerate(chunks):
                name = T.sanitize_filename(texts[i], 40) or f"tts_{i + 1}"
                zf.writestr(f"{name}.wav", T.wav_file_bytes(c, sr))
        return Response(zbuf.getvalue(), media_type="application/zip",
                        headers={"Content-Disposition": 'attachment; filename="tts_outputs.zip"'})

    return app


def _ascii(name: str) -> str:
    """HTTP header values are latin-1: non-ASCII characters kept by sanitize_filename (Hangul, accents) become '_' in the header
    (the ZIP member names keep them)."""
    return "".join(ch if ord(ch) < 128 else "_" for ch in name)


_app = None


def get_app():
    """The service configured from the environment like the reference's (py/service.py:19-24): TTS_ONNX_DIR, plus
    TTS_DEVICE (one GPU) or TTS_DEVICES (e.g. "0-7": one engine + host thread per GPU, every coalesced launch group dealt out over
    them — tts.MultiGpuTextToSpeech; a device may be named twice, "0,0" = two handles on GPU 0, whose launch groups then overlap on the
    device), TTS_LANES (handles on the one GPU of TTS_DEVICE, dealt launch groups alternately), TTS_MAX_BATCH, TTS_MAX_WAIT_MS.
    TTS_USE_GPU=0 is refused: there is no CPU path here."""
    global _app
    if _app is None:
        from . import tts as T
        if os.getenv("TTS_USE_GPU", "1").strip().lower() in {"0", "false", "no", "n", "off"}:
            raise RuntimeError("supertonic_b200.service is GPU-only (TTS_USE_GPU=0 requested)")
        onnx_dir = os.getenv("TTS_ONNX_DIR", "assets/onnx")
        devs = os.getenv("TTS_DEVICES", "").strip()
        tt = (T.MultiGpuTextToSpeech(onnx_dir, T.parse_devices(devs)) if devs
              else T.load_text_to_speech(onnx_dir, True, int(os.getenv("TTS_DEVICE", "0")), lanes=int(os.getenv("TTS_LANES", "1"))))
        _app = create_app(tt, max_batch=int(os.getenv("TTS_MAX_BATCH", "32")), max_wait_ms=float(os.getenv("TTS_MAX_WAIT_MS", "2")))
    return _app


def __getattr__(name):          # `uvicorn supertonic_b200.service:app` builds the engine on first access, not at import
    if name == "app":
        return get_app()
    raise AttributeError(name)
