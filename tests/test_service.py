"""HTTP serving wrapper (supertonic_b200/service.py) against the reference service's contract (py/service.py:28-136):
validation and status codes, WAV / ZIP responses and their names, trimming to int(sr * duration), and the dynamic batcher
underneath (concurrent requests coalesce into one launch; results are routed back to the right request). CPU-only: the engine is
a stand-in with TextToSpeech's interface — the real one is exercised by tests/test_service_gpu.py."""
import io
import struct
import threading
import time
import zipfile

import numpy as np
import pytest
from starlette.testclient import TestClient

from supertonic_b200 import service
from supertonic_b200.tts import Style, SynthesisResult

SR = 44100


class FakeTTS:
    """duration = 0.01 s per character; waveform = constant (len(text) % 100) / 100 — enough to see routing and trimming."""
    sample_rate = SR

    def __init__(self, delay=0.0):
        self.calls, self.delay = [], delay

    def _one(self, text):
        d = np.float32(0.01 * len(text))
        n = int(SR * float(d)) + 3072                     # untrimmed, like the engine's whole latent frames
        return np.full(n, (len(text) % 100) / 100.0, np.float32), float(d)

    def synthesize_many(self, texts, langs, style, total_step, speed=1.05, max_batch=128, seed=0, copy=False, **kw):
        time.sleep(self.delay)
        assert style.ttl.shape[0] == len(texts) == len(langs)
        self.calls.append(("many", len(texts), total_step, speed))
        return [self._one(t) for t in texts]

    def call(self, text, lang, style, total_step, speed=1.05, silence_duration=0.3):
        self.calls.append(("call", 1, total_step, speed))
        w, d = self._one(text)
        return SynthesisResult(w, np.asarray([d], np.float32))


def fake_styles(paths):
    for p in paths:
        if "missing" in p:
            raise RuntimeError(f"Failed to open voice style file: {p}")
    return Style(np.zeros((len(paths), 50, 256), np.float32), np.zeros((len(paths), 8, 16), np.float32))


@pytest.fixture()
def rig():
    tts = FakeTTS()
    app = service.create_app(tts, fake_styles, max_batch=8, max_wait_ms=1.0)
    with TestClient(app) as c:
        yield c, tts, app
    app.state.batcher.close()


def parse_wav(b):
    assert b[:4] == b"RIFF" and b[8:16] == b"WAVEfmt "
    fmt, ch, sr, _, _, bits = struct.unpack("<hhiihh", b[20:36])
    assert (fmt, ch, sr, bits) == (1, 1, SR, 16) and b[36:40] == b"data"
    n = struct.unpack("<i", b[40:44])[0]
    return np.frombuffer(b[44:44 + n], "<i2")


def test_health(rig):
    c, _, _ = rig
    r = c.get("/health")
    assert r.status_code == 200 and r.json() == {"status": "ok"}


def test_single_request_returns_trimmed_pcm16_wav(rig):
    c, tts, _ = rig
    text = "Hello there, world"
    r = c.post("/tts", json={"text": text})
    assert r.status_code == 200 and r.headers["content-type"] == "audio/wav"
    assert r.headers["content-disposition"] == 'attachment; filename="Hello_there__world.wav"'
    pcm = parse_wav(r.content)
    assert len(pcm) == int(SR * float(np.float32(0.01 * len(text))))           # py/service.py:61-69
    assert np.all(pcm == int(np.float32(0.18) * np.float32(32767)))            # clamp * 32767, truncated (cpp/helper.cpp:985-988)
    assert tts.calls == [("many", 1, 5, 1.05)]


def test_batch_request_returns_zip_in_order(rig):
    c, _, _ = rig
    texts = ["first one", "the second text", "", "네 번째"]
    r = c.post("/tts", json={"text": texts, "lang": ["en", "en", "en", "ko"], "voice_style": ["a.json"] * 4, "batch": True,
                             "total_step": 3, "speed": 1.2})
    assert r.status_code == 200 and r.headers["content-type"] == "application/zip"
    assert r.headers["content-disposition"] == 'attachment; filename="tts_outputs.zip"'
    zf = zipfile.ZipFile(io.BytesIO(r.content))
    assert zf.namelist() == ["first_one.wav", "the_second_text.wav", "tts_3.wav", "네_번째.wav"]
    for name, t in zip(zf.namelist(), texts):
        assert len(parse_wav(zf.read(name))) == int(SR * float(np.float32(0.01 * len(t))))


def test_validation_matches_the_reference_service(rig):
    c, _, _ = rig
    r = c.post("/tts", json={"text": ["a", "b"]})
    assert r.status_code == 400 and r.json()["detail"] == "Non-batch mode requires single text, lang, and voice_style."
    r = c.post("/tts", json={"text": ["a", "b"], "lang": ["en"], "voice_style": ["x", "y"], "batch": True})
    assert r.status_code == 400 and r.json()["detail"] == "text, lang, and voice_style must have the same length."
    r = c.post("/tts", json={"text": "a", "lang": "xx"})
    assert r.status_code == 400 and r.json()["detail"] == "Invalid language(s): xx"
    r = c.post("/tts", json={"text": "a", "voice_style": "missing.json"})
    assert r.status_code == 400 and "Failed to open voice style file" in r.json()["detail"]
    assert c.post("/tts", json={"text": "a", "total_step": 0}).status_code == 422
    assert c.post("/tts", json={"text": "a", "total_step": 51}).status_code == 422
    assert c.post("/tts", json={"text": "a", "speed": 0}).status_code == 422
    assert c.post("/tts", json={"text": "a", "silence_duration": -1}).status_code == 422
    assert c.post("/tts", json={}).status_code == 422


def test_long_text_keeps_call_semantics(rig):
    c, tts, _ = rig
    text = ("This sentence is repeated to pass the three hundred byte chunk limit. " * 6).strip()
    r = c.post("/tts", json={"text": text, "silence_duration": 0.5})
    assert r.status_code == 200 and tts.calls[-1][0] == "call"


def test_concurrent_requests_are_coalesced_and_routed_back():
    tts = FakeTTS(delay=0.05)
    app = service.create_app(tts, fake_styles, max_batch=16, max_wait_ms=30.0)
    with TestClient(app) as c:
        out = {}

        def one(i):
            text = "x" * (10 + i)
            r = c.post("/tts", json={"text": text, "total_step": 5 if i % 4 else 7})
            out[i] = (r.status_code, len(parse_wav(r.content)), int(SR * float(np.float32(0.01 * len(text)))))

        th = [threading.Thread(target=one, args=(i,)) for i in range(12)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert all(v[0] == 200 and v[1] == v[2] for v in out.values()), out
        st = c.get("/stats").json()
    app.state.batcher.close()
    assert st["utterances"] == 12 and st["launches"] < 12 and st["max_coalesced"] > 1, st
    assert all(k == "many" for k, *_ in tts.calls)
    assert {ts for _, _, ts, _ in tts.calls} == {5, 7}            # different total_step never share a launch
