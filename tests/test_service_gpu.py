"""The HTTP serving wrapper on the real engine (B200 box): /tts responses equal what the library returns for the same
utterance, whatever the request was coalesced with (reference contract py/service.py:84-136)."""
import io
import os
import threading
import zipfile

import numpy as np
import pytest

from tests import _util as U
from tests.test_service import parse_wav

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rig():
    from starlette.testclient import TestClient
    from supertonic_b200 import service, surrogate, tts as T
    root = surrogate.ensure_assets("tiny")
    tt = T.load_text_to_speech(root + "/onnx")
    app = service.create_app(tt, max_batch=16, max_wait_ms=20.0)
    with TestClient(app) as c:
        yield dict(c=c, tt=tt, root=root, T=T, app=app)
    app.state.batcher.close()
    tt.engine.close()


def _style(rig, name):
    return os.path.join(rig["root"], "voice_styles", name + ".json")


def test_single_request_equals_the_library_call(rig):
    T, tt = rig["T"], rig["tt"]
    text = "The quick brown fox jumps over a lazy dog."
    r = rig["c"].post("/tts", json={"text": text, "voice_style": _style(rig, "M1")})
    assert r.status_code == 200 and r.headers["content-type"] == "audio/wav"
    pcm = parse_wav(r.content)
    style = T.load_voice_style([_style(rig, "M1")])
    (w, d), = tt.synthesize_many([text], ["en"], style, 5, 1.05, seed=1, copy=True)      # the batcher's first launch uses seed 1
    want = parse_wav(T.wav_file_bytes(w[: int(tt.sample_rate * d)], tt.sample_rate))
    assert len(pcm) == len(want) and np.array_equal(pcm, want)
    assert np.abs(pcm.astype(np.int32)).max() > 0


def test_concurrent_and_batch_requests(rig):
    T, tt, c = rig["T"], rig["tt"], rig["c"]
    texts, langs = U.make_batch(5, 10, 20, 120)
    voices = [("M1", "F1", "M2", "F2")[i % 4] for i in range(len(texts))]
    style = T.load_voice_style([_style(rig, v) for v in voices])
    ref = tt.synthesize_many(texts, langs, style, 4, 1.0, seed=3, copy=True)
    want = [int(tt.sample_rate * d) for _, d in ref]                                     # durations do not depend on the noise
    got = {}

    def one(i):
        r = c.post("/tts", json={"text": texts[i], "lang": langs[i], "voice_style": _style(rig, voices[i]), "total_step": 4, "speed": 1.0})
        got[i] = (r.status_code, len(parse_wav(r.content)) if r.status_code == 200 else -1)

    th = [threading.Thread(target=one, args=(i,)) for i in range(len(texts))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert all(got[i] == (200, want[i]) for i in range(len(texts))), (got, want)
    st = c.get("/stats").json()
    assert st["max_coalesced"] > 1, st
    r = c.post("/tts", json={"text": texts[:3], "lang": langs[:3], "voice_style": [_style(rig, v) for v in voices[:3]], "batch": True,
                             "total_step": 4, "speed": 1.0})
    assert r.status_code == 200 and r.headers["content-type"] == "application/zip"
    zf = zipfile.ZipFile(io.BytesIO(r.content))
    assert [len(parse_wav(zf.read(n))) for n in zf.namelist()] == want[:3]


def test_long_text_goes_through_call(rig):
    text = " ".join(["This sentence is repeated until the text needs more than one chunk of three hundred bytes."] * 5)
    r = rig["c"].post("/tts", json={"text": text, "voice_style": _style(rig, "F1"), "silence_duration": 0.25})
    assert r.status_code == 200
    assert len(parse_wav(r.content)) > 0
