"""The drop-in boundary, exercised on the B200 box from the C++ side.

1. `supertonic_b200/bin/example_cuda` — the C++ host API (csrc/tts_host.h, the reference's cpp/helper.h surface over the
   fast layer) — must write the same PCM bytes as the Python mirror given the same noise seed, for `call`, `--batch`
   and `--many`.
2. `oracle/_ref/example_onnx_stc` — the UNMODIFIED reference cpp/example_onnx.cpp + cpp/helper.cpp compiled against
   include/ort_shim/onnxruntime_cxx_api.h and linked to libsupertonic_cuda.so (built in the CPU container by
   oracle/build_ref.sh; /root/reference itself does not exist on the GPU box) — must run end to end, and because its host code
   derives the output length from OUR duration_predictor, every file must hold exactly int(sr * duration) samples with
   the duration the library computes (the reference's noise is unseedable, so sample values are not compared).
"""
import os
import struct
import subprocess

import numpy as np
import pytest

from tests import _util as U

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "supertonic_b200", "bin", "example_cuda")
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "example_onnx_stc")
TEXTS = ["A first short sentence for the drop-in test.", "And a second one, slightly longer than the first."]


def _read_wav(path):
    b = open(path, "rb").read()
    assert b[:4] == b"RIFF" and b[8:16] == b"WAVEfmt " and b[36:40] == b"data"
    fmt, ch, sr, _, _, bits = struct.unpack("<hhiihh", b[20:36])
    assert (fmt, ch, bits) == (1, 1, 16)
    n = struct.unpack("<i", b[40:44])[0]
    assert len(b) == 44 + n
    return sr, np.frombuffer(b[44:], "<i2")


@pytest.fixture(scope="module")
def rig():
    from supertonic_b200 import surrogate, tts
    root = surrogate.ensure_assets("tiny")
    t = tts.load_text_to_speech(root + "/onnx")
    yield dict(root=root, tts=t, mod=tts)
    t.engine.close()


def _run(cmd, cwd):
    p = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p.stdout


def test_cpp_host_cli_equals_python_mirror(rig, tmp_path):
    assert os.path.exists(CLI), "build first: make -C supertonic_b200/csrc"
    root, T = rig["root"], rig["mod"]
    voices = [f"{root}/voice_styles/M1.json", f"{root}/voice_styles/F1.json"]
    # --- call(): single text, sequential chunks
    long_text = " ".join(U.make_text(np.random.default_rng(i), 120) + "." for i in range(4))     # > 300 bytes -> 2+ chunks
    out = _run([CLI, "--onnx-dir", root + "/onnx", "--voice-style", voices[0], "--text", long_text, "--n-test", "1", "--total-step", "2",
                "--save-dir", str(tmp_path / "a"), "--seed", "0"], ROOT)
    assert "Saved:" in out
    (f,) = os.listdir(tmp_path / "a")
    assert f == T.sanitize_filename(long_text, 20) + "_1.wav"
    sr, pcm = _read_wav(tmp_path / "a" / f)
    t = rig["tts"]; t.noise_seed = 0; t._calls = 0
    r = t.call(long_text, "en", T.load_voice_style(voices[:1]), 2, 1.05)
    assert len(T.chunk_text(long_text, 300)) >= 2
    want = T.wav_file_bytes(r.wav[:int(np.float32(sr) * r.duration[0])], sr)
    assert open(tmp_path / "a" / f, "rb").read() == want
    # --- batch(): two texts, two voices, one padded rectangle
    _run([CLI, "--onnx-dir", root + "/onnx", "--voice-style", ",".join(voices), "--text", "|".join(TEXTS), "--lang", "en,en", "--batch",
          "--n-test", "1", "--total-step", "2", "--save-dir", str(tmp_path / "b"), "--seed", "5"], ROOT)
    t.noise_seed = 5; t._calls = 0
    r = t.batch(TEXTS, ["en", "en"], T.load_voice_style(voices), 2, 1.05)
    row = len(r.wav) // 2
    for b in range(2):
        f = T.sanitize_filename(TEXTS[b], 20) + "_1.wav"
        want = T.wav_file_bytes(r.wav[b * row:b * row + int(np.float32(sr) * r.duration[b])], sr)
        assert open(tmp_path / "b" / f, "rb").read() == want
    # --- many(): packed rows; same durations as batch(), sample counts = int(d*sr) computed on the device
    _run([CLI, "--onnx-dir", root + "/onnx", "--voice-style", ",".join(voices), "--text", "|".join(TEXTS), "--lang", "en,en", "--many",
          "--n-test", "1", "--total-step", "2", "--save-dir", str(tmp_path / "c"), "--seed", "5"], ROOT)
    for b in range(2):
        _, pcm = _read_wav(tmp_path / "c" / (T.sanitize_filename(TEXTS[b], 20) + "_1.wav"))
        assert len(pcm) == int(np.float32(r.duration[b]) * np.float32(sr))


def test_unmodified_reference_cli_runs_on_the_shim(rig, tmp_path):
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref/example_onnx_stc not built (needs /root/reference at build time)")
    root, T = rig["root"], rig["mod"]
    voices = [f"{root}/voice_styles/M1.json", f"{root}/voice_styles/F2.json"]
    out = _run([REF_CLI, "--onnx-dir", root + "/onnx", "--voice-style", ",".join(voices), "--text", "|".join(TEXTS), "--lang", "en,en",
                "--batch", "--n-test", "2", "--total-step", "3", "--save-dir", str(tmp_path)], ROOT)
    assert "Synthesis completed successfully" in out
    eng = rig["tts"].engine
    ids, mask = eng.text_to_ids(TEXTS, ["en", "en"])
    st = T.load_voice_style(voices)
    dur = eng.duration(ids, st.dp, mask) / np.float32(1.05)
    for n in (1, 2):
        for b in range(2):
            sr, pcm = _read_wav(tmp_path / (T.sanitize_filename(TEXTS[b], 20) + f"_{n}.wav"))
            assert sr == 44100 and len(pcm) == int(np.float32(sr) * dur[b])
            assert np.abs(pcm.astype(np.int32)).max() > 0
    # single-text long-form path of the reference (TextToSpeech::call -> chunkText -> sequential _infer)
    long_text = " ".join(U.make_text(np.random.default_rng(i), 110) + "." for i in range(4))
    _run([REF_CLI, "--onnx-dir", root + "/onnx", "--voice-style", voices[0], "--text", long_text, "--n-test", "1", "--total-step", "2",
          "--save-dir", str(tmp_path / "long")], ROOT)
    (f,) = os.listdir(tmp_path / "long")
    sr, pcm = _read_wav(tmp_path / "long" / f)
    chunks = T.chunk_text(long_text, 300)
    total = np.float32(0)
    for i, c in enumerate(chunks):
        ci, cm = eng.text_to_ids([c], ["en"])
        d = (eng.duration(ci, st.dp[:1], cm) / np.float32(1.05))[0]
        total = d if i == 0 else np.float32(total + np.float32(d + np.float32(0.3)))
    assert len(pcm) == int(np.float32(sr) * total)
