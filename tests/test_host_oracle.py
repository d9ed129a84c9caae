"""Pins oracle/host_ref.py (the CPU restatement of the reference's host arithmetic) against
golden vectors produced by RUNNING the unmodified reference cpp/helper.cpp
(tests/golden/host_golden.json ← oracle/make_golden.py)."""
import numpy as np
import pytest

from oracle import host_ref as hr
from supertonic_b200 import surrogate


def _b(s):
    return s.encode("utf-8", "surrogateescape")


def _by_kind(golden, kind):
    return [r for r in golden if r["case"]["kind"] == kind]


def test_text_frontend_matches_reference(host_golden):
    indexer = surrogate.build_indexer()
    n = 0
    for r in _by_kind(host_golden, "text"):
        c = r["case"]
        if "error" in r:
            with pytest.raises(RuntimeError) as e:
                hr.unicode_processor_call(indexer, c["texts"], c["langs"])
            assert str(e.value) == r["error"]
            continue
        ids, mask = hr.unicode_processor_call(indexer, c["texts"], c["langs"])
        np.testing.assert_array_equal(ids, np.asarray(r["text_ids"], np.int64), err_msg=str(c))
        np.testing.assert_array_equal(mask, np.asarray(r["text_mask"], np.float32), err_msg=str(c))
        n += 1
    assert n >= 14


def test_default_sentence_is_154_tokens(host_golden):
    r = _by_kind(host_golden, "text")[0]            # cpp/example_onnx.cpp:17, SURVEY.md App. G
    assert len(r["text_ids"][0]) == 154


def test_chunk_text_matches_reference(host_golden):
    rs = _by_kind(host_golden, "chunk")
    assert len(rs) >= 9
    for r in rs:
        got = hr.chunk_text(r["case"]["text"], r["case"]["max_len"])
        assert got == [_b(x) for x in r["chunks"]], r["case"]
    assert [len(_b(x)) for x in rs[0]["chunks"]] == [199, 265, 187]


def test_masks_match_reference(host_golden):
    for r in _by_kind(host_golden, "latent_mask"):
        c = r["case"]
        got = hr.get_latent_mask(c["wav_lengths"], c["base_chunk_size"], c["chunk_compress_factor"])
        np.testing.assert_array_equal(got, np.asarray(r["mask"], np.float32))
    for r in _by_kind(host_golden, "length_mask"):
        c = r["case"]
        got = hr.length_to_mask(c["lengths"], c.get("max_len", -1))
        np.testing.assert_array_equal(got, np.asarray(r["mask"], np.float32).reshape(got.shape))


def test_sanitize_and_wav_match_reference(host_golden):
    for r in _by_kind(host_golden, "sanitize"):
        assert hr.sanitize_filename(r["case"]["text"], r["case"]["max_len"]) == _b(r["name"])
    for r in _by_kind(host_golden, "wav"):
        assert hr.wav_bytes(r["case"]["samples"], r["case"]["sample_rate"]) == bytes(r["bytes"])


def test_cfg_and_style_loading(host_golden, tiny_assets):
    import json, os
    (r,) = _by_kind(host_golden, "cfg")
    assert r["cfg"] == [44100, 512, 6, 24]
    (r,) = _by_kind(host_golden, "style")
    ttl = []
    for p in r["case"]["paths"]:
        j = json.load(open(os.path.join(tiny_assets, "voice_styles", p)))
        ttl.append(np.asarray(j["style_ttl"]["data"], np.float32))
    ttl = np.concatenate(ttl)
    assert list(ttl.shape) == r["ttl_shape"]
    np.testing.assert_allclose(ttl.astype(np.float64).sum(), r["ttl_sum"], rtol=1e-6)
    np.testing.assert_array_equal(ttl.reshape(-1)[:4], np.asarray(r["ttl_head"], np.float32))


def test_latent_geometry_matches_reference(host_golden):
    """`latent_geometry` (the restatement of TextToSpeech::sampleNoisyLatent's length arithmetic, cpp/helper.cpp:424-467) against the
    UNMODIFIED reference run on float32 durations — around multiples of the 3072-sample latent frame, random ones, and the
    float32-vs-float64 case below: latent length, latent channel count and mask; the reference's noise is zero exactly where the mask is."""
    rs = _by_kind(host_golden, "noisy_latent")
    assert len(rs) >= 17
    for r in rs:
        d = np.asarray(r["case"]["duration"], np.float32)
        wl, L, mask = hr.latent_geometry(d, 44100, 512, 6)
        assert r["latent_shape"] == [len(d), 24 * 6, L], r["case"]
        np.testing.assert_array_equal(mask, np.asarray(r["mask"], np.float32))
        assert r["masked_zero"] and r["live_nonzero"]


def test_latent_geometry_float32_semantics():
    """cpp/helper.cpp:430-438: float32 formula == integer formula of :767 on every draw (App. G)."""
    rng = np.random.default_rng(0)
    for _ in range(2000):
        d = rng.uniform(0.5, 25.0, size=rng.integers(1, 6)).astype(np.float32)
        wl, L, mask = hr.latent_geometry(d, 44100, 512, 6)
        assert mask.shape[2] == L
        assert wl.dtype == np.int64 and (wl == (d * np.float32(44100)).astype(np.int64)).all()
    # a case where float64 arithmetic would disagree with the float32 formula must follow float32
    d = np.asarray([2.0897958], np.float32)
    _, L, _ = hr.latent_geometry(d, 44100, 512, 6)
    assert L == int((np.float32(d[0] * np.float32(44100)) + np.float32(3072) - np.float32(1)) / np.float32(3072))
