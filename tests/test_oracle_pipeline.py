"""CPU checks of the oracle pipeline itself (tiny surrogate graphs, seconds): the properties the packed-row CUDA path relies on
must already hold for the reference semantics (oracle/host_ref.ReferenceTTS over the ONNX interpreter):

  * an utterance synthesised alone equals the same utterance inside a padded batch on its valid region — i.e. the reference's
    padding to the batch maximum (cpp/helper.cpp:376, 430) carries no information, which is what makes packing parity-neutral;
  * padded token / frame positions are exactly zero after every masked stage;
  * durations are invariant under batch composition bit for bit (fp64 duration predictor, rounded once).
"""
import numpy as np
import pytest

from oracle import host_ref
from oracle.pipeline import OraclePipeline, make_noise
from tests import _util as U


@pytest.fixture(scope="module")
def ora(tiny_assets):
    return OraclePipeline(tiny_assets)


def test_padding_carries_no_information(ora):
    texts, langs = U.make_batch(7, 3, 20, 90)
    ids, mask = host_ref.unicode_processor_call(ora.indexer, texts, langs)
    ttl, dp = ora.style(["M1", "F1", "M2"])
    nz = np.random.default_rng(1).standard_normal((3, 144, 200)).astype(np.float32)
    tr = {}
    wav, dur = ora.infer_ids(ids, mask, ttl, dp, 2, np.float32(1.05), lambda B, D, L: nz[:, :, :L], tr)
    L = tr["latent_len"]
    wav = wav.reshape(3, L * 3072)
    # masked stages leave exact zeros in the padding
    assert np.all(tr["text_emb"] * (1 - mask) == 0)
    assert np.all(tr["xs"][-1] * (1 - tr["latent_mask"]) == 0)
    for b in range(3):
        t = int(mask[b].sum())
        tr1 = {}
        wav1, dur1 = ora.infer_ids(ids[b:b + 1, :t], mask[b:b + 1, :, :t], ttl[b:b + 1], dp[b:b + 1], 2, np.float32(1.05),
                                   lambda B, D, L1, b=b: nz[b:b + 1, :, :L1], tr1)
        np.testing.assert_array_equal(dur1, dur[b:b + 1])                       # bit-exact durations
        np.testing.assert_array_equal(tr1["wav_lengths"], tr["wav_lengths"][b:b + 1])
        Lb = tr1["latent_len"]
        assert Lb == int(tr["latent_mask"][b].sum())
        assert np.abs(tr1["xs"][-1][0] - tr["xs"][-1][b, :, :Lb]).max() < 2e-5   # fp32 reduction-order noise only
        n = int(tr1["wav_lengths"][0])
        assert U.snr_db(wav1[:n], wav[b, :n]) > 80.0


def test_call_concatenates_untrimmed_chunks_with_silence(ora):
    """TextToSpeech::call (cpp/helper.cpp:685-723): sequential chunks, 0.3 s of zeros between, durations summed in float32."""
    text = " ".join(U.make_text(np.random.default_rng(i), 130) + "." for i in range(4))
    chunks = host_ref.chunk_text(text, 300)
    assert len(chunks) >= 2
    ttl, dp = ora.style(["M1"])
    wav, dur = ora.call(text, "en", ttl, dp, 2, 1.05, 0.3, make_noise(4))
    total, n_samples = np.float32(0), 0
    for i, c in enumerate(chunks):
        w, d = ora._infer([c], ["en"], ttl, dp, 2, np.float32(1.05), make_noise(4))
        n_samples += len(w) + (int(np.float32(0.3) * np.float32(44100)) if i else 0)
        total = d[0] if i == 0 else np.float32(total + np.float32(d[0] + np.float32(0.3)))
    assert len(wav) == n_samples and dur[0] == total
    with pytest.raises(RuntimeError):
        ora.call(text, "en", *ora.style(["M1", "F1"]), 2)
    with pytest.raises(RuntimeError):
        ora.batch(["a", "b"], ["en", "en"], ttl, dp, 2, noise=make_noise(0))


# ---------------------------------------------------------------------------------------------------------------------------------
# The restatement of the reference's ORCHESTRATION pinned to the reference itself: tests/golden/pipeline_golden.json holds what the
# UNMODIFIED TextToSpeech::call / batch (cpp/helper.cpp:469-734, compiled where it lies as oracle/_ref/ref_pipe) did over the
# closed-form stand-ins of oracle/ref_stub_fake/onnxruntime_cxx_api.h; the same stand-ins in numpy drive host_ref.ReferenceTTS here.
from oracle.fake_graphs import fake_runs as _fake_runs  # noqa: E402  (the same formulas as oracle/ref_stub_fake/onnxruntime_cxx_api.h)


def test_orchestration_restatement_matches_the_unmodified_reference(tiny_assets):
    import json, os
    from supertonic_b200 import surrogate
    with open(os.path.join(os.path.dirname(__file__), "golden", "pipeline_golden.json"), encoding="utf-8") as fh:
        golden = json.load(fh)["results"]
    cfg = json.load(open(os.path.join(tiny_assets, "onnx", "tts.json")))
    chunk = int(cfg["ae"]["base_chunk_size"]) * int(cfg["ttl"]["chunk_compress_factor"])
    rng = np.random.default_rng(0)
    noise = lambda B, D, L: rng.standard_normal((B, D, L)).astype(np.float32)      # the stand-ins ignore its values, as the golden run did
    assert len(golden) >= 10
    ok = errs = 0
    for r in golden:
        c = r["case"]
        ttl, dp_style = [], []
        for name in c["styles"]:
            j = json.load(open(os.path.join(tiny_assets, "voice_styles", name)))
            ttl.append(np.asarray(j["style_ttl"]["data"], np.float32).reshape(j["style_ttl"]["dims"]))
            dp_style.append(np.asarray(j["style_dp"]["data"], np.float32).reshape(j["style_dp"]["dims"]))
        ttl, dp_style = np.concatenate(ttl), np.concatenate(dp_style)
        trace = []
        tts = host_ref.ReferenceTTS(cfg, surrogate.build_indexer(), *_fake_runs(trace, chunk))
        try:
            if c["kind"] == "call":
                wav, dur = tts.call(c["text"], c["lang"], ttl, dp_style, c["total_step"], c["speed"], c["silence_duration"], noise)
            else:
                wav, dur = tts.batch(c["texts"], c["langs"], ttl, dp_style, c["total_step"], c["speed"], noise)
        except (RuntimeError, ValueError) as e:
            assert "error" in r and str(e) == r["error"], (c, str(e), r.get("error"))
            errs += 1
            continue
        assert "error" not in r, r.get("error")
        assert len(wav) == r["wav_len"], c
        np.testing.assert_array_equal(np.asarray(dur, np.float32), np.asarray(r["duration"], np.float32))
        samp = np.concatenate([wav[::1009], wav[-1:]])
        np.testing.assert_array_equal(samp, np.asarray(r["wav_samples"], np.float32))
        assert float(np.cumsum(wav, dtype=np.float64)[-1]) == r["wav_sum"]            # sequential double sum, as the driver adds
        # every Session::Run of the reference, in order: graph, input names as passed, shapes, the scalar tensors of the Euler loop
        assert len(trace) == len(r["trace"]), (len(trace), len(r["trace"]))
        for got, want in zip(trace, r["trace"]):
            assert got["graph"] == want["graph"] and got["inputs"] == want["inputs"] and got["shapes"] == want["shapes"], (got["graph"], got["inputs"], want)
            if want["graph"] == "vector_estimator.onnx":
                assert got["total_step"] == want["total_step"] and got["current_step"] == want["current_step"]
                assert want["masked_zero"] and want["is_prev_output"] and got["masked_zero"] and got["is_prev_output"]
        ok += 1
    assert ok >= 7 and errs >= 3


def test_the_two_reference_ports_agree_except_where_recorded():
    """pipeline_golden.json also holds what the UNMODIFIED Python port (py/helper.py, run at golden-generation time with a fake
    `onnxruntime` computing the same stand-ins) returned for the same cases. The ports agree — waveform, durations, which tensors each
    run receives (the Python port passes current_step before total_step; inputs are matched by name) — except for Korean long-form
    text: chunkText counts BYTES in C++ (std::string::length, cpp/helper.cpp:1117-1186) and characters in Python, so `call()` cuts a
    Korean text into more, shorter chunks in C++. The library follows the C++ port (north star: the C++ API is the drop-in surface)."""
    import json, os
    with open(os.path.join(os.path.dirname(__file__), "golden", "pipeline_golden.json"), encoding="utf-8") as fh:
        golden = json.load(fh)["results"]
    agree = differ = 0
    for r in golden:
        py = r.get("py")
        assert py is not None, "regenerate with oracle/make_golden.py where /root/reference is present"
        if "error" in r:
            assert "error" in py and r["error"].split(".")[0] in py["error"], (r["error"], py["error"])      # AssertionError / ValueError texts
            continue
        if r["case"]["kind"] == "call" and r["case"]["lang"] == "ko":
            n_cpp = sum(t["graph"] == "vocoder.onnx" for t in r["trace"]); n_py = sum(t["graph"] == "vocoder.onnx" for t in py["trace"])
            assert n_cpp > n_py >= 2 and py["wav_len"] != r["wav_len"], (n_cpp, n_py)                            # byte vs character chunk limit
            differ += 1
            continue
        assert len(r["trace"]) == len(py["trace"])
        for a, b in zip(r["trace"], py["trace"]):
            assert a["graph"] == b["graph"] and sorted(a["inputs"]) == sorted(b["inputs"])
            assert dict(zip(a["inputs"], map(tuple, a["shapes"]))) == dict(zip(b["inputs"], map(tuple, b["shapes"])))
        assert py["wav_len"] == r["wav_len"] and py["wav_sum"] == r["wav_sum"] and py["wav_samples"] == r["wav_samples"]
        assert py["duration"] == r["duration"]
        agree += 1
    assert agree >= 6 and differ == 1
