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
