"""Multi-GPU engine (north_star: "a request batch is length-bucketed and partitioned across the 8 B200s of one box ... no collective
runs on the hot path"; the reference entry point that fans out is TextToSpeech::batch, cpp/helper.cpp:725-734).

CPU part: the planning / dealing / re-assembly logic of tts.MultiGpuTextToSpeech over fake engines. GPU part: two engines (on two
GPUs when the box has them, else two handles on one GPU) against one engine for the same request — durations bit-equal, waveforms
equal up to what batch composition is allowed to change (SNR >= 60 dB with the fp16 vocoder)."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import _util as U


class _FakeEngine:
    """Stands in for capi.Engine: real text front-end (host-only), synthetic 'audio' that depends only on the utterance's token
    count and its noise_index, so any mistake in grouping / sharding / re-assembly shows up in the output."""

    class _Cfg:
        chunk_size, sample_rate, latent_channels = 3072, 44100, 144

    def __init__(self, fe, name):
        self.fe, self.name, self.cfg, self.calls, self.waits = fe, name, self._Cfg(), [], 0

    def text_to_ids(self, texts, langs):
        return self.fe(texts, langs)

    def synthesize_joined(self, ids, mask, ttl, dp, total_step, speed, seed=0, noise=None, noise_index=None, pcm16=False, pinned=False,
                          wait=True, gap_samples=0):
        B = ids.shape[0]
        lens = mask.reshape(B, -1).sum(1).astype(np.int64)
        assert (np.diff(np.asarray(noise_index)) > 0).all(), "a launch group holds request indices in ascending order"
        self.calls.append((self.name, [int(i) for i in noise_index], pinned))
        frames = (lens + 9) // 10
        off = np.concatenate([[0], np.cumsum(frames * 3072)])
        out = np.zeros(int(off[-1]), np.float32)
        wl = frames * 3072 - 7
        for b in range(B):
            out[off[b]:off[b] + wl[b]] = float(noise_index[b]) + 0.001 * float(lens[b]) + float(ttl[b].sum())
        return dict(out=out, offsets=off, duration=(lens * 0.07).astype(np.float32), wav_lengths=wl, frames=frames)

    def wait(self):
        self.waits += 1


def _request(n, seed=3):
    texts, langs = U.make_batch(seed, n, 20, 300)
    return texts, langs


def test_parse_devices():
    from supertonic_b200 import tts as T
    assert T.parse_devices("0-3") == [0, 1, 2, 3]
    assert T.parse_devices("0,2, 5") == [0, 2, 5]
    assert T.parse_devices("0-1,4") == [0, 1, 4]
    with pytest.raises(ValueError):
        T.parse_devices(" ")


@pytest.mark.parametrize("ndev", [1, 2, 3, 8])
def test_requests_are_dealt_out_and_reassembled_in_input_order(tiny_assets, ndev):
    from supertonic_b200 import capi, tts as T
    fe = capi.Frontend(os.path.join(tiny_assets, "onnx", "unicode_indexer.json"))
    texts, langs = _request(150)
    rng = np.random.default_rng(0)
    style = T.Style(rng.standard_normal((150, 2, 3)).astype(np.float32), rng.standard_normal((150, 1, 2)).astype(np.float32))
    engs = [_FakeEngine(fe, f"e{k}") for k in range(ndev)]
    multi = T.MultiGpuTextToSpeech(None, list(range(ndev)), engines=engs)
    got = multi.synthesize_many(texts, langs, style, 5, 1.05, max_batch=16, seed=4)
    one = T.TextToSpeech.__new__(T.TextToSpeech)          # the single-engine path over the same fake
    one.engine = _FakeEngine(fe, "solo")
    want = T.TextToSpeech.synthesize_many(one, texts, langs, style, 5, 1.05, max_batch=16, seed=4)
    assert len(got) == len(want) == 150
    for (w1, d1), (w2, d2) in zip(got, want):
        np.testing.assert_array_equal(w1, w2)
        assert d1 == d2
    # every utterance went to exactly one engine, exactly once; every engine that got work was waited for
    seen = sorted(i for e in engs for c in e.calls for i in c[1])
    assert seen == list(range(150))
    assert all(e.waits == 1 for e in engs if e.calls)
    if ndev > 1:
        loads = [sum(len(c[1]) for c in e.calls) for e in engs]
        assert max(loads) - min(loads) <= 32, loads          # LPT over ~10 groups of 16: no engine is left idle


def test_style_count_mismatch_is_the_reference_error(tiny_assets):
    from supertonic_b200 import capi, tts as T
    fe = capi.Frontend(os.path.join(tiny_assets, "onnx", "unicode_indexer.json"))
    multi = T.MultiGpuTextToSpeech(None, [0, 1], engines=[_FakeEngine(fe, "a"), _FakeEngine(fe, "b")])
    with pytest.raises(RuntimeError, match="Number of texts must match number of style vectors"):
        multi.synthesize_many(["a", "b"], ["en", "en"], T.Style(np.zeros((1, 2, 3), np.float32), np.zeros((1, 1, 2), np.float32)), 5)


@pytest.mark.gpu
def test_two_engines_equal_one_engine_per_utterance(full_assets):
    import torch
    from supertonic_b200 import tts as T
    devs = [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]
    texts, langs = _request(40, seed=8)
    voices = [("M1", "F1", "M2", "F2")[i % 4] for i in range(40)]
    style = T.load_voice_style([os.path.join(full_assets, "voice_styles", v + ".json") for v in voices])
    one = T.load_text_to_speech(os.path.join(full_assets, "onnx"), True, 0)
    multi = T.MultiGpuTextToSpeech(os.path.join(full_assets, "onnx"), devs)
    try:
        want = one.synthesize_many(texts, langs, style, 3, 1.05, max_batch=12, seed=21, copy=True)
        got = multi.synthesize_many(texts, langs, style, 3, 1.05, max_batch=12, seed=21, copy=True)
        again = multi.synthesize_many(texts, langs, style, 3, 1.05, max_batch=12, seed=21, copy=True)
        for (w1, d1), (w2, d2), (w3, d3) in zip(got, want, again):
            assert d1 == d2 == d3 and len(w1) == len(w2)
            np.testing.assert_array_equal(w1, w3)                      # the same plan on the same engines: bit-identical
            assert U.snr_db(w1, w2) >= 60.0
        # another grouping of the same request (other launch groups, other row-tile counts): same noise streams, same audio
        other = one.synthesize_many(texts, langs, style, 3, 1.05, max_batch=40, seed=21, copy=True)
        for (w1, d1), (w2, d2) in zip(other, want):
            assert d1 == d2 and U.snr_db(w1, w2) >= 60.0
        pcm = multi.synthesize_many(texts, langs, style, 3, 1.05, max_batch=12, seed=21, copy=True, pcm16=True)
        for (q, _), (w, _) in zip(pcm, got):
            np.testing.assert_array_equal(q, (np.clip(w, np.float32(-1), np.float32(1)) * np.float32(32767)).astype(np.int16))
    finally:
        multi.close(); one.engine.close()


@pytest.mark.gpu
def test_a_repeated_request_replays_its_cuda_graphs(full_assets):
    """The launch groups of a request must not depend on the frames-per-token ratio the engines measured on the previous call
    (each engine of a multi-device object sees other groups, so its ratio jitters): the groups are balanced on integer token
    counts. A third identical request captures no new CUDA graph on any engine — a re-capture costs ~100 ms and is serialised
    across the engines of a process (8 devices: 6 k instead of 200 k audio-s/s when the plan was unstable)."""
    from supertonic_b200 import tts as T
    texts, langs = _request(96, seed=4)
    voices = [("M1", "F1", "M2", "F2")[i % 4] for i in range(96)]
    style = T.load_voice_style([os.path.join(full_assets, "voice_styles", v + ".json") for v in voices])
    multi = T.MultiGpuTextToSpeech(os.path.join(full_assets, "onnx"), [0, 0, 0])
    try:
        for k in range(2):
            multi.synthesize_many(texts, langs, style, 2, 1.05, max_batch=16, seed=k)
        caps = [e.kernel_variants()["graph_captures"] for e in multi.engines]
        plan = T.plan_many(multi.engine, texts, langs, 16)
        multi.synthesize_many(texts, langs, style, 2, 1.05, max_batch=16, seed=7)
        assert [e.kernel_variants()["graph_captures"] for e in multi.engines] == caps
        assert T.plan_many(multi.engine, texts, langs, 16).groups == plan.groups
    finally:
        multi.close()


@pytest.mark.gpu
def test_request_lanes_give_the_single_handle_result(full_assets):
    """TextToSpeech(lanes=...): launch groups dealt alternately to two handles on one GPU (consecutive requests then overlap on the
    device). Same weights, same plan, noise keyed by the index in the request: bit-identical audio, also for a request stream
    issued with wait=False."""
    from supertonic_b200 import tts as T
    texts, langs = _request(40, seed=9)
    voices = [("M1", "F1", "M2", "F2")[i % 4] for i in range(40)]
    style = T.load_voice_style([os.path.join(full_assets, "voice_styles", v + ".json") for v in voices])
    one = T.load_text_to_speech(os.path.join(full_assets, "onnx"), True, 0)
    two = T.load_text_to_speech(os.path.join(full_assets, "onnx"), True, 0, lanes=2)
    try:
        assert len(two.lanes) == 2 and two.lanes[0] is two.engine
        for k in range(3):          # the first request also settles both planners' frames-per-token estimate
            want = one.synthesize_many(texts, langs, style, 3, 1.05, max_batch=12, seed=30 + k, copy=True)
            got = two.synthesize_many(texts, langs, style, 3, 1.05, max_batch=12, seed=30 + k, copy=True)
        for (w1, d1), (w2, d2) in zip(got, want):
            assert d1 == d2
            np.testing.assert_array_equal(w1, w2)
        assert all(e.launches > 0 for e in two.lanes)          # both handles took launch groups
        # request stream: three requests in flight over the two lanes, results read after wait()
        outs = [two.synthesize_many(texts[:10], langs[:10], T.Style(style.ttl[:10], style.dp[:10]), 3, 1.05, max_batch=12, seed=50, wait=False)
                for _ in range(2)]
        two.wait()
        ref = one.synthesize_many(texts[:10], langs[:10], T.Style(style.ttl[:10], style.dp[:10]), 3, 1.05, max_batch=12, seed=50, copy=True)
        for o in outs:
            for (w1, d1), (w2, d2) in zip(o, ref):
                assert d1 == d2
                np.testing.assert_array_equal(w1, w2)
    finally:
        one.engine.close()
        for e in two.lanes:
            e.close()


@pytest.mark.gpu
def test_cpp_multi_gpu_many_equals_single_engine(full_assets):
    """supertonic::MultiGpuTextToSpeech::many (csrc/tts_host.cc: std::thread per device) against TextToSpeech::many."""
    import torch
    from supertonic_b200 import capi
    lib = C.CDLL(os.path.join(os.path.dirname(capi.LIB_PATH), "libsupertonic_host.so"))
    n = 24
    texts, langs = _request(n, seed=12)
    styles = [os.path.join(full_assets, "voice_styles", ("M1", "F1", "M2", "F2")[i % 4] + ".json") for i in range(n)]
    devs = [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]
    arr = lambda xs: (C.c_char_p * len(xs))(*[x.encode() for x in xs])       # noqa: E731
    dur = np.zeros((2, n), np.float32); ns = np.zeros((2, n), np.int64); asum = np.zeros((2, n), np.float64)
    rc = lib.stc_host_many_check(os.path.join(full_assets, "onnx").encode(), (C.c_int * 2)(*devs), 2, arr(texts), arr(langs), n, arr(styles),
                                 C.c_int(3), C.c_float(1.05), C.c_int(7), dur.ctypes.data_as(C.c_void_p), ns.ctypes.data_as(C.c_void_p),
                                 asum.ctypes.data_as(C.c_void_p))
    assert rc == 0
    np.testing.assert_array_equal(dur[0], dur[1])
    np.testing.assert_array_equal(ns[0], ns[1])
    assert (ns[0] > 0).all()
    np.testing.assert_allclose(asum[0], asum[1], rtol=2e-3)
