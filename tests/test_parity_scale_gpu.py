"""GPU parity tests AT BENCHMARK SCALE (run on the B200 box).

The kernels `bench.py` times are picked by size-dependent dispatchers (fused ConvNeXt MLP form by row-tile count, one-SM vs
two-SM tcgen05 GEMM by tile count, register vs shared-memory-ring depthwise conv by chain length, tcgen05 vs CUDA-core
attention by key count). The small-batch tests in test_parity_gpu.py never reach the large-batch variants, so every case here
is sized to land in one of them, is compared with the ORACLE (oracle/pipeline.py — the CPU restatement of the reference's
`_infer`, cpp/helper.cpp:469-683) on identical inputs with injected noise, and asserts through `stc_kernel_variants` that the
variant it is meant to cover actually ran.

Tolerances are those of test_parity_gpu.py: durations / wav_lengths / frame counts bit-exact; latents max-abs <= 2e-4
(north-star bound 1e-3); waveform SNR >= 60 dB with the default single-pass fp16 vocoder (north-star bound 40 dB).
"""
import numpy as np
import pytest

from tests import _util as U

pytestmark = pytest.mark.gpu

LAT_TOL_NORTH_STAR, LAT_TOL_EXPECTED = 1e-3, 2e-4
SNR_NORTH_STAR, SNR_EXPECTED, SNR_EXACT = 40.0, 60.0, 80.0


@pytest.fixture(scope="module")
def rig():
    from oracle.pipeline import OraclePipeline
    from supertonic_b200 import capi, surrogate
    root = surrogate.ensure_assets("full")
    eng = capi.Engine(root + "/onnx")
    yield dict(root=root, eng=eng, ora=OraclePipeline(root), capi=capi)
    eng.close()


def _delta(before, after):
    return {k: after.get(k, 0) - before.get(k, 0) for k in after if after.get(k, 0) != before.get(k, 0)}


def _bench_batch(rig, n, seed):
    """Exactly what bench.py builds for --batch n (seed 1234 is configs[1])."""
    import bench
    from oracle import host_ref
    texts, langs, voices = bench.workload(n, seed)
    ids, mask = rig["eng"].text_to_ids(texts, langs)
    ids_ref, mask_ref = host_ref.unicode_processor_call(rig["ora"].indexer, texts, langs)
    np.testing.assert_array_equal(ids, ids_ref); np.testing.assert_array_equal(mask, mask_ref)
    ttl, dp = U.styles(rig["root"], voices)
    return ids, mask, ttl, dp


def _packed_vs_oracle(rig, ids, mask, ttl, dp, steps, noise_seed, check_all=True):
    """stc_synthesize_packed (the throughput entry point bench.py's e2e leg and synthesize_many use) against the oracle's padded
    `_infer` of the same batch with the same noise: utterances are independent on both sides, so every utterance's valid region
    must agree. Returns (total latent frames, worst latent error, worst waveform SNR)."""
    B = ids.shape[0]
    nz = np.random.default_rng(noise_seed).standard_normal((B, 144, 420)).astype(np.float32)
    tr = {}
    wav_ref, dur_ref = rig["ora"].infer_ids(ids, mask, ttl, dp, steps, np.float32(1.05), lambda b, d, L: nz[:, :, :L], tr)
    assert tr["latent_len"] <= nz.shape[2]
    out = rig["eng"].synthesize_packed(ids, mask, ttl, dp, steps, 1.05, noise=nz, want_latent=True)
    np.testing.assert_array_equal(out["duration"], dur_ref)
    np.testing.assert_array_equal(out["wav_lengths"], tr["wav_lengths"])
    np.testing.assert_array_equal(out["frames"], tr["latent_mask"].reshape(B, -1).sum(1).astype(np.int64))     # getLatentMask row sums
    wav_ref = wav_ref.reshape(B, -1)
    worst_err, worst_snr = 0.0, 1e9
    for b in range(B):
        Lb, n = int(out["frames"][b]), int(tr["wav_lengths"][b])
        err = float(np.abs(out["latent"][b].T - tr["xs"][-1][b, :, :Lb]).max())
        snr = U.snr_db(out["wavs"][b], wav_ref[b, :n])
        assert len(out["wavs"][b]) == n
        worst_err, worst_snr = max(worst_err, err), min(worst_snr, snr)
        if check_all:
            assert err <= LAT_TOL_NORTH_STAR and err <= LAT_TOL_EXPECTED, (b, err)
            assert snr >= SNR_NORTH_STAR and snr >= SNR_EXPECTED, (b, snr)
    return int(out["frames"].sum()), worst_err, worst_snr


def test_configs1_batch_matches_the_oracle(rig):
    """configs[1] exactly as bench.py builds it: 32 utterances, 4 621 latent frames = 37 row tiles, 5 858 text tokens, total_step 5.
    Reaches the 4-slice fused MLP in its CTA-pair form, the two-SM fp16 vocoder GEMMs, the BN = 128/256 one-SM tiles and the ring depthwise conv."""
    ids, mask, ttl, dp = _bench_batch(rig, 32, 1234)
    v0 = rig["eng"].kernel_variants()
    frames, err, snr = _packed_vs_oracle(rig, ids, mask, ttl, dp, 5, 77)
    assert frames == 4621
    d = _delta(v0, rig["eng"].kernel_variants())
    assert d.get("gemm2_f16", 0) > 0, d                       # vocoder pw2 / conv_in / head on CTA pairs
    assert d.get("gemm2_f16_astat", 0) > 0, d                 # vocoder pw1: A rows resident in shared memory (gemm2_astat.cuh)
    assert d.get("dwconv_ln_chain", 0) >= 10, d               # vocoder depthwise conv + LayerNorm (long chains)
    assert d.get("mlp_stream2_x4", 0) == 160 and d.get("mlp_stream2_x3", 0) == 12, d   # 37 latent row tiles (18 CTA pairs + 1) / 46 text row tiles (23 pairs)
    assert d.get("dp_convnext_fused", 0) == 4, d              # fp64 duration predictor, one kernel per block
    print(f"configs[1]: latent max-abs {err:.2e}, worst wav SNR {snr:.1f} dB, variants {d}")


@pytest.mark.parametrize("n,seed,tiles", [(16, 5, 16), (24, 9, None), (32, 4234, 38), (44, 2, 47), (64, 3, 65), (96, 4, None)])
def test_other_row_tile_counts_match_the_oracle(rig, n, seed, tiles):
    """The fused ConvNeXt MLP changes its hidden-slice count with the number of 128-row tiles (16 slices of 64 units at <= 9
    tiles ... 4 at 37, 3 at 38-49, 2 at 50-74, 1 from ~100; it used to take a second wave at 38): batches on both sides of every
    boundary, each against the oracle."""
    ids, mask, ttl, dp = _bench_batch(rig, n, seed)
    frames, err, snr = _packed_vs_oracle(rig, ids, mask, ttl, dp, 2, 100 + n)
    if tiles is not None:
        assert -(-frames // 128) == tiles, frames
    print(f"B={n}: {frames} frames = {-(-frames // 128)} row tiles, latent max-abs {err:.2e}, worst wav SNR {snr:.1f} dB")


def test_programmatic_dependent_launch_changes_no_bit(rig):
    """Every kernel goes out with programmatic stream serialization (launch_k, STC_PDL): a kernel may start while its predecessor
    still runs and waits for it itself (griddepcontrol.wait before its first global access). A missing or late wait would be a
    race between consecutive kernels, so: configs[1] three times with PDL and once on a handle without it — all four results
    bit-identical (latents, waveforms, durations)."""
    import os
    ids, mask, ttl, dp = _bench_batch(rig, 32, 1234)
    nz = np.random.default_rng(5).standard_normal((32, 144, 420)).astype(np.float32)
    runs = [rig["eng"].synthesize_packed(ids, mask, ttl, dp, 5, 1.05, noise=nz, want_latent=True) for _ in range(3)]
    os.environ["STC_PDL"] = "0"
    try:
        plain = rig["capi"].Engine(rig["root"] + "/onnx")
    finally:
        del os.environ["STC_PDL"]
    try:
        runs.append(plain.synthesize_packed(ids, mask, ttl, dp, 5, 1.05, noise=nz, want_latent=True))
    finally:
        plain.close()
    for r in runs[1:]:
        np.testing.assert_array_equal(r["duration"], runs[0]["duration"])
        for b in range(32):
            np.testing.assert_array_equal(r["latent"][b], runs[0]["latent"][b])
            np.testing.assert_array_equal(r["wavs"][b], runs[0]["wavs"][b])


def test_vocoder_at_two_sm_gemm_scale(rig):
    """stc_vocode on 16 x 200 latent frames = 19 200 vocoder rows: both pointwise projections of every block take the two-SM
    (cta_group::2) fp16 GEMM and the depthwise conv its shared-memory ring; also in the split-bf16 mode (>= 80 dB)."""
    import os
    rng = np.random.default_rng(21)
    lat = rng.standard_normal((16, 144, 200)).astype(np.float32)
    lat[3, :, 120:] = 0
    want = rig["ora"].voc(dict(latent=lat))
    v0 = rig["eng"].kernel_variants()
    got = rig["eng"].vocode(lat)
    d = _delta(v0, rig["eng"].kernel_variants())
    assert d.get("gemm2_f16", 0) >= 10 and d.get("gemm2_f16_astat", 0) >= 10, d
    assert d.get("dwconv_ln_chain", 0) >= 10, d
    snr = U.snr_db(got, want)
    assert snr >= SNR_EXPECTED, snr
    os.environ["STC_VOC"] = "bf16x3"
    try:
        exact = rig["capi"].Engine(rig["root"] + "/onnx")
    finally:
        del os.environ["STC_VOC"]
    try:
        got2 = exact.vocode(lat)
        assert exact.kernel_variants().get("gemm2_bf16x3", 0) >= 20
    finally:
        exact.close()
    snr2 = U.snr_db(got2, want)
    assert snr2 >= SNR_EXACT and snr2 > snr, (snr2, snr)


@pytest.mark.parametrize("steps", [2, 5, 10, 20])
def test_configs2_multilingual_step_sweep(rig, steps):
    """configs[2]: en/ko/es/pt/fr in one batch, one voice style per utterance, total_step 2/5/10/20, end to end through both
    `stc_synthesize` (the reference's padded rectangle) and `stc_synthesize_packed`."""
    from oracle import host_ref
    texts = ["This morning, I took a walk in the park, and the sound of the birds and the breeze was so pleasant.", U.KO, U.ES, U.PT, U.FR]
    langs = ["en", "ko", "es", "pt", "fr"]
    ids, mask = rig["eng"].text_to_ids(texts, langs)
    ids_ref, mask_ref = host_ref.unicode_processor_call(rig["ora"].indexer, texts, langs)
    np.testing.assert_array_equal(ids, ids_ref); np.testing.assert_array_equal(mask, mask_ref)
    ttl, dp = U.styles(rig["root"], ["M1", "F1", "M2", "F2", "M1"])
    nz = np.random.default_rng(steps).standard_normal((5, 144, 300)).astype(np.float32)
    tr = {}
    wav_ref, dur_ref = rig["ora"].infer_ids(ids, mask, ttl, dp, steps, np.float32(1.05), lambda b, d, L: nz[:, :, :L], tr)
    L = tr["latent_len"]
    out = rig["eng"].synthesize(ids, mask, ttl, dp, steps, 1.05, noise=nz, want_latent=True)
    assert out["L"] == L
    np.testing.assert_array_equal(out["duration"], dur_ref)
    np.testing.assert_array_equal(out["wav_lengths"], tr["wav_lengths"])
    err = np.abs(out["latent"] - tr["xs"][-1]).max()
    assert err <= LAT_TOL_NORTH_STAR and err <= LAT_TOL_EXPECTED, err
    snr = U.snr_db(out["wav"].reshape(-1), wav_ref)
    assert snr >= SNR_NORTH_STAR and snr >= SNR_EXPECTED, snr
    _packed_vs_oracle(rig, ids, mask, ttl, dp, steps, steps)


def test_device_resident_graph_keys_do_not_alias(rig):
    """Two device-resident calls whose four input pointers differ only by a permutation the old folded key could not see
    (ids / mask swapped by 512-byte-granular offsets) must each run on their own inputs."""
    import torch
    eng = rig["eng"]
    ids, mask, ttl, dp = _bench_batch(rig, 4, 77)
    B, T = ids.shape
    lens = mask.reshape(B, -1).sum(1).astype(np.int32)
    pool = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")

    def place(arr, off):
        t = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
        view = pool[off:off + t.numel() * t.element_size()].view(t.dtype).view(t.shape)
        view.copy_(t)
        return view
    cs = eng.cfg.chunk_size
    cap = int(lens.sum() * 0.12 * eng.cfg.sample_rate) + (B + 8) * cs
    wav = torch.empty(cap, dtype=torch.float32, device="cuda"); dur = torch.empty(B, dtype=torch.float32, device="cuda")
    res = []
    ids2 = ids.copy(); ids2[:, 5:9] = ids2[:, 9:13]                       # a second, different input set
    for k, (i_off, m_off, which) in enumerate([(0x0000, 0x8000, ids), (0x0400, 0x8200, ids2), (0x0000, 0x8000, ids)]):
        pool.zero_()
        di, dm = place(which, i_off), place(mask, m_off)
        dt, dd = place(ttl, 0x40000), place(dp, 0xC0000)
        off = eng.synthesize_packed_device(di.data_ptr(), dm.data_ptr(), dt.data_ptr(), dd.data_ptr(), B, T, 2, 1.05, 5, wav.data_ptr(), cap,
                                           dur.data_ptr(), text_lens=lens)
        torch.cuda.synchronize()
        res.append((off.copy(), dur.cpu().numpy().copy(), wav[:int(off[-1])].cpu().numpy().copy()))
    want = [eng.duration(ids, dp, mask) / np.float32(1.05), eng.duration(ids2, dp, mask) / np.float32(1.05)]
    np.testing.assert_array_equal(res[0][1], want[0])
    np.testing.assert_array_equal(res[1][1], want[1])
    np.testing.assert_array_equal(res[2][1], want[0])
    np.testing.assert_array_equal(res[0][2], res[2][2])
    assert not np.array_equal(res[0][1], res[1][1])


def test_style_shapes_are_checked_before_the_c_abi(rig):
    """A voice style with other dims than the graphs' (or a style batch that does not match the texts) is rejected with the
    reference's shape error instead of being read out of bounds."""
    capi, eng = rig["capi"], rig["eng"]
    ids, mask, ttl, dp = _bench_batch(rig, 3, 5)
    for bad_ttl, bad_dp in [(ttl[:2], dp), (ttl, dp[:1]), (ttl[:, :40], dp), (ttl, dp[:, :, :8]), (ttl[:, :, :128], dp)]:
        with pytest.raises(capi.StcError) as e:
            eng.synthesize_packed(ids, mask, bad_ttl, bad_dp, 2, 1.05)
        assert e.value.code == -1 and "invalid dimensions" in str(e.value)
    with pytest.raises(capi.StcError):
        eng.duration(ids, dp[:, :4], mask)
    with pytest.raises(capi.StcError):
        eng.text_encode(ids, ttl[:, :10], mask)


def test_device_pcm16_matches_the_reference_wav_goldens(rig, host_golden):
    """stc_out_opts.pcm16: the device quantiser against the bytes the UNMODIFIED reference writeWavFile produced
    (tests/golden/host_golden.json 'wav' cases, generated by oracle/make_golden.py from cpp/helper.cpp:943-990), byte for byte."""
    seen = 0
    for g in host_golden:
        if g["case"]["kind"] != "wav" or not g["case"]["samples"]:
            continue
        x = np.asarray(g["case"]["samples"], np.float32)
        want = np.frombuffer(bytes(g["bytes"][44:]), dtype="<i2")
        np.testing.assert_array_equal(rig["eng"].debug_pcm16(x), want)
        seen += 1
    assert seen >= 1
    # truncation toward zero, clamping, and 100k random samples against the host restatement
    x = np.concatenate([np.asarray([0.99999, -0.99999, 1.0, -1.0, 7.0, -7.0, 3.0517578e-05, -3.0517578e-05, 0.0], np.float32),
                        (np.random.default_rng(3).standard_normal(100000) * 0.7).astype(np.float32)])
    want = (np.clip(x, np.float32(-1), np.float32(1)) * np.float32(32767)).astype(np.int16)
    np.testing.assert_array_equal(rig["eng"].debug_pcm16(x), want)


def test_joined_long_form_output_on_the_device(rig):
    """configs[3] / SURVEY.md §8 f3: the chunks of one text as a packed batch, joined with silence and quantised on the device —
    equal to the host-side join (cpp/helper.cpp:706-714) and writeWavFile quantisation (cpp/helper.cpp:985-988) of the same floats."""
    from supertonic_b200 import tts as T
    eng = rig["eng"]
    text = " ".join(U.make_text(np.random.default_rng(i), 150) + "." for i in range(6))
    chunks = T.chunk_text(text, 300)
    n = len(chunks)
    assert n >= 3
    ids, mask = eng.text_to_ids(chunks, ["en"] * n)
    ttl, dp = U.styles(rig["root"], ["M1"] * n)
    gap = int(np.float32(0.3) * np.float32(44100))
    plain = eng.synthesize_joined(ids, mask, ttl, dp, 3, 1.05, seed=9)
    ref = eng.synthesize_packed(ids, mask, ttl, dp, 3, 1.05, seed=9)
    cs = eng.cfg.chunk_size
    for b in range(n):                                                     # no options: the packed layout itself
        np.testing.assert_array_equal(plain["out"][plain["offsets"][b]:plain["offsets"][b] + ref["wav_lengths"][b]], ref["wavs"][b])
    joined = eng.synthesize_joined(ids, mask, ttl, dp, 3, 1.05, seed=9, gap_samples=gap)
    want = np.concatenate([np.concatenate([plain["out"][plain["offsets"][b]:plain["offsets"][b + 1]], np.zeros(gap if b + 1 < n else 0, np.float32)])
                           for b in range(n)])
    np.testing.assert_array_equal(joined["out"], want)
    np.testing.assert_array_equal(joined["frames"], plain["frames"])
    assert joined["offsets"][n] == plain["offsets"][n] + (n - 1) * gap and (plain["frames"] * cs == np.diff(plain["offsets"])).all()
    pcm = eng.synthesize_joined(ids, mask, ttl, dp, 3, 1.05, seed=9, gap_samples=gap, pcm16=True, pinned="pcm_test")
    assert pcm["out"].dtype == np.int16
    np.testing.assert_array_equal(pcm["out"], (np.clip(want, np.float32(-1), np.float32(1)) * np.float32(32767)).astype(np.int16))
    # asynchronous form (pinned result, stc_wait) and graph replay deliver the same bytes
    a1 = eng.synthesize_joined(ids, mask, ttl, dp, 3, 1.05, seed=9, gap_samples=gap, pcm16=True, pinned="pcm_async", wait=False)
    eng.wait()
    np.testing.assert_array_equal(a1["out"], pcm["out"])
    # the Python mirror of the reference API
    tt = T.TextToSpeech(eng)
    r = tt.call_batched(text, "en", T.Style(ttl[:1], dp[:1]), 3, 1.05, 0.3, seed=9)
    np.testing.assert_array_equal(r.wav, want)
    assert abs(float(r.duration[0]) - (float(plain["duration"].sum()) + 0.3 * (n - 1))) < 1e-3
    assert T.wav_file_bytes_pcm16(pcm["out"], 44100) == T.wav_file_bytes(want, 44100)
