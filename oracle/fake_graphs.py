"""ORACLE — test infrastructure, never the product path.

The closed-form stand-ins for the four graphs in numpy: the same formulas as oracle/ref_stub_fake/onnxruntime_cxx_api.h (the fake ONNX
Runtime the unmodified C++ reference orchestration is compiled against, oracle/_ref/ref_pipe). Every constant is a power of two, so
float32 arithmetic is exact on both sides. Used by tests/test_oracle_pipeline.py (to drive host_ref.ReferenceTTS) and by
oracle/make_golden.py (to drive the unmodified PYTHON reference, py/helper.py, through a fake `onnxruntime` module).

    duration[b]       = tokens_b / 16 + (text_ids[b][0] % 5) / 32
    text_emb[b][c][t] = text_mask[b][0][t] * (c + 1) / 32                       (4 channels)
    denoised[b][d][t] = latent_mask[b][0][t] * (1/2 + current_step[b] / 16 + total_step[b] / 128 + (d % 16) / 1024)
    wav_tts[b][i]     = latent[b][i % D][i // chunk] / 2 + ((i % 97) - 48) / 128
"""
import numpy as np

f = np.float32


def fake_runs(trace: list, chunk: int):
    """-> (dp, te, ve, voc): callables feed-dict -> ndarray that also append one record per call to `trace`."""
    def note(graph, feed, **extra):
        trace.append(dict(graph=graph, inputs=list(feed), shapes=[list(np.asarray(v).shape) for v in feed.values()], **extra))

    def dp(feed):
        note("duration_predictor.onnx", feed)
        ids, mask = np.asarray(feed["text_ids"]), np.asarray(feed["text_mask"], f)
        return (mask[:, 0, :].sum(1).astype(f) * f(0.0625) + (ids[:, 0] % 5).astype(f) * f(0.03125)).astype(f)

    def te(feed):
        note("text_encoder.onnx", feed)
        return (np.asarray(feed["text_mask"], f) * (np.arange(1, 5, dtype=f) * f(0.03125))[None, :, None]).astype(f)

    def ve(feed):
        x, lm = np.asarray(feed["noisy_latent"], f), np.asarray(feed["latent_mask"], f)
        prev = trace[-1].get("_out") if trace and trace[-1]["graph"] == "vector_estimator.onnx" else None
        cur, tot = np.asarray(feed["current_step"], f), np.asarray(feed["total_step"], f)
        D = x.shape[1]
        # every term is a multiple of 2^-10 below 4: exact in float32 in any association
        out = (lm * (f(0.5) + cur * f(0.0625) + tot * f(0.0078125))[:, None, None]
               + lm * ((np.arange(D) % 16).astype(f) * f(0.0009765625))[None, :, None]).astype(f)
        note("vector_estimator.onnx", feed, total_step=[float(v) for v in tot], current_step=[float(v) for v in cur],
             masked_zero=bool(np.all(x * (1 - lm) == 0)),
             is_prev_output=bool(cur[0] == 0 or (prev is not None and np.array_equal(prev, x))), _out=out)
        return out

    def voc(feed):
        note("vocoder.onnx", feed)
        x = np.asarray(feed["latent"], f)
        B, D, L = x.shape
        i = np.arange(L * chunk)
        return (x[:, i % D, i // chunk] * f(0.5) + (((i % 97) - 48).astype(f) * f(0.0078125))[None]).astype(f)
    return dp, te, ve, voc


def fake_onnxruntime_module(trace: list, chunk: int):
    """A module object that can stand in for `onnxruntime` under py/helper.py: InferenceSession(path, sess_options=, providers=).run(None, feed)."""
    import os
    import types
    dp, te, ve, voc = fake_runs(trace, chunk)
    by_file = {"duration_predictor.onnx": dp, "text_encoder.onnx": te, "vector_estimator.onnx": ve, "vocoder.onnx": voc}

    class SessionOptions:
        pass

    class InferenceSession:
        def __init__(self, path, sess_options=None, providers=None):
            self._fn = by_file[os.path.basename(path)]

        def run(self, output_names, feed):
            return [self._fn(feed)]
    m = types.ModuleType("onnxruntime")
    m.SessionOptions, m.InferenceSession = SessionOptions, InferenceSession
    return m
