"""ctypes binding of libsupertonic_cuda.so (include/supertonic_cuda.h).

There is no fallback: if the shared library has not been built (``__graft_entry__.build()`` or
``make -C supertonic_b200/csrc``) importing this module raises, and without a CUDA device
``Engine(...)`` raises with the library's message.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsupertonic_cuda.so")
if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} is missing — build it first (python -c 'import __graft_entry__ as g; g.build()'); "
                      "supertonic_b200 has no CPU fallback for the neural path")
lib = C.CDLL(LIB_PATH)

STC_OK = 0
PREC_DEFAULT, PREC_BF16X3, PREC_FP32_SIMT = 0, 1, 2
ERR_CAPACITY = -5


class StcOutOpts(C.Structure):
    _fields_ = [("pcm16", C.c_int32), ("gap_samples", C.c_int64), ("noise_index", C.c_void_p)]


class StcConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "sample_rate", "base_chunk_size", "chunk_compress_factor", "latent_dim", "latent_channels", "chunk_size",
        "text_emb_channels", "style_ttl_tokens", "style_ttl_dim", "style_dp_tokens", "style_dp_dim", "vocab_size")]


_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_pf = C.POINTER(C.c_float)
_pi64 = C.POINTER(C.c_int64)
_SIGS = {
    "stc_create": (_i, [C.c_char_p, _i, _i, C.POINTER(_vp)]),
    "stc_destroy": (None, [_vp]),
    "stc_last_error": (C.c_char_p, [_vp]),
    "stc_get_config": (_i, [_vp, C.POINTER(StcConfig)]),
    "stc_validate_style": (_i, [_vp, _i, _vp, _vp]),
    "stc_duration": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "stc_text_encode": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "stc_vector_step": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "stc_vocode": (_i, [_vp, _vp, _i, _i, _vp]),
    "stc_synthesize": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _i64, C.c_uint64, _vp, _i64, _vp, _vp, _vp, _vp]),
    "stc_synthesize_device": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, C.c_uint64, _vp, _i64, _vp, _vp]),
    "stc_synthesize_packed": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _i64, C.c_uint64, _vp, _i64, _vp, _vp, _vp, _vp]),
    "stc_synthesize_packed_async": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, C.c_uint64, _vp, _i64, _vp, _vp, _vp]),
    "stc_synthesize_packed_ex": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _i64, C.c_uint64, _vp, _vp, _i64, _vp, _vp, _vp, _i]),
    "stc_debug_pcm16": (_i, [_vp, _vp, _i64, _vp]),
    "stc_derive_arch": (_i, [C.c_char_p, C.c_char_p, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "stc_wait": (_i, [_vp]),
    "stc_synthesize_packed_device": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, C.c_uint64, _vp, _i64, _vp, _vp]),
    "stc_pinned_alloc": (_i, [C.c_size_t, C.POINTER(_vp)]),
    "stc_pinned_free": (None, [_vp]),
    "stc_text_to_ids": (_i, [_vp, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), _i, _vp, _vp, _i64, _vp]),
    "stc_frontend_open": (_i, [C.c_char_p, C.POINTER(_vp)]),
    "stc_frontend_close": (None, [_vp]),
    "stc_frontend_text_to_ids": (_i, [_vp, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), _i, _vp, _vp, _i64, _vp]),
    "stc_chunk_text": (_i, [C.c_char_p, _i, _vp, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(_i)]),
    "stc_launch_count": (C.c_uint64, [_vp]),
    "stc_kernel_variants": (_i, [_vp, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "stc_set_graphs": (_i, [_vp, _i]),
    "stc_stream": (_vp, [_vp]),
    "stc_set_profile": (_i, [_vp, _i]),
    "stc_last_stage_ms": (_i, [_vp, _vp]),
    "stc_kernel_profile": (_i, [_vp, _i, _vp]),
    "stc_debug_mlp": (_i, [_vp, _i, _i, _pf, _pf, _pf]),
    "stc_debug_dwconv": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _pf, _pf, _pf]),
    "stc_debug_gemm": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _pf, _pf]),
}
for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)          # AttributeError here == header and library out of sync
    _fn.restype, _fn.argtypes = _res, _args

EXPORTED = sorted(_SIGS)


class StcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[stc {code}] {msg}")
        self.code = code


def _err(h) -> str:
    """The library's message for the last failure; bytes taken from a graph file (node / tensor names of a corrupted .onnx) may not be UTF-8."""
    m = lib.stc_last_error(h)
    return m.decode("utf-8", "replace") if m else ""


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_vp)


def _cf(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


def _b(s) -> bytes:
    return s if isinstance(s, bytes) else s.encode("utf-8", "surrogateescape")


def _texts_to_ids(fn, handle, texts: Sequence, langs: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
    n = len(texts)
    raw = [_b(t) for t in texts]
    ta = (C.c_char_p * n)(*raw)
    la = (C.c_char_p * n)(*[_b(l) for l in langs])
    # one pass with a guessed width (bytes + tags + slack; the rare expanding rules — '@' -> ' at ', 'e.g.,' -> 'for example, ' —
    # can exceed it: the library then reports the exact token count and the call is repeated once)
    cap = max((len(r) for r in raw), default=0) + 40
    T = C.c_int64(0)
    for _ in range(2):
        ids = np.empty((n, cap), np.int64)
        mask = np.empty((n, 1, cap), np.float32)
        rc = fn(handle, ta, la, n, _ptr(ids), _ptr(mask), cap, C.byref(T))
        if rc != STC_OK and T.value > cap:
            cap = T.value
            continue
        break
    if rc != STC_OK:
        raise StcError(rc, _err(None))
    return np.ascontiguousarray(ids[:, :T.value]), np.ascontiguousarray(mask[:, :, :T.value])


def chunk_text(text, max_len: int) -> List[bytes]:
    """chunkText of the C++ reference (cpp/helper.cpp:1117-1186) → list of UTF-8 byte strings."""
    raw = _b(text)
    if b"\0" in raw:
        raise ValueError("text contains NUL")
    need, n = C.c_size_t(0), C.c_int(0)
    cap = len(raw) + 64
    while True:
        buf = C.create_string_buffer(cap)
        rc = lib.stc_chunk_text(raw, max_len, buf, cap, C.byref(need), C.byref(n))
        if rc == STC_OK:
            parts = buf.raw[:need.value].split(b"\0")[:-1]
            return parts[:n.value] if n.value else [b""]
        if rc != ERR_CAPACITY:
            raise StcError(rc, _err(None))
        cap = need.value + 16


def derive_arch(onnx_path: str, kind: str) -> dict:
    """The layer plan the library derives from a graph's NODES (csrc/graph_plan.h) — host only, no GPU."""
    import json
    need = C.c_size_t(0)
    lib.stc_derive_arch(_b(onnx_path), _b(kind), None, 0, C.byref(need))
    if not need.value:
        raise StcError(-4, _err(None))
    buf = C.create_string_buffer(need.value + 16)
    rc = lib.stc_derive_arch(_b(onnx_path), _b(kind), buf, need.value + 16, C.byref(need))
    if rc != STC_OK:
        raise StcError(rc, _err(None))
    return json.loads(buf.value.decode())


class Frontend:
    """Host-only text front-end (no GPU needed): UnicodeProcessor::call of the C++ reference."""

    def __init__(self, unicode_indexer_json: str):
        self._h = _vp()
        rc = lib.stc_frontend_open(_b(unicode_indexer_json), C.byref(self._h))
        if rc != STC_OK:
            raise StcError(rc, _err(None))

    def __call__(self, texts, langs):
        return _texts_to_ids(lib.stc_frontend_text_to_ids, self._h, texts, langs)

    text_to_ids = __call__

    def close(self):
        if getattr(self, "_h", None):
            lib.stc_frontend_close(self._h)
            self._h = None

    __del__ = close


class Engine:
    """One handle == one GPU replica of the four graphs (reference: loadOnnxAll, cpp/helper.cpp:784-795)."""

    def __init__(self, onnx_dir: str, device: int = 0, precision: int = PREC_DEFAULT):
        self._h = _vp()
        rc = lib.stc_create(_b(onnx_dir), device, precision, C.byref(self._h))
        if rc != STC_OK:
            self._h = None
            raise StcError(rc, _err(None))
        self.cfg = StcConfig()
        lib.stc_get_config(self._h, C.byref(self.cfg))
        self.onnx_dir = onnx_dir

    def close(self):
        if getattr(self, "_h", None):
            for arr, ptr in getattr(self, "_pinned", {}).values():
                lib.stc_pinned_free(ptr)
            self._pinned = {}
            lib.stc_destroy(self._h)
            self._h = None

    __del__ = close

    def pinned(self, name: str, count: int, dtype) -> np.ndarray:
        """A page-locked host array of at least `count` elements, owned by the engine and reused between calls
        (contents are overwritten by the next call that asks for the same `name`)."""
        if not hasattr(self, "_pinned"):
            self._pinned = {}
        dt = np.dtype(dtype)
        cur = self._pinned.get(name)
        if cur is None or cur[0].size < count or cur[0].dtype != dt:
            if cur is not None:
                if getattr(self, "_pending", None):
                    self.wait()                            # an outstanding copy may still target the buffer being replaced
                lib.stc_pinned_free(cur[1])
            n = int(count * 1.25) + 1024
            ptr = _vp()
            rc = lib.stc_pinned_alloc(n * dt.itemsize, C.byref(ptr))
            if rc != STC_OK:
                raise StcError(rc, _err(None))
            buf = (C.c_char * (n * dt.itemsize)).from_address(ptr.value)
            cur = (np.frombuffer(buf, dtype=dt, count=n), ptr)
            self._pinned[name] = cur
        return cur[0][:count]

    def _chk(self, rc):
        if rc != STC_OK:
            raise StcError(rc, _err(self._h))

    def _check(self, B, T=None, ttl=None, dp=None, mask=None, ids=None):
        """Shapes against the engine's geometry before bare pointers cross the C ABI (ORT raises at the same boundary)."""
        def shp(a):
            return None if a is None else (C.c_int64 * 3)(*(tuple(a.shape) + (-1, -1, -1))[:3])
        if (ttl is not None and ttl.ndim != 3) or (dp is not None and dp.ndim != 3):
            raise StcError(-1, "Got invalid dimensions for input: style tensors must have rank 3")
        if ttl is not None or dp is not None:
            self._chk(lib.stc_validate_style(self._h, B, shp(ttl), shp(dp)))
        if ids is not None and ids.shape != (B, T):
            raise StcError(-1, f"Got invalid dimensions for input: text_ids {ids.shape}, expected {(B, T)}")
        if mask is not None and mask.size != B * T:
            raise StcError(-1, f"Got invalid dimensions for input: text_mask {mask.shape}, expected {(B, 1, T)}")

    # ---- parity layer (1:1 with the reference's four Session::Run calls)
    def duration(self, text_ids, style_dp, text_mask) -> np.ndarray:
        ids, sty, m = _cf(text_ids, np.int64), _cf(style_dp, np.float32), _cf(text_mask, np.float32)
        B, T = ids.shape
        self._check(B, T, dp=sty, mask=m)
        out = np.empty((B,), np.float32)
        self._chk(lib.stc_duration(self._h, _ptr(ids), _ptr(sty), _ptr(m), B, T, _ptr(out)))
        return out

    def text_encode(self, text_ids, style_ttl, text_mask) -> np.ndarray:
        ids, sty, m = _cf(text_ids, np.int64), _cf(style_ttl, np.float32), _cf(text_mask, np.float32)
        B, T = ids.shape
        self._check(B, T, ttl=sty, mask=m)
        out = np.empty((B, self.cfg.text_emb_channels, T), np.float32)
        shp = np.zeros(3, np.int64)
        self._chk(lib.stc_text_encode(self._h, _ptr(ids), _ptr(sty), _ptr(m), B, T, _ptr(out), _ptr(shp)))
        assert tuple(shp) == out.shape
        return out

    def vector_step(self, noisy_latent, text_emb, style_ttl, text_mask, latent_mask, total_step, current_step) -> np.ndarray:
        x, te, sty = _cf(noisy_latent, np.float32), _cf(text_emb, np.float32), _cf(style_ttl, np.float32)
        tm, lm = _cf(text_mask, np.float32), _cf(latent_mask, np.float32)
        tot, cur = _cf(total_step, np.float32), _cf(current_step, np.float32)
        B, D, L = x.shape
        T = te.shape[2]
        self._check(B, T, ttl=sty, mask=tm)
        if D != self.cfg.latent_channels or te.shape[:2] != (B, self.cfg.text_emb_channels) or lm.size != B * L or tot.size != B or cur.size != B:
            raise StcError(-1, f"Got invalid dimensions for input: noisy_latent {x.shape}, text_emb {te.shape}, latent_mask {lm.shape}, "
                               f"total_step {tot.shape}, current_step {cur.shape}")
        out = np.empty_like(x)
        self._chk(lib.stc_vector_step(self._h, _ptr(x), _ptr(te), _ptr(sty), _ptr(tm), _ptr(lm), _ptr(tot), _ptr(cur),
                                      B, L, T, _ptr(out)))
        return out

    def vocode(self, latent) -> np.ndarray:
        x = _cf(latent, np.float32)
        B, D, L = x.shape
        if D != self.cfg.latent_channels:
            raise StcError(-1, f"Got invalid dimensions for input: latent {x.shape}, expected [B,{self.cfg.latent_channels},L]")
        out = np.empty((B, L * self.cfg.chunk_size), np.float32)
        self._chk(lib.stc_vocode(self._h, _ptr(x), B, L, _ptr(out)))
        return out

    # ---- fast layer
    def synthesize(self, text_ids, text_mask, style_ttl, style_dp, total_step: int, speed: float = 1.05,
                   noise: Optional[np.ndarray] = None, seed: int = 0, want_latent: bool = False, wav_cap: Optional[int] = None):
        """Whole `_infer` body (cpp/helper.cpp:488-682). Returns dict(wav[B, L*cs], duration[B], wav_lengths[B], L, latent?)."""
        ids, m = _cf(text_ids, np.int64), _cf(text_mask, np.float32)
        sttl, sdp = _cf(style_ttl, np.float32), _cf(style_dp, np.float32)
        B, T = ids.shape
        self._check(B, T, ttl=sttl, dp=sdp, mask=m)
        cs = self.cfg.chunk_size
        nz, nld = None, 0
        if noise is not None:
            nz = _cf(noise, np.float32)
            if nz.ndim != 3 or nz.shape[:2] != (B, self.cfg.latent_channels):
                raise StcError(-1, f"Got invalid dimensions for input: noise {nz.shape}, expected [{B},{self.cfg.latent_channels},>=L]")
            nld = nz.shape[2]
        # duration is data dependent; start from a per-token guess and retry once with the exact size on overflow
        cap = wav_cap or max(int(T * 0.12 * self.cfg.sample_rate / cs) + 8, 16) * cs
        dur = np.empty((B,), np.float32)
        wl = np.empty((B,), np.int64)
        L = C.c_int64(0)
        for _ in range(2):
            wav = np.empty((B, cap), np.float32)
            lat = np.empty((B, self.cfg.latent_channels, cap // cs), np.float32) if want_latent else None
            rc = lib.stc_synthesize(self._h, _ptr(ids), _ptr(m), _ptr(sttl), _ptr(sdp), B, T, int(total_step), float(speed),
                                    _ptr(nz), nld, seed, _ptr(wav), cap, _ptr(dur), _ptr(wl), C.byref(L), _ptr(lat))
            if rc == ERR_CAPACITY and L.value * cs > cap:
                cap = L.value * cs
                continue
            self._chk(rc)
            break
        Lv = L.value
        res = dict(wav=wav[:, :Lv * cs], duration=dur, wav_lengths=wl, L=Lv)
        if want_latent:
            res["latent"] = lat.reshape(-1)[:B * self.cfg.latent_channels * Lv].reshape(B, self.cfg.latent_channels, Lv)
        return res

    def synthesize_packed(self, text_ids, text_mask, style_ttl, style_dp, total_step: int, speed: float = 1.05,
                          noise: Optional[np.ndarray] = None, seed: int = 0, want_latent: bool = False, pinned=False,
                          wait: bool = True):
        """Throughput path: packed latent rows, no padded frames. Returns dict(wavs=[B trimmed arrays], duration[B],
        wav_lengths[B], frames[B], latent? (list of [frames_b, D] arrays)).
        pinned=True (or a buffer name): the waveforms are VIEWS into an engine-owned page-locked buffer (D2H at PCIe speed, no
        copy); they stay valid until the next call that uses the same buffer name on this engine.
        wait=False (needs a pinned buffer, no injected noise): returns once the sizes are known and the work is enqueued; the
        waveform views hold data only after `Engine.wait()`. The next call may be issued before that (its compute overlaps this
        call's copy) as long as it uses a different pinned buffer name."""
        ids, m = _cf(text_ids, np.int64), _cf(text_mask, np.float32)
        sttl, sdp = _cf(style_ttl, np.float32), _cf(style_dp, np.float32)
        B, T = ids.shape
        self._check(B, T, ttl=sttl, dp=sdp, mask=m)
        cs, D = self.cfg.chunk_size, self.cfg.latent_channels
        nz, nld = None, 0
        if noise is not None:
            nz = _cf(noise, np.float32)
            if nz.ndim != 3 or nz.shape[:2] != (B, D):
                raise StcError(-1, f"Got invalid dimensions for input: noise {nz.shape}, expected [{B},{D},>=frames]")
            nld = nz.shape[2]
        cap = int(m.sum() * 0.12 * self.cfg.sample_rate) + (B + 8) * cs
        dur = np.empty((B,), np.float32); wl = np.empty((B,), np.int64); off = np.zeros((B + 1,), np.int64)
        use_async = (not wait) and bool(pinned) and noise is None and not want_latent
        pname = pinned if isinstance(pinned, str) else "wav_packed"
        # (re-using a buffer whose earlier copy is still in flight is ordered by the library — copies run in issue order on one
        #  stream — the caller just must have consumed the older result by then)
        for _ in range(2):
            wav = self.pinned(pname, cap, np.float32) if pinned else np.empty((cap,), np.float32)
            lat = np.empty((cap // cs, D), np.float32) if want_latent else None
            if use_async:
                rc = lib.stc_synthesize_packed_async(self._h, _ptr(ids), _ptr(m), _ptr(sttl), _ptr(sdp), B, T, int(total_step), float(speed),
                                                     seed, _ptr(wav), cap, _ptr(off), _ptr(dur), _ptr(wl))
                if rc == ERR_CAPACITY and off[B] > cap:
                    cap = int(off[B])
                    continue
                self._chk(rc)
                self._pending = getattr(self, "_pending", set()) | {pname}
                break
            rc = lib.stc_synthesize_packed(self._h, _ptr(ids), _ptr(m), _ptr(sttl), _ptr(sdp), B, T, int(total_step), float(speed),
                                           _ptr(nz), nld, seed, _ptr(wav), cap, _ptr(off), _ptr(dur), _ptr(wl), _ptr(lat))
            if rc == ERR_CAPACITY and off[B] > cap:
                cap = int(off[B])
                continue
            self._chk(rc)
            break
        frames = ((off[1:] - off[:-1]) // cs).astype(np.int64)
        res = dict(wavs=[wav[off[b]:off[b] + wl[b]] for b in range(B)], duration=dur, wav_lengths=wl, frames=frames)
        if want_latent:
            fo = np.concatenate([[0], np.cumsum(frames)])
            res["latent"] = [lat[fo[b]:fo[b + 1]] for b in range(B)]
        return res

    def synthesize_joined(self, text_ids, text_mask, style_ttl, style_dp, total_step: int, speed: float = 1.05,
                          noise: Optional[np.ndarray] = None, seed: int = 0, pcm16: bool = False, gap_samples: int = 0,
                          pinned=False, wait: bool = True, noise_index=None):
        """stc_synthesize_packed_ex: the batch as ONE output array — utterance b's untrimmed frames_b*cs samples at offsets[b], with
        `gap_samples` zeros after every utterance but the last (TextToSpeech::call's silence join, cpp/helper.cpp:706-714), as
        float32 or, with pcm16=True, int16 quantised on the device like writeWavFile (cpp/helper.cpp:985-988).
        noise_index (int64[B], optional): utterance b draws the device noise stream (seed, noise_index[b]) instead of (seed, b) — a
        request split over several calls / GPUs then gets the noise one call would give it.
        Returns dict(out=array[total], offsets[B+1], duration[B], wav_lengths[B], frames[B])."""
        ids, m = _cf(text_ids, np.int64), _cf(text_mask, np.float32)
        sttl, sdp = _cf(style_ttl, np.float32), _cf(style_dp, np.float32)
        B, T = ids.shape
        self._check(B, T, ttl=sttl, dp=sdp, mask=m)
        cs = self.cfg.chunk_size
        nz, nld = None, 0
        if noise is not None:
            nz = _cf(noise, np.float32)
            if nz.ndim != 3 or nz.shape[:2] != (B, self.cfg.latent_channels):
                raise StcError(-1, f"Got invalid dimensions for input: noise {nz.shape}")
            nld = nz.shape[2]
        dt = np.int16 if pcm16 else np.float32
        ni = None if noise_index is None else _cf(noise_index, np.int64)
        if ni is not None and ni.shape != (B,):
            raise StcError(-1, f"noise_index must have shape ({B},)")
        opts = StcOutOpts(int(bool(pcm16)), int(gap_samples), None if ni is None else ni.ctypes.data)
        cap = int(m.sum() * 0.12 * self.cfg.sample_rate) + (B + 8) * cs + (B - 1) * int(gap_samples)
        dur = np.empty((B,), np.float32); wl = np.empty((B,), np.int64); off = np.zeros((B + 1,), np.int64)
        use_async = (not wait) and bool(pinned) and noise is None
        pname = pinned if isinstance(pinned, str) else "wav_joined"
        for _ in range(2):
            out = self.pinned(pname, cap, dt) if pinned else np.empty((cap,), dt)
            rc = lib.stc_synthesize_packed_ex(self._h, _ptr(ids), _ptr(m), _ptr(sttl), _ptr(sdp), B, T, int(total_step), float(speed),
                                              _ptr(nz), nld, seed, C.byref(opts), _ptr(out), cap, _ptr(off), _ptr(dur), _ptr(wl),
                                              int(use_async))
            if rc == ERR_CAPACITY and off[B] > cap:
                cap = int(off[B])
                continue
            self._chk(rc)
            if use_async:
                self._pending = getattr(self, "_pending", set()) | {pname}
            break
        gaps = np.concatenate([np.full(B - 1, int(gap_samples), np.int64), [0]])
        frames = ((off[1:] - off[:-1] - gaps) // cs).astype(np.int64)
        return dict(out=out[:int(off[B])], offsets=off, duration=dur, wav_lengths=wl, frames=frames)

    def debug_pcm16(self, samples) -> np.ndarray:
        x = _cf(samples, np.float32).reshape(-1)
        out = np.empty(x.shape, np.int16)
        self._chk(lib.stc_debug_pcm16(self._h, _ptr(x), x.size, _ptr(out)))
        return out

    def wait(self):
        """Deliver every outstanding wait=False call (their waveform views hold data afterwards)."""
        self._pending = set()
        self._chk(lib.stc_wait(self._h))

    def synthesize_packed_device(self, ids_ptr: int, mask_ptr: int, sttl_ptr: int, sdp_ptr: int, B: int, T: int, total_step: int,
                                 speed: float, seed: int, wav_ptr: int, wav_cap: int, dur_ptr: int,
                                 text_lens: Optional[np.ndarray] = None) -> np.ndarray:
        """Device-resident packed variant; returns wav offsets [B+1] (floats). StcError(code=-5).need holds the size to retry with.
        text_lens: host int32[B] token counts (packs the text side too)."""
        off = np.zeros((B + 1,), np.int64)
        tl = None if text_lens is None else _cf(text_lens, np.int32)
        rc = lib.stc_synthesize_packed_device(self._h, ids_ptr, mask_ptr, sttl_ptr, sdp_ptr, _ptr(tl), B, T, int(total_step), float(speed),
                                              seed, wav_ptr, wav_cap, _ptr(off), dur_ptr)
        if rc != STC_OK:
            e = StcError(rc, _err(self._h))
            e.need = int(off[B])
            raise e
        return off

    def synthesize_device(self, ids_ptr: int, mask_ptr: int, sttl_ptr: int, sdp_ptr: int, B: int, T: int, total_step: int,
                          speed: float, seed: int, wav_ptr: int, wav_ld: int, dur_ptr: int) -> int:
        """Device-resident variant (raw device pointers, e.g. torch `.data_ptr()`); returns L. Rows of wav are dense
        [B][L*cs] from wav_ptr. Raises StcError(code=-5) when wav_ld < L*cs (retry with a larger buffer)."""
        L = C.c_int64(0)
        rc = lib.stc_synthesize_device(self._h, ids_ptr, mask_ptr, sttl_ptr, sdp_ptr, B, T, int(total_step), float(speed),
                                       seed, wav_ptr, wav_ld, dur_ptr, C.byref(L))
        if rc != STC_OK:
            e = StcError(rc, _err(self._h))
            e.L = L.value
            raise e
        return L.value

    def text_to_ids(self, texts, langs):
        return _texts_to_ids(lib.stc_text_to_ids, self._h, texts, langs)

    # ---- introspection
    @property
    def launches(self) -> int:
        return int(lib.stc_launch_count(self._h))

    def kernel_variants(self) -> dict:
        """{variant name: launches issued or captured since creation} of the size-dependent kernel dispatchers."""
        need = C.c_size_t(0)
        lib.stc_kernel_variants(self._h, None, 0, C.byref(need))
        buf = C.create_string_buffer(need.value + 64)
        self._chk(lib.stc_kernel_variants(self._h, buf, need.value + 64, C.byref(need)))
        return {k: int(v) for k, v in (ln.split("=") for ln in buf.value.decode().splitlines() if ln)}

    def set_graphs(self, on: bool):
        lib.stc_set_graphs(self._h, int(on))

    def set_profile(self, level: int):
        lib.stc_set_profile(self._h, int(level))

    def kernel_profile(self):
        """Per-kernel-class totals of the last synthesize call (profile level 2)."""
        res = {}
        for cls, name in enumerate(("gemm_tc", "dwconv_ln", "attention", "fused_mlp", "gemm_f16", "dwconv_ln_hbm")):
            out = np.zeros(4, np.float64)
            lib.stc_kernel_profile(self._h, cls, _ptr(out))
            res[name] = dict(ms=out[0], flops=out[1], bytes=out[2], launches=int(out[3]))
        return res

    def debug_gemm(self, M, N, K, bn=0, cm=1, cn=1, epilogue=0, iters=20):
        """One tcgen05 GEMM with a forced launch configuration -> (microseconds per launch, max-abs error vs fp32 CUDA cores)."""
        ms, err = C.c_float(0), C.c_float(0)
        self._chk(lib.stc_debug_gemm(self._h, M, N, K, bn, cm, cn, epilogue, iters, C.byref(ms), C.byref(err)))
        return ms.value * 1000.0, err.value

    def debug_mlp(self, M, iters=20):
        """Fused ConvNeXt MLP vs two GEMMs on M rows -> (us fused, us unfused, max-abs difference)."""
        a, b, e = C.c_float(0), C.c_float(0), C.c_float(0)
        self._chk(lib.stc_debug_mlp(self._h, M, iters, C.byref(a), C.byref(b), C.byref(e)))
        return a.value * 1000.0, b.value * 1000.0, e.value

    def debug_dwconv(self, rows, C_, K, dil, causal=False, B=8, rt=0, iters=20):
        """Depthwise conv + LayerNorm, sliding-window vs tiled kernel -> (us slide, us tile, max-abs difference)."""
        a, b, e = C.c_float(0), C.c_float(0), C.c_float(0)
        self._chk(lib.stc_debug_dwconv(self._h, rows, C_, K, dil, int(causal), B, rt, iters, C.byref(a), C.byref(b), C.byref(e)))
        return a.value * 1000.0, b.value * 1000.0, e.value

    def stage_ms(self):
        out = np.zeros(5, np.float32)
        lib.stc_last_stage_ms(self._h, _ptr(out))
        return dict(zip(("dp", "te", "ve", "vocoder", "whole"), out.tolist()))

    @property
    def stream(self) -> int:
        return int(lib.stc_stream(self._h) or 0)
