"""ORACLE — test infrastructure, never the product path.

Line-for-line CPU restatement (numpy float32 / Python ints / bytes) of the reference's HOST
arithmetic around the four Session::Run calls. C++ behaviour is canonical for the drop-in
(SURVEY.md App. E), so strings are handled as UTF-8 *bytes* exactly like cpp/helper.cpp does.

Pinned against tests/golden/host_golden.json, which is produced by running the unmodified
reference cpp/helper.cpp (oracle/make_golden.py) — see tests/test_host_oracle.py — and, for the
orchestration (ReferenceTTS: _infer / call / batch), against tests/golden/pipeline_golden.json:
what the unmodified TextToSpeech::call / batch did, Run by Run, over the closed-form stand-ins of
oracle/ref_stub_fake/onnxruntime_cxx_api.h (tests/test_oracle_pipeline.py). What stays UNPINNED is
the arithmetic inside the four Session::Run calls (oracle/onnx_interp.py stands in for ONNX Runtime).
"""
from __future__ import annotations

import re
import struct
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

AVAILABLE_LANGS = ["en", "ko", "es", "pt", "fr"]           # cpp/helper.cpp:15
_WS = b" \t\n\v\f\r"                                        # std::isspace, "C" locale


def _b(s) -> bytes:
    return s if isinstance(s, bytes) else s.encode("utf-8", "surrogateescape")


def trim(b: bytes) -> bytes:                                # cpp/helper.cpp:30-42
    return b.strip(_WS)


def _replace_all(s: bytes, frm: bytes, to: bytes) -> bytes:
    """find/replace loop of cpp/helper.cpp:89-95 (scan resumes after the inserted text)."""
    pos = 0
    while True:
        pos = s.find(frm, pos)
        if pos < 0:
            return s
        s = s[:pos] + to + s[pos + len(frm):]
        pos += len(to)


_REPLACEMENTS = [(_b(a), _b(c)) for a, c in [                # cpp/helper.cpp:69-87
    ("–", "-"), ("‑", "-"), ("—", "-"), ("_", " "), ("“", '"'), ("”", '"'), ("‘", "'"),
    ("’", "'"), ("´", "'"), ("`", "'"), ("[", " "), ("]", " "), ("|", " "), ("/", " "), ("#", " "),
    ("→", " "), ("←", " ")]]
_SPECIAL = [_b(x) for x in ["♥", "☆", "♡", "©", "\\"]]      # cpp/helper.cpp:105
_EXPR = [(b"@", b" at "), (b"e.g.,", b"for example, "), (b"i.e.,", b"that is, ")]   # :114-118
_EMOJI = re.compile(rb"\xF0\x9F[\x80-\xBF][\x80-\xBF]")     # :99-101
_PUNCT_FIX = [(re.compile(rb" ,"), b","), (re.compile(rb" \."), b"."), (re.compile(rb" !"), b"!"),
              (re.compile(rb" \?"), b"?"), (re.compile(rb" ;"), b";"), (re.compile(rb" :"), b":"),
              (re.compile(rb" '"), b"'")]                   # :129-135
_SPACES = re.compile(rb"[ \t\n\v\f\r]+")                    # :152
_END3 = [_b(x) for x in ["…", "。", "」", "』", "】", "〉", "》", "›", "»", "“", "”", "‘", "’"]]


def preprocess_text(text, lang: str) -> bytes:               # cpp/helper.cpp:52-200
    s = _b(text)
    for a, c in _REPLACEMENTS:
        s = _replace_all(s, a, c)
    s = _EMOJI.sub(b"", s)
    for sym in _SPECIAL:
        s = s.replace(sym, b"")
    for a, c in _EXPR:
        s = _replace_all(s, a, c)
    for rx, to in _PUNCT_FIX:
        s = rx.sub(to, s)
    for dup, one in ((b'""', b'"'), (b"''", b"'"), (b"``", b"`")):
        while dup in s:
            i = s.find(dup)
            s = s[:i] + one + s[i + 2:]
    s = trim(_SPACES.sub(b" ", s))
    if s:
        ends = s[-1:] in (b".", b"!", b"?", b";", b":", b",", b"'", b'"', b")", b"]", b"}", b">")
        if not ends and len(s) >= 3 and s[-3:] in _END3:     # NB: 2-byte » and › never match 3 bytes unless aligned
            ends = True
        if not ends:
            s += b"."
    if lang not in AVAILABLE_LANGS:
        raise RuntimeError(f"Invalid language: {lang}. Available: en, ko, es, pt, fr")
    return b"<" + lang.encode() + b">" + s + b"</" + lang.encode() + b">"


_LATIN = {}
for _base, _mark, _chars in [                                 # cpp/helper.cpp:214-269
    (0x0301, None, "ÁÉÍÓÚáéíóú"), (0x0300, None, "ÀÈÌÒÙàèìòù"), (0x0302, None, "ÂÊÎÔÛâêîôû"),
    (0x0303, None, "ÃÑÕãñõ"), (0x0308, None, "ÄËÏÖÜäëïöü"), (0x0327, None, "Çç")]:
    import unicodedata as _ud
    for _ch in _chars:
        _LATIN[ord(_ch)] = [ord(_ud.normalize("NFD", _ch)[0]), _base]


def decompose(cp: int, out: List[int]) -> None:               # cpp/helper.cpp:272-300
    if 0xAC00 <= cp < 0xAC00 + 11172:
        s = cp - 0xAC00
        out.append(0x1100 + s // 588)
        out.append(0x1161 + (s % 588) // 28)
        if s % 28 > 0:
            out.append(0x11A7 + s % 28)
        return
    if cp in _LATIN:
        out.extend(_LATIN[cp])
        return
    out.append(cp & 0xFFFF)


def text_to_unicode_values(t: bytes) -> List[int]:            # cpp/helper.cpp:302-347
    vals: List[int] = []
    i, n = 0, len(t)
    while i < n:
        c = t[i]
        if c & 0x80 == 0:
            cp, i = c, i + 1
        elif c & 0xE0 == 0xC0 and i + 1 < n:
            cp, i = ((c & 0x1F) << 6) | (t[i + 1] & 0x3F), i + 2
        elif c & 0xF0 == 0xE0 and i + 2 < n:
            cp, i = ((c & 0x0F) << 12) | ((t[i + 1] & 0x3F) << 6) | (t[i + 2] & 0x3F), i + 3
        elif c & 0xF8 == 0xF0 and i + 3 < n:
            cp = ((c & 0x07) << 18) | ((t[i + 1] & 0x3F) << 12) | ((t[i + 2] & 0x3F) << 6) | (t[i + 3] & 0x3F)
            i += 4
        else:
            i += 1
            continue
        decompose(cp, vals)
    return vals


def length_to_mask(lengths: Sequence[int], max_len: int = -1) -> np.ndarray:   # cpp/helper.cpp:740-757
    if max_len == -1:
        max_len = int(max(lengths))
    ar = np.arange(max_len)[None, None, :]
    return (ar < np.asarray(lengths, np.int64)[:, None, None]).astype(np.float32)


def get_latent_mask(wav_lengths: Sequence[int], base_chunk_size: int, ccf: int) -> np.ndarray:  # :759-770
    ls = base_chunk_size * ccf
    return length_to_mask([(int(w) + ls - 1) // ls for w in wav_lengths])


def unicode_processor_call(indexer: Sequence[int], texts, langs) -> Tuple[np.ndarray, np.ndarray]:
    """cpp/helper.cpp:355-390 → text_ids[B,T] int64 (pad 0; out-of-table stays 0), text_mask[B,1,T]."""
    vals = [text_to_unicode_values(preprocess_text(t, l)) for t, l in zip(texts, langs)]
    lens = [len(v) for v in vals]
    T = max(lens)
    ids = np.zeros((len(vals), T), np.int64)
    n = len(indexer)
    for i, v in enumerate(vals):
        for j, u in enumerate(v):
            if u < n:
                ids[i, j] = indexer[u]
    return ids, length_to_mask(lens)


_PARA = re.compile(rb"\n[ \t\n\v\f\r]*\n+")                   # cpp/helper.cpp:1121
_SENT = re.compile(rb"[.!?][ \t\n\v\f\r]+")                   # :1135


def chunk_text(text, max_len: int) -> List[bytes]:            # cpp/helper.cpp:1117-1186
    text = _b(text)
    chunks: List[bytes] = []
    paragraphs = [p for p in (trim(x) for x in _split_tokens(_PARA, text)) if p]
    for para in paragraphs:
        sentences: List[bytes] = []
        starts, pos = [], 0
        for m in _SENT.finditer(para):                      # sregex_token_iterator(..., -1) pieces
            starts.append((pos, m.start()))
            pos = m.end()
        if pos < len(para) or not starts:
            starts.append((pos, len(para)))
        for a, e in starts:
            if e > a:
                # "add back the punctuation" (:1148-1153): first delimiter match at/after the piece's start
                m2 = _SENT.search(para, a)
                sentences.append(para[a:e] + (m2.group(0) if m2 else b""))
        cur = b""
        for s in sentences:
            if len(cur) + len(s) + 1 <= max_len:
                if cur:
                    cur += b" "
                cur += s
            else:
                if cur:
                    chunks.append(trim(cur))
                cur = s
        if cur:
            chunks.append(trim(cur))
    if not chunks:
        chunks.append(trim(text))
    return chunks


def _split_tokens(rx, s: bytes) -> List[bytes]:
    """std::sregex_token_iterator(..., -1): pieces between matches; no trailing empty piece."""
    out, pos = [], 0
    for m in rx.finditer(s):
        out.append(s[pos:m.start()])
        pos = m.end()
    if pos < len(s) or not out and not s:
        out.append(s[pos:])
    return out


def sanitize_filename(text, max_len: int) -> bytes:           # cpp/helper.cpp:1070-1111
    t = _b(text)
    out = b""
    cnt = i = 0
    while i < len(t) and cnt < max_len:
        c = t[i]
        if (48 <= c <= 57) or (65 <= c <= 90) or (97 <= c <= 122) or c == 95:
            out += t[i:i + 1]; i += 1
        elif c & 0xE0 == 0xC0 and i + 1 < len(t):
            out += t[i:i + 2]; i += 2
        elif c & 0xF0 == 0xE0 and i + 2 < len(t):
            out += t[i:i + 3]; i += 3
        elif c & 0xF8 == 0xF0 and i + 3 < len(t):
            out += t[i:i + 4]; i += 4
        else:
            out += b"_"; i += 1
        cnt += 1
    return out


def wav_bytes(samples: Sequence[float], sample_rate: int) -> bytes:   # cpp/helper.cpp:943-990
    x = np.asarray(samples, np.float32)
    q = (np.clip(x, np.float32(-1), np.float32(1)) * np.float32(32767)).astype(np.int16)  # truncation toward 0
    data = q.astype("<i2").tobytes()
    hdr = b"RIFF" + struct.pack("<i", 36 + len(data)) + b"WAVE" + b"fmt " + struct.pack(
        "<ihhiihh", 16, 1, 1, sample_rate, sample_rate * 2, 2, 16) + b"data" + struct.pack("<i", len(data))
    return hdr + data


# --------------------------------------------------------------------------- latent-length math
def latent_geometry(duration: np.ndarray, sample_rate: int, base_chunk_size: int, ccf: int):
    """float32 semantics of cpp/helper.cpp:430-438 (== py/helper.py:165-168): never 'improve' to f64.
    Returns (wav_lengths int64[B], latent_len int, latent_mask f32[B,1,max(latent_lengths)])."""
    d = np.asarray(duration, np.float32)
    sr = np.float32(sample_rate)                              # int → float promotion of `d * sample_rate_`
    wav_len_max = np.float32(d.max() * sr)
    wav_lengths = (d * sr).astype(np.int64)                   # static_cast<int64_t>: truncation
    cs = base_chunk_size * ccf
    latent_len = int(np.float32(np.float32(wav_len_max + np.float32(cs)) - np.float32(1)) / np.float32(cs))
    return wav_lengths, latent_len, get_latent_mask(wav_lengths, base_chunk_size, ccf)


# --------------------------------------------------------------------------- _infer / call / batch
class ReferenceTTS:
    """Restatement of TextToSpeech::_infer / call / batch (cpp/helper.cpp:469-734) over four `run`
    callables with Session::Run semantics: run(dict name→ndarray) → ndarray."""

    def __init__(self, cfg: Dict, indexer: Sequence[int], dp: Callable, te: Callable, ve: Callable, voc: Callable):
        self.sr = int(cfg["ae"]["sample_rate"])
        self.bcs = int(cfg["ae"]["base_chunk_size"])
        self.ccf = int(cfg["ttl"]["chunk_compress_factor"])
        self.ldim = int(cfg["ttl"]["latent_dim"])
        self.indexer = indexer
        self.dp, self.te, self.ve, self.voc = dp, te, ve, voc

    def infer_ids(self, text_ids, text_mask, style_ttl, style_dp, total_step: int, speed: float,
                  noise: Callable[[int, int, int], np.ndarray], trace: Optional[dict] = None):
        B = text_ids.shape[0]
        dur = np.asarray(self.dp(dict(text_ids=text_ids, style_dp=style_dp, text_mask=text_mask)), np.float32).reshape(-1)[:B]
        dur = (dur / np.float32(speed)).astype(np.float32)                      # :529-531
        text_emb = self.te(dict(text_ids=text_ids, style_ttl=style_ttl, text_mask=text_mask))
        wav_lengths, L, latent_mask = latent_geometry(dur, self.sr, self.bcs, self.ccf)
        D = self.ldim * self.ccf
        xt = np.asarray(noise(B, D, L), np.float32).reshape(B, D, L)           # injected N(0,1) (:442-455)
        # NB the reference multiplies [B,D,latent_len] by a mask of width max(latent_lengths) (:460-466);
        # the two widths agree (0 mismatches in 2e6 draws) — asserted here.
        assert latent_mask.shape[2] == L, (latent_mask.shape, L)
        xt = xt * latent_mask
        tot = np.full((B,), total_step, np.float32)
        if trace is not None:
            trace.update(duration=dur, wav_lengths=wav_lengths, latent_len=L, latent_mask=latent_mask,
                         text_emb=text_emb, x0=xt.copy(), xs=[])
        for step in range(total_step):                                          # :590-659
            cur = np.full((B,), step, np.float32)
            xt = np.asarray(self.ve(dict(noisy_latent=xt, text_emb=text_emb, style_ttl=style_ttl,
                                         text_mask=text_mask, latent_mask=latent_mask,
                                         total_step=tot, current_step=cur)), np.float32)
            if trace is not None:
                trace["xs"].append(xt.copy())
        wav = np.asarray(self.voc(dict(latent=xt)), np.float32)                 # :662-682
        return wav.reshape(-1), dur

    def _infer(self, texts, langs, style_ttl, style_dp, total_step, speed, noise, trace=None):
        if len(texts) != style_ttl.shape[0]:
            raise RuntimeError("Number of texts must match number of style vectors")       # :479-481
        ids, mask = unicode_processor_call(self.indexer, texts, langs)
        return self.infer_ids(ids, mask, style_ttl, style_dp, total_step, speed, noise, trace)

    def batch(self, texts, langs, style_ttl, style_dp, total_step, speed=1.05, noise=None):
        return self._infer(texts, langs, style_ttl, style_dp, total_step, np.float32(speed), noise)

    def call(self, text, lang, style_ttl, style_dp, total_step, speed=1.05, silence_duration=0.3, noise=None):
        if style_ttl.shape[0] != 1:
            raise RuntimeError("Single speaker text to speech only supports single style")  # :694-696
        wav_cat: Optional[np.ndarray] = None
        dur_cat = np.float32(0)
        for chunk in chunk_text(text, 120 if lang == "ko" else 300):                         # :698-699
            wav, dur = self._infer([chunk], [lang], style_ttl, style_dp, total_step, np.float32(speed), noise)
            if wav_cat is None:
                wav_cat, dur_cat = wav, dur[0]
            else:
                sil = np.zeros(int(np.float32(silence_duration) * np.float32(self.sr)), np.float32)   # :710
                wav_cat = np.concatenate([wav_cat, sil, wav])                                # untrimmed (:706-714)
                dur_cat = np.float32(dur_cat + np.float32(dur[0] + np.float32(silence_duration)))    # :714
        return wav_cat, np.asarray([dur_cat], np.float32)
