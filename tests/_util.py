"""Shared helpers for the parity tests: seeded synthetic inputs + oracle sessions."""
import os

import numpy as np

WORDS = ("the quick brown fox jumps over a lazy dog while seven silver ships sail south toward quiet harbours and "
         "nobody knows why morning light feels warmer after rain or how distant thunder rolls across open fields "
         "yesterday we walked along the river talking about music science and old friends from school").split()

KO = "오늘 아침에 공원을 산책했는데, 새소리와 바람 소리가 너무 기분 좋았어요."                       # reference test_all.sh:64
ES = "El niño comió piñas en la montaña, ¿verdad? ¡Sí, señor!"
PT = "A ação e o coração não são fáceis de explicar, mas vovô tentou."
FR = "Où est l'hôtel? Ça coûte très cher, naïve Zoë préfère le café déjà."


def make_text(rng, n_chars):
    out = []
    while sum(len(w) + 1 for w in out) < n_chars:
        out.append(WORDS[rng.integers(len(WORDS))])
    s = " ".join(out)[:max(n_chars, 2)].strip()
    return s[0].upper() + s[1:]


def make_batch(seed, n, lo=20, hi=120):
    rng = np.random.default_rng(seed)
    return [make_text(rng, int(rng.integers(lo, hi + 1))) for _ in range(n)], ["en"] * n


def styles(root, names):
    from oracle.pipeline import load_style
    return load_style([os.path.join(root, "voice_styles", n + ".json") for n in names])


def snr_db(x, ref):
    x = np.asarray(x, np.float64); ref = np.asarray(ref, np.float64)
    err = ((x - ref) ** 2).sum()
    return 200.0 if err == 0 else 10 * np.log10((ref ** 2).sum() / err)
