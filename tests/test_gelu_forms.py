"""Error bounds of the two erf-GELU forms the tcgen05 epilogues use (csrc/gemm_tc.cuh), restated operation by operation in float32 numpy
against the exact (x (erf(x / sqrt2) + 1)) / 2 the graphs define (reference: Erf-GELU node chain of every ConvNeXt block, run by ORT at
cpp/helper.cpp:643, :668). The device versions use MUFU.RCP / MUFU.EX2 (2 ulp approximations); the numpy restatement uses the correctly
rounded operations, so this pins the FORMULAS' error, which dominates: the stated bounds in the kernel comments are these numbers."""
import math

import numpy as np
from scipy.special import erf

f = np.float32


def _exact(x):
    x = x.astype(np.float64)
    return (x * (erf(x / math.sqrt(2.0)) + 1.0)) * 0.5


def gelu_two_mufu(x):
    """gelu_erf_mufu / gelu_erf_mufu2: Abramowitz-Stegun 7.1.26, reciprocal + exp2."""
    ax = np.abs(x)
    t = (f(1) / (f(0.3275911 * 0.70710678) * ax + f(1))).astype(f)
    poly = (t * f(1.061405429) + f(-1.453152027)).astype(f)
    for c in (1.421413741, -0.284496736, 0.254829592):
        poly = (poly * t + f(c)).astype(f)
    poly = (poly * t).astype(f)
    e = np.exp2(((x * x).astype(f) * f(-0.5 * 1.4426950408889634)).astype(f)).astype(f)
    erf_abs = (f(1) - (poly * e).astype(f)).astype(f)
    h = (f(0.5) * x).astype(f)
    return (np.abs(h) * erf_abs + h).astype(f)


def gelu_one_mufu(x):
    """gelu_erf_rcp2: Abramowitz-Stegun 7.1.28, degree-6 polynomial, four squarings, one reciprocal."""
    a = (0.0705230784, 0.0422820123, 0.0092705272, 0.0001520143, 0.0002765672, 0.0000430638)
    c = [f(a[k] / math.sqrt(2.0) ** (k + 1)) for k in range(6)]
    ax = np.abs(x)
    with np.errstate(over="ignore"):
        q = (ax * c[5] + c[4]).astype(f)
        for k in (3, 2, 1, 0):
            q = (q * ax + c[k]).astype(f)
        q = (q * ax + f(1)).astype(f)
        for _ in range(4):
            q = (q * q).astype(f)
        r = (f(1) / q).astype(f)
    h = (f(0.5) * x).astype(f)
    ah = (f(0.5) * ax).astype(f)
    return ((h + ah).astype(f) - (ah * r).astype(f)).astype(f)


def test_erf_gelu_forms_stay_within_their_stated_error():
    x = np.linspace(-12.0, 12.0, 1_000_001).astype(f)
    want = _exact(x)
    assert np.abs(gelu_two_mufu(x) - want).max() <= 6e-7
    assert np.abs(gelu_one_mufu(x) - want).max() <= 9e-7
    # the one-MUFU form feeds an fp16 operand (vocoder pw1): from |y| = 2e-3 up its error is below half an fp16 ulp of the output; below
    # that (x < -3.2, where GELU is a ~1e-4 tail) it is the same few 1e-7 in absolute terms, as for the two-MUFU form
    big = np.abs(want) >= 2e-3
    assert (np.abs(gelu_one_mufu(x) - want)[big] / np.abs(want)[big]).max() <= 2.0 ** -11
    assert (np.abs(gelu_two_mufu(x) - want)[big] / np.abs(want)[big]).max() <= 2.0 ** -11


def test_one_mufu_form_saturates_cleanly():
    """Every finite input gives a finite, correct limit (an infinite accumulator would give inf * 0 = NaN; fp32 accumulations of fp16
    operands over K <= 2048 cannot reach it)."""
    x = np.array([40.0, 100.0, 1e4, 3e38, -40.0, -100.0, -1e4, -3e38, 0.0, -0.0], dtype=f)
    got = gelu_one_mufu(x)
    assert np.all(np.isfinite(got[:3])) and np.array_equal(got[:3], x[:3])            # p^16 overflows to +inf, 1/inf = 0, result x
    assert np.array_equal(got[4:7], np.zeros(3, f))
    assert got[8] == 0 and got[9] == 0
