"""CPU model of `score_rows_fx` (conformal_bayesopt_b200/csrc/score.cuh): the soft-rank sigmoid of every q = 1 kernel.

The kernel evaluates rho = sigmoid(dd / temp) = 1 / (1 + 2^(-x/16)), x = 16 dd / (temp ln 2), from the two 32-bit words of
t = M - x (M = 1.5 * 2^20): fraction -> cubic, integer part -> table index and exponent, bit 19 -> hard count.  This test
restates those bit manipulations in numpy WITH THE CONSTANTS PARSED FROM THE HEADER, so that an edit of a constant or of the
word layout that breaks the approximation fails here, on the CPU, before any GPU run (reference semantics of the soft rank:
conformalbo/helpers.py:283-300 -- torch.sigmoid of the score differences over temp)."""
import os
import re

import numpy as np

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "conformal_bayesopt_b200", "csrc", "score.cuh")


def _consts():
    src = open(HDR).read()

    def num(name):
        m = re.search(r"\b" + name + r"\s*=\s*([-+0-9.eE]+)", src)
        assert m, name
        return float(m.group(1))

    return dict(M=num("kFxMagic"), C3=num("kFxC3"), B2=num("kFxB2"), B1=num("kFxB1"), B0=num("kFxB0"))


def _words(t):
    u = t.view(np.uint64)
    return (u >> np.uint64(32)).astype(np.uint32), (u & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def _double(hi, lo):
    return ((hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)).view(np.float64)


def fx_sigmoid(x, c):
    """rho and the hard-count bit exactly as the kernel forms them (numpy has no fma: differences ~1e-16)."""
    t = np.float64(c["M"]) - x                                    # the kernel's second fma delivers this, rounded once
    hi, lo = _words(t)
    count_bit = (hi >> np.uint32(19)).astype(np.int64) - 0x826    # [x <= 0] after rounding to 2^-32
    w = _double(np.uint32(0x3FF00000) + (lo >> np.uint32(12)), lo << np.uint32(20))       # 1 + z 2^-32
    pl = ((w + c["B2"]) * w + c["B1"]) * w + c["B0"]              # 2^(z/16) / c3
    j = np.arange(16)
    tab = c["C3"] * np.exp2(j / 16.0)
    thi, tlo = _words(tab)
    thi = thi - (j.astype(np.uint32) << np.uint32(16))            # exp_tabfx_init: pattern minus (j << 16)
    idx = (hi & np.uint32(15)).astype(np.int64)
    ts = _double(thi[idx] + (hi << np.uint32(16)), tlo[idx])      # exp_tab_scaled: + (n << 16) restores j and adds kk << 20
    return 1.0 / (ts * pl + 1.0), count_bit


def test_constants_reproduce_the_sigmoid_over_the_whole_unclamped_range():
    c = _consts()
    assert c["M"] == 1.5 * 2 ** 20
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-946.0, 946.0, 400000), rng.uniform(-40.0, 40.0, 400000), rng.normal(0, 1e-3, 1000),
                        np.array([-946.0, 946.0, -945.999999, 945.999999, 0.0, 1.0, -1.0, 16.0, -16.0, 15.9999999999, 2.0 ** -33])])
    rho, cnt = fx_sigmoid(x, c)
    want = 1.0 / (1.0 + np.exp2(-x / 16.0))
    assert np.all(np.isfinite(rho))
    assert np.abs(rho - want).max() < 4e-10                       # the kernels' documented bound (|err| < 3e-10 + rounding)
    rel = np.abs(rho / want - 1.0)
    assert rel.max() < 2e-9                                       # also relative: the tails keep their digits
    # hard count: bit 19 of the high word is [x <= 0] up to the 2^-32 rounding of x
    far = np.abs(x) > 2.0 ** -31
    assert np.array_equal(cnt[far], (x[far] < 0).astype(np.int64))
    assert set(np.unique(cnt)) <= {0, 1}


def test_clamp_words_are_the_high_words_of_the_clamp_bounds():
    """CLAMP variant: the integer clamp constants must be the high words of M -+ 946 (zero low words)."""
    src = open(HDR).read()
    m = re.search(r"kLo\s*=\s*(0x[0-9a-fA-F]+),\s*kHi\s*=\s*(0x[0-9a-fA-F]+)", src)
    assert m
    c = _consts()
    for word, v in ((int(m.group(1), 16), c["M"] - 946.0), (int(m.group(2), 16), c["M"] + 946.0)):
        hi, lo = _words(np.array([v]))
        assert int(hi[0]) == word and int(lo[0]) == 0
    # a clamped t (any low word) still lands inside the table's exponent range: |x| < 947
    hi = np.array([int(m.group(1), 16), int(m.group(2), 16)], dtype=np.uint32)
    t = _double(hi, np.array([0xFFFFFFFF, 0xFFFFFFFF], dtype=np.uint32))
    assert np.all(np.abs(c["M"] - t) < 947.0)
    rho, _ = fx_sigmoid(c["M"] - t, c)
    assert np.all(np.isfinite(rho)) and rho[0] > 1 - 1e-16 and rho[1] < 1e-17      # t = M - 946: x = +946, rho = 1; t = M + 947: rho = 0
