"""Accuracy of the kernels' own elementary functions (dhj_fastmath.cuh), compiled for the host, against
mpmath at 40 digits and against libm.  On the device only the 20-bit MUFU seeds differ (the Newton steps
remove the difference), so these bounds carry over; tests/test_gpu_parity.py checks the end result."""
import ctypes
import os
import subprocess

import mpmath as mp
import numpy as np
import pytest
from numpy.ctypeslib import ndpointer

from conftest import PKG, ROOT

mp.mp.dps = 40


@pytest.fixture(scope="module")
def fmlib():
    d = os.path.join(ROOT, "tests", "host_emu")
    out = os.path.join(d, "_build", "libfastmath_emu.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-mfma", "-fPIC", "-shared", "-I", os.path.join(PKG, "csrc"),
                    "-x", "c++", os.path.join(d, "fastmath_emu.cpp"), "-o", out], check=True)
    lib = ctypes.CDLL(out)
    D = ndpointer(np.float64, flags="C")
    lib.fm_sincos.argtypes = [D, ctypes.c_int, D, D]
    lib.fm_exp.argtypes = [D, ctypes.c_int, D]
    lib.fm_exp_tab.argtypes = [D, ctypes.c_int, D]
    lib.fm_log_ratio.argtypes = [D, D, ctypes.c_int, D]
    lib.fm_atan2.argtypes = [D, D, ctypes.c_int, D]
    lib.fm_atan2_tab.argtypes = [D, D, ctypes.c_int, D]
    lib.fm_log_tab.argtypes = [D, ctypes.c_int, D]
    lib.fm_div.argtypes = [D, D, ctypes.c_int, D]
    lib.fm_rcp.argtypes = [D, ctypes.c_int, D]
    lib.fm_sqrt.argtypes = [D, ctypes.c_int, D, D]
    return lib


def ulp_err(got, exact):
    """|got - exact| in ulps of exact (exact: list of mpf)."""
    out = np.empty(len(got))
    for i, (g, e) in enumerate(zip(got, exact)):
        u = np.spacing(abs(float(e))) if e != 0 else np.spacing(0.0)
        out[i] = float(abs(mp.mpf(float(g)) - e) / mp.mpf(float(u)))
    return out


def test_sincos(fmlib):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-900, 900, 3000), rng.uniform(-1e5, 1e5, 1000), rng.uniform(-1, 1, 1000),
                        np.arange(0, 200) * np.pi * (1 + 1e-16), [0.0, -0.0, 1e-300, np.pi / 4, 100 * np.pi]])
    x = np.ascontiguousarray(x)
    s, c = np.empty_like(x), np.empty_like(x)
    fmlib.fm_sincos(x, x.size, s, c)
    es = ulp_err(s, [mp.sin(mp.mpf(float(v))) for v in x])
    ec = ulp_err(c, [mp.cos(mp.mpf(float(v))) for v in x])
    # near zeros of sin/cos the reduction error (<= 1 ulp of r) is relative to a tiny value: bound it in
    # absolute terms there, in ulps elsewhere
    big_s, big_c = np.abs(s) > 1e-3, np.abs(c) > 1e-3
    assert es[big_s].max() <= 1.5 and ec[big_c].max() <= 1.5
    assert np.abs(s - np.sin(x)).max() <= 3e-16 and np.abs(c - np.cos(x)).max() <= 3e-16
    assert np.isnan(_one(fmlib, "sincos", np.nan)[0])


def _one(lib, name, *args):
    if name == "sincos":
        x = np.array([args[0]]); s, c = np.empty(1), np.empty(1)
        lib.fm_sincos(x, 1, s, c); return s[0], c[0]
    if name in ("exp", "exp_tab"):
        x = np.array([args[0]]); o = np.empty(1); getattr(lib, "fm_" + name)(x, 1, o); return o[0]
    if name in ("atan2", "atan2_tab"):
        o = np.empty(1); getattr(lib, "fm_" + name)(np.array([args[0]]), np.array([args[1]]), 1, o); return o[0]
    if name == "log_ratio":
        o = np.empty(1); lib.fm_log_ratio(np.array([args[0]]), np.array([args[1]]), 1, o); return o[0]


@pytest.mark.parametrize("name", ["exp", "exp_tab"])
def test_exp(fmlib, name):
    rng = np.random.default_rng(2)
    x = np.ascontiguousarray(np.concatenate([rng.uniform(-700, 709, 3000), rng.uniform(-40, 5, 3000),
                                             rng.uniform(-1e-3, 1e-3, 500), [0.0, -745.0, -708.5, 709.0, 1.0, -1.0]]))
    o = np.empty_like(x)
    getattr(fmlib, "fm_" + name)(x, x.size, o)
    e = ulp_err(o, [mp.exp(mp.mpf(float(v))) for v in x])
    normal = x > -708
    assert e[normal].max() <= 1.0, e[normal].max()
    assert (o[~normal] == 0.0).all()                             # below the smallest normal: flushed to zero
    assert _one(fmlib, name, -np.inf) == 0.0 and _one(fmlib, name, -800.0) == 0.0
    assert _one(fmlib, name, np.inf) == np.inf and _one(fmlib, name, 710.0) == np.inf
    assert _one(fmlib, name, -707.9) > 0.0
    assert np.isnan(_one(fmlib, name, np.nan))


def test_log_ratio(fmlib):
    rng = np.random.default_rng(3)
    a = np.ascontiguousarray(np.exp(rng.uniform(-30, 30, 4000)))
    b = np.ascontiguousarray(np.concatenate([a[:2000] * rng.uniform(0.2, 5.0, 2000), np.exp(rng.uniform(-30, 30, 2000))]))
    o = np.empty_like(a)
    fmlib.fm_log_ratio(a, b, a.size, o)
    exact = [mp.log(mp.mpf(float(x)) / mp.mpf(float(y))) for x, y in zip(a, b)]
    err = np.array([float(abs(mp.mpf(float(g)) - e)) for g, e in zip(o, exact)])
    scale = np.maximum(1.0, np.abs(o))
    assert (err / scale).max() <= 2.3e-16, (err / scale).max()
    # ratios next to 1: the result keeps full RELATIVE accuracy (a - b' is exact)
    a2 = np.ascontiguousarray(1.0 + rng.uniform(-1e-6, 1e-6, 1000)); b2 = np.ones_like(a2)
    o2 = np.empty_like(a2)
    fmlib.fm_log_ratio(a2, b2, a2.size, o2)
    assert ulp_err(o2, [mp.log(mp.mpf(float(x))) for x in a2]).max() <= 2.0
    assert np.isnan(_one(fmlib, "log_ratio", np.nan, 1.0)) and np.isnan(_one(fmlib, "log_ratio", 1.0, np.nan))
    assert _one(fmlib, "log_ratio", 3.0, 3.0) == 0.0


def test_log_tab(fmlib):
    rng = np.random.default_rng(6)
    w = np.ascontiguousarray(np.concatenate([np.exp(rng.uniform(-40, 40, 4000)), rng.uniform(0.2, 2.0, 4000),
                                             1.0 + rng.uniform(-1e-3, 1e-3, 500), [1.0, 0.5, 2.0, 1.0 - 2 ** -53]]))
    o = np.empty_like(w)
    fmlib.fm_log_tab(w, w.size, o)
    exact = [mp.log(mp.mpf(float(v))) for v in w]
    err = np.array([float(abs(mp.mpf(float(g)) - e)) for g, e in zip(o, exact)])
    assert (err / np.maximum(1.0, np.abs(o))).max() <= 2.3e-16, (err / np.maximum(1.0, np.abs(o))).max()
    assert abs(o[-4]) <= 5e-18                 # log(1): the table entry and the polynomial cancel to rounding
    bad = np.array([np.nan, np.inf]); ob = np.empty(2)
    fmlib.fm_log_tab(bad, 2, ob)
    assert np.isnan(ob).all()


@pytest.mark.parametrize("name", ["atan2", "atan2_tab"])
def test_atan2(fmlib, name):
    rng = np.random.default_rng(4)
    n = 6000
    y = rng.standard_normal(n) * np.exp(rng.uniform(-10, 10, n))
    x = rng.standard_normal(n) * np.exp(rng.uniform(-10, 10, n))
    y[:8] = [0.0, -0.0, 1.0, -1.0, 0.0, -0.0, 1.0, 1.0]
    x[:8] = [1.0, 1.0, 0.0, 0.0, -1.0, -1.0, 1.0, -1.0]
    y, x = np.ascontiguousarray(y), np.ascontiguousarray(x)
    o = np.empty(n)
    getattr(fmlib, "fm_" + name)(y, x, n, o)
    e = ulp_err(o[8:], [mp.atan2(mp.mpf(float(a)), mp.mpf(float(b))) for a, b in zip(y[8:], x[8:])])
    assert e.max() <= 2.0, e.max()
    assert np.array_equal(o[:8], np.arctan2(y[:8], x[:8])) or np.abs(o[:8] - np.arctan2(y[:8], x[:8])).max() <= 5e-16
    assert np.signbit(o[1]) and o[5] == -np.pi
    assert np.isnan(_one(fmlib, name, np.nan, 1.0)) and np.isnan(_one(fmlib, name, 1.0, np.nan))


def test_div_rcp_sqrt(fmlib):
    rng = np.random.default_rng(5)
    n = 20000
    a = np.ascontiguousarray(rng.standard_normal(n) * np.exp(rng.uniform(-50, 50, n)))
    b = np.ascontiguousarray(rng.standard_normal(n) * np.exp(rng.uniform(-50, 50, n)))
    o = np.empty(n)
    fmlib.fm_div(a, b, n, o)
    want = a / b
    assert (np.abs(o - want) <= np.spacing(np.abs(want))).all()
    assert (o == want).mean() > 0.99                                # nearly always correctly rounded
    fmlib.fm_rcp(b, n, o)
    assert (np.abs(o - 1.0 / b) <= np.spacing(np.abs(1.0 / b))).all()
    p = np.ascontiguousarray(np.abs(a))
    s, y = np.empty(n), np.empty(n)
    fmlib.fm_sqrt(p, n, s, y)
    assert (np.abs(s - np.sqrt(p)) <= np.spacing(np.sqrt(p))).all() and (s == np.sqrt(p)).mean() > 0.99
    assert ulp_err(y[:3000], [1 / mp.sqrt(mp.mpf(float(v))) for v in p[:3000]]).max() <= 1.5
    z = np.zeros(1); fmlib.fm_sqrt(z, 1, s, y); assert s[0] == 0.0
