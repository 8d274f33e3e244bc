#!/usr/bin/env python3
"""Developer probe (CPU only): accuracy of the kernel ARITHMETIC in csrc/dhj_math.cuh, through the test-only host
emulation (tests/host_emu/emu.cpp, g++), against the C restatement of the reference on a large random sample.
Run before and after an algebra change: the error statistics must not move.
usage: python scripts/emu_check.py [n_sets]"""
import ctypes
import os
import subprocess
import sys

import numpy as np
from numpy.ctypeslib import ndpointer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cos_oracle as O  # noqa: E402

out = os.path.join(ROOT, "tests", "host_emu", "_build", "libdhj_emu_check.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-fopenmp", "-I",
                os.path.join(ROOT, "option-pricing-ffn-lbfgs_b200", "csrc"), "-x", "c++",
                os.path.join(ROOT, "tests", "host_emu", "emu.cpp"), "-o", out], check=True)
lib = ctypes.CDLL(out)
D, I = ndpointer(np.float64, flags="C"), ndpointer(np.int32, flags="C")
lib.emu_price_list.argtypes = [D, D, ctypes.c_int, D, ctypes.c_int, D, I, ctypes.c_double, ctypes.c_double,
                               ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_double, D, D]


def emu(params, S0, strike, mat, call, r, q, N):
    P, M = params.shape[0], mat.size
    outp, ab = np.empty((P, M)), np.empty((P, M, 2))
    lib.emu_price_list(np.ascontiguousarray(params), np.ascontiguousarray(S0), 1, np.ascontiguousarray(strike), M,
                       np.ascontiguousarray(mat), np.ascontiguousarray(call.astype(np.int32)), r, q, P, M, N, 10.0, outp, ab)
    return outp


n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
rng = np.random.default_rng(99)
for tag, lo, hi in (("generator ranges", O.GENERATOR_RANGES[:, 0], O.GENERATOR_RANGES[:, 1]),
                    ("wide ranges", np.array([0.005, 0.1, 0.005, 0.05, -0.99, 0.005, 0.1, 0.005, 0.05, -0.99, 0.0, -0.3, 0.01]),
                     np.array([0.25, 10.0, 0.25, 1.0, 0.5, 0.25, 10.0, 0.25, 1.0, 0.5, 2.0, 0.2, 0.4]))):
    params = rng.uniform(lo, hi, size=(n, 13))
    spots = rng.uniform(70, 140, size=n)
    K = np.tile(O.GENERATOR_STRIKES_REL[None, :] * spots[:, None] / 100.0, (1, 3))
    T = np.repeat(O.GENERATOR_MATURITIES, 5)
    call = (np.arange(15) % 4 != 3)
    for N in (128, 256):
        got = emu(params, spots, K, T, call, 0.03, 0.01, N)
        want = O.c_price_batch(params, spots, K, T, call, 0.03, 0.01, N)
        scale = np.maximum(np.abs(want), 1e-3 * spots[:, None])
        err = np.abs(got - want) / scale
        err = np.where(np.isnan(want) & np.isnan(got), 0.0, err)
        print(f"{tag:17s} N={N}: max {np.nanmax(err):.3e}  p99.9 {np.nanquantile(err, 0.999):.3e}  median {np.nanmedian(err):.3e}  "
              f"rms {np.sqrt(np.nanmean(err ** 2)):.3e}  nan-mismatch {int((np.isnan(got) != np.isnan(want)).sum())}")
