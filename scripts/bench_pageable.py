#!/usr/bin/env python3
"""Developer probe: host-buffer entry point (dhj_price_grid) with PAGEABLE NumPy arrays vs pinned torch buffers on
the C2 shape.  Usage (GPU box): python scripts/bench_pageable.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "option-pricing-ffn-lbfgs_b200"))
import torch  # noqa: E402

import dhj  # noqa: E402

ctx = dhj.Context(0)
R = np.array([(0.025, 0.080), (1.5, 4.5), (0.025, 0.065), (0.20, 0.50), (-0.85, -0.40), (0.020, 0.070), (0.30, 1.20),
              (0.025, 0.070), (0.10, 0.35), (-0.70, -0.20), (0.05, 0.25), (-0.08, -0.01), (0.03, 0.12)])
P = 1 << 20
Ks, Ts = np.array([90., 95, 100, 105, 110]), np.array([.25, .5, 1.])
params = np.random.default_rng(0).uniform(R[:, 0], R[:, 1], size=(P, 13))
for name, pin in (("pageable", False), ("pinned", True)):
    if pin:
        pin_in = torch.from_numpy(params).pin_memory()
        pin_out = torch.empty((P, 3, 5), dtype=torch.float64).pin_memory()
        a, out = pin_in.numpy(), pin_out.numpy()
    else:
        a, out = params, np.empty((P, 3, 5))
    for _ in range(2):
        ctx.price_grid(a, 100.0, Ks, Ts, 0.03, out=out)
    t0 = time.perf_counter()
    for _ in range(5):
        ctx.price_grid(a, 100.0, Ks, Ts, 0.03, out=out)
    dt = (time.perf_counter() - t0) / 5
    print(f"{name} host buffers: {dt * 1e3:.2f} ms per 1Mi sets -> {P * 15 / dt:.4g} prices/s")
