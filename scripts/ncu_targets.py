#!/usr/bin/env python3
"""Small drivers for ncu captures of the kernels that quick_perf.py does not isolate (run via gpurun, under ncu):
    loss   k_loss_batch, the fused loss / forward-difference kernel at a mid-size round (500 optimiser states =
           7 000 loss evaluations = 21 000 items: one wave of blocks, the regime of C5's later rounds)
    gen    k_gen_draws / k_price_batch / k_gen_market: one dataset sweep of 4 Mi samples (dhj_generate_dev)
usage: python scripts/ncu_targets.py loss|gen"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "option-pricing-ffn-lbfgs_b200"))
sys.path.insert(0, ROOT)
import dhj  # noqa: E402
from bench import GEN, C4_SEED, GRID_K, GRID_T  # noqa: E402

which = sys.argv[1]
ctx = dhj.default_context()
if which == "loss":
    n = 500
    data = ctx.generate(C4_SEED, 0, n, **GEN)
    spots, market = data["spots"], data["market"]
    K = np.tile(np.array(GRID_K)[None, :] * spots[:, None] / 100.0, (1, 3)); T = np.repeat(np.array(GRID_T), 5)
    mk = ctx.market(spots, 0.03, K, T, np.ones(15), market)
    x = dhj.initial_guesses(spots, K, T, market, 3)[:, 2, :]
    idx = np.arange(n, dtype=np.int32)
    import time
    for rep in range(6):
        t0 = time.perf_counter()
        f, g = mk.loss_fd(x, 1e-8, idx)
        dt = time.perf_counter() - t0
    print(f"loss_fd, {n} states (fused kernel, {14 * n} evaluations): {dt * 1e6:.0f} us per call incl. copies")
else:
    import torch
    n = 1 << 22
    dev = torch.device("cuda", 0)
    bufs = {k: torch.empty(s, dtype=torch.float64, device=dev)
            for k, s in (("params", (n, 13)), ("spots", (n,)), ("model", (n, 15)), ("market", (n, 15)), ("loss", (n,)))}
    args = [GEN[k] for k in ("path_len", "lo", "hi", "persistence", "spot0", "ret_mean", "ret_sd", "noise_sd",
                             "strikes_rel", "maturities", "r")]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(3):
        e0.record()
        ctx.generate_dev(C4_SEED, 0, n, *args, bufs["params"].data_ptr(), bufs["spots"].data_ptr(), bufs["model"].data_ptr(),
                         bufs["market"].data_ptr(), bufs["loss"].data_ptr(), torch.cuda.current_stream().cuda_stream)
        e1.record(); torch.cuda.synchronize()
    print(f"generate_dev, {n} samples: {e0.elapsed_time(e1):.2f} ms")
