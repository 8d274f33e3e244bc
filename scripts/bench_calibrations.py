#!/usr/bin/env python3
"""Calibration benchmarks (second half of BASELINE.json's metric: "s per 15-option calib"):
  C1  README configuration: one 15-option market, calibrate(maxiter=300, multi_start=3) through the drop-in class
  C5  n independent markets x 3 starts in lock-step (dhj.calibrate_many), optionally sharded over ranks
usage: python scripts/bench_calibrations.py [n_markets]     (torchrun for several GPUs)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "option-pricing-ffn-lbfgs_b200")
sys.path.insert(0, PKG)
for sub in ("models", "calibration", "data"):
    sys.path.insert(0, os.path.join(PKG, "src", sub))

import dhj  # noqa: E402
from lbfgs_calibrator import DoubleHestonJumpCalibrator  # noqa: E402
from synthetic_generator import PARAM_RANGES, STRIKES, MATURITIES  # noqa: E402

n_markets = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
device = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    device = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=device)
ctx = dhj.default_context()

# ---- C1 -----------------------------------------------------------------------------------------------------
true_p = np.array([0.04, 2.0, 0.04, 0.3, -0.5, 0.04, 1.5, 0.04, 0.2, -0.3, 0.1, 0.0, 0.1])    # tests/test_suite.py:274-279
K1 = np.tile([90.0, 95.0, 100.0, 105.0, 110.0], 3)
T1 = np.repeat([0.25, 0.5, 1.0], 5)
mkt = ctx.price_list(true_p, 100.0, K1, T1, np.ones(15), 0.05)[0]
opts = [{"strike": K1[j], "maturity": T1[j], "price": mkt[j], "option_type": "call"} for j in range(15)]
cal = DoubleHestonJumpCalibrator(100.0, 0.05, opts)
cal.calibrate(maxiter=3, multi_start=1)
times = []
for rep in range(5):
    np.random.seed(0)
    t0 = time.perf_counter()
    r1 = cal.calibrate(maxiter=300, multi_start=3)
    times.append(time.perf_counter() - t0)
c1 = {"seconds_median": float(np.median(times)), "seconds_min": float(min(times)), "final_loss": float(r1.final_loss),
      "iterations": int(r1.iterations), "message": r1.message}

# ---- C5 -----------------------------------------------------------------------------------------------------
rng = np.random.default_rng(7)
lo = np.array([v[0] for v in PARAM_RANGES.values()]); hi = np.array([v[1] for v in PARAM_RANGES.values()])
params = rng.uniform(lo, hi, size=(n_markets, 13))
spots = 100.0 * np.exp(0.05 * rng.standard_normal(n_markets))
strikes = np.tile(STRIKES[None, :] * spots[:, None] / 100.0, (1, 3))
mats = np.repeat(MATURITIES, 5)
from dhj.shard import shard_bounds  # noqa: E402
model = ctx.price_grid(params, spots, STRIKES.astype(float), MATURITIES, 0.03, scale_by_spot=True).reshape(n_markets, 15)
market = model * (1 + 0.02 * rng.standard_normal((n_markets, 15)))                         # generator's 2 % noise
# warm-up at full size: device / pinned buffers of the final shapes, NCCL communicator (its first collective costs ~1 s)
np.random.seed(1)
dhj.calibrate_many_sharded(spots, 0.03, strikes, mats, np.ones(15), market, maxiter=2, multi_start=3, device=device)
if world > 1:
    dist.barrier()
np.random.seed(1)
launches0 = ctx.launch_count
t0 = time.perf_counter()
res = dhj.calibrate_many_sharded(spots, 0.03, strikes, mats, np.ones(15), market, maxiter=300, multi_start=3,
                                 device=device)
wall = time.perf_counter() - t0
if rank == 0:
    fl = res["final_loss"]
    out = {"c1": c1,
           "c5": {"markets": n_markets, "starts": 3, "n_gpus": world, "seconds": wall,
                  "calibrations_per_s": n_markets / wall, "rounds": int(res["rounds"]),
                  "launches_rank0": int(ctx.launch_count - launches0),
                  "final_loss_median": float(np.median(fl)), "final_loss_p95": float(np.percentile(fl, 95)),
                  "frac_below_1pct": float(np.mean(fl * 100 < 1.0)), "success_rate": float(np.mean(res["success"])),
                  "iterations_median": float(np.median(res["iterations"]))}}
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
