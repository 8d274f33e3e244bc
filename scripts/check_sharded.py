#!/usr/bin/env python3
"""Multi-GPU consistency check on real hardware (NCCL): the sharded paths must give the bits of the unsharded ones.
  C4  generate_synthetic_arrays(seed=, sharded=True): the ranks' shards, written to disk and read back, concatenate to
      the dataset one GPU generates in one piece
  C5  calibrate_many_sharded == calibrate_many on every rank's own GPU
  C2  price_grid_sharded == price_grid
usage: torchrun --nproc-per-node N scripts/check_sharded.py   (rank 0 prints one line per check)"""
import os
import shutil
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "option-pricing-ffn-lbfgs_b200")
sys.path.insert(0, PKG)
sys.path.insert(0, ROOT)
for sub in ("models", "calibration", "data"):
    sys.path.insert(0, os.path.join(PKG, "src", sub))

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
os.environ["DHJ_DEVICE"] = str(local)
import dhj  # noqa: E402
import synthetic_generator as gen  # noqa: E402
from dhj.shard import price_grid_sharded  # noqa: E402
from bench import GEN, GRID_K, GRID_T  # noqa: E402

ctx = dhj.default_context()
ok = True

# ---- C4 ---------------------------------------------------------------------------------------------------
n = 20_300                                                    # 41 histories: ragged over the ranks, partial last history
tmp = os.path.join(tempfile.gettempdir(), "dhj_check_sharded")
if rank == 0:
    shutil.rmtree(tmp, ignore_errors=True)
dist.barrier()
mine = gen.generate_synthetic_arrays(n, seed=17, path_len=500, sharded=True, save_path=tmp)
dist.barrier()
if rank == 0:
    whole = gen.generate_synthetic_arrays(n, seed=17, path_len=500)
    shards = gen.load_sharded(tmp)
    same = len(shards) == world
    for key in ("params", "spots", "model_prices", "market_prices", "losses", "strikes"):
        cat = np.concatenate([np.asarray(s.data[key]) for s in shards], axis=0)
        same = same and bool(np.array_equal(cat, whole[key]))
    same = same and shards[-1][len(shards[-1]) - 1].spot == whole["spots"][-1]
    print(f"C4 sharded dataset ({world} shards, {n} samples): union of the shards == one-piece dataset: {same}")
    ok = ok and same

# ---- C5 ---------------------------------------------------------------------------------------------------
m = 600
data = ctx.generate(7, 0, m, **GEN)
spots, market = data["spots"], data["market"]
K = np.tile(np.array(GRID_K)[None, :] * spots[:, None] / 100.0, (1, 3)); T = np.repeat(np.array(GRID_T), 5)
np.random.seed(1)
x0 = dhj.initial_guesses(spots, K, T, market, 3)
sh = dhj.calibrate_many_sharded(spots, 0.03, K, T, np.ones(15), market, maxiter=60, multi_start=3, x0=x0, device=dev)
one = dhj.calibrate_many(spots, 0.03, K, T, np.ones(15), market, maxiter=60, multi_start=3, x0=x0, pipelines=1)
same = all(np.array_equal(sh[k], one[k], equal_nan=True) for k in ("x", "final_loss", "iterations", "status", "best_start", "model_prices"))
flags = torch.tensor([1.0 if same else 0.0], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"C5 calibrate_many_sharded over {world} ranks == calibrate_many on one GPU ({m} markets x 3 starts): {bool(flags.item())}")
    ok = ok and bool(flags.item())

# ---- C2 ---------------------------------------------------------------------------------------------------
rng = np.random.default_rng(5)
params = rng.uniform(GEN["lo"], GEN["hi"], size=(100_003, 13))
full = price_grid_sharded(ctx.price_grid, params, 100.0, np.array(GRID_K), np.array(GRID_T), 0.03, device=dev)
want = ctx.price_grid(params, 100.0, np.array(GRID_K), np.array(GRID_T), 0.03)
flags = torch.tensor([1.0 if np.array_equal(full, want) else 0.0], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"C2 price_grid_sharded over {world} ranks == price_grid (100 003 sets, gathered over NCCL): {bool(flags.item())}")
    ok = ok and bool(flags.item())
    print("ALL OK" if ok else "MISMATCH")
dist.destroy_process_group()
