#!/usr/bin/env python3
"""Developer probe: wall time of the synthetic generator's flat-array form at scale, split into
the host draw stream (dhj_generator_draws) and the batched pricing.  Usage (GPU box): python scripts/bench_generator.py [n]"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "option-pricing-ffn-lbfgs_b200"))
from src.data import synthetic_generator as gen  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
gen.generate_synthetic_arrays(1000)                                   # context, pinned slots, caches
np.random.seed(42)
t0 = time.perf_counter()
gen._draw_inputs(n)
t_draw = time.perf_counter() - t0
np.random.seed(42)
t0 = time.perf_counter()
data = gen.generate_synthetic_arrays(n)
t_all = time.perf_counter() - t0
with tempfile.TemporaryDirectory() as d:
    t0 = time.perf_counter()
    gen._save_arrays(data, os.path.join(d, "flat"))
    t_save = time.perf_counter() - t0
    t0 = time.perf_counter()
    ds = gen.SyntheticCalibrationSet.load(os.path.join(d, "flat"))
    first = ds[n // 2]
    t_load = time.perf_counter() - t0
print(f"generator, {n} samples ({15 * n} prices): draws {t_draw:.3f} s, total {t_all:.3f} s "
      f"({15 * n / t_all:.3e} prices/s end to end), .npy sink {t_save:.3f} s, mmap load + one record {t_load * 1e3:.1f} ms; "
      f"record {n // 2}: spot {first.spot:.4f} loss {first.final_loss:.3e}")
