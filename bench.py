#!/usr/bin/env python3
"""bench.py — COS option prices per second (N=128, FP64) on B200, and the reference CPU path beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], "C2"): 1 048 576 random parameter sets in the generator's ranges
(src/data/synthetic_generator.py:75-89 of the reference) x the 15-option grid (5 strikes x 3 maturities),
S0 = 100, r = 0.03, calls, COS N = 128, float64.  One "step" = one pass of the pricing kernel over that
batch.  With N GPUs every rank prices its own, differently seeded, batch of the same size (weak
scaling; the path shards by parameter set with no data-path collective — SURVEY §8e); the per-rank
price checksums are gathered with one NCCL all_gather after the timed steps.

Numbers on the JSON line:
  value        prices/s over all ranks, inputs resident in HBM, CUDA-event time of the K steps (max
               over ranks), L2 flushed between steps (untimed);
  e2e          the same metric through the host-buffer C-ABI call (`dhj_price_grid` via
               `dhj.Context.price_grid`): pinned host params in, pinned host prices out, H2D and D2H
               inside the timed region;
  roofline     FP64-pipe roofline of k_price: algorithmic FLOP (55 731 FLOP per price for
               this grid) / kernel time, against the DFMA rate measured on this GPU by `dhj_fp64_peak`
               (MEASURED_PEAKS.json has no FP64 number; nominal 148 SM x 64 FMA/clk x 2 x 1.965 GHz =
               37.2 TFLOP/s is reported too).  `hbm` carries the (irrelevant, compute-bound) HBM figure;
  cpu_baseline the oracle's scalar port of the reference algorithm (the reference's execution model:
               one option at a time, Python loop over k) on all host cores, bounded sample.
`--impl reference` times only that CPU path and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "option-pricing-ffn-lbfgs_b200"))

METRIC = "COS option prices/sec (N=128, FP64)"
UNIT = "prices/s"
# FLOP per price = N * (F_CF / nK + F_PAY), F_CF = 1707, F_PAY = 94
F_CF, F_PAY = 1707.0, 94.0
NOMINAL_FP64_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12

WORKLOADS = {
    # name: (sets per GPU, strikes, maturities, N, r, description)
    "c2": dict(P=1 << 20, strikes=[90.0, 95.0, 100.0, 105.0, 110.0], maturities=[0.25, 0.5, 1.0], N=128, r=0.03,
               name="C2: 1Mi random parameter sets x 15-option grid (5K x 3T), N=128, S0=100, r=0.03, calls"),
    "c3": dict(P=1024, strikes=list(np.linspace(80.0, 120.0, 200)), maturities=list(np.linspace(0.25, 2.0, 20)),
               N=256, r=0.03,
               name="C3: dense surface 200K x 20T, N=256, 1024 parameter sets per launch"),
    # C4: the 100 M-set dataset, STRONG scaling: the sets are divided over the ranks, parameters generated on the device
    "c4": dict(P=100_000_000, strikes=[90.0, 95.0, 100.0, 105.0, 110.0], maturities=[0.25, 0.5, 1.0], N=128, r=0.03,
               strong=True, device_params=True,
               name="C4: synthetic_generator dataset, 100M parameter sets x 15 options (K = K_rel*spot/100), N=128, "
                    "sharded by parameter set"),
}


def flop_per_price(nK: int, N: int) -> float:
    return N * (F_CF / nK + F_PAY)


# sampling ranges of the reference's generator (src/data/synthetic_generator.py:75-89)
PARAM_RANGES = np.array([
    (0.025, 0.080), (1.5, 4.5), (0.025, 0.065), (0.20, 0.50), (-0.85, -0.40),
    (0.020, 0.070), (0.30, 1.20), (0.025, 0.070), (0.10, 0.35), (-0.70, -0.20),
    (0.05, 0.25), (-0.08, -0.01), (0.03, 0.12)])


def gen_params(P: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.uniform(PARAM_RANGES[:, 0], PARAM_RANGES[:, 1], size=(P, 13))


# ------------------------------------------------------------------------------------------------
# CPU path (oracle scalar port = the reference's execution model), all cores
# ------------------------------------------------------------------------------------------------
def _cpu_chunk(job):
    params, strikes, maturities, N, r = job
    from oracle import cos_oracle as O
    t0 = time.perf_counter()
    acc = 0.0
    for p in params:
        for T in maturities:
            for K in strikes:
                acc += O.price_scalar(p, 100.0, K, T, r, True, 0.0, N)
    return len(params) * len(strikes) * len(maturities), acc, time.perf_counter() - t0


def cpu_throughput(wl, sets_per_core: int, pool, cores: int, seed: int):
    """prices/s of the scalar port on `cores` processes; returns (value, n_prices, seconds)."""
    params = gen_params(sets_per_core * cores, seed)
    jobs = [(params[i::cores], wl["strikes"], wl["maturities"], wl["N"], wl["r"]) for i in range(cores)]
    t0 = time.perf_counter()
    done = pool.map(_cpu_chunk, jobs)
    dt = time.perf_counter() - t0
    n = sum(d[0] for d in done)
    return n / dt, n, dt


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def make_pool(cores):
    import multiprocessing as mp
    return mp.get_context("spawn").Pool(cores)


def run_reference_arm(args, wl):
    """--impl reference: the reference algorithm's CPU path, bounded sample per step, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    per_price = 6.1e-3 if wl["N"] <= 128 else 11.3e-3          # BASELINE.md §2, s per price per core
    n_opt = len(wl["strikes"]) * len(wl["maturities"])
    # ~2 s of work per step per core
    sets_per_core = max(1, int(round(2.0 / (per_price * n_opt))))
    with make_pool(cores) as pool:
        for w in range(args.warmup):
            cpu_throughput(wl, sets_per_core, pool, cores, 1000 + w)
        n_tot, t_tot = 0, 0.0
        for k in range(args.steps):
            _, n, dt = cpu_throughput(wl, sets_per_core, pool, cores, 2000 + k)
            n_tot += n
            t_tot += dt
    value = n_tot / t_tot
    sample = f"{sets_per_core * cores} parameter sets x {n_opt} options per step ({sets_per_core} per core)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "oracle/cos_oracle.py price_scalar: the reference's per-option Python/NumPy-scalar "
                                 "algorithm (src/models/double_heston.py:160-192), one process per core"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, device: int):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[device]) if vis and vis.split(",")[device].isdigit() else device
            self._h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:       # noqa: BLE001
            self.error = repr(e)
        self._thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:        # noqa: BLE001
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.ok:
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self.ok:
            self._thread.join()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
def run_b200_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    # CPU baseline first (rank 0, N=1 only), before CUDA is initialised in this process
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        per_price = 6.1e-3 if wl["N"] <= 128 else 11.3e-3
        n_opt = len(wl["strikes"]) * len(wl["maturities"])
        sets_per_core = max(1, int(round(12.0 / (per_price * n_opt))))       # ~12 s of work per core
        with make_pool(cores) as pool:
            v, n, dt = cpu_throughput(wl, sets_per_core, pool, cores, 4242)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{sets_per_core * cores} seeded parameter sets x {n_opt} options "
                                  f"({n} prices, {dt:.1f} s wall) of the same workload",
                        "per_core": v / cores,
                        "note": "oracle/cos_oracle.py price_scalar = the reference's per-option algorithm "
                                "(src/models/double_heston.py:160-192), one process per host core"}
        try:     # best-effort CPU line: the same arithmetic restated in C (oracle/cos_oracle.c), OpenMP over options
            from oracle import cos_oracle as O
            lib = O.c_library()
            pc = gen_params(2500 * cores, 777)
            Kc = np.tile(np.array(wl["strikes"]), len(wl["maturities"]))
            Tc = np.repeat(np.array(wl["maturities"]), len(wl["strikes"]))
            O.c_price_batch(pc[:cores], 100.0, Kc, Tc, np.ones(Kc.size), wl["r"], 0.0, wl["N"])      # warm-up
            t0c = time.perf_counter()
            O.c_price_batch(pc, 100.0, Kc, Tc, np.ones(Kc.size), wl["r"], 0.0, wl["N"])
            dtc = time.perf_counter() - t0c
            cpu_baseline["c_port"] = {"value": pc.shape[0] * Kc.size / dtc, "unit": UNIT,
                                      "threads": int(lib.oracle_threads()), "kind": "port",
                                      "sample": f"{pc.shape[0]} parameter sets x {Kc.size} options, {dtc:.1f} s",
                                      "note": "NOT the reference: oracle/cos_oracle.c, the reference's arithmetic in C99 "
                                              "with OpenMP (a best-effort CPU implementation)"}
        except Exception as e:      # noqa: BLE001
            cpu_baseline["c_port"] = {"unavailable": repr(e)}

    import torch
    import torch.distributed as dist
    import dhj

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = dhj.Context(local)

    P, N, r = wl["P"], wl["N"], wl["r"]
    if wl.get("strong"):
        P = -(-P // world)                                   # this rank's block of the fixed total
    strikes, mats = np.array(wl["strikes"]), np.array(wl["maturities"])
    nK, nT = strikes.size, mats.size
    n_prices = P * nK * nT

    # synthetic inputs: pinned host copy (for the e2e arm) and HBM-resident copy (kernel arm)
    if wl.get("device_params"):
        # too large to stage through NumPy comfortably: uniform draws on the device (torch's Philox), same ranges
        gen = torch.Generator(device=dev); gen.manual_seed(7 + rank)
        lo_t = torch.tensor(PARAM_RANGES[:, 0], device=dev); hi_t = torch.tensor(PARAM_RANGES[:, 1], device=dev)
        d_big = torch.rand((P, 13), dtype=torch.float64, device=dev, generator=gen) * (hi_t - lo_t) + lo_t
        P_e2e = 1 << 20
        h_params = d_big[:P_e2e].cpu().pin_memory()
    else:
        d_big, P_e2e = None, P
        h_params = torch.from_numpy(gen_params(P, 20260101 + rank)).pin_memory()
    scaled = bool(wl.get("device_params"))                  # C4: per-set spot, strikes scale with it
    if scaled:
        d_s0_big = 100.0 * torch.exp(0.05 * torch.randn((P,), dtype=torch.float64, device=dev, generator=gen))
        h_s0 = d_s0_big[:P_e2e].cpu().pin_memory()
    else:
        d_s0_big = None
        h_s0 = torch.full((1,), 100.0, dtype=torch.float64).pin_memory()
    h_out = torch.empty((P_e2e, nT, nK), dtype=torch.float64).pin_memory()
    d_params = d_big if d_big is not None else h_params.to(dev)
    d_s0 = d_s0_big if d_s0_big is not None else h_s0.to(dev)
    d_out = torch.empty((P, nT, nK), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)        # > 126 MB L2
    # (checksum, count) of every rank: gathered once, after the timed steps (the path has no exchange step; a gather
    # inside the loop only measures how late NCCL's kernel gets an SM next to 112 k pricing blocks)
    gathered = torch.zeros((world, 2), dtype=torch.float64, device=dev)
    mine = torch.tensor([[0.0, float(n_prices)]], dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step_kernel():
        ctx.price_grid_dev(d_params.data_ptr(), P, d_s0.data_ptr(), 1 if scaled else 0, strikes, mats, r, 0.0, N, 10.0,
                           scaled, True, d_out.data_ptr(), stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # FP64 roofline denominator measured on this GPU
    fp64_peak, _ = ctx.fp64_peak(8192)

    # ---- kernel arm -------------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_kernel()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = ctx.launch_count
    with ClockSampler(local) as clocks:
        barrier()
        for e0, e1 in ev:
            flush.zero_()                      # L2 flush, outside the timed events
            e0.record()
            step_kernel()
            e1.record()
        barrier()
    if world > 1:                              # results stay sharded; their checksums travel over NVLink (NCCL)
        mine[0, 0].copy_(d_out.sum())
        dist.all_gather_into_tensor(gathered, mine)
    launches = ctx.launch_count - launches0
    ms_steps = [e0.elapsed_time(e1) for e0, e1 in ev]
    t_ms = torch.tensor([sum(ms_steps)], dtype=torch.float64, device=dev)
    per_rank_ms = [float(t_ms.item()) / args.steps]
    if world > 1:
        all_ms = torch.zeros(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(all_ms, t_ms)
        per_rank_ms = [float(v) / args.steps for v in all_ms.tolist()]
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    total_ms = float(t_ms.item())
    value = world * n_prices * args.steps / (total_ms * 1e-3)
    checksum = float(d_out.sum().item())

    # ---- end-to-end arm: host buffers through the C-ABI -------------------------------------------
    np_params, np_s0, np_out = h_params.numpy(), h_s0.numpy(), h_out.numpy()
    for _ in range(max(1, min(args.warmup, 3))):
        ctx.price_grid(np_params, np_s0, strikes, mats, r, 0.0, N, 10.0, scaled, True, out=np_out)
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.price_grid(np_params, np_s0, strikes, mats, r, 0.0, N, 10.0, scaled, True, out=np_out)
        _ = float(np_out[0, 0, 0])             # the result is on the host when the call returns
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * P_e2e * nK * nT * e2e_steps / float(e2e_s.item())
    e2e_match = bool(np.array_equal(np_out, d_out[:P_e2e].cpu().numpy()))

    if rank == 0:
        per_gpu = value / world
        fpp = flop_per_price(nK, N)
        achieved = per_gpu * fpp / 1e12
        kernel_ms = total_ms / args.steps
        bytes_alg = P * (13 * 8 + nK * nT * 8)
        traffic, ncu = None, None
        tpath = os.path.join(ROOT, "profiles", "ncu_summary.json")
        if os.path.exists(tpath):
            try:
                ncu = json.load(open(tpath)).get(args.workload)
                traffic = ncu.get("dram_bytes_per_launch") if ncu else None
            except Exception:      # noqa: BLE001
                traffic, ncu = None, None
        hbm_peak = None
        ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(ppath):
            hbm_peak = json.load(open(ppath)).get("hbm_gbs")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": kernel_ms, "higher_is_better": True,
            "scaling": "strong" if wl.get("strong") else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "sets_per_gpu": P, "options_per_set": nK * nT, "N": N,
                       "l2": "256 MiB buffer written between timed steps (untimed); inputs+outputs = "
                             f"{bytes_alg / 2**20:.0f} MiB per step",
                       "sharding": "by parameter set, one batch per rank, no data-path collective; NCCL all_gather of the ranks' checksums after the timed steps"
                                   if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(P_e2e * 13 * 8 + (P_e2e * 8 if scaled else 8)),
                    "d2h_bytes_per_step": int(P_e2e * nK * nT * 8), "steps": e2e_steps, "sets_per_step": P_e2e,
                    "api": "dhj.Context.price_grid -> dhj_price_grid (pinned host in/out)",
                    "bit_identical_to_kernel_arm": e2e_match},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak, "traffic": traffic,
                         "peak_source": "dhj_fp64_peak DFMA-chain probe on this GPU (MEASURED_PEAKS.json has no FP64 "
                                        "figure)", "nominal_peak": NOMINAL_FP64_TFLOPS,
                         "frac_of_nominal": achieved / NOMINAL_FP64_TFLOPS,
                         "flop_per_price": fpp, "kernel": "k_price_batch" if nK <= 8 else "k_price_dense",
                         "kernel_ms": kernel_ms,
                         "note": "achieved = prices/s x SURVEY §8d contract FLOP/price (the reference's formulas: one "
                                 "sincos per strike and k). The kernel contracts strikes by rotation recurrences and "
                                 "executes fewer flops, so frac can exceed 1; see `executed` and DESIGN.md §5",
                         "executed": None if not ncu else {
                             "flop_per_price": ncu["executed_flop_per_price"],
                             "tflops": per_gpu * ncu["executed_flop_per_price"] / 1e12,
                             "frac_of_peak": per_gpu * ncu["executed_flop_per_price"] / 1e12 / fp64_peak,
                             "ncu_fp64_pipe_active_pct": ncu["fp64_pipe_active_pct"],
                             "ncu_issue_active_pct": ncu["issue_active_pct"],
                             "source": "profiles/ncu_summary.json (ncu --set full of the same kernel)"},
                         "hbm": {"achieved_gbs": bytes_alg / (kernel_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                 "note": "compute-bound path: algorithmic bytes / kernel time, for information"}},
            "ms_per_step_per_rank": per_rank_ms,
            "clocks": clocks.summary(),
            "checksum": checksum, "checksum_all_ranks": (float(gathered[:, 0].sum().item()) if world > 1 else checksum),
        }
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        if args.workload == "c2":
            line["calibration"] = time_readme_calibration(ctx)
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def time_readme_calibration(ctx, repeats=5):
    """Second half of BASELINE.json's metric: seconds per 15-option calibration, the README configuration (C1):
    DoubleHestonJumpCalibrator(...).calibrate(maxiter=300, multi_start=3) through the drop-in class on the
    reference suite's market (tests/test_suite.py:274-302).  Reference: 425 s in the build container, 117.8 s
    published (README.md:173)."""
    for sub in ("models", "calibration"):
        sys.path.insert(0, os.path.join(ROOT, "option-pricing-ffn-lbfgs_b200", "src", sub))
    from lbfgs_calibrator import DoubleHestonJumpCalibrator
    from dhj import default_context
    cal_ctx = default_context()                                              # the context the drop-in classes use
    true_p = np.array([0.04, 2.0, 0.04, 0.3, -0.5, 0.04, 1.5, 0.04, 0.2, -0.3, 0.1, 0.0, 0.1])
    K = np.tile([90.0, 95.0, 100.0, 105.0, 110.0], 3)
    T = np.repeat([0.25, 0.5, 1.0], 5)
    mkt = ctx.price_list(true_p, 100.0, K, T, np.ones(15), 0.05)[0]
    opts = [{"strike": K[j], "maturity": T[j], "price": mkt[j], "option_type": "call"} for j in range(15)]
    cal = DoubleHestonJumpCalibrator(100.0, 0.05, opts)
    cal.calibrate(maxiter=3, multi_start=1)                                   # warm-up
    times, res = [], None
    for _ in range(repeats):
        np.random.seed(0)
        launches0 = cal_ctx.launch_count
        t0 = time.perf_counter()
        res = cal.calibrate(maxiter=300, multi_start=3)
        times.append(time.perf_counter() - t0)
        launches = cal_ctx.launch_count - launches0
    return {"metric": "s per 15-option calibration (maxiter=300, multi_start=3)", "value": float(np.median(times)),
            "min": float(min(times)), "unit": "s", "higher_is_better": False, "repeats": repeats,
            "final_loss": float(res.final_loss), "iterations": int(res.iterations), "gpu_launches": int(launches),
            "api": "DoubleHestonJumpCalibrator.calibrate -> dhj_loss_fd (one launch per optimiser round)",
            "reference_seconds": {"build_container": 425.0, "published_m1": 117.8}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, wl)
    else:
        run_b200_arm(args, wl)


if __name__ == "__main__":
    main()
