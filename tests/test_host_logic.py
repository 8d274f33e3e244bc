"""Host-side logic of the drop-in modules and of the sharding helpers (CPU; the GPU context is stubbed)."""
import os
import sys

import numpy as np
import pytest

from conftest import PKG, ROOT
from oracle import cos_oracle as O


@pytest.fixture()
def dropins(monkeypatch):
    """Import the drop-in modules the way the reference's users do (top-level names via sys.path)."""
    import dhj
    import dhj._native as native

    class NoDevice:
        def __getattr__(self, name):
            raise AssertionError(f"host-logic test touched the device ({name})")

    monkeypatch.setattr(native, "default_context", lambda: NoDevice())
    monkeypatch.setattr(dhj, "default_context", lambda: NoDevice())
    for sub in ("models", "calibration", "data"):
        monkeypatch.syspath_prepend(os.path.join(PKG, "src", sub))
    for name in ("double_heston", "lbfgs_calibrator", "synthetic_generator"):
        sys.modules.pop(name, None)
    import double_heston
    import lbfgs_calibrator
    import synthetic_generator
    yield double_heston, lbfgs_calibrator, synthetic_generator
    for name in ("double_heston", "lbfgs_calibrator", "synthetic_generator"):
        sys.modules.pop(name, None)


def c1_options(g):
    return [{"strike": float(k), "maturity": float(t), "price": float(p), "option_type": "call"}
            for k, t, p in zip(g["strike"], g["maturity"], g["market"])]


def test_signatures_match_reference(dropins):
    import inspect
    dh, cal, gen = dropins
    sig = inspect.signature(dh.DoubleHeston.__init__)
    assert list(sig.parameters) == ["self", "S0", "K", "T", "r", "v01", "kappa1", "theta1", "sigma1", "rho1", "v02",
                                    "kappa2", "theta2", "sigma2", "rho2", "lambda_j", "mu_j", "sigma_j",
                                    "option_type", "q"]
    assert sig.parameters["option_type"].default == "C" and sig.parameters["q"].default == 0.0
    assert inspect.signature(dh.DoubleHeston.pricing).parameters["N"].default == 128
    assert inspect.signature(dh.DoubleHeston.truncationRange).parameters["L"].default == 10
    for m in ("characteristic_function", "chi_k", "psi_k"):
        assert hasattr(dh.DoubleHeston, m)
    c = inspect.signature(cal.DoubleHestonJumpCalibrator.calibrate)
    assert c.parameters["maxiter"].default == 300 and c.parameters["multi_start"].default == 3
    assert list(inspect.signature(cal.DoubleHestonJumpCalibrator.__init__).parameters) == \
        ["self", "spot", "risk_free_rate", "market_options"]
    g = inspect.signature(gen.generate_synthetic_calibrations)
    assert g.parameters["n_samples"].default == 500
    assert g.parameters["save_path"].default == "lbfgs_calibrations_synthetic.pkl"
    fields = [f.name for f in cal.CalibrationResult.__dataclass_fields__.values()]
    assert fields == ["date", "spot", "risk_free", "parameters", "market_prices", "model_prices", "market_options",
                      "final_loss", "calibration_time", "success", "iterations", "message"]
    assert cal.CalibrationResult.__module__ == "lbfgs_calibrator"       # pickles carry the top-level name


def test_package_style_import(monkeypatch):
    """README spelling of the reference: from src.calibration.lbfgs_calibrator import ... (README.md:64)."""
    import dhj
    import dhj._native as native
    monkeypatch.setattr(native, "default_context", lambda: None)
    monkeypatch.setattr(dhj, "default_context", lambda: None)
    monkeypatch.syspath_prepend(PKG)
    for name in list(sys.modules):
        if name == "src" or name.startswith("src.") or name in ("double_heston", "lbfgs_calibrator",
                                                                "synthetic_generator"):
            sys.modules.pop(name)
    from src.calibration.lbfgs_calibrator import DoubleHestonJumpCalibrator, CalibrationResult   # noqa: F401
    from src.models.double_heston import DoubleHeston                                              # noqa: F401
    from src.data.synthetic_generator import generate_synthetic_calibrations                       # noqa: F401
    for name in list(sys.modules):
        if name == "src" or name.startswith("src.") or name in ("double_heston", "lbfgs_calibrator",
                                                                "synthetic_generator"):
            sys.modules.pop(name)


def test_transforms_and_guesses_match_reference(dropins, golden):
    _, cal, _ = dropins
    g = golden("initial_guess.npz")
    c = cal.DoubleHestonJumpCalibrator(float(g["spot"]), float(g["r"]), c1_options(g))
    assert c.n_calls == 0 and c.best_loss == np.inf and len(c.param_names) == 13
    assert np.array_equal(c.market_prices, g["market"])
    np.random.seed(0)
    assert np.array_equal(c.get_initial_guess(0), g["g0"])
    assert np.array_equal(c.get_initial_guess(1), g["g1"])
    assert np.array_equal(c.get_initial_guess(2), g["g2"])
    assert np.array_equal(c.get_initial_guess(1), g["g1_second_draw"])
    x = g["g1"]
    p = c.transform_params(x)
    assert list(p) == c.param_names
    assert np.array_equal(np.array([p[n] for n in c.param_names]), O.transform_params(x))
    assert np.allclose(c.inverse_transform_params(p), x, rtol=1e-14, atol=1e-15)
    assert c.compute_feller_penalty(p) == O.feller_penalty(O.transform_params(x))
    big = dict(p); big["sigma1"] = 5.0
    second = max(0, p["sigma2"] ** 2 - 2 * p["kappa2"] * p["theta2"])
    assert c.compute_feller_penalty(big) == 1000.0 * ((25.0 - 2 * p["kappa1"] * p["theta1"]) + second)
    nanp = dict(p); nanp["sigma1"] = float("nan")
    assert c.compute_feller_penalty(nanp) == 1000.0 * second     # Python max(0, nan) == 0


def test_generator_host_recurrence(dropins, golden):
    _, _, gen = dropins
    g = golden("generator_seed42.npz")
    np.random.seed(42)
    names, params, spots, noise = gen._draw_inputs(20)
    assert names == list(O.PARAM_NAMES)
    assert np.array_equal(params, g["params"]) and np.array_equal(spots, g["spots"])
    assert np.array_equal(g["model_prices"] + noise * g["model_prices"], g["market_prices"])
    assert gen._trading_dates(20) == list(g["dates"])
    assert gen._trading_dates(0) == []
    assert [gen._trading_date(i) for i in range(20)] == list(g["dates"])


def test_generator_draws_match_numpy_stream(dropins):
    """dhj_generator_draws (csrc/dhj_draws.cpp) against the reference's own per-sample np.random calls
    (synthetic_generator.py:98-142): same values bit for bit, same generator state afterwards — also when the
    state carries a cached gaussian, across an MT19937 refill, and for n = 0 / 1."""
    _, _, gen = dropins
    lo = np.array([v[0] for v in gen.PARAM_RANGES.values()])
    hi = np.array([v[1] for v in gen.PARAM_RANGES.values()])

    def numpy_stream(n):
        params, spots, noise = np.empty((n, 13)), np.empty(n), np.empty((n, 15))
        for i in range(n):
            fresh = np.array([np.random.uniform(a, b) for a, b in zip(lo, hi)])      # 13 scalar draws, as the reference
            if i > 0:
                fresh = 0.9 * params[i - 1] + (1 - 0.9) * fresh
                spots[i] = spots[i - 1] * (1 + np.random.normal(0.0003, 0.01))
            else:
                spots[i] = 100.0
            params[i] = fresh
            noise[i] = [np.random.normal(0, 0.02) for _ in range(15)]
        return params, spots, noise

    for seed, n, odd in ((42, 20, False), (7, 400, True), (3, 1, True), (5, 0, False)):
        np.random.seed(seed)
        if odd:
            np.random.normal()                      # leaves the pair's second variate cached in the state
        start = np.random.get_state()
        want = numpy_stream(n)
        end = np.random.get_state()
        np.random.set_state(start)
        _, *got = gen._draw_inputs(n)
        now = np.random.get_state()
        assert all(np.array_equal(w, g) for w, g in zip(want, got)), (seed, n)
        assert np.array_equal(end[1], now[1]) and end[2:] == now[2:]
        assert np.random.random() == (np.random.set_state(end), np.random.random())[1]


def test_flat_dataset_sinks_and_lazy_view(dropins, tmp_path):
    """SURVEY §8f N1 on the host side: .npz / directory-of-.npy sinks and the lazy CalibrationResult view
    (no pricing involved: the arrays are made up)."""
    _, cal, gen = dropins
    n = 12
    rng = np.random.default_rng(3)
    data = {"param_names": list(gen.PARAM_RANGES), "params": rng.uniform(0.01, 1.0, (n, 13)),
            "spots": 100.0 + rng.standard_normal(n), "strikes": rng.uniform(80, 120, (n, 15)),
            "maturities": np.repeat(gen.MATURITIES, 5), "model_prices": rng.uniform(1, 20, (n, 15)),
            "market_prices": rng.uniform(1, 20, (n, 15)), "losses": rng.uniform(0, 1e-3, n)}
    gen._save_arrays(data, tmp_path / "flat")
    gen._save_arrays(data, str(tmp_path / "flat.npz"))
    for ds, mapped in ((gen.SyntheticCalibrationSet.load(tmp_path / "flat"), True),
                       (gen.SyntheticCalibrationSet.load(str(tmp_path / "flat.npz")), False),
                       (gen.SyntheticCalibrationSet(data), False)):
        assert len(ds) == n and isinstance(ds.data["params"], np.memmap) == mapped
        r = ds[5]
        assert isinstance(r, cal.CalibrationResult) and r.date == gen._trading_dates(6)[5]
        assert r.spot == data["spots"][5] and r.risk_free == gen.RISK_FREE and r.final_loss == data["losses"][5]
        assert list(r.parameters) == list(gen.PARAM_RANGES) and r.parameters["kappa1"] == data["params"][5, 1]
        assert np.array_equal(r.market_prices, data["market_prices"][5]) and len(r.market_options) == 15
        assert r.market_options[7] == {"strike": data["strikes"][5, 7], "maturity": 0.5,
                                       "price": data["market_prices"][5, 7], "option_type": "call"}
        assert r.calibration_time is None and r.iterations is None and r.success is True
        assert ds[-1].date == gen._trading_dates(n)[-1]
        sub = ds[3:9]
        assert len(sub) == 6 and sub[2].date == r.date and sub[2].spot == r.spot       # dates follow the slice
        assert [x.spot for x in ds] == list(data["spots"])
        with pytest.raises(IndexError):
            ds[n]
        # NumPy integer indices (np.random.permutation / np.arange: the usual training access pattern)
        assert ds[np.int64(5)].spot == r.spot and ds[np.int32(-1)].date == ds[-1].date
        for idx in np.random.default_rng(0).permutation(n)[:4]:
            assert ds[idx].spot == data["spots"][idx]
    # a sliced view written as .npz / directory keeps its offset (dates continue) and indexes with NumPy integers
    sub = gen.SyntheticCalibrationSet(data)[4:10]
    sub.save(str(tmp_path / "sub.npz"))
    sub.save(tmp_path / "sub_dir")
    for back in (gen.SyntheticCalibrationSet.load(str(tmp_path / "sub.npz")),
                 gen.SyntheticCalibrationSet.load(tmp_path / "sub_dir")):
        assert len(back) == 6 and back[np.int64(1)].date == gen._trading_dates(6)[5]
        assert back[1:][0].spot == data["spots"][5]
    # counter-stream datasets: the date is the step inside the history
    hist = dict(data, _first=495, _path_len=500)
    hs = gen.SyntheticCalibrationSet(hist)
    assert hs[4].date == gen._trading_dates(500)[499] and hs[5].date == "2022-01-03" and hs[6:][0].date == "2022-01-04"


def test_lockstep_evaluator_batches_requests(dropins):
    """The multi-start driver answers one request per running optimiser with ONE launch, also when
    optimisers retire at different times."""
    import threading
    _, cal, _ = dropins

    class FakeMarket:
        def __init__(self):
            self.calls = []

        def loss_fd(self, xs, h, want_all=False):
            xs = np.asarray(xs)
            self.calls.append(xs.shape[0])
            f = (xs ** 2).sum(axis=1)
            return f, 2 * xs, np.tile(f[:, None], (1, 14))

    mk = FakeMarket()
    ev = cal._LockstepEvaluator(mk, 3)
    rounds = {0: 2, 1: 5, 2: 3}
    got = {}

    def worker(i):
        acc = []
        for k in range(rounds[i]):
            f, g, f_all = ev.request(i, np.full(13, float(i + k)))
            acc.append(f)
        ev.retire(i)
        got[i] = acc

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(3)]
    [t.start() for t in ts]
    [t.join(10) for t in ts]
    assert all(not t.is_alive() for t in ts)
    assert mk.calls == [3, 3, 2, 1, 1]                       # 5 launches instead of 10
    for i in range(3):
        assert got[i] == [13.0 * (i + k) ** 2 for k in range(rounds[i])]


def test_shard_bounds():
    from dhj.shard import shard_bounds
    for n in (0, 1, 7, 8, 9, 1000003):
        for w in (1, 2, 3, 8):
            blocks = [shard_bounds(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in blocks) == (-(-n // w) if n else 0)
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_history_shards_cover_the_dataset():
    """Counter-stream datasets are split by whole histories: the ranks' ranges are disjoint, ordered, cover [0, n) and cut
    only at history boundaries (except the dataset's end)."""
    from dhj.shard import history_shard
    for n, path_len in ((100_000_000, 500), (1500, 500), (1499, 500), (7, 500), (0, 500), (1000, 1), (12345, 7)):
        for w in (1, 2, 3, 8):
            blocks = [history_shard(n, path_len, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            assert all(lo % path_len == 0 for lo, hi in blocks if lo < n)
            assert all(hi % path_len == 0 or hi == n for lo, hi in blocks)
    assert history_shard(100_000_000, 500, 8, 3) == (37_500_000, 50_000_000)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, PKG)
    from dhj.shard import price_grid_sharded, shard_bounds
    from oracle import cos_oracle as O2
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    P = 11                                                    # ragged: 6 + 5
    params = rng.uniform(O2.GENERATOR_RANGES[:, 0], O2.GENERATOR_RANGES[:, 1], size=(P, 13))
    spots = rng.uniform(90, 110, size=P)
    K = np.tile(O2.GENERATOR_STRIKES_REL, 3); T = np.repeat(O2.GENERATOR_MATURITIES, 5)

    def fake_ctx_price_grid(p, s0, strikes, mats, r):       # stands in for ctx.price_grid (tests may use the oracle)
        kk = np.tile(strikes[None, :] * s0[:, None] / 100.0, (1, len(mats)))
        return O2.price_batch(p, s0, kk, np.repeat(mats, len(strikes)), np.ones(kk.shape[1]), r).reshape(len(p), len(mats), -1)

    full = price_grid_sharded(fake_ctx_price_grid, params, spots, O2.GENERATOR_STRIKES_REL, O2.GENERATOR_MATURITIES, 0.03)
    local, (lo, hi) = price_grid_sharded(fake_ctx_price_grid, params, spots, O2.GENERATOR_STRIKES_REL,
                                         O2.GENERATOR_MATURITIES, 0.03, gather=False)
    want = fake_ctx_price_grid(params, spots, O2.GENERATOR_STRIKES_REL, O2.GENERATOR_MATURITIES, 0.03)
    # the dataset sweep's sharding (C4): every rank draws ITS histories of the counter stream (the oracle's restatement
    # stands in for the device here), the gathered union equals the stream drawn in one piece
    from dhj.shard import gather_rows, history_shard
    n_ds, path_len = 1300, 500                                # 3 histories over 2 ranks: 1000 + 300 samples
    s_lo, s_hi = history_shard(n_ds, path_len, world, rank)
    mine = O2.counter_draws(11, s_lo, s_hi - s_lo, path_len)[0]
    per = -(-3 // world) * path_len                           # padded block of the gather
    padded = np.zeros((per, 13)); padded[:mine.shape[0]] = mine
    gathered = gather_rows(padded, per * world)
    whole = O2.counter_draws(11, 0, n_ds, path_len)[0]
    union = np.concatenate([gathered[r * per:r * per + (hi_r - lo_r)]
                            for r in range(world) for lo_r, hi_r in [history_shard(n_ds, path_len, world, r)]])
    q.put((rank, bool(np.array_equal(full, want)), bool(np.array_equal(local, want[lo:hi])),
           (lo, hi) == shard_bounds(P, world, rank) and bool(np.array_equal(union, whole))))
    dist.destroy_process_group()


def test_sharded_pricing_gloo_world2():
    """world_size-2 gloo run of the N>1 paths: block sharding by parameter set + all_gather of prices, and the dataset
    sweep's sharding by whole histories of the counter stream."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(30) for p in procs]
    assert res == [(0, True, True, True), (1, True, True, True)]


def test_vectorised_initial_guesses_match_calibrator(dropins):
    """dhj.initial_guesses (many markets at once) = get_initial_guess market by market, same global RNG stream."""
    _, cal, _ = dropins
    from dhj.calibrate_many import initial_guesses
    rng = np.random.default_rng(0)
    n = 7
    spots = rng.uniform(90, 110, n)
    K = np.tile(np.array([90.0, 95, 100, 105, 110])[None, :] * spots[:, None] / 100, (1, 3))
    K[3] *= 1.5                                             # a market without ATM options: implied_var falls back to 0.04
    T = np.repeat([0.25, 0.5, 1.0], 5)
    prices = rng.uniform(2, 15, (n, 15))
    np.random.seed(5)
    got = initial_guesses(spots, K, T, prices, 5)
    np.random.seed(5)
    for i in range(n):
        c = cal.DoubleHestonJumpCalibrator(spots[i], 0.03, [
            {"strike": K[i, j], "maturity": T[j], "price": prices[i, j], "option_type": "call"} for j in range(15)])
        for s in range(5):
            assert np.array_equal(got[i, s], c.get_initial_guess(s % 3)), (i, s)
