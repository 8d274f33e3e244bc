"""Drop-in replacement for the reference's `src/data/synthetic_generator.py`: synthetic calibration data
for FFN training, with all option prices computed in ONE batched launch on a B200.

IMPORTANT (as in the reference): this produces SYNTHETIC data — randomly sampled parameters, synthetic
noisy "market" prices, no L-BFGS optimisation; `calibration_time` and `iterations` are None.

`generate_synthetic_calibrations(n_samples=500, save_path='lbfgs_calibrations_synthetic.pkl')` keeps the
reference's signature, return type (list of `CalibrationResult`), pickle format and — for seeded runs —
its values: the global NumPy RNG is consumed in the reference's order (per sample: 13 uniforms in the
order of the parameter ranges, one normal for the spot return when i > 0, then 15 normals of price
noise; /root/reference/src/data/synthetic_generator.py:98-142).  The draws do not depend on the prices,
so they are hoisted out of the pricing loop and the n x 15 prices are then computed by one call of
`dhj.Context.price_grid` (strikes scaled by spot as in :125).

The draw stream itself (sequential: one MT19937 generator, AR(1) smoothing, spot walk) runs in the library
too (`dhj_generator_draws`, bit-identical to the per-sample `np.random` calls): 0.7 s per million samples
instead of 19 s of Python loop.

For datasets too large for a list of Python objects: `generate_synthetic_arrays` returns flat
arrays (optionally written as .npz or as a directory of .npy files that can be memory-mapped), and
`SyntheticCalibrationSet` presents such arrays as a lazy sequence of `CalibrationResult` objects.

For sweeps that must not depend on one sequential host stream (BASELINE config 4: 100 M samples sharded over
1/2/4/8 GPUs) `generate_synthetic_arrays(n, seed=...)` switches to the library's COUNTER STREAM
(`dhj_generate`, csrc/dhj_generate.cuh): parameters, spots and noise are drawn ON THE DEVICE as Philox functions of
(seed, sample index), priced, noised and reduced to the per-sample loss there; samples are grouped in independent
reference-style histories of `path_len` steps.  With `sharded=True` every rank of the process group produces and
writes its own block of histories (`shard-RRRRR-of-WWWWW/` under `save_path`), and the union of the shards is the
same dataset whatever the number of ranks.
"""
import json
import operator
import pickle
import sys
from datetime import datetime, timedelta
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).parent.parent / 'calibration'))
sys.path.insert(0, str(Path(__file__).parent.parent / 'models'))
sys.path.insert(0, str(Path(__file__).parent.parent.parent))

from lbfgs_calibrator import CalibrationResult  # noqa: E402
from double_heston import DoubleHeston  # noqa: E402,F401
from dhj import default_context, generator_draws  # noqa: E402

# parameter ranges, in RNG draw order (synthetic_generator.py:75-89)
PARAM_RANGES = {
    'v1_0': (0.025, 0.080), 'kappa1': (1.5, 4.5), 'theta1': (0.025, 0.065), 'sigma1': (0.20, 0.50),
    'rho1': (-0.85, -0.40),
    'v2_0': (0.020, 0.070), 'kappa2': (0.30, 1.20), 'theta2': (0.025, 0.070), 'sigma2': (0.10, 0.35),
    'rho2': (-0.70, -0.20),
    'lambda_j': (0.05, 0.25), 'mu_j': (-0.08, -0.01), 'sigma_j': (0.03, 0.12),
}
STRIKES = np.array([90, 95, 100, 105, 110])     # % of spot (:91)
MATURITIES = np.array([0.25, 0.5, 1.0])         # (:92)
SPOT_BASE = 100.0
RISK_FREE = 0.03
PERSISTENCE = 0.9                               # AR(1) coefficient (:107)


def _trading_dates(n):
    """n consecutive weekdays from 2022-01-03 (:59-67)."""
    day, out = datetime(2022, 1, 3), []
    while len(out) < n:
        if day.weekday() < 5:
            out.append(day.strftime('%Y-%m-%d'))
        day += timedelta(days=1)
    return out


def _draw_inputs(n):
    """Host recurrence: parameters (AR(1)-smoothed), spots (random walk) and price-noise factors, drawn from the
    global NumPy RNG in the reference's order (per sample 13 uniforms, [1 normal], 15 normals) by the library's
    restatement of NumPy's legacy generator: bit-identical values and final RNG state (SURVEY §8a row 15;
    tests/test_host_logic.py compares it with the per-sample np.random calls)."""
    names = list(PARAM_RANGES)
    lo = np.array([v[0] for v in PARAM_RANGES.values()])
    hi = np.array([v[1] for v in PARAM_RANGES.values()])
    params, spots, noise = generator_draws(n, lo, hi, PERSISTENCE, SPOT_BASE, 0.0003, 0.01, 0.02,
                                           MATURITIES.size * STRIKES.size)
    return names, params, spots, noise


_ARRAY_KEYS = ('params', 'spots', 'strikes', 'maturities', 'model_prices', 'market_prices', 'losses')


def _save_arrays(data, save_path):
    """`x.npz` -> one archive; anything else -> a directory of .npy files (memory-mappable)."""
    save_path = str(save_path)
    if save_path.endswith('.npz'):
        np.savez(save_path, **{k: np.asarray(v) for k, v in data.items()})
        return
    Path(save_path).mkdir(parents=True, exist_ok=True)
    for k in _ARRAY_KEYS:
        np.save(Path(save_path) / f'{k}.npy', data[k])
    np.save(Path(save_path) / 'param_names.npy', np.asarray(data['param_names']))
    for k in ('_first', '_path_len'):
        if k in data:
            np.save(Path(save_path) / f'{k}.npy', np.asarray(int(data[k])))


class SyntheticCalibrationSet:
    """Lazy sequence view of a flat synthetic dataset: `len(ds)`, `ds[i]` -> `CalibrationResult` built on demand
    (same fields as the objects `generate_synthetic_calibrations` returns), slicing -> another view.  Holds only
    the arrays — O(1) Python objects however many samples; with `load(path, mmap=True)` not even those."""

    def __init__(self, data):
        self.data = data
        self.param_names = [str(s) for s in data['param_names']]

    @classmethod
    def load(cls, path, mmap=True):
        path = str(path)
        if path.endswith('.npz'):
            z = np.load(path)
            return cls({k: (int(z[k]) if k.startswith('_') else z[k]) for k in z.files})
        d = {k: np.load(Path(path) / f'{k}.npy', mmap_mode='r' if mmap else None) for k in _ARRAY_KEYS}
        d['param_names'] = np.load(Path(path) / 'param_names.npy')
        for k in ('_first', '_path_len'):
            if (Path(path) / f'{k}.npy').exists():
                d[k] = int(np.load(Path(path) / f'{k}.npy'))
        return cls(d)

    def save(self, path):
        _save_arrays(self.data, path)

    def __len__(self):
        return int(self.data['spots'].shape[0])

    def _history_step(self, i):
        """Trading-day number of local sample i: its global index, modulo the history length for counter-stream
        datasets (every history starts on 2022-01-03 like the reference's single one)."""
        g = int(self.data.get('_first', 0)) + i
        path_len = int(self.data.get('_path_len', 0))
        return g % path_len if path_len > 0 else g

    def __getitem__(self, i):
        d = self.data
        if isinstance(i, slice):
            sub = {k: (d[k][i] if k != 'maturities' else d[k]) for k in _ARRAY_KEYS}
            sub['param_names'] = d['param_names']
            sub['_first'] = int(d.get('_first', 0)) + (i.indices(len(self))[0] if len(self) else 0)
            if '_path_len' in d:
                sub['_path_len'] = int(d['_path_len'])
            assert i.step in (None, 1), "contiguous slices only"
            return SyntheticCalibrationSet(sub)
        n = len(self)
        i = operator.index(i)                 # NumPy integer indices (np.random.permutation, np.arange) included
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError(i)
        mats = d['maturities']
        options = [{'strike': d['strikes'][i, j], 'maturity': mats[j], 'price': d['market_prices'][i, j],
                    'option_type': 'call'} for j in range(mats.size)]
        return CalibrationResult(
            date=_trading_date(self._history_step(i)), spot=d['spots'][i], risk_free=RISK_FREE,
            parameters={name: d['params'][i, k] for k, name in enumerate(self.param_names)},
            market_prices=np.array(d['market_prices'][i]), model_prices=np.array(d['model_prices'][i]),
            market_options=options, final_loss=d['losses'][i],
            calibration_time=None, success=True, iterations=None,
            message='Synthetic data (not from real calibration)')

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def _trading_date(i):
    """The i-th weekday on or after 2022-01-03 (a Monday), in closed form (:59-67).  Python's datetime ends in
    year 9999 (sample ~2.08 M; the reference's own date loop raises there): later samples are labelled by index."""
    try:
        return (datetime(2022, 1, 3) + timedelta(days=7 * (i // 5) + i % 5)).strftime('%Y-%m-%d')
    except OverflowError:
        return f'2022-01-03+{i}bd'


def generate_synthetic_arrays(n_samples, ctx=None, save_path=None, *, seed=None, path_len=500, first=0,
                              sharded=False, group=None):
    """Flat-array form of the generator: dict of params[n,13], spots[n], strikes[n,15], maturities[15],
    model_prices[n,15], market_prices[n,15], losses[n] (same values as the object form).  With `save_path` the
    arrays are also written — `*.npz`: one archive; any other path: a directory of .npy files that
    `SyntheticCalibrationSet.load` memory-maps — the on-disk form for datasets too large for a pickle of Python
    objects.

    `seed=None` (default): the reference's stream — the global NumPy RNG, one history of n samples, bit-compatible
    with `generate_synthetic_calibrations` (prices on the device, draws / noise / loss on the host).
    `seed=<int>`: the library's counter stream (`dhj_generate`): samples [first, first + n) of the dataset `seed`,
    drawn, priced, noised and reduced to losses on the device; independent histories of `path_len` samples.
    `sharded=True` (counter stream only): this rank's block of whole histories of the n-sample dataset
    (`dhj.shard.shard_bounds` over histories), written to `save_path/shard-RRRRR-of-WWWWW/` when `save_path` is
    given; the returned dict carries `_first` (global index of its first sample).  Without a process group the
    single shard is the whole dataset."""
    ctx = ctx or default_context()
    if seed is None:
        if sharded or first:
            raise ValueError("the reference's NumPy stream is sequential: `first` / `sharded` need `seed=`")
        return _generate_numpy_stream(n_samples, ctx, save_path)
    n_samples, path_len, first = int(n_samples), int(path_len), int(first)
    rank = world = None
    if sharded:
        from dhj.shard import _dist, history_shard
        dist = _dist()
        world = dist.get_world_size(group) if dist else 1
        rank = dist.get_rank(group) if dist else 0
        lo_i, hi_i = history_shard(n_samples, path_len, world, rank)
        first, n_samples = first + lo_i, hi_i - lo_i
    lo = np.array([v[0] for v in PARAM_RANGES.values()])
    hi = np.array([v[1] for v in PARAM_RANGES.values()])
    g = ctx.generate(seed, first, n_samples, path_len, lo, hi, PERSISTENCE, SPOT_BASE, 0.0003, 0.01, 0.02,
                     STRIKES.astype(np.float64), MATURITIES, RISK_FREE)
    spots = g['spots']
    data = {'param_names': list(PARAM_RANGES), 'params': g['params'], 'spots': spots,
            'strikes': np.tile(STRIKES[None, :] * spots[:, None] / 100.0, (1, MATURITIES.size)),
            'maturities': np.repeat(MATURITIES, STRIKES.size), 'model_prices': g['model'],
            'market_prices': g['market'], 'losses': g['loss'], '_first': first, '_path_len': path_len}
    if save_path is not None:
        if sharded:
            shard_dir = Path(save_path) / f'shard-{rank:05d}-of-{world:05d}'
            _save_arrays(data, shard_dir)
            with open(shard_dir / 'manifest.json', 'w') as f:
                json.dump({'seed': int(seed), 'first': first, 'n': n_samples, 'path_len': path_len, 'rank': rank,
                           'world': world, 'stream': 'counter (Philox4x32-10, csrc/dhj_generate.cuh)'}, f)
        else:
            _save_arrays(data, save_path)
    return data


def load_sharded(save_path):
    """The shards a `generate_synthetic_arrays(..., sharded=True, save_path=...)` run wrote, in dataset order:
    list of memory-mapped `SyntheticCalibrationSet`s (each knows the global index of its first sample)."""
    dirs = sorted(Path(save_path).glob('shard-*-of-*'))
    return [SyntheticCalibrationSet.load(d, mmap=True) for d in dirs]


def _generate_numpy_stream(n_samples, ctx, save_path):
    names, params, spots, noise = _draw_inputs(n_samples)
    model = ctx.price_grid(params, spots, STRIKES.astype(np.float64), MATURITIES, RISK_FREE,
                           scale_by_spot=True, is_call=True).reshape(n_samples, MATURITIES.size * STRIKES.size)
    market = model + noise * model                                              # (:141-142)
    rel = (model - market) / market
    strikes = np.tile(STRIKES[None, :] * spots[:, None] / 100.0, (1, MATURITIES.size))
    data = {'param_names': names, 'params': params, 'spots': spots, 'strikes': strikes,
            'maturities': np.repeat(MATURITIES, STRIKES.size), 'model_prices': model,
            'market_prices': market, 'losses': np.mean(rel ** 2, axis=1)}
    if save_path is not None:
        _save_arrays(data, save_path)
    return data


def generate_synthetic_calibrations(n_samples: int = 500,
                                    save_path: str = 'lbfgs_calibrations_synthetic.pkl'):
    """Generate `n_samples` synthetic CalibrationResult objects, pickle them to `save_path`, return them."""
    bar = "=" * 70
    print(f"{bar}\nGENERATING SYNTHETIC HISTORICAL CALIBRATIONS (B200 batched pricing)\n{bar}")
    print("WARNING: synthetic data — sampled parameters, noisy model prices, no optimisation;")
    print("         timing/iteration fields are None.  For FFN training only, not for benchmarks.")
    print(f"  calibrations: {n_samples}   save path: {save_path}")

    dates = _trading_dates(n_samples)
    data = generate_synthetic_arrays(n_samples)
    names = data['param_names']

    calibrations = []
    for i, date in enumerate(dates):
        spot = data['spots'][i]
        options = [{'strike': data['strikes'][i, j], 'maturity': data['maturities'][j],
                    'price': data['market_prices'][i, j], 'option_type': 'call'}
                   for j in range(data['maturities'].size)]
        calibrations.append(CalibrationResult(
            date=date, spot=spot, risk_free=RISK_FREE,
            parameters={name: data['params'][i, k] for k, name in enumerate(names)},
            market_prices=data['market_prices'][i].copy(), model_prices=data['model_prices'][i].copy(),
            market_options=options, final_loss=data['losses'][i],
            calibration_time=None, success=True, iterations=None,
            message='Synthetic data (not from real calibration)'))

    print(f"Saving to {save_path}...")
    with open(save_path, 'wb') as f:
        pickle.dump(calibrations, f)

    if n_samples:
        losses, spots = data['losses'], data['spots']
        print(f"{bar}\nGENERATION COMPLETE\n{bar}")
        print(f"  total {len(calibrations)}; loss mean {np.mean(losses):.6f} median {np.median(losses):.6f} "
              f"min {np.min(losses):.6f} max {np.max(losses):.6f}")
        print(f"  spot start ${spots[0]:.2f} end ${spots[-1]:.2f} ({(spots[-1] / spots[0] - 1) * 100:+.2f}%) "
              f"min ${spots.min():.2f} max ${spots.max():.2f}")
        for k, name in enumerate(names):
            col = data['params'][:, k]
            print(f"  {name:10s}: mean={col.mean():.4f}, std={col.std():.4f}, min={col.min():.4f}, max={col.max():.4f}")
        err = np.abs((data['model_prices'] - data['market_prices']) / data['market_prices']).ravel() * 100
        print(f"  pricing error: mean {err.mean():.2f}% median {np.median(err):.2f}% "
              f"p95 {np.percentile(err, 95):.2f}% max {err.max():.2f}%")
    print(f"Synthetic calibrations saved to: {save_path}")
    return calibrations


if __name__ == "__main__":
    n_samples = 500
    if len(sys.argv) > 1:
        try:
            n_samples = int(sys.argv[1])
        except ValueError:
            print(f"Invalid number of samples: {sys.argv[1]}")
            print("Usage: python synthetic_generator.py [n_samples]")
            sys.exit(1)
    generate_synthetic_calibrations(n_samples=n_samples, save_path='lbfgs_calibrations_synthetic.pkl')
