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

For datasets too large for a list of Python objects use `generate_synthetic_arrays`, which returns /
fills flat arrays.
"""
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
from dhj import default_context  # noqa: E402

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
    """Host recurrence: parameters (AR(1)-smoothed), spots (random walk) and price-noise factors.

    Three global-RNG calls per sample instead of the reference's 29 scalar ones, consuming the stream in the
    same order (13 uniforms, [1 normal], 15 normals): bit-identical values."""
    names = list(PARAM_RANGES)
    lo = np.array([v[0] for v in PARAM_RANGES.values()])
    hi = np.array([v[1] for v in PARAM_RANGES.values()])
    n_opt = MATURITIES.size * STRIKES.size
    params = np.empty((n, len(names)))
    spots = np.empty(n)
    noise = np.empty((n, n_opt))
    for i in range(n):
        fresh = np.random.uniform(lo, hi)
        if i > 0:
            fresh = PERSISTENCE * params[i - 1] + (1 - PERSISTENCE) * fresh
            spots[i] = spots[i - 1] * (1 + np.random.normal(0.0003, 0.01))
        else:
            spots[i] = SPOT_BASE
        params[i] = fresh
        noise[i] = np.random.normal(0, 0.02, size=n_opt)
    return names, params, spots, noise


def generate_synthetic_arrays(n_samples, ctx=None, save_path=None):
    """Flat-array form of the generator: dict of params[n,13], spots[n], strikes[n,15], maturities[15],
    model_prices[n,15], market_prices[n,15], losses[n] (same values as the object form).  With `save_path` the
    arrays are also written as one .npz — the on-disk form for datasets too large for a pickle of Python
    objects."""
    ctx = ctx or default_context()
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
        np.savez(save_path, **{k: np.asarray(v) for k, v in data.items()})
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
