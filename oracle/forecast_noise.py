"""TEST INFRASTRUCTURE -- numpy restatement of the reference's GaussianNoiseForecaster.

Only tests/ may import this.  It follows the reference line by line, INCLUDING its use of the legacy global numpy RNG
(`np.random.normal`), so that with the same `np.random.seed` it reproduces the reference's noisy observations bit for bit
(tests/golden/noisy_forecast.npz pins that).  The CUDA path draws from a counter-based generator instead, so it is
compared with this restatement in distribution only ("distributional parity only").

reference: src/pymgrid/forecast/forecaster.py
  :91-170   Forecaster: fill value (high + low) / 2, _pad, _clip, __call__ = forecast -> pad -> clip
  :220-262  GaussianNoiseForecaster: noise_std (relative_noise: * |mean(time_series)|; increase_uncertainty:
            * (1 + log(1 + k)) for forecast row k), val_c_n + normal(scale=noise_std)
src/pymgrid/modules/base/timeseries/base_timeseries_module.py
  :99-122   _update_step -> forecast(): window ts[t+1 : t+1+H]; t past the end -> full_pad
  :233-240  set_forecaster: the series handed to the forecaster is time_series[initial_step:final_step]
"""
import numpy as np


def series_bounds(ts, pull_zero):
    """per-column (low, high) of a [T, C] series; load / pv pull the bounds to include 0 (base_timeseries_module.py:81-88),
    the grid's four columns use their own min / max (grid_module.py:125-132)."""
    low, high = ts.min(axis=0), ts.max(axis=0)
    if pull_zero:
        low, high = np.minimum(low, 0.0), np.maximum(high, 0.0)
    return low, high


def noise_std(std, ts_window, horizon, n_cols, increase_uncertainty, relative_noise):
    """GaussianNoiseForecaster._get_noise_std, forecaster.py:237-250.  `ts_window` = time_series[initial_step:final_step]."""
    scalar = std
    if relative_noise:
        scalar = scalar * np.abs(ts_window.mean())
    if increase_uncertainty:
        return scalar * np.outer(1 + np.log(1 + np.arange(horizon)), np.ones(n_cols))
    return scalar


def _get_noise(std, size):
    """forecaster.py:252-260: full-shape draw, or the leading rows of the std table when the window is short."""
    try:
        return np.random.normal(scale=std, size=size)
    except ValueError:
        return np.random.normal(scale=std[:size[0], :], size=size)


def noisy_state(ts, t, horizon, low, high, std):
    """Unnormalised module state [current (C), forecast (H x C)] with a Gaussian-noise forecaster, drawing from the global
    numpy RNG exactly where the reference does.  ts [T, C]; low / high per column; std from `noise_std`."""
    T, C = ts.shape
    fill = (high + low) / 2                                          # forecaster.py:95
    if t >= T:                                                       # base_timeseries_module.py:113-116, :140-143
        return np.concatenate([fill, np.tile(fill, horizon)])
    window = ts[t + 1:t + 1 + horizon]
    forecast = window + _get_noise(std, window.shape).reshape(window.shape)     # forecaster.py:262
    if forecast.shape[0] < horizon:                                  # _pad, forecaster.py:120-132: pad rows carry no noise
        forecast = np.concatenate([forecast, np.tile(fill, (horizon - forecast.shape[0], 1))])
    forecast = np.minimum(np.maximum(forecast, low), high)           # _clip, forecaster.py:139-149
    return np.concatenate([ts[t], forecast.reshape(-1)])


def normalise(state, low, high, horizon):
    spread = high - low
    spread = np.where(spread == 0, 1.0, spread)                      # utils/space.py:204-205
    return (state - np.tile(low, 1 + horizon)) / np.tile(spread, 1 + horizon)


class NoisyModule:
    """One time-series module (load, pv or grid) with its forecaster settings."""

    def __init__(self, ts, horizon, pull_zero, std, increase_uncertainty, relative_noise, initial_step, final_step):
        self.ts = ts if ts.ndim == 2 else ts.reshape(-1, 1)
        self.horizon = horizon
        self.low, self.high = series_bounds(self.ts, pull_zero)
        stop = final_step if final_step > 0 else len(self.ts)
        self.std = noise_std(std, self.ts[initial_step:stop], horizon, self.ts.shape[1], increase_uncertainty, relative_noise)

    def observe(self, t):
        """normalised observation of the module at step t: one forecast() call = one draw (base_timeseries_module.py:99-101)"""
        return normalise(noisy_state(self.ts, t, self.horizon, self.low, self.high, self.std), self.low, self.high, self.horizon)

    def sigma_normalised(self):
        """per (forecast row, column) standard deviation in normalised units; 0 where the column is constant (the clip
        pins a constant column to its bound) -- what the CUDA path is parameterised with."""
        std = np.broadcast_to(np.asarray(self.std, dtype=np.float64), (self.horizon, self.ts.shape[1])) if np.ndim(self.std) \
            else np.full((self.horizon, self.ts.shape[1]), float(self.std))
        spread = self.high - self.low
        return np.where(spread > 0, std / np.where(spread > 0, spread, 1.0), 0.0)
