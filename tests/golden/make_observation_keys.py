"""BaseMicrogridEnv(observation_keys=...) (envs/base/base.py:109-163, 211-218) on the LIVE reference
-> tests/golden/observation_keys.npz.  Build container only:  python tests/golden/make_observation_keys.py

The env's observation becomes `state_series(normalized=True).loc[:, :, keys]`: for every key in the order given, the
modules that have such a field in listing order.  Recorded: pymgrid25 scenario 1 (fused module set) and a two-battery,
two-load grid (composed path): observation_space shape, reset observation, a few discrete steps.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
warnings.simplefilter("ignore")

FUSED_KEYS = ['soc', 'load_current', 'grid_status_current', 'load_forecast_3', 'current_status', 'renewable_forecast_0']
COMPOSED_KEYS = ['soc', 'load_current', 'renewable_forecast_2', 'current_charge', 'import_price_current']
FUSED_ACTIONS, COMPOSED_ACTIONS = (3, 0, 7, 11, 5, 2), (1, 0, 4, 5, 2, 3)


def composed_modules(ns):
    rng = np.random.default_rng(3)
    load, pv = 100 + 100 * rng.random(60), 200 * rng.random(60)
    ts = dict(forecaster="oracle")
    return [ns.BatteryModule(10, 100, 50, 50, 0.9, init_soc=0.2), ns.BatteryModule(10, 1000, 10, 10, 0.7, init_soc=0.3),
            ("pv", ns.RenewableModule(time_series=pv, forecast_horizon=4, **ts)), ns.LoadModule(time_series=load, forecast_horizon=2, **ts),
            ns.LoadModule(time_series=0.5 * load, forecast_horizon=2, **ts), ns.GridModule(100, 100, [0.2, 0.1, 0.5] * np.ones((60, 3)))]


def flow(env, actions):
    rows, rewards = [env.reset()], []
    for a in actions:
        o, r, _, _ = env.step(a)
        rows.append(o)
        rewards.append(r)
    return np.array(rows), np.array(rewards)


def main():
    from oracle.ref_loader import load_reference
    load_reference()
    import pymgrid.modules as R
    from pymgrid.envs import DiscreteMicrogridEnv
    data = {}
    data["fused_obs"], data["fused_rewards"] = flow(DiscreteMicrogridEnv.from_scenario(1, observation_keys=FUSED_KEYS), FUSED_ACTIONS)
    data["composed_obs"], data["composed_rewards"] = flow(DiscreteMicrogridEnv(composed_modules(R), observation_keys=COMPOSED_KEYS),
                                                          COMPOSED_ACTIONS)
    # the log a (single) env keeps: the microgrid's columns + the action it was given (envs/discrete/discrete.py:141)
    import json
    for label, make in (("s0", lambda: DiscreteMicrogridEnv.from_scenario(0)), ("s1", lambda: DiscreteMicrogridEnv.from_scenario(1)),
                        ("s2", lambda: DiscreteMicrogridEnv.from_scenario(2)), ("composed", lambda: DiscreteMicrogridEnv(composed_modules(R)))):
        env = make()
        rng = np.random.default_rng(5)
        acts = [int(rng.integers(0, env.action_space.n)) for _ in range(8)]
        for a in acts:
            env.step(a)
        log = env.log
        data[f"envlog_{label}_actions"] = np.array(acts)
        data[f"envlog_{label}_columns"] = np.array(json.dumps([list(c) for c in log.columns]))
        data[f"envlog_{label}_values"] = log.to_numpy(dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "observation_keys.npz"), **data)
    print({k: v.shape for k, v in data.items()})


if __name__ == "__main__":
    main()
