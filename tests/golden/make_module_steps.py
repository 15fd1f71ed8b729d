"""BaseMicrogridModule.step on modules used WITHOUT a Microgrid (the reference's operator API; its module-level tests use it
this way) from the LIVE reference -> tests/golden/module_steps.npz.  Build container only.

Twelve modules (two of each kind, corner parameters included), 30 steps each alternating normalised and unnormalised
actions that reach beyond the modules' limits: observation, reward, done, info (provided / absorbed / extra), battery and
genset state after every step, and where the reference raises, the exception type.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
warnings.simplefilter("ignore")
N_STEPS = 30


def specs():
    """[(label, class name, kwargs)] -- plain data, shared by the recorder (reference classes) and the tests (ours)"""
    rng = np.random.default_rng(77)
    T = 24
    load, pv = 50 * rng.random(T) + 1, 80 * np.clip(rng.random(T) - 0.3, 0, None)
    grid4 = np.stack([rng.uniform(0.05, 0.9, T), rng.uniform(0, 0.4, T), rng.uniform(0, 0.6, T), (rng.random(T) > 0.3).astype(float)], axis=1)
    return [
        ("load_h0", "LoadModule", dict(time_series=load)),
        ("load_h3", "LoadModule", dict(time_series=load, forecaster="oracle", forecast_horizon=3, final_step=20)),
        ("pv_h0", "RenewableModule", dict(time_series=pv)),
        ("pv_h5_raises", "RenewableModule", dict(time_series=pv, forecaster="oracle", forecast_horizon=5, raise_errors=True)),
        ("battery", "BatteryModule", dict(min_capacity=10.0, max_capacity=100.0, max_charge=40.0, max_discharge=35.0, efficiency=0.9,
                                         battery_cost_cycle=0.02, init_soc=0.5)),
        ("battery_lossless", "BatteryModule", dict(min_capacity=0.0, max_capacity=55.5, max_charge=60.0, max_discharge=11.0, efficiency=1.0,
                                                  battery_cost_cycle=0.0, init_charge=20.0)),
        ("genset", "GensetModule", dict(running_min_production=10.0, running_max_production=50.0, genset_cost=0.5, co2_per_unit=2.0,
                                       cost_per_unit_co2=0.1)),
        ("genset_slow", "GensetModule", dict(running_min_production=0.0, running_max_production=33.0, genset_cost=0.7, start_up_time=2,
                                            wind_down_time=3, allow_abortion=False, init_start_up=False)),
        ("grid_weak", "GridModule", dict(max_import=30.0, max_export=20.0, time_series=grid4, cost_per_unit_co2=0.2,
                                        forecaster="oracle", forecast_horizon=2)),
        ("grid_3col", "GridModule", dict(max_import=25.0, max_export=0.0, time_series=grid4[:, :3])),
        ("slack", "UnbalancedEnergyModule", dict(raise_errors=False, loss_load_cost=7.5, overgeneration_cost=1.25)),
        ("slack_default", "UnbalancedEnergyModule", dict(raise_errors=False)),
    ]


def actions_for(label, cls, kwargs, rng):
    """[(action, normalized)] per step: plain data again"""
    out = []
    for k in range(N_STEPS):
        normalized = k % 2 == 0
        if cls == "LoadModule":
            a = np.array([])
        elif cls == "GensetModule":
            hi = kwargs["running_max_production"]
            a = np.array([rng.random(), rng.random() if normalized else rng.uniform(0, 1.4 * hi)])
        elif cls == "UnbalancedEnergyModule":
            normalized = False                                   # (-inf, inf) bounds: a normalised action has no meaning
            a = float(rng.uniform(-50, 50))
        elif cls == "RenewableModule":
            a = float(rng.random() if normalized else rng.uniform(0, 90))
            if kwargs.get("raise_errors") and k < 6:
                a, normalized = 0.0, False                       # inside the limits: the clip that raises comes later
        elif cls == "BatteryModule":
            a = float(rng.random() if normalized else rng.uniform(-80, 80))
        else:
            a = float(rng.random() if normalized else rng.uniform(-40, 40))
        out.append((a, normalized))
    return out


def run(module, cls, actions):
    rows, rewards, dones, infos, states = [], [], [], [], []
    raised_at, raised = -1, ""
    for k, (a, normalized) in enumerate(actions):
        try:
            obs, reward, done, info = module.step(a, normalized=normalized)
        except Exception as exc:      # noqa: BLE001
            raised_at, raised = k, type(exc).__name__
            break
        rows.append(np.asarray(obs, dtype=np.float64).ravel()); rewards.append(reward); dones.append(bool(done))
        infos.append([info.get("provided_energy", 0.0), info.get("absorbed_energy", 0.0),
                      info.get("co2_production", info.get("curtailment", 0.0)), float("absorbed_energy" in info)])
        if cls == "BatteryModule":
            states.append([module.current_charge, module.soc])
        elif cls == "GensetModule":
            states.append([float(x) for x in module.state])
        else:
            states.append([float(module.current_step)])
    n = len(rewards)
    width = len(rows[0]) if rows else 0
    return dict(obs=np.array(rows).reshape(n, width), rewards=np.array(rewards), dones=np.array(dones, dtype=np.uint8),
                info=np.array(infos).reshape(n, 4), states=np.array(states, dtype=np.float64).reshape(n, len(states[0]) if states else 0),
                raised_at=np.array(raised_at), raised=np.array(raised), reset_obs=np.asarray(module.reset(), dtype=np.float64).ravel())


def main():
    from oracle.ref_loader import load_reference
    load_reference()
    import pymgrid.modules as R
    data = {}
    rng = np.random.default_rng(5)
    for label, cls, kwargs in specs():
        acts = actions_for(label, cls, kwargs, rng)
        rec = run(getattr(R, cls)(**kwargs), cls, acts)
        for k, v in rec.items():
            data[f"{label}_{k}"] = v
        print(label, len(rec["rewards"]), int(rec["raised_at"]), rec["raised"])
    np.savez_compressed(os.path.join(HERE, "module_steps.npz"), **data)


if __name__ == "__main__":
    main()
