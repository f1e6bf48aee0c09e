"""Regenerates tests/golden/ref_controller.npz by driving the REFERENCE's own controller
(prius_controller/src/controller.cpp compiled unmodified into oracle/_ref/libref_controller.so, see
oracle/ref_controller_harness.cpp) with scripted /mp profiles, model states and clock values.
Run in the development container: python tests/golden/make_controller_golden.py
Event rows: (kind, t, a, b) with kind 0 = model state (a = speed, b = yaw rate), 1 = control pass at t,
2 = new /mp profile number a at t.  Output rows (after every event): ok, m_control_it, throttle, brake,
steer, shift_gears, messages published."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))
SEEDS = (21, 22, 23)
N_EVENTS = 1500


def scenario(seed):
    """(profiles, events): profiles = list of (yaw_rates, durations_ms, speeds, max_speed, max_accel)."""
    rng = np.random.default_rng(seed)
    profiles, events, t = [], [], float(rng.uniform(0, 100))
    t0 = t

    def new_profile():
        n = int(rng.integers(1, 40))
        profiles.append((rng.uniform(-0.6, 0.6, n), rng.choice([12.5, 25.0, 50.0, 100.0, 250.0], n) * rng.uniform(0.5, 2.0),
                         rng.uniform(-0.5, 3.0, n) * (rng.random(n) > 0.1), float(rng.uniform(0.5, 3.0)), float(rng.uniform(0.5, 3.0))))
        events.append((2, t, len(profiles) - 1, 0.0))

    events.append((0, t, float(rng.uniform(0, 2)), float(rng.uniform(-0.3, 0.3))))
    new_profile()
    while len(events) < N_EVENTS:
        r = rng.random()
        if r < 0.25:
            events.append((0, t, float(rng.uniform(-0.5, 3.5)), float(rng.uniform(-0.7, 0.7))))
        elif r < 0.985:
            if rng.random() > 0.05:                     # now and then two passes share a clock value (dt = 0)
                t += 0.01 * float(rng.uniform(0.5, 1.5))
            events.append((1, t, 0.0, 0.0))
        else:
            t += 0.003
            new_profile()
    return t0, profiles, events


def drive(ctrl_factory, t0, profiles, events):
    """Run the script through a controller with the RefController / rollout.PriusController interface."""
    c = ctrl_factory(t0)
    rows = []
    for kind, t, a, b in events:
        ok = True
        if kind == 0:
            c.model_state(a, b)
        elif kind == 1:
            ok = c.step(t)
        else:
            c.profile(t, *profiles[int(a)])
        rows.append((float(ok),) + tuple(float(v) for v in c.state()))
    return np.array(rows)


def main():
    from oracle import ref_binding as ref
    assert ref.controller_available(), "oracle/_ref/libref_controller.so could not be built"
    blob = {}
    for seed in SEEDS:
        t0, profiles, events = scenario(seed)
        out = drive(ref.RefController, t0, profiles, events)
        blob[f"s{seed}_out"] = out
        blob[f"s{seed}_events"] = np.array(events, dtype=np.float64)
        print(f"seed {seed}: {len(events)} events, {len(profiles)} profiles, {int((out[:, 0] == 0).sum())} refused passes, "
              f"cursor up to {int(out[:, 1].max())}")
    np.savez_compressed(os.path.join(OUT, "ref_controller.npz"), **blob)


if __name__ == "__main__":
    main()
