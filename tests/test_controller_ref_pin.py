"""Row N4 (profile follower): autonomouscar_b200/rollout.py's PriusController / ProfileFollower against the
REFERENCE's own controller.cpp, compiled unmodified against the ROS test doubles
(oracle/_ref/libref_controller.so, oracle/ref_controller_harness.cpp):
  * against the committed outputs of the reference object (tests/golden/ref_controller.npz) -- everywhere;
  * live, on fresh random scripts -- where oracle/_ref is present.
Cursor, throttle, brake, steer, gear and the publish count are compared exactly after every event
(NaN / inf from the reference's dt = 0 divisions included)."""
import os
import sys

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import make_controller_golden as mkg  # noqa: E402

from autonomouscar_b200 import rollout  # noqa: E402


class Mirror:
    """rollout.PriusController behind the interface make_controller_golden.drive() expects."""

    def __init__(self, t0):
        self.c = rollout.PriusController(t0)

    def model_state(self, speed, yaw_rate):
        self.c.on_model_state(speed, yaw_rate)

    def profile(self, t, *p):
        self.c.on_profile(t, *p)

    def step(self, t):
        return self.c.step(t)

    def state(self):
        c = self.c
        return (c.control_it, c.throttle, c.brake, c.steer, c.shift_gears, c.published)


def assert_same(got, want):
    assert got.shape == want.shape
    np.testing.assert_array_equal(got[:, 0], want[:, 0], err_msg="which passes the reference can execute")
    np.testing.assert_array_equal(got[:, 2:], want[:, 2:], err_msg="throttle / brake / steer / gear / publish count")
    live = np.ones(len(want), dtype=bool)          # the cursor is compared until a pass is refused; the harness
    for k in range(len(want)):                     # refuses BEFORE the reference moves it, the mirror after
        if want[k, 0] == 0:
            live[k:] = False
            nxt = np.nonzero(np.arange(len(want)) > k)[0]
            for j in nxt:                          # a new profile resets both
                if want[j, 1] == 0 and want[j, 0] == 1 and got[j, 1] == 0:
                    live[j:] = True
                    break
    np.testing.assert_array_equal(got[live, 1], want[live, 1], err_msg="m_control_it")


def test_mirror_reproduces_reference_controller_fixture():
    z = np.load(os.path.join(GOLD, "ref_controller.npz"))
    for seed in mkg.SEEDS:
        t0, profiles, events = mkg.scenario(seed)
        np.testing.assert_array_equal(np.array(events, dtype=np.float64), z[f"s{seed}_events"])   # no hand edits
        want = z[f"s{seed}_out"]
        assert_same(mkg.drive(Mirror, t0, profiles, events), want)
        assert want[:, 1].max() >= 10 and (want[:, 0] == 0).any()          # deep cursors and refused passes occur


def test_mirror_matches_live_reference_controller():
    from oracle import ref_binding as ref
    if not ref.controller_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    for seed in (201, 202, 203, 204):
        t0, profiles, events = mkg.scenario(seed)
        assert_same(mkg.drive(Mirror, t0, profiles, events), mkg.drive(ref.RefController, t0, profiles, events))


def test_profile_follower_is_the_controllers_cursor():
    """The roll-outs' ProfileFollower (and through test_native_rollout_equals_python_rollout the C++ one) walks
    the profile exactly like the reference object's m_control_it."""
    from oracle import ref_binding as ref
    if not ref.controller_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(9)
    for trial in range(20):
        n = int(rng.integers(2, 30))
        prof = (rng.uniform(-0.5, 0.5, n), rng.uniform(20, 300, n), rng.uniform(0, 2, n))
        c, f = ref.RefController(0.0), rollout.ProfileFollower()
        c.profile(0.0, *prof, 1.5, 1.5)
        f.new_profile(0.0, *prof)
        f.command(0.0)                                    # motionPlanningCallback runs one pass itself (:182)
        t = 0.0
        while True:
            t += 0.01
            ok = c.step(t)
            cmd = f.command(t)
            assert ok == (cmd is not None)
            if not ok:
                break
            assert f.it == int(c.state()[0])
            assert cmd == (prof[0][f.it], prof[2][f.it])
        assert f.it == n
