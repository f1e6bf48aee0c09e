"""Row N2 (local costmap): the oracle restatement (oracle/costmap_ref.h) against the REFERENCE's own
local_costmap.cpp (oracle/_ref/libref_costmap.so, compiled unmodified against ROS test doubles):
  * against the committed outputs of the reference binary (tests/golden/ref_costmap.npz) -- everywhere;
  * live, on fresh random scenes and car poses, byte for byte -- where oracle/_ref is present."""
import os
import sys

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import make_costmap_golden as mcg  # noqa: E402


def test_oracle_reproduces_reference_costmaps(oracle):
    z = np.load(os.path.join(GOLD, "ref_costmap.npz"))
    p = mcg.params()
    for k, (seed, car, offset) in enumerate(mcg.CASES):
        sc = mcg.scene_from_fixture(z, k)
        g = mcg.global_grid(p, car, offset)
        got, _ = oracle.costmap_build(p, sc.inputs, g, math_mode=oracle.MATH_LIBM)
        want = z[f"c{k}_grid"]
        np.testing.assert_array_equal(got, want)
        assert (want == 100).sum() > 1000 and (want == -1).sum() >= 1
        # the deterministic-math policy (what the GPU reproduces) paints the same picture
        det, _ = oracle.costmap_build(p, sc.inputs, g, math_mode=oracle.MATH_DET)
        assert (det != want).mean() < 1e-3


def test_oracle_matches_live_reference_costmap(oracle):
    from autonomouscar_b200 import worlds
    from oracle import ref_binding as ref
    if not ref.costmap_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(77)
    p = mcg.params()
    for trial in range(5):
        car = (float(rng.uniform(-150, 150)), float(rng.uniform(-150, 150)), float(rng.uniform(-3.1, 3.1)))
        ref.costmap_set_tfs(worlds.costmap_tfs(*car))
        node = ref.RefCostmap(p)
        sc = worlds.costmap_scene(100 + trial, max_reach=float(rng.uniform(12, 29)))
        ref.costmap_derive_inputs(sc)
        g = mcg.global_grid(p, car, int(rng.integers(0, 160))) if trial % 2 == 0 else None
        if g is not None:
            node.set_global(g)
        want, _ = node.build(sc)
        got, _ = oracle.costmap_build(node.params, sc.inputs, g, math_mode=oracle.MATH_LIBM)
        np.testing.assert_array_equal(got, want)
        # a second scan on the same node: clearMap (LC:349-358) must leave no state behind
        sc2 = worlds.costmap_scene(200 + trial)
        ref.costmap_derive_inputs(sc2)
        want2, _ = node.build(sc2)
        got2, _ = oracle.costmap_build(node.params, sc2.inputs, g, math_mode=oracle.MATH_LIBM)
        np.testing.assert_array_equal(got2, want2)
        node.close()


# ---- row N3: the global road map ---------------------------------------------------------------
import make_global_costmap_golden as mgg  # noqa: E402


def test_oracle_reproduces_reference_global_costmaps(oracle):
    z = np.load(os.path.join(GOLD, "ref_global_costmap.npz"))
    for k, (seed, lanes, width) in enumerate(mgg.SMALL_CASES):
        got, _ = oracle.global_costmap_build(mgg.small_params(lanes, width), z[f"s{k}_xy"], z[f"s{k}_off"],
                                             math_mode=oracle.MATH_LIBM)
        want = z[f"s{k}_grid"]
        np.testing.assert_array_equal(got, want)
        assert ((want >= 0) & (want < 99)).sum() > 100000 and (want == -1).sum() == mgg.SMALL + 1   # GC:83-89 never writes [0, H]
    # the launch-file map from the reference's own road polylines: same bytes as the reference published
    from autonomouscar_b200.costmap_capi import default_global_costmap_params
    xy, off = mgg.road_edges()
    got, _ = oracle.global_costmap_build(default_global_costmap_params(), xy, off, math_mode=oracle.MATH_LIBM)
    sha, rows, cols, total = mgg.full_summary(got, 12000)
    np.testing.assert_array_equal(sha, z["full_sha"])
    np.testing.assert_array_equal(rows, z["full_rows"])


def test_oracle_matches_live_reference_global_costmap(oracle, tmp_path):
    from oracle import ref_binding as ref
    if not ref.global_costmap_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    for seed, lanes, width in ((11, 1, 5.0), (12, 2, 11.1), (13, 4, 14.0)):
        xy, off = mgg.synth_polylines(seed, n_lines=9)
        p = mgg.small_params(lanes, width)
        want, _ = ref.global_costmap_build(p, xy, off, tmp_path)
        got, _ = oracle.global_costmap_build(p, xy, off, math_mode=oracle.MATH_LIBM)
        np.testing.assert_array_equal(got, want)


def test_oracle_matches_live_reference_costmap_over_parameters(oracle):
    """Other resolutions, non-square windows (the reference divides by the HEIGHT where it should use the
    width, LC:279 -- restated, not repaired), other buffers / ranges / inflation radii."""
    from autonomouscar_b200 import worlds
    from autonomouscar_b200.costmap_capi import default_costmap_params
    from oracle import ref_binding as ref
    if not ref.costmap_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    cases = [  # res, height_cells, width_cells, z buffer, obstacle range, max inflation, global map?
        (0.25, 300.0, 260.0, 0.075, 2.0, 5.0, False),
        (0.25, 260.0, 300.0, 0.2, 1.0, 8.0, False),
        (0.5, 150.0, 150.0, 0.075, 2.0, 5.0, True),
        (0.2, 400.0, 400.0, 0.0, 3.5, 2.0, True),
        (0.125, 520.0, 500.0, 0.075, 2.0, 10.0, False),
    ]
    for k, (res, hc, wc, zb, orange, infl, with_global) in enumerate(cases):
        G = 1200
        p = default_costmap_params(res=res, height_cells=hc, width_cells=wc, z_ground_buffer=zb, obstacle_range=orange,
                                   max_inflation_radius=infl, global_height_cells=float(G), global_width_cells=float(G))
        car = (20.0 + 7 * k, -15.0 + 3 * k, 0.3 * k - 0.5)
        ref.costmap_set_tfs(worlds.costmap_tfs(*car))
        node = ref.RefCostmap(p)
        assert (node.params.height_cells, node.params.width_cells) == (hc, wc)
        sc = worlds.costmap_scene(500 + k, max_reach=28.0)
        ref.costmap_derive_inputs(sc)
        g = None
        if with_global:
            g = worlds.road_like_global_grid(G, 11 * k, zero_at=worlds.costmap_corner_global_cells(p, *car))
            node.set_global(g)
        want, _ = node.build(sc)
        got, _ = oracle.costmap_build(node.params, sc.inputs, g, math_mode=oracle.MATH_LIBM)
        np.testing.assert_array_equal(got, want)
        assert (want == 100).sum() > 200
        node.close()
