"""GPU parity: batched checkCollision through the C ABI vs the CPU oracle, bit-exact.

Every call goes through libprius_planner_b200.so (pp_collide_batch / _dev); the oracle is
only the checker.
"""
import numpy as np
import pytest
import torch

from autonomouscar_b200 import (PP_COLLISION_AS_SHIPPED, PP_COLLISION_ORIENTED, PP_COLLISION_REF_AABB,
                                PP_GRID_CSPACE, PP_GRID_OCCUPANCY_MSG, default_params, worlds)

pytestmark = pytest.mark.gpu


def _planner(params, grid, layout=PP_GRID_CSPACE):
    from autonomouscar_b200 import Planner
    pl = Planner(params, device=0)
    pl.set_grid(grid, layout)
    return pl


CASES = [
    # (res, W, H) : launch-file grid, BASELINE 0.1 m grids, ragged non-multiple-of-32 sizes
    (0.25, 300, 300),
    (0.1, 1024, 1024),
    (0.1, 517, 333),
    (0.1, 64, 40),
]


@pytest.mark.parametrize("res,W,H", CASES)
@pytest.mark.parametrize("mode", [PP_COLLISION_REF_AABB, PP_COLLISION_ORIENTED])
def test_collide_matches_oracle(oracle, cuda_device, res, W, H, mode):
    p = default_params(costmap_res=res, costmap_width=W, costmap_height=H, collision_mode=mode)
    grid = worlds.block_grid(W, H, frac=0.03, block=8, seed=W + H)
    poses = worlds.random_poses(p, 60000, seed=7, inset_m=-8.0)  # negative inset: poses beyond the border too
    pl = _planner(p, grid)
    got = pl.collide(poses)
    want, _ = oracle.collide_batch(p, grid, poses)
    assert got.dtype == np.uint8 and got.shape == want.shape
    assert 0.0 < want.mean() < 1.0
    np.testing.assert_array_equal(got, want)
    # device-pointer entry point, and broad phase off ("direct"), must agree bit for bit
    d = torch.from_numpy(poses).cuda()
    f1 = pl.collide_dev(d); pl.sync()
    pl.set_broad_phase(False)
    f2 = pl.collide_dev(d); pl.sync()
    np.testing.assert_array_equal(f1.cpu().numpy(), want)
    np.testing.assert_array_equal(f2.cpu().numpy(), want)
    if mode == PP_COLLISION_ORIENTED:
        assert pl.last_narrow_count() >= int(want.sum())
    pl.close()


@pytest.mark.parametrize("res,W,H", [(0.1, 700, 520), (0.25, 300, 300), (0.2, 400, 333), (0.125, 512, 512)])
@pytest.mark.parametrize("world", ["cells", "solid", "mixed"])
def test_oriented_proofs_stress(oracle, cuda_device, res, W, H, world):
    """The batch kernel decides most poses by proofs (yaw-binned dilated maps, oriented 4x4-block hull and core,
    perimeter pass, cell-level hull): poses chosen to sit on their edges -- yaws on the bin borders (k pi / 16) and
    one ulp either side, axis-aligned yaws (a band of the footprint drops out), |yaw| beyond the binned range, poses
    hugging single-cell obstacles (no 4x4 block is full) and solid areas (nearly every block is), lattice shapes other
    than the two templated ones -- must agree with the oracle bit for bit, with and without the broad phase."""
    p = default_params(costmap_res=res, costmap_width=W, costmap_height=H, collision_mode=PP_COLLISION_ORIENTED)
    rng = np.random.default_rng(W * 7 + H + len(world))
    if world == "cells":
        grid = worlds.cell_noise_grid(W, H, p=0.004, seed=W + 1)
    elif world == "solid":
        grid = np.zeros((W, H), dtype=np.int8)
        for _ in range(6):
            i, j = rng.integers(0, W - 60), rng.integers(0, H - 60)
            grid[i:i + rng.integers(20, 60), j:j + rng.integers(20, 60)] = 100
    else:
        grid = worlds.block_grid(W, H, frac=0.02, block=5, seed=H + 3)
        grid[rng.random(grid.shape) < 0.001] = 100
    n = 40000
    poses = worlds.random_poses(p, n, seed=11, inset_m=-5.0)
    k = rng.integers(-40, 40, size=n)
    border = (k * (np.pi / 16)).astype(np.float32)
    kind = rng.integers(0, 6, size=n)
    yaw = poses[:, 2].copy()
    yaw = np.where(kind == 0, border, yaw)
    yaw = np.where(kind == 1, np.nextafter(border, np.float32(10.0)), yaw)
    yaw = np.where(kind == 2, np.nextafter(border, np.float32(-10.0)), yaw)
    yaw = np.where(kind == 3, (rng.integers(-3, 4, size=n) * (np.pi / 2)).astype(np.float32), yaw)
    yaw = np.where((kind == 4) & (rng.random(n) < 0.2), (yaw * 60.0).astype(np.float32), yaw)   # |yaw| up to ~190 rad
    poses[:, 2] = yaw.astype(np.float32)
    # half of the poses hug an obstacle cell: within ~3 m of one, where the proofs are undecided most often
    occ = np.argwhere(grid == 100)
    pick = occ[rng.integers(0, len(occ), size=n // 2)]
    poses[: n // 2, 1] = ((W // 2 - pick[:, 0]) * res + rng.uniform(-3.0, 3.0, n // 2)).astype(np.float32)
    poses[: n // 2, 0] = ((H // 2 - pick[:, 1]) * res + rng.uniform(-3.0, 3.0, n // 2)).astype(np.float32)
    pl = _planner(p, grid)
    want, _ = oracle.collide_batch(p, grid, poses)
    assert 0.02 < want.mean() < 0.98
    d = torch.from_numpy(poses).cuda()
    torch.cuda.synchronize()
    f1 = pl.collide_dev(d); pl.sync()
    np.testing.assert_array_equal(f1.cpu().numpy(), want)
    pl.set_broad_phase(False)
    f2 = pl.collide_dev(d); pl.sync()
    np.testing.assert_array_equal(f2.cpu().numpy(), want)
    pl.close()


def test_as_shipped_is_always_false(cuda_device):
    # LP:445: the shipped reference returns false unconditionally
    p = default_params(collision_mode=PP_COLLISION_AS_SHIPPED)
    pl = _planner(p, np.full((300, 300), 100, dtype=np.int8))
    assert pl.collide(worlds.random_poses(p, 5000)).sum() == 0
    pl.close()


@pytest.mark.parametrize("mode", [PP_COLLISION_REF_AABB, PP_COLLISION_ORIENTED])
def test_empty_and_full_grids(oracle, cuda_device, mode):
    p = default_params(costmap_res=0.1, costmap_width=512, costmap_height=512, collision_mode=mode)
    poses = worlds.random_poses(p, 20000, seed=3, inset_m=3.0)
    pl = _planner(p, np.zeros((512, 512), dtype=np.int8))
    assert pl.collide(poses).sum() == 0
    pl.set_grid(np.full((512, 512), 99, dtype=np.int8))   # 99 is road cost, not an obstacle (LP:451)
    assert pl.collide(poses).sum() == 0
    pl.set_grid(np.full((512, 512), 100, dtype=np.int8))
    assert pl.collide(poses).sum() == len(poses)
    # far outside the grid / absurd coordinates: no cell can be touched
    far = np.array([[1e4, 0, 0], [0, -1e4, 1], [3e9, 3e9, 0], [-1e30, 5, 2]], dtype=np.float32)
    np.testing.assert_array_equal(pl.collide(far), oracle.collide_batch(p, np.full((512, 512), 100, np.int8), far)[0])
    assert pl.collide(far).sum() == 0
    # empty batch
    assert pl.collide(np.zeros((0, 3), np.float32)).shape == (0,)
    pl.close()


def test_permutation_invariance_and_single_cell(oracle, cuda_device):
    p = default_params(costmap_res=0.1, costmap_width=700, costmap_height=700, collision_mode=PP_COLLISION_ORIENTED)
    grid = np.zeros((700, 700), dtype=np.int8)
    grid[350, 350] = 100  # one occupied cell at the vehicle origin
    poses = worlds.random_poses(p, 40000, seed=11, inset_m=30.0)
    poses[:, :2] *= 0.15  # concentrate around the obstacle
    pl = _planner(p, grid)
    f = pl.collide(poses)
    want, _ = oracle.collide_batch(p, grid, poses)
    np.testing.assert_array_equal(f, want)
    assert 0 < f.sum() < len(f)
    perm = np.random.default_rng(0).permutation(len(poses))
    np.testing.assert_array_equal(pl.collide(poses[perm]), f[perm])
    pl.close()


def test_occupancy_msg_layout_flip(oracle, cuda_device):
    # LP:425-441: cspace[i][j] = data[H*W - i*H - j - 1]
    W, H = 96, 70
    p = default_params(costmap_res=0.25, costmap_width=W, costmap_height=H, collision_mode=PP_COLLISION_REF_AABB)
    msg = np.random.default_rng(5).integers(-1, 101, size=W * H).astype(np.int8)
    pl = _planner(p, msg, PP_GRID_OCCUPANCY_MSG)
    want = oracle.cspace_from_msg(msg, W, H)
    np.testing.assert_array_equal(pl.get_cspace(), want)
    poses = worlds.random_poses(p, 5000, seed=2, inset_m=-2.0)
    np.testing.assert_array_equal(pl.collide(poses), oracle.collide_batch(p, want, poses)[0])
    pl.close()


def test_footprint_transform_tolerance(oracle, cuda_device):
    """north_star: footprint-transform floats agree within 1e-5 m.  The GPU lattice is
    bit-identical to the oracle's and both sit within 1e-5 m of a float64 libm transform."""
    p = default_params(costmap_res=0.1, costmap_width=4096, costmap_height=4096, collision_mode=PP_COLLISION_ORIENTED)
    from autonomouscar_b200 import Planner
    pl = Planner(p, device=0)
    nu, nv, ok = oracle.lattice_dims(p)
    assert (nu, nv, ok) == (48, 17, True)
    rng = np.random.default_rng(9)
    for _ in range(40):
        x, y = rng.uniform(-200, 200, 2)
        yaw = rng.uniform(-7, 7)
        got = pl.footprint_points(x, y, yaw)
        want = oracle.footprint_points(p, x, y, yaw)
        np.testing.assert_array_equal(got, want)  # bit-exact vs oracle
        u = np.linspace(-p.car_length / 2, p.car_length / 2, nu)
        v = np.linspace(-p.car_width / 2, p.car_width / 2, nv)
        U, V = np.meshgrid(u, v, indexing="ij")
        ref = np.stack([(x + U * np.cos(yaw) - V * np.sin(yaw)).ravel(),
                        (y + U * np.sin(yaw) + V * np.cos(yaw)).ravel()], 1)
        assert np.abs(got - ref).max() < 1e-5  # tolerance stated by BASELINE.json:north_star
    pl.close()


def test_edges_match_oracle(oracle, cuda_device):
    for mode in (PP_COLLISION_REF_AABB, PP_COLLISION_ORIENTED):
        p = default_params(costmap_res=0.1, costmap_width=800, costmap_height=800, collision_mode=mode)
        grid = worlds.block_grid(800, 800, frac=0.03, block=8, seed=21)
        a = worlds.random_poses(p, 20000, seed=4, inset_m=1.0)
        rng = np.random.default_rng(1)
        ln = rng.uniform(0, 3.0, len(a)).astype(np.float32)
        edges = np.empty((len(a), 5), dtype=np.float32)
        edges[:, 0:2] = a[:, 0:2]
        edges[:, 2] = a[:, 0] + ln * np.cos(a[:, 2])
        edges[:, 3] = a[:, 1] + ln * np.sin(a[:, 2])
        edges[:, 4] = a[:, 2]
        edges[:50, 2:4] = edges[:50, 0:2]  # zero-length edges
        pl = _planner(p, grid)
        got = pl.collide_edges_dev(torch.from_numpy(edges).cuda()); pl.sync()
        want = oracle.collide_edges(p, grid, edges)
        np.testing.assert_array_equal(got.cpu().numpy(), want)
        assert 0 < want.sum() < len(want)
        pl.close()


def test_errors_are_loud(cuda_device):
    from autonomouscar_b200 import Planner, PlannerError
    p = default_params(collision_mode=PP_COLLISION_REF_AABB)
    pl = Planner(p, device=0)
    with pytest.raises(PlannerError):  # no grid yet
        pl.collide(np.zeros((4, 3), np.float32))
    with pytest.raises(AssertionError):  # wrong grid size
        pl.set_grid(np.zeros((10, 10), np.int8))
    pl.close()
    with pytest.raises(PlannerError):  # footprint wider than the tile window supports
        Planner(default_params(costmap_res=0.01), device=0)


def test_full_size_config3_properties(oracle, cuda_device):
    """BASELINE configs[2] at full size (2^24 poses, 4096^2 grid): size-independent properties.
    (a) the pruned product path and the 'direct' path (full lattice for every pose) agree bit for
    bit; (b) the host-pointer entry point (chunked H2D / D2H pipeline) agrees with the device one;
    (c) a random sub-sample agrees with the oracle; (d) REF_AABB likewise."""
    from autonomouscar_b200 import Planner
    n = 1 << 24
    p = default_params(costmap_res=0.1, costmap_width=4096, costmap_height=4096, collision_mode=PP_COLLISION_ORIENTED)
    grid = worlds.block_grid(4096, 4096, frac=0.01, block=16, seed=20181)
    pl = Planner(p, device=0)
    pl.set_grid(grid)
    box = worlds.pose_box(p)
    poses = pl.sample_poses_dev(2018, 0, 0, n, *box)
    f_broad = pl.collide_dev(poses)
    pl.sync()
    narrow_broad = pl.last_narrow_count()
    pl.set_broad_phase(False)
    f_direct = pl.collide_dev(poses)
    pl.sync()
    assert pl.last_narrow_count() == n and narrow_broad < n // 4
    assert torch.equal(f_broad, f_direct)
    pl.set_broad_phase(True)
    host = poses.cpu()
    f_host = pl.collide(host.numpy())
    assert np.array_equal(f_host, f_broad.cpu().numpy())
    rng = np.random.default_rng(1)
    pick = rng.choice(n, 60000, replace=False)
    sub = host.numpy()[pick]
    want, _ = oracle.collide_batch(p, grid, sub, n_threads=8)
    np.testing.assert_array_equal(f_host[pick], want)
    assert 0.05 < want.mean() < 0.15
    pl.set_collision_mode(PP_COLLISION_REF_AABB)
    p.collision_mode = PP_COLLISION_REF_AABB
    f_aabb = pl.collide_dev(poses)
    pl.sync()
    want, _ = oracle.collide_batch(p, grid, sub, n_threads=8)
    np.testing.assert_array_equal(f_aabb.cpu().numpy()[pick], want)
    pl.close()


def test_8192_grid_matches_oracle(oracle, cuda_device):
    """BASELINE configs[3] grid size (8192^2 @0.1 m, 64 MiB of cells)."""
    from autonomouscar_b200 import Planner
    for mode in (PP_COLLISION_REF_AABB, PP_COLLISION_ORIENTED):
        p = default_params(costmap_res=0.1, costmap_width=8192, costmap_height=8192, collision_mode=mode)
        grid = worlds.block_grid(8192, 8192, frac=0.01, block=16, seed=4, noise=False)
        poses = worlds.random_poses(p, 40000, seed=12, inset_m=-5.0)
        pl = Planner(p, device=0)
        pl.set_grid(grid)
        got = pl.collide(poses)
        want, _ = oracle.collide_batch(p, grid, poses, n_threads=8)
        np.testing.assert_array_equal(got, want)
        assert 0 < want.sum() < len(want)
        pl.close()


@pytest.mark.parametrize("res,W,H", [(0.1, 4096, 4096), (0.25, 300, 300), (0.1, 517, 333), (0.1, 8192, 2048)])
def test_ref_aabb_stream_kernel_cell_boundaries(oracle, cuda_device, res, W, H):
    """REF_AABB batch kernel (TMA-staged pose stream, fast base-cell arithmetic with an exact fallback): poses ON
    cell boundaries (k * res as float32 and its float32 neighbours, where int(W/2 - y/res) of LP:471-472 flips),
    at the map border, far away, NaN / inf -- against the oracle's float64 division, bit for bit; the aligned
    streaming variant, the misaligned plain variant and a ragged tail."""
    p = default_params(costmap_res=res, costmap_width=W, costmap_height=H, collision_mode=PP_COLLISION_REF_AABB)
    grid = worlds.block_grid(W, H, frac=0.02, block=8, seed=W ^ H)
    rng = np.random.default_rng(99)
    n = 200003
    poses = worlds.random_poses(p, n, seed=5, inset_m=-6.0).copy()
    k = rng.integers(-W, W, n // 2)
    on = (k * res).astype(np.float32)
    on = np.nextafter(on, np.where(rng.integers(0, 3, on.size) == 0, np.float32(-1e9), np.where(rng.integers(0, 2, on.size) == 0, np.float32(1e9), on)).astype(np.float32))
    poses[: n // 2, 1] = on                                              # y decides the first index (LP:471)
    poses[n // 4: 3 * n // 4, 0] = (rng.integers(-H, H, 3 * n // 4 - n // 4) * res).astype(np.float32)
    poses[-8:, 0] = [np.nan, np.inf, -np.inf, 3e9, -3e9, 1e30, 0.0, -0.0]
    poses[-16:-8, 1] = [np.nan, np.inf, -np.inf, 3e9, -3e9, 1e30, 0.0, -0.0]
    want, _ = oracle.collide_batch(p, grid, poses)
    assert 0.0 < want.mean() < 1.0
    pl = _planner(p, grid)
    d = torch.from_numpy(poses).cuda()
    got = pl.collide_dev(d); pl.sync()
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    # misaligned views (pose pointer 12 B off a 16 B boundary, flag pointer 1 B off): the plain kernel
    out = torch.zeros(n + 1, dtype=torch.uint8, device="cuda")
    pl.collide_dev(d[1:], out=out[1:n]); pl.sync()
    np.testing.assert_array_equal(out[1:n].cpu().numpy(), want[1:])
    # host-pointer entry (chunked pipeline) agrees
    np.testing.assert_array_equal(pl.collide(poses), want)
    pl.close()


def test_pack_unpack_flags(cuda_device):
    """pp_pack_flags_dev / pp_unpack_flags_dev: word k bit b = flag 32k + b, ragged lengths, unused bits zero."""
    from autonomouscar_b200 import Planner
    pl = Planner(default_params(), device=0)
    rng = np.random.default_rng(3)
    for n in (1, 15, 16, 31, 32, 33, 1000, 65536, 100003):
        f = (rng.random(n) < 0.3).astype(np.uint8)
        d = torch.from_numpy(f).cuda()
        bits = pl.pack_flags_dev(d); pl.sync()
        want = np.packbits(np.pad(f, (0, (-n) % 32)), bitorder="little").view(np.uint32)
        np.testing.assert_array_equal(bits.cpu().numpy().view(np.uint32), want)
        back = pl.unpack_flags_dev(bits, n); pl.sync()
        np.testing.assert_array_equal(back.cpu().numpy(), f)
    pl.close()


@pytest.mark.parametrize("mode", [PP_COLLISION_REF_AABB, PP_COLLISION_ORIENTED])
def test_bit_results_equal_packed_flags(cuda_device, mode):
    """pp_collide_batch_bits_dev (in ORIENTED mode the batch kernel ORs the bits in itself): the words equal
    pp_pack_flags_dev of pp_collide_batch_dev's flags, for ragged lengths, a reused (dirty) output buffer, with and
    without the broad phase, and poses off the grid."""
    p = default_params(costmap_res=0.1, costmap_width=1024, costmap_height=1024, collision_mode=mode)
    grid = worlds.block_grid(1024, 1024, frac=0.02, block=8, seed=31)
    pl = _planner(p, grid)
    for n in (1, 31, 32, 127, 128, 129, 4099, 300007):
        poses = torch.from_numpy(worlds.random_poses(p, n, seed=n, inset_m=-6.0)).cuda()
        torch.cuda.synchronize()                         # (the upload runs on torch's stream, the kernels on the handle's)
        for broad in (1, 0) if n <= 4099 else (1,):
            pl.set_broad_phase(bool(broad))
            want = pl.pack_flags_dev(pl.collide_dev(poses))
            out = torch.full(((n + 31) // 32,), -1, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()                     # (the fill runs on torch's stream, the kernel on the handle's)
            got = pl.collide_bits_dev(poses, out=out)
            pl.sync()
            assert torch.equal(got, want), (n, broad)
    pl.set_broad_phase(True)
    pl.close()


@pytest.mark.parametrize("mode", [PP_COLLISION_REF_AABB, PP_COLLISION_ORIENTED])
def test_compact_int16_pose_format(oracle, cuda_device, mode):
    """pp_collide_batch_q16 (6 B per pose over PCIe): the device reconstructs float32 poses with two separately rounded
    float32 operations per coordinate; the flags equal pp_collide_batch -- and the oracle -- on exactly those poses."""
    from autonomouscar_b200.capi import quantise_poses
    p = default_params(costmap_res=0.1, costmap_width=1024, costmap_height=1024, collision_mode=mode)
    grid = worlds.block_grid(1024, 1024, frac=0.02, block=8, seed=31)
    pl = _planner(p, grid)
    poses = worlds.random_poses(p, 300007, seed=17, inset_m=-4.0)
    box = worlds.pose_box(p, inset_m=-4.0)
    q, quant, back = quantise_poses(poses, *box)
    assert np.abs(back[:, :2] - poses[:, :2]).max() < 2e-3 and np.abs(back[:, 2] - poses[:, 2]).max() < 1e-4
    got = pl.collide_q16(q, quant)
    np.testing.assert_array_equal(got, pl.collide(back))
    want, _ = oracle.collide_batch(p, grid, back)
    np.testing.assert_array_equal(got, want)
    assert 0.0 < want.mean() < 1.0
    pl.close()
