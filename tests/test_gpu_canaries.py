"""compute-sanitizer is closed on this GPU pool, so the kernels that write caller-provided device buffers are run
between canary regions instead: every byte before and after the output must still hold its pattern afterwards.
Covers the round-2 kernels (TMA-staged REF_AABB stream kernel and its ragged tails, ORIENTED kernel with deferred
flag stores, bit packing / unpacking, the int16 pose path, device-resident plan records) at awkward sizes."""
import numpy as np
import pytest
import torch

from autonomouscar_b200 import PP_COLLISION_ORIENTED, PP_COLLISION_REF_AABB, default_params, make_query, worlds
from autonomouscar_b200.capi import pp_query
from autonomouscar_b200.results import record_stride

pytestmark = pytest.mark.gpu
PAD = 4096


def _guarded(n_bytes, dtype=torch.uint8, align=16):
    """(whole buffer, view of n_bytes in the middle, aligned to `align`): canaries 0xA5 on both sides."""
    whole = torch.full((n_bytes + 2 * PAD + align,), 0xA5, dtype=torch.uint8, device="cuda")
    off = PAD + (-(whole.data_ptr() + PAD)) % align
    return whole, whole[off: off + n_bytes], off


def _intact(whole, off, n_bytes):
    w = whole.cpu().numpy()
    return bool((w[:off] == 0xA5).all() and (w[off + n_bytes:] == 0xA5).all())


@pytest.mark.parametrize("mode", [PP_COLLISION_REF_AABB, PP_COLLISION_ORIENTED])
def test_collision_kernels_stay_inside_their_flag_buffers(oracle, cuda_device, mode):
    from autonomouscar_b200 import Planner
    p = default_params(costmap_res=0.1, costmap_width=600, costmap_height=520, collision_mode=mode)
    grid = worlds.block_grid(600, 520, frac=0.02, block=8, seed=2)
    pl = Planner(p, device=0)
    pl.set_grid(grid)
    for n in (1, 3, 4, 31, 2047, 2048, 8191, 8192, 8193, 10241, 16384 + 5, 50001):
        poses = worlds.random_poses(p, n, seed=n, inset_m=-3.0)
        d = torch.from_numpy(poses).cuda()
        whole, flags, off = _guarded(n)
        pl.collide_dev(d, out=flags)
        pl.sync()
        assert _intact(whole, off, n), (mode, n)
        want, _ = oracle.collide_batch(p, grid, poses)
        np.testing.assert_array_equal(flags.cpu().numpy(), want)
        # bits: (n + 31) // 32 words, nothing beyond
        nw = (n + 31) // 32
        whole_b, bits, off_b = _guarded(4 * nw)
        pl.pack_flags_dev(flags, out=bits.view(torch.int32))
        pl.sync()
        assert _intact(whole_b, off_b, 4 * nw), (mode, n)
        whole_u, back, off_u = _guarded(n)
        pl.unpack_flags_dev(bits.view(torch.int32), n, out=back)
        pl.sync()
        assert _intact(whole_u, off_u, n) and torch.equal(back, flags)
    pl.close()


def test_plan_records_stay_inside_their_buffer(cuda_device):
    from autonomouscar_b200 import Planner
    p = default_params(collision_mode=PP_COLLISION_ORIENTED, max_nodes=5000)
    grid = worlds.block_grid(300, 300, frac=0.01, block=4, seed=9)
    worlds.clear_disc(grid, p, 0.0, 0.0, 4.0)
    pl = Planner(p, device=0)
    pl.set_grid(grid)
    for n, cap in ((1, 2), (7, 33), (150, 96), (300, 17)):
        qs = worlds.forward_queries(n, seed=n)
        stride = record_stride(cap)
        whole, rec, off = _guarded(n * stride)
        pl.plan_batch_dev((pp_query * n)(*qs), n, cap_path=cap, out=rec.view(n, stride))
        pl.sync()
        assert _intact(whole, off, n * stride), (n, cap)
        hdr = rec.view(n, stride)[:, :64].cpu().numpy().view(np.int32)
        assert (hdr[:, 10] == cap).all() and set(np.unique(hdr[:, 0])) <= {0, 1}
    pl.close()
