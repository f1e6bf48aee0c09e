"""GPU parity for rows S1 (Philox sampler), S2 (nearest neighbour) and A6 (exact-match
lookups, the reference's std::find at LP:56-57)."""
import numpy as np
import pytest
import torch

from autonomouscar_b200 import default_params

pytestmark = pytest.mark.gpu

# Random123 known-answer vectors for philox4x32-10 (counter, key) -> output
PHILOX_KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


@pytest.fixture(scope="module")
def pl(cuda_device):
    from autonomouscar_b200 import Planner
    p = Planner(default_params(), device=0)
    yield p
    p.close()


def test_philox_known_answers(pl, oracle):
    for ctr, key, want in PHILOX_KAT:
        assert pl.philox(ctr, key) == want
        assert oracle.philox4x32_10(ctr, key) == want


def test_sampler_matches_oracle(pl, oracle):
    box = (-202.3, 202.3, -150.0, 99.5)
    n = 100003
    for seed, stream, first in ((1, 0, 0), (0xDEADBEEFCAFE, 7, (1 << 33) + 5)):
        got = pl.sample_poses_dev(seed, stream, first, n, *box); pl.sync()
        want = oracle.sample_poses(seed, stream, first, n, *box)
        np.testing.assert_array_equal(got.cpu().numpy(), want)  # bit-exact float32
    a = got.cpu().numpy()
    assert a[:, 0].min() >= box[0] and a[:, 0].max() <= box[1]
    assert abs(a[:, 2].mean()) < 0.05 and a[:, 2].min() >= -np.pi - 1e-6 and a[:, 2].max() <= np.pi + 1e-6
    # a sub-range of the stream equals the same indices of the full stream
    part = pl.sample_poses_dev(1, 0, 1000, 50, *box); pl.sync()
    full = oracle.sample_poses(1, 0, 0, 2000, *box)
    np.testing.assert_array_equal(part.cpu().numpy(), full[1000:1050])


@pytest.mark.parametrize("n_nodes,n_query", [(1, 5), (31, 33), (1000, 257), (20011, 1024), (150001, 2500)])
def test_nearest_matches_oracle(pl, oracle, n_nodes, n_query):
    rng = np.random.default_rng(n_nodes)
    nodes = rng.uniform(-50, 50, (n_nodes, 2)).astype(np.float32)
    q = rng.uniform(-60, 60, (n_query, 2)).astype(np.float32)
    # force exact ties: duplicate nodes, and queries equidistant from mirrored nodes
    if n_nodes > 20:
        nodes[10] = nodes[3]
        nodes[n_nodes - 1] = nodes[3]
        q[0] = nodes[3]
        nodes[5] = (1.0, 2.0); nodes[6] = (-1.0, 2.0); q[1] = (0.0, 2.0)
    idx, d2 = pl.nearest_dev(torch.from_numpy(nodes).cuda(), torch.from_numpy(q).cuda()); pl.sync()
    want_idx, want_d2 = oracle.nearest(nodes, q)
    np.testing.assert_array_equal(idx.cpu().numpy(), want_idx)   # ties -> lowest index
    np.testing.assert_array_equal(d2.cpu().numpy(), want_d2)     # same float32 bits


def test_find_exact_matches_oracle(pl, oracle):
    rng = np.random.default_rng(3)
    keys = rng.normal(size=(5000, 6))
    keys[100] = keys[40]          # duplicate: first match (40) must win
    keys[4999] = keys[40]
    q = np.concatenate([keys[[40, 100, 0, 4998, 4999]], rng.normal(size=(200, 6))])
    q[7] = keys[9]; q[7, 5] = np.nextafter(q[7, 5], 1.0)  # one ulp off: not equal
    got = pl.find_exact_dev(torch.from_numpy(keys).cuda(), torch.from_numpy(q).cuda()); pl.sync()
    want = oracle.find_exact(keys, q)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    assert list(want[:5]) == [40, 40, 0, 4998, 40] and want[7] == -1
