"""Host logic of the ORIENTED batch kernel's first test (csrc/pp_near_spans.h), checked on the CPU against the oracle's
lattice geometry: the structuring elements must cover every 4x4 block a lattice sample can land in."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("near_spans") / "test_near_spans")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", os.path.join(ROOT, "tests", "cpp", "test_near_spans.cpp"),
                           "-o", out])
    return out


@pytest.mark.parametrize("res,bins", [(0.1, 16), (0.25, 16), (0.2, 16), (0.125, 8), (0.1, 32), (0.5, 4)])
def test_spans_cover_every_sample_block(exe, res, bins):
    out = subprocess.run([exe, str(res), str(bins), "6000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "outside 0" in out.stdout
