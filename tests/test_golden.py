"""Frozen oracle outputs (tests/golden/*.npz, made by tests/golden/make_golden.py): the oracle must still reproduce them
bit for bit (CPU), and so must the device path (GPU)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden import CASES, CASES_MIXED, run  # noqa: E402
from oracle.oracle_api import OracleApi  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def check(api, name, exact):
    ref = np.load(os.path.join(GOLDEN, name + ".npz"))
    got = run(api, name)
    for k in ref.files:
        if exact:
            assert np.array_equal(got[k], ref[k]), (name, k)
        else:
            tol = 1e-12 if "entropy" in name else 1e-13     # log() differs in the last bit between CUDA and glibc
            assert np.abs(got[k] - ref[k]).max() <= tol * max(np.abs(ref[k]).max(), 1e-300), (name, k)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_golden(name):
    check(OracleApi(), name, exact=True)


@pytest.mark.parametrize("name", list(CASES_MIXED))
def test_oracle_and_emulated_kernels_reproduce_golden_mixed(name):
    """p-nonconforming fixtures: the oracle and the device functors on the host-loop backend (tests/emu), bit for bit.  The device
    itself is checked against the same fixtures in tests/test_zz_gpu_mixed.py."""
    from emu.emu_api import EmuApi
    check(OracleApi(), name, exact=True)
    check(EmuApi(), name, exact=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_device_reproduces_golden(gpu_api_cls, name):
    check(gpu_api_cls(), name, exact=False)
