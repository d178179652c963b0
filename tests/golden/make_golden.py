"""Regenerates tests/golden/*.npz: residuals of the CPU oracle (oracle/h3d_oracle.cpp, itself pinned to the reference's
regression values by tests/test_oracle_pins.py) on small generated meshes.  The Fortran reference cannot run in this
container, so these are oracle outputs, frozen so that neither the oracle nor the device path can drift unnoticed.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from horses3d_b200.dgsem import DGSem  # noqa: E402
from horses3d_b200.hostmesh import GAUSS, GAUSSLOBATTO  # noqa: E402
from horses3d_b200.physics import make_physics  # noqa: E402
from oracle.oracle_api import OracleApi  # noqa: E402
from parity import channel_state, get_mesh, perturbed_tgv  # noqa: E402

CASES = {
    "tgv_ns_p3": dict(ne=2, N=3, nodes=GAUSS, bc=None, kw=dict(flow="NS", mach=0.08, reynolds=1600.0)),
    "tgv_ns_p7": dict(ne=2, N=7, nodes=GAUSS, bc=None, kw=dict(flow="NS", mach=0.08, reynolds=1600.0)),
    "tgv_euler_split_pirozzoli_p4": dict(ne=2, N=4, nodes=GAUSSLOBATTO, bc=None, kw=dict(flow="Euler", mach=0.3, inviscid="split-form", averaging="pirozzoli")),
    "channel_ns_smagorinsky_p3": dict(ne=3, N=3, nodes=GAUSS, bc="channel", kw=dict(flow="NS", mach=0.3, reynolds=200.0, les="smagorinsky")),
    "channel_ns_br2_p2": dict(ne=2, N=2, nodes=GAUSS, bc="channel", kw=dict(flow="NS", mach=0.3, reynolds=200.0, viscous="BR2")),
    "channel_ns_ip_p2": dict(ne=2, N=2, nodes=GAUSS, bc="channel", kw=dict(flow="NS", mach=0.3, reynolds=200.0, viscous="IP")),
    "channel_ns_entropy_split_p3": dict(ne=2, N=3, nodes=GAUSSLOBATTO, bc="channel", kw=dict(flow="NS", mach=0.3, reynolds=200.0, inviscid="split-form", averaging="chandrasekar", riemann="central", gradient_variables="Entropy")),
    "tgv_ns_energy_p4": dict(ne=2, N=4, nodes=GAUSS, bc=None, kw=dict(flow="NS", mach=0.3, reynolds=200.0, gradient_variables="Energy")),
}


# p-nonconforming meshes: random anisotropic element orders on the curved, re-oriented box (tests/mixed_cases.py)
CASES_MIXED = {
    "box_ns_mixed_p2to4": dict(bc=None, lo=2, hi=4, kw=dict(flow="NS", mach=0.3, reynolds=200.0, riemann="roe")),
    "channel_ns_mixed_p2to4": dict(bc="channel", lo=2, hi=4, kw=dict(flow="NS", mach=0.3, reynolds=150.0, riemann="roe")),
    "box_euler_mixed_p1to5": dict(bc=None, lo=1, hi=5, kw=dict(flow="Euler", mach=0.3, riemann="standard roe")),
    "box_euler_split_pirozzoli_mixed_p2to5": dict(bc=None, lo=2, hi=5, nodes=GAUSSLOBATTO, kw=dict(flow="Euler", mach=0.3, inviscid="split-form", averaging="pirozzoli", riemann="roe")),
}


def run_mixed(api, name):
    import mixed_cases as MC
    c = CASES_MIXED[name]
    phys = make_physics(**c["kw"])
    mesh = MC.channel(phys, 2, c["lo"], c["hi"], seed=3) if c["bc"] else MC.periodic_box(2, c["lo"], c["hi"], seed=5, nodes=c.get("nodes", GAUSS))
    sem = DGSem(api, mesh, phys)
    sem.set_Q(MC.smooth_state(sem, phys.Mach))
    sem.ComputeTimeDerivative(0.0)
    out = sem.download(Q=True, QDot=True, gradients=bool(phys.computeGradients))
    sem.TakeRK3Step(0.0, 1.0e-3)
    out["Q_after_rk3"] = sem.Q()
    return out


def run(api, name):
    if name in CASES_MIXED:
        return run_mixed(api, name)
    c = CASES[name]
    phys = make_physics(**c["kw"])
    mesh = get_mesh(c["ne"], c["N"], c["nodes"], 0.1, True, bc=c["bc"], phys=phys)
    sem = DGSem(api, mesh, phys)
    sem.set_initial_condition((lambda x: channel_state(x, phys)) if c["bc"] else perturbed_tgv)
    sem.ComputeTimeDerivative(0.0)
    out = sem.download(Q=True, QDot=True, gradients=True)
    sem.TakeRK3Step(0.0, 1.0e-3)
    out["Q_after_rk3"] = sem.Q()
    return out


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name in list(CASES) + list(CASES_MIXED):
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        out = run(OracleApi(), name)
        np.savez_compressed(os.path.join(here, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})
