"""Throughput of the p-nonconforming device path (h3d_set_mesh_p) on one B200: DOF-updates/s of RK3 steps on a curved periodic box
whose elements have random anisotropic orders in [lo, hi].  Not the headline bench (bench.py): a measurement aid for DESIGN 5b.

    python scripts/bench_mixed.py --ne 24 --lo 3 --hi 7 --steps 20 --warmup 3
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

from horses3d_b200.capi import GpuApi                      # noqa: E402
from horses3d_b200.dgsem import DGSem                      # noqa: E402
from horses3d_b200.physics import make_physics             # noqa: E402
import mixed_cases as MC                                   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ne", type=int, default=24)
    ap.add_argument("--lo", type=int, default=3)
    ap.add_argument("--hi", type=int, default=7)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    phys = make_physics(flow="NS", mach=0.08, reynolds=1600.0, riemann="roe")
    mesh = MC.periodic_box(a.ne, a.lo, a.hi, seed=7, amp=0.1)
    api = GpuApi()
    sem = DGSem(api, mesh, phys)
    sem.set_Q(MC.smooth_state(sem, phys.Mach))
    dt = 0.2 * min(sem.MaxTimeStep(0.4, 0.4))
    for _ in range(a.warmup):
        sem.TakeRK3Step(0.0, dt)
    api.call("synchronize")
    l0 = api.kernel_launches()
    api.call("timer_begin")
    for _ in range(a.steps):
        sem.TakeRK3Step(0.0, dt)
    ms = C.c_double()
    api.call("timer_end", C.byref(ms))
    assert not sem.checkForNan()
    n_face_nodes = int(np.prod(sem.face_orders[:, :2] + 1, axis=1).sum())
    print(json.dumps({"metric": "DOF-updates/s (p-nonconforming mesh, NS/BR1/Roe RK3)", "value": sem.NDOF * 3 * a.steps / (ms.value * 1e-3),
                      "unit": "DOF-updates/s", "ms_per_step": ms.value / a.steps, "ndof": sem.NDOF, "face_nodes": n_face_nodes,
                      "elements": sem.nElem, "orders": [a.lo, a.hi], "gpu_launches": api.kernel_launches() - l0, "dtype": "f64"}))


if __name__ == "__main__":
    main()
