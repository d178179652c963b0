"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
   compute-sanitizer --tool racecheck python scripts/sanitize_case.py [N]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from horses3d_b200.capi import GpuApi  # noqa: E402
from horses3d_b200.dgsem import DGSem  # noqa: E402
from horses3d_b200.hostmesh import GAUSS, GAUSSLOBATTO  # noqa: E402
from horses3d_b200.physics import make_physics  # noqa: E402
from parity import channel_state, get_mesh, perturbed_tgv  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 7
for kw, nodes, bc in [(dict(flow="NS", mach=0.3, reynolds=200.0, les="smagorinsky"), GAUSS, "channel"),
                      (dict(flow="NS", mach=0.08, reynolds=1600.0), GAUSS, None),
                      (dict(flow="Euler", mach=0.3, inviscid="split-form", averaging="pirozzoli"), GAUSSLOBATTO, None),
                      (dict(flow="NS", mach=0.3, reynolds=200.0, viscous="BR2"), GAUSS, "channel"),
                      (dict(flow="NS", mach=0.3, reynolds=200.0, viscous="IP", gradient_variables="Energy"), GAUSS, "channel"),
                      (dict(flow="NS", mach=0.3, reynolds=200.0, inviscid="split-form", averaging="chandrasekar", riemann="central",
                            gradient_variables="Entropy"), GAUSSLOBATTO, None)]:
    phys = make_physics(**kw)
    mesh = get_mesh(3, N, nodes, 0.1, True, bc=bc, phys=phys)      # 27 elements: several tiles per CTA never happen, 1 tile each
    sem = DGSem(GpuApi(), mesh, phys)
    sem.set_initial_condition((lambda x: channel_state(x, phys)) if bc else perturbed_tgv)
    sem.TakeRK3Step(0.0, 1e-3)
    sem.TakeRK3Step(1e-3, 1e-3)
    if kw.get("viscous") == "BR2":          # the stage limiter and the second volume-integral pass ride along
        sem.enable_limiter(True, 1.0e-3)
        sem.integrate(1, dt=1e-3, scheme="SSPRK33", monitors=False)
        print("entropy rate", sem.volume_monitor("entropy rate"), "balance", sem.volume_monitor("entropy balance"))
    r = sem.ComputeMaxResiduals()
    print(kw.get("inviscid", "standard"), bc, "residuals", r, "dt", sem.MaxTimeStep(0.4, 0.4), "KE", sem.volume_monitors()["kinetic energy"])
    assert np.isfinite(r).all()
print("done")
