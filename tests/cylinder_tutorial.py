"""BASELINE configs[4]: tutorials/Cylinder (Cylinder_Re100.control:1-40) on its own GMSH mesh MESH/cyl_circ.msh, P=3, M 0.2, Re 100,
Roe, BR1, RK3, cfl = dcfl = 0.6, no-slip cylinder, free-slip z-planes (the control file's `left__right`), inflow, outflow -- with
the Smagorinsky model configs[4] asks for.  Shared by the CPU, single-GPU and multi-GPU tests.  The mesh file is the reference's,
copied to tests/golden/ (the GPU box has no /root/reference)."""
import math
import os

import numpy as np

from horses3d_b200.hostmesh import GAUSS, HostMesh
from horses3d_b200.physics import bc_parameters, make_physics

MESH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cyl_circ.msh")
ZONES = [("cylinder", "noslipwall"), ("left", "freeslipwall"), ("right", "freeslipwall"), ("inlet", "inflow"), ("outlet", "outflow")]


def physics(les=True):
    kw = dict(les="smagorinsky") if les else {}
    return make_physics(flow="NS", mach=0.2, reynolds=100.0, riemann="roe", **kw)


def mesh(phys):
    p_in = 1.0 / phys.gammaM2
    v_in = phys.Mach * math.sqrt(phys.gamma * p_in / 1.0)              # InflowBC.f90:172-186
    params = []
    for _, t in ZONES:
        if t == "inflow":
            params.append(bc_parameters("inflow", phys, rho=1.0, v=v_in, aoa_theta=0.0, aoa_phi=0.0, p=p_in))
        elif t == "outflow":
            params.append(bc_parameters("outflow", phys, p=p_in))
        else:
            params.append(bc_parameters(t, phys))
    return HostMesh.read(MESH).connect([(z, t, None) for z, t in ZONES], np.array(params))


def initial_condition(x, phys):
    """Uniform flow along x (AOA theta = phi = 0), the state the tutorial's Re 40 precursor starts from."""
    Q = np.zeros(x.shape[:-1] + (5,))
    Q[..., 0], Q[..., 1] = 1.0, 1.0
    Q[..., 4] = (1.0 / phys.gammaM2) / (phys.gamma - 1.0) + 0.5
    return Q


def run(sem, steps, cfl=0.3):
    """`steps` CFL-limited RK3 steps; returns the residuals and the drag / lift monitors of the control file.  The tutorial runs
    cfl = dcfl = 0.6 from the converged Re 40 field (restart = .true.); the impulsive start used here (no restart file) needs 0.3."""
    sem.set_Q(initial_condition(sem.node_coordinates(), sem.physics))
    res = sem.integrate(steps, cfl=cfl, dcfl=cfl, monitors=False)[-1]["residuals"]
    cd = sem.surface_monitor("cylinder", "drag", [1.0, 0.0, 0.0], reference_surface=1.0)
    cl = sem.surface_monitor("cylinder", "lift", [0.0, 1.0, 0.0], reference_surface=1.0)
    return np.array(res), cd, cl
