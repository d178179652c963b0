"""Cases of the p-nonconforming path shared by the CPU (emulated kernels) and GPU parity tests."""
import numpy as np

from horses3d_b200 import physics as P
from horses3d_b200.dgsem import DGSem, taylor_green_ic
from horses3d_b200.hostmesh import GAUSS, GAUSSLOBATTO, HostMesh
from horses3d_b200.physics import bc_parameters, make_physics


def random_orders(nElem, lo, hi, seed, anisotropic=True):
    rng = np.random.default_rng(seed)
    o = rng.integers(lo, hi + 1, size=(nElem, 3 if anisotropic else 1)).astype(np.int32)
    return o if anisotropic else np.repeat(o, 3, axis=1)


def periodic_box(ne=3, lo=2, hi=5, seed=7, nodes=GAUSS, anisotropic=True, amp=0.15):
    """Curved, randomly re-oriented periodic box (all eight face rotations) with random element orders: every projection type on
    both sides of the faces."""
    m = HostMesh.box(ne, amp=amp, shuffle=True, seed=seed).connect()
    return m.geometry_p(random_orders(m.nElem, lo, hi, seed, anisotropic), nodes)


CHANNEL_BCS = [("front", "periodic", "back"), ("back", "periodic", "front"), ("bottom", "noslipwall", None), ("top", "freeslipwall", None),
               ("left", "inflow", None), ("right", "outflow", None)]


def channel(phys, ne=3, lo=2, hi=4, seed=3, nodes=GAUSS, uniform=None):
    """Box with the four boundary conditions of the reference's cylinder cases and random element orders (uniform=N: the same mesh
    through the uniform-order geometry).  With the LES wall model the wall distances are computed too."""
    params = []
    for _, t, _c in CHANNEL_BCS:
        if t == "inflow":
            params.append(bc_parameters("inflow", phys, rho=1.0, v=1.0, aoa_theta=0.0, aoa_phi=0.0, p=1.0 / phys.gammaM2))
        elif t == "outflow":
            params.append(bc_parameters("outflow", phys, p=1.0 / phys.gammaM2))
        elif t == "periodic":
            params.append(np.zeros(16))
        else:
            params.append(bc_parameters(t, phys))
    m = HostMesh.box(ne, amp=0.1, shuffle=True, seed=seed).connect(CHANNEL_BCS, np.array(params))
    m = m.geometry(uniform, nodes, reference_order=True) if uniform else m.geometry_p(random_orders(m.nElem, lo, hi, seed), nodes)
    if phys.les_wall_model:
        m.wall_distances()
    return m


def smooth_state(sem, mach):
    x = sem.node_coordinates().reshape(-1, 3)
    Q = taylor_green_ic(x, p0=1.0 / (1.4 * mach ** 2)).reshape(-1, 5)
    Q[:, 0] += 0.05 * np.sin(x[:, 0]) * np.cos(x[:, 1] + 0.3) * np.cos(x[:, 2])
    Q[:, 3] += 0.1 * np.cos(x[:, 0]) * np.sin(x[:, 2] + 0.2)
    return Q


def run_case(api, mesh, phys, scheme="rk3", dt=2.0e-3, source=False, zone=None, state_from=None):
    """One residual, one RK step with a residual after it, the reductions; returns everything that is compared.  state_from =
    (single-domain DGSem, global element ids of this partition): the initial state is taken from the global one, element by element
    (periodic images of a node have different coordinates in different elements' patches only through the element that owns them, so
    evaluating the state function on the partition gives the same values -- this just makes it explicit)."""
    sem = DGSem(api, mesh, phys)
    if state_from is None:
        sem.set_Q(smooth_state(sem, phys.Mach))
    else:
        sem0, ge = state_from
        Q0 = smooth_state(sem0, phys.Mach)
        sem.set_Q(np.concatenate([Q0[sem0.elem_offset[e]:sem0.elem_offset[e + 1]] for e in ge]))
    if source:
        rng = np.random.default_rng(11)
        sem.set_source(0.01 * rng.standard_normal((sem.NDOF, 5)))
    out = {}
    sem.ComputeTimeDerivative(0.0)
    d = sem.download(QDot=True, gradients=bool(phys.computeGradients))
    out["QDot0"] = d["QDot"]
    if phys.computeGradients:
        out["Ux0"], out["Uy0"], out["Uz0"] = d["U_x"], d["U_y"], d["U_z"]
    sem._rk_step(sem.SCHEMES[scheme], 0.0, dt, True, None)
    d = sem.download(Q=True, QDot=True)
    out["Q1"], out["QDot1"] = d["Q"], d["QDot"]
    out["residuals"] = sem.ComputeMaxResiduals()
    out["dt"] = np.array(sem.MaxTimeStep(0.3, 0.2))
    kinds = [P.INT_VOLUME, P.INT_KINETIC_ENERGY, P.INT_KINETIC_ENERGY_RATE, P.INT_VELOCITY, P.INT_INTERNAL_ENERGY, P.INT_ENTROPY, P.INT_MATH_ENTROPY,
             P.INT_ENTROPY_RATE]
    if phys.computeGradients:
        kinds += [P.INT_ENSTROPHY, P.INT_ENTROPY_BALANCE, P.INT_KINETIC_ENERGY_BALANCE]
    out["integrals"] = np.array([sem.ScalarVolumeIntegral(k) for k in kinds])
    out["nan"] = np.array([sem.checkForNan()])
    if zone is not None:
        ks = [P.SURF_SURFACE, P.SURF_MASS_FLOW, P.SURF_FLOW_RATE, P.SURF_PRESSURE, P.SURF_VEC_SURFACE, P.SURF_PRESSURE_FORCE]
        if phys.computeGradients:
            ks += [P.SURF_TOTAL_FORCE, P.SURF_VISCOUS_FORCE]
        out["surface"] = np.concatenate([np.atleast_1d(sem.SurfaceIntegral(zone, k)) for k in ks])
    return sem, out


def compare(a, b, tol=0.0):
    """Largest difference relative to the field's max-norm, per key."""
    worst = {}
    for k in a:
        scale = max(np.abs(a[k]).max(), 1e-300)
        worst[k] = float(np.abs(np.asarray(a[k], dtype=float) - np.asarray(b[k], dtype=float)).max() / scale)
    bad = {k: v for k, v in worst.items() if v > tol}
    return worst, bad


def limiter_and_statistics_case(api, limited=True, minimum=0.05, scheme="ssprk33"):
    """Gauss-Lobatto mesh with random orders; a deep density pit and a pressure pit at corner nodes of a few elements (face nodes: the
    traces stay positive).  One SSPRK stage and two steps with the stage limiter, then three samples of the statistics monitor."""
    phys = make_physics(flow="NS", mach=0.3, reynolds=200.0, riemann="roe")
    sem = DGSem(api, periodic_box(3, 2, 4, seed=5, nodes=GAUSSLOBATTO, amp=0.1), phys)
    Q = smooth_state(sem, phys.Mach)
    for e in (0, 5, 13):
        a, b = sem.elem_offset[e], sem.elem_offset[e + 1] - 1          # first and last node of the element: corners
        vel, pr = Q[a, 1:4] / Q[a, 0], 0.4 * (Q[a, 4] - 0.5 * (Q[a, 1:4] ** 2).sum() / Q[a, 0])
        Q[a, :] = [1.0e-3, *(1.0e-3 * vel), pr / 0.4 + 0.5 * 1.0e-3 * (vel ** 2).sum()]
        Q[b, 4] = 0.5 * (Q[b, 1:4] ** 2).sum() / Q[b, 0] + 1.0e-4 / 0.4
    sem.set_Q(Q)
    if limited:
        sem.enable_limiter(True, minimum)
    code = sem.SCHEMES[scheme]
    sem.api.call("rk_stage", code, 0, 0.0, 1.0e-5)
    out = {"Q_one_stage": sem.Q()}
    if not limited:
        return sem, out          # the pits are not survivable without the limiter
    sem.set_Q(Q)
    sem.integrate(2, dt=1.0e-5, scheme=scheme, monitors=False)
    out["Q_two_steps"] = sem.Q()
    for k in range(3):
        sem.UpdateStatistics(reset=(k == 0))
        if k == 1:
            sem.snapshot_begin()                       # autosave beside the time loop: the state at this point ...
            before = sem.Q()
        sem.TakeRK3Step(0.0, 1.0e-5)
        if k == 1:
            out["snapshot"] = sem.snapshot_end()       # ... delivered after the loop has moved on
            assert np.array_equal(out["snapshot"], before) and not np.array_equal(out["snapshot"], sem.Q())
    data, ns = sem.Statistics()
    out["statistics"] = data
    out["samples"] = np.array([ns])
    return sem, out
