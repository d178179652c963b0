"""Parity of the CUDA path (through the C ABI) with the CPU oracle on identical inputs.  Needs a B200.

Tolerance: the north-star bar is 1e-12 relative for the per-DOF time derivative in FP64.  "Relative" is
measured per equation against the max-norm of that equation's field (a per-DOF ratio is meaningless where the
field crosses zero).  The library is compiled without FMA contraction (as the reference's gfortran build) and keeps
the reference's accumulation order, so fields are expected to be BIT-IDENTICAL to the oracle; the tests assert a
ten-times tighter bound than the north star (1e-13) and scripts/parity_report.py records bit-equality.
"""
import numpy as np
import pytest

from horses3d_b200 import physics as P
from horses3d_b200.dgsem import DGSem, taylor_green_ic
from horses3d_b200.hostmesh import GAUSS, GAUSSLOBATTO
from horses3d_b200.physics import make_physics
from oracle.oracle_api import OracleApi
from parity import channel_state, get_mesh, perturbed_tgv, rel_err

pytestmark = pytest.mark.gpu

TOL_QDOT = 1.0e-13
# Entropy gradient variables take log(p) and log(rho) at every node (VariableConversion_NS.f90:233): CUDA's log and glibc's
# differ in the last bit for some arguments, and the derivative matrix amplifies that by O(N^2).  Those cases are held to the
# north-star bound itself.
TOL_ENTROPY = 1.0e-12


def tol_of(kw):
    # the same holds for WALE, which takes three real powers per node (pow differs in the last bits between CUDA and glibc)
    loose = kw.get("gradient_variables", "State").lower() == "entropy" or kw.get("les", "none").lower() == "wale"
    return TOL_ENTROPY if loose else TOL_QDOT


def run_pair(gpu_api_cls, mesh, phys, ic=perturbed_tgv):
    out = []
    for api in (OracleApi(), gpu_api_cls()):
        sem = DGSem(api, mesh, phys)
        sem.set_initial_condition(ic)
        sem.ComputeTimeDerivative(0.0)
        out.append((sem, sem.download(Q=True, QDot=True, gradients=True)))
    return out


CASES = [
    # (ne, N, nodes, amp, shuffle, physics kwargs)
    (3, 3, GAUSS, 0.0, False, dict(flow="NS", mach=0.08, reynolds=1600.0)),
    (3, 3, GAUSS, 0.1, True, dict(flow="NS", mach=0.08, reynolds=1600.0)),
    (2, 7, GAUSS, 0.1, True, dict(flow="NS", mach=0.08, reynolds=1600.0)),
    (3, 2, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=100.0)),
    (3, 1, GAUSS, 0.0, True, dict(flow="NS", mach=0.3, reynolds=100.0)),
    (2, 4, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=100.0, riemann="lax-friedrichs")),
    (2, 5, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=100.0, riemann="central")),
    (2, 6, GAUSS, 0.1, True, dict(flow="Euler", mach=0.3)),
    (2, 8, GAUSS, 0.1, True, dict(flow="NS", mach=0.1, reynolds=500.0)),
    (2, 9, GAUSS, 0.1, True, dict(flow="NS", mach=0.1, reynolds=500.0)),
    (3, 3, GAUSSLOBATTO, 0.1, True, dict(flow="Euler", mach=0.3, inviscid="split-form", averaging="pirozzoli")),
    (2, 7, GAUSSLOBATTO, 0.1, True, dict(flow="Euler", mach=0.3, inviscid="split-form", averaging="pirozzoli")),
    (2, 4, GAUSSLOBATTO, 0.1, True, dict(flow="Euler", mach=0.3, inviscid="split-form", averaging="kennedy-gruber", riemann="lax-friedrichs")),
    (2, 5, GAUSSLOBATTO, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, inviscid="split-form", averaging="standard")),
    (2, 9, GAUSSLOBATTO, 0.1, True, dict(flow="Euler", mach=0.3, inviscid="split-form", averaging="pirozzoli")),
    (2, 3, GAUSSLOBATTO, 0.0, False, dict(flow="NS", mach=0.3, reynolds=200.0)),
    # the remaining Riemann solvers and averages (SURVEY 8f "next"): K2 / K4 ingredients
    (3, 3, GAUSSLOBATTO, 0.1, True, dict(flow="Euler", mach=0.08, inviscid="split-form", averaging="chandrasekar", riemann="central", compute_gradients=True)),
    (2, 7, GAUSSLOBATTO, 0.1, True, dict(flow="Euler", mach=0.3, inviscid="split-form", averaging="chandrasekar", riemann="central")),
    (3, 3, GAUSSLOBATTO, 0.1, True, dict(flow="Euler", mach=0.3, inviscid="split-form", averaging="pirozzoli", riemann="standard roe")),
    (2, 4, GAUSSLOBATTO, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, inviscid="split-form", averaging="entropy conserving", riemann="standard roe")),
    (2, 5, GAUSSLOBATTO, 0.1, True, dict(flow="Euler", mach=0.3, inviscid="split-form", averaging="ducros", riemann="u-diss")),
    (2, 3, GAUSSLOBATTO, 0.1, True, dict(flow="Euler", mach=0.3, inviscid="split-form", averaging="morinishi", riemann="lax-friedrichs")),
    (2, 7, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, riemann="rusanov")),
    (3, 3, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, riemann="standard roe", averaging="kennedy-gruber")),
    (2, 4, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, riemann="roe-pike")),
    (2, 3, GAUSSLOBATTO, 0.1, True, dict(flow="Euler", mach=0.3, inviscid="split-form", averaging="pirozzoli", riemann="low dissipation roe")),
    (2, 5, GAUSSLOBATTO, 0.1, True, dict(flow="Euler", mach=0.3, inviscid="split-form", averaging="chandrasekar", riemann="matrix dissipation")),
    # BR2 and interior-penalty viscous discretizations: K7 / K8 / K9 ingredients
    (3, 3, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, viscous="BR2")),
    (2, 7, GAUSS, 0.1, True, dict(flow="NS", mach=0.08, reynolds=1600.0, viscous="BR2")),
    (3, 1, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=100.0, viscous="BR2", penalty_parameter=3.0)),
    (2, 9, GAUSS, 0.1, True, dict(flow="NS", mach=0.1, reynolds=500.0, viscous="BR2")),
    (2, 4, GAUSSLOBATTO, 0.1, True, dict(flow="NS", mach=0.08, reynolds=1600.0, inviscid="split-form", averaging="kennedy-gruber", riemann="standard roe", viscous="BR2")),
    (3, 3, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, viscous="IP")),
    (2, 7, GAUSS, 0.1, True, dict(flow="NS", mach=0.08, reynolds=1600.0, viscous="IP")),
    (2, 5, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=100.0, viscous="IP", ip_variant="NIPG", penalty_parameter=2.0)),
    (2, 2, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=100.0, viscous="IP", ip_variant="IIPG")),
    (3, 3, GAUSSLOBATTO, 0.1, True, dict(flow="NS", mach=0.08, reynolds=1600.0, inviscid="split-form", averaging="chandrasekar", riemann="roe-pike", viscous="IP")),
    # entropy / energy gradient variables: K10 / K11 ingredients
    (2, 7, GAUSSLOBATTO, 0.1, True, dict(flow="NS", mach=0.3, reynolds=10.0, inviscid="split-form", averaging="pirozzoli", gradient_variables="Energy")),
    (2, 7, GAUSSLOBATTO, 0.1, True, dict(flow="NS", mach=0.3, reynolds=10.0, inviscid="split-form", averaging="chandrasekar", riemann="matrix dissipation", gradient_variables="Entropy")),
    (3, 3, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, gradient_variables="Energy")),
    (3, 4, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, gradient_variables="Entropy")),
    (2, 9, GAUSS, 0.1, True, dict(flow="NS", mach=0.1, reynolds=500.0, gradient_variables="Entropy")),
    (2, 5, GAUSSLOBATTO, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, inviscid="split-form", averaging="chandrasekar", riemann="central", gradient_variables="Entropy", viscous="BR2")),
    (2, 3, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, gradient_variables="Energy", viscous="IP")),
]


@pytest.mark.parametrize("ne,N,nodes,amp,shuffle,kw", CASES)
def test_time_derivative_matches_oracle(gpu_api_cls, ne, N, nodes, amp, shuffle, kw):
    mesh = get_mesh(ne, N, nodes, amp, shuffle)
    (so, o), (sg, g) = run_pair(gpu_api_cls, mesh, make_physics(**kw))
    assert np.array_equal(o["Q"], g["Q"])                       # upload/download round trip is exact
    if kw.get("flow", "NS") != "Euler" or kw.get("compute_gradients"):
        for k in ("U_x", "U_y", "U_z"):
            assert rel_err(g[k], o[k]) < tol_of(kw), k
    assert rel_err(g["QDot"], o["QDot"]) < tol_of(kw)
    assert sg.api.kernel_launches() > 0


BC_CASES = [
    (3, 3, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0)),
    (2, 7, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0)),
    (3, 4, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, les="smagorinsky")),
    (3, 2, GAUSS, 0.0, False, dict(flow="Euler", mach=0.3, riemann="lax-friedrichs")),
    (3, 3, GAUSSLOBATTO, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, inviscid="split-form", averaging="pirozzoli", les="smagorinsky")),
    (3, 4, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, les="smagorinsky", les_wall_model="linear")),
    (2, 7, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, les="smagorinsky", les_wall_model="linear")),
    (3, 3, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, viscous="BR2")),
    (2, 7, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, viscous="BR2")),
    (3, 4, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, viscous="IP", les="smagorinsky")),
    (2, 7, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, viscous="IP")),
    (3, 3, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, les="wale")),
    (2, 7, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, les="vreman")),
    (3, 4, GAUSSLOBATTO, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, inviscid="split-form", averaging="pirozzoli", les="vreman", gradient_variables="Energy")),
    (3, 3, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, gradient_variables="Energy")),
    (3, 5, GAUSSLOBATTO, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, inviscid="split-form", averaging="chandrasekar", riemann="central", gradient_variables="Entropy")),
    (3, 4, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, gradient_variables="Entropy", les="smagorinsky")),
    (2, 7, GAUSS, 0.1, True, dict(flow="NS", mach=0.3, reynolds=200.0, gradient_variables="Energy", les="smagorinsky", les_wall_model="linear")),
]


@pytest.mark.parametrize("ne,N,nodes,amp,shuffle,kw", BC_CASES)
def test_boundary_conditions_and_les_match_oracle(gpu_api_cls, ne, N, nodes, amp, shuffle, kw):
    """Config-5 ingredients: inflow, outflow, no-slip (moving/fixed), free-slip, Smagorinsky."""
    phys = make_physics(**kw)
    mesh = get_mesh(ne, N, nodes, amp, shuffle, bc="channel", phys=phys)
    assert (mesh.array("faceType") == P.FACE_BOUNDARY).sum() == 6 * ne * ne
    if phys.les_wall_model:
        dw = mesh.wall_distances().array("dWall")
        assert dw.min() >= 0.0 and (0.4 * dw.min() < 0.2 * (mesh.array("volume").max() / (N + 1) ** 3) ** (1 / 3))   # the limiter is active
    (so, o), (sg, g) = run_pair(gpu_api_cls, mesh, phys, ic=lambda x: channel_state(x, phys))
    if kw.get("flow", "NS") != "Euler":
        for k in ("U_x", "U_y", "U_z"):
            assert rel_err(g[k], o[k]) < tol_of(kw), k
    assert np.abs(o["QDot"]).max() > 1e-3
    assert rel_err(g["QDot"], o["QDot"]) < tol_of(kw)
    for api in (so, sg):
        api.TakeRK3Step(0.0, 1e-3)
    assert rel_err(sg.Q(), so.Q()) < tol_of(kw)


@pytest.mark.parametrize("dims,N", [((3, 1, 2), 3), ((1, 1, 1), 7), ((2, 2, 1), 2), ((1, 2, 1), 4)])
def test_elements_that_are_their_own_periodic_neighbours_match_oracle(gpu_api_cls, dims, N):
    """Smallest meshes: with one element along a periodic direction a face has the SAME element on both sides (the
    reference's CylinderNSpol3_1elem_y.mesh situation); down to a single element that is its own neighbour six times."""
    from horses3d_b200.hostmesh import HostMesh
    mesh = HostMesh.box(dims[0], ney=dims[1], nez=dims[2], amp=0.05, bFaceOrder=2).connect().geometry(N, GAUSS)
    assert mesh.nElem == dims[0] * dims[1] * dims[2] and mesh.nFaces == 3 * mesh.nElem
    (so, o), (sg, g) = run_pair(gpu_api_cls, mesh, make_physics(flow="NS", mach=0.3, reynolds=100.0))
    for k in ("U_x", "U_y", "U_z", "QDot"):
        assert rel_err(g[k], o[k]) < TOL_QDOT, k
    for api in (so, sg):
        api.TakeRK3Step(0.0, 1e-3)
    assert rel_err(sg.Q(), so.Q()) < TOL_QDOT


def test_surface_integrals_match_oracle(gpu_api_cls):
    """ScalarSurfaceIntegral / VectorSurfaceIntegral of every kind on every zone of the channel mesh, after an RK step
    (prolonged updated state, gradients of the last stage as the reference has them)."""
    phys = make_physics(flow="NS", mach=0.3, reynolds=200.0)
    mesh = get_mesh(3, 4, GAUSS, 0.1, True, bc="channel", phys=phys)
    sems = []
    for api in (OracleApi(), gpu_api_cls()):
        sem = DGSem(api, mesh, phys)
        sem.set_initial_condition(lambda x: channel_state(x, phys))
        sem.TakeRK3Step(0.0, 1.0e-3)
        sems.append(sem)
    so, sg = sems
    nonzero = 0
    for zone in range(len(mesh.bcs)):
        for kind in range(8):
            a, b = np.atleast_1d(so.SurfaceIntegral(zone, kind)), np.atleast_1d(sg.SurfaceIntegral(zone, kind))
            assert np.abs(a - b).max() <= 1e-12 * max(np.abs(a).max(), 1.0), (zone, kind, a, b)
            nonzero += int(np.abs(a).max() > 1e-3)
    assert nonzero > 20
    assert abs(so.surface_monitor(0, "drag", [1.0, 0.0, 0.0], reference_surface=1.0) - sg.surface_monitor(0, "drag", [1.0, 0.0, 0.0], reference_surface=1.0)) < 1e-11
    e = gpu_api_cls()
    with pytest.raises(RuntimeError):
        DGSem(e, get_mesh(2, 2, GAUSS), make_physics(flow="Euler")).SurfaceIntegral(0, P.SURF_TOTAL_FORCE)


def test_asynchronous_snapshot(gpu_api_cls):
    """h3d_snapshot_begin / _end: the snapshot holds the state at the point of the call although the loop kept stepping."""
    phys = make_physics(flow="NS", mach=0.08, reynolds=1600.0)
    mesh = get_mesh(4, 3, GAUSS, 0.1, True)
    sem = DGSem(gpu_api_cls(), mesh, phys)
    sem.set_initial_condition(taylor_green_ic)
    sem.TakeRK3Step(0.0, 1e-3)
    Q1 = sem.Q()
    sem.snapshot_begin()
    with pytest.raises(RuntimeError):
        sem.snapshot_begin()
    for _ in range(3):
        sem.TakeRK3Step(0.0, 1e-3)
    snap = sem.snapshot_end()
    assert np.array_equal(snap, Q1) and not np.array_equal(sem.Q(), Q1)
    with pytest.raises(RuntimeError):
        sem.snapshot_end()


def test_statistics_match_oracle(gpu_api_cls):
    """StatisticsMonitor_UpdateValues: running averages over five steps, with a reset in between."""
    phys = make_physics(flow="NS", mach=0.3, reynolds=200.0)
    mesh = get_mesh(3, 3, GAUSS, 0.1, True)
    out = []
    for api in (OracleApi(), gpu_api_cls()):
        sem = DGSem(api, mesh, phys)
        sem.set_initial_condition(perturbed_tgv)
        sem.TakeRK3Step(0.0, 1.0e-2); sem.UpdateStatistics()
        sem.TakeRK3Step(0.0, 1.0e-2); sem.UpdateStatistics(reset=True)
        samples = []
        for _ in range(4):
            sem.TakeRK3Step(0.0, 1.0e-2); sem.UpdateStatistics()
            Q = sem.Q(); samples.append(Q[..., 1] / Q[..., 0])
        data, ns = sem.Statistics()
        assert ns == 5 and data.shape[-1] == 29
        out.append((data, np.mean(samples[-4:], axis=0)))
    (do, _), (dg, _) = out
    assert np.abs(do - dg).max() <= 1e-13 * np.abs(do).max()
    for v in range(29):
        assert np.abs(do[..., v]).max() > 0 or v in ()


def test_probes_match_oracle(gpu_api_cls):
    """Probe_Update: every variable at points inside, on a face of and at a corner of curved, re-oriented elements."""
    from horses3d_b200 import probes
    phys = make_physics(flow="NS", mach=0.3, reynolds=200.0)
    mesh = get_mesh(3, 4, GAUSS, 0.1, True)
    vals = []
    for api in (OracleApi(), gpu_api_cls()):
        sem = DGSem(api, mesh, phys)
        sem.set_initial_condition(perturbed_tgv)
        sem.TakeRK3Step(0.0, 1.0e-3)
        pts = [[1.0, 2.0, 3.0], [2.0 * np.pi / 3.0, 1.0, 1.0], [0.3, 5.9, 4.4], [3.1, 3.2, 0.05]]
        pr = [probes.Probe(sem, x, v) for x in pts for v in probes.VARIABLES]
        assert all(p.active for p in pr)
        # the located point maps back onto the requested position
        X = sem.node_coordinates()
        for p, x in zip(pr[::len(probes.VARIABLES)], pts):
            assert np.abs(np.einsum("kjic,i,j,k->c", X[p.eID], p.l[0], p.l[1], p.l[2]) - np.array(x)).max() < 1e-11
        vals.append(probes.evaluate(sem, pr))
    assert np.abs(vals[0]).max() > 0.1
    assert np.abs(vals[0] - vals[1]).max() <= 1e-13 * np.abs(vals[0]).max()


@pytest.mark.parametrize("scheme", ["RK3", "RK5", "Euler", "LSERK14-4", "SSPRK33", "SSPRK43"])
def test_rk_steps_and_monitors_match_oracle(gpu_api_cls, scheme):
    mesh = get_mesh(4, 3, GAUSS, 0.1, True)
    phys = make_physics(flow="NS", mach=0.08, reynolds=1600.0)
    recs = []
    for api in (OracleApi(), gpu_api_cls()):
        sem = DGSem(api, mesh, phys)
        sem.set_initial_condition(taylor_green_ic)
        recs.append((sem.integrate(10, cfl=0.4, dcfl=0.4, scheme=scheme), sem.Q()))
    (ro, Qo), (rg, Qg) = recs
    for a, b in zip(ro, rg):
        assert abs(a["dt"] - b["dt"]) <= 1e-13 * max(a["dt"], 1e-300)
        assert np.abs(a["residuals"] - b["residuals"]).max() < 1e-10 * np.abs(a["residuals"]).max()
        for k in ("kinetic energy", "kinetic energy rate", "enstrophy"):
            assert abs(a[k] - b[k]) < 1e-10 * max(abs(a[k]), 1e-300), k
    assert rel_err(Qg, Qo) < 1e-12


@pytest.mark.parametrize("scheme", ["SSPRK33", "SSPRK43"])
def test_stage_limiter_matches_oracle(gpu_api_cls, scheme):
    """stage_limiter (ExplicitMethods.f90:1755-1847) after every SSPRK stage: same limited elements, same values."""
    from test_oracle_pins import limiter_case
    (so, Q1o), (sg, Q1g) = limiter_case(OracleApi(), scheme), limiter_case(gpu_api_cls(), scheme)
    (_, Q1n) = limiter_case(gpu_api_cls(), scheme, limited=False)
    assert np.abs(Q1g - Q1n).max() > 1e-3                          # the limiter acted on the device
    assert rel_err(Q1g, Q1o) < TOL_QDOT
    assert rel_err(sg.Q(), so.Q()) < TOL_QDOT


@pytest.mark.parametrize("kw", [dict(flow="NS", mach=0.3, reynolds=200.0),
                                dict(flow="NS", mach=0.3, reynolds=200.0, gradient_variables="Energy", les="smagorinsky"),
                                dict(flow="NS", mach=0.3, reynolds=200.0, gradient_variables="Entropy", inviscid="split-form", averaging="chandrasekar", riemann="central"),
                                dict(flow="Euler", mach=0.3)])
def test_volume_monitors_match_oracle(gpu_api_cls, kw):
    """Every ScalarVolumeIntegral kind behind the volume monitors (VolumeMonitor.f90:297-330), after an RK3 step on a curved
    mesh with boundaries: volumes, kinetic energy (rate), enstrophy, entropy (rate, balance), math entropy, internal energy,
    mean velocity."""
    phys = make_physics(**kw)
    nodes = GAUSSLOBATTO if kw.get("inviscid") == "split-form" else GAUSS
    mesh = get_mesh(3, 4, nodes, 0.1, True, bc="channel", phys=phys)
    vals = []
    for api in (OracleApi(), gpu_api_cls()):
        sem = DGSem(api, mesh, phys)
        sem.set_initial_condition(lambda x: channel_state(x, phys))
        sem.TakeRK3Step(0.0, 1.0e-3)
        sem.ComputeTimeDerivative(1.0e-3)
        names = [k for k in sem.VOLUME_MONITORS if phys.flowIsNavierStokes or k not in ("enstrophy", "entropy balance", "kinetic energy balance")]
        vals.append(np.array([sem.volume_monitor(k) for k in names]))
    assert np.abs(vals[0]).min() > 0.0
    assert (np.abs(vals[1] - vals[0]) <= 1e-11 * np.abs(vals[0])).all(), (names, vals[0], vals[1] - vals[0])


@pytest.mark.parametrize("scheme", ["RK3", "RK5"])
def test_stagewise_step_with_time_dependent_source_matches_oracle(gpu_api_cls, scheme):
    """h3d_rk_stage with the source updated at every stage time (the K3 manufactured source of the reference's
    NavierStokes/Convergence case), P=7; and rk_step == its stages in a row, bit for bit."""
    from convergence_case import state_source_in_point
    mesh = get_mesh(2, 7, GAUSS, 0.1, True)
    phys = make_physics(flow="NS", mach=0.3, reynolds=10.0)
    args = (phys.gammaMinus1, phys.gammaM2, phys.mu, phys.kappa)
    out = []
    for api in (OracleApi(), gpu_api_cls()):
        sem = DGSem(api, mesh, phys)
        X = sem.node_coordinates() / np.pi            # the manufactured state stays positive on [0, 2]^3
        exact = lambda t: state_source_in_point(X[..., 0], X[..., 1], X[..., 2], t, *args)
        sem.set_Q(exact(0.0)[0])
        rec = sem.integrate(3, dt=1.0e-3, scheme=scheme, source=lambda t: exact(t)[1], ctd_after_step=True, monitors=False)
        out.append((rec, sem.download(Q=True, QDot=True)))
    (ro, o), (rg, g) = out
    assert np.abs(o["QDot"]).max() > 1.0
    assert rel_err(g["Q"], o["Q"]) < TOL_QDOT and rel_err(g["QDot"], o["QDot"]) < TOL_QDOT
    assert np.abs(ro[-1]["residuals"] - rg[-1]["residuals"]).max() < 1e-12 * np.abs(ro[-1]["residuals"]).max()
    # constant source: one call per step equals the stages one by one
    S = exact(0.0)[1]
    res = []
    for staged in (False, True):
        sem = DGSem(gpu_api_cls(), mesh, phys)
        sem.set_Q(exact(0.0)[0])
        sem.set_source(S)
        step = sem.TakeRK3Step if scheme == "RK3" else sem.TakeRK5Step
        step(0.0, 1.0e-3, source=(lambda t: S) if staged else None)
        res.append(sem.Q())
    assert np.array_equal(res[0], res[1])


def test_k1_taylor_green_on_gpu(gpu_api_cls):
    """The reference's own TaylorGreen regression (K1) run on the device path."""
    mesh = get_mesh(32, 3, GAUSS)
    sem = DGSem(gpu_api_cls(), mesh, make_physics(flow="NS", mach=0.08, reynolds=1600.0, riemann="roe"))
    sem.set_initial_condition(taylor_green_ic)
    rec = sem.integrate(5, cfl=0.4, dcfl=0.4)[-1]
    res = np.array([1.6417830052388520E-05, 1.2677577061211545E-01, 1.2677577048633804E-01, 2.4981129585617484E-01, 6.2174425106488129E-01])
    assert np.abs(rec["residuals"] - res).max() < 1.0e-7
    assert abs(rec["kinetic energy"] - 1.2499879367819486E-01) < 1.0e-11
    assert abs(rec["kinetic energy rate"] - (-4.2807806718622574E-04)) < 1.0e-11
    assert abs(rec["enstrophy"] - 3.7499683882517909E-01) < 1.0e-11


def test_k5_cylinder_regressions_on_gpu(gpu_api_cls):
    """The reference's Cylinder and CylinderSmagorinsky regressions (K5, K5b: the curved CylinderNSpol3.mesh with no-slip, free-slip,
    inflow and outflow zones, 100 CFL-limited RK3 steps, drag / lift monitors, wake probe) run on the DEVICE path against the values
    and tolerances of Solver/test/NavierStokes/Cylinder*/SETUP/ProblemFile.f90 (mesh copied to tests/golden/)."""
    from test_oracle_pins import _cylinder_100_steps
    got, cd, cl, wake_u = _cylinder_100_steps(api=gpu_api_cls())
    res = np.array([8.8131248889811715E+00, 1.7608838068776613E+01, 1.9037533106262516E-01, 2.4301352846288605E+01, 2.4063786464536835E+02])
    assert np.abs((got - res) / res).max() < 1.0e-11
    assert abs(cd - 3.4573345486345943E+01) < 1.0e-11 * 35.0 and abs(cl - (-4.6800322917661674E-04)) < 1.0e-11
    assert abs(wake_u - 1.0965307794823676E-08) < 1.0e-11
    got, cd, cl, wake_u = _cylinder_100_steps(api=gpu_api_cls(), les="smagorinsky", les_wall_model="linear")
    res = np.array([7.58705681758851, 15.5542852761418, 0.231394835496677, 20.0848567943827, 207.594579145771])
    assert np.abs(got - res).max() < 1.0e-7
    assert abs(cd - 34.9438869828619) < 1.0e-11 * 35.0 and abs(cl - (-1.582092121135137E-004)) < 1.0e-11
    assert abs(wake_u - 9.867445291005896E-009) < 1.0e-11


def test_k4_box_around_circle_regressions_on_gpu(gpu_api_cls):
    """The reference's Euler/BoxAroundCircle (StandardDG) and BoxAroundCirclePirozzoli (SplitDG) regressions (K4b, K4: 1000 steps on
    the curved mesh around a cylinder) on the DEVICE path: final time, residuals, force and pressure monitors, wake probe."""
    from test_oracle_pins import k4_case, k4b_case
    k4b_case(gpu_api_cls())
    k4_case(gpu_api_cls())


@pytest.mark.parametrize("case", ["state", "energy", "entropy"])
def test_k3_k10_convergence_p7_on_gpu(gpu_api_cls, case):
    """The reference's Convergence, Convergence_energy and Convergence_entropy regressions (K3, K10: P=7, manufactured solution
    with a time-dependent source, t = 1) run on the device path, on the generated equivalent of UnitCube4x4.mesh: residuals,
    L2 state errors, L2 QDot errors and the entropy-rate monitor against the reference's values at its 1e-11 tolerance."""
    from test_oracle_pins import convergence_case
    kw, nodes, cfl, res, e0, q0, er = {
        "state": (dict(), GAUSS, 0.5,
                  [6.2801762330611588E-01, 1.8889640334957627E+00, 2.5256897695536247E+00, 4.4142472296503827E+00, 2.5163928650671146E+00],
                  [1.0983475326313417E-06, 1.4788256133056976E-06, 4.5499827613507929E-07, 9.0819927730318800E-07, 2.5402026557722347E-06],
                  [1.1342700947907287E-05, 1.1638989807665964E-05, 3.5549224957856481E-06, 1.0769093706709006E-05, 2.0658954210939997E-05],
                  8.7517056213126665E-08),
        "energy": (dict(inviscid="split-form", averaging="pirozzoli", riemann="roe", gradient_variables="Energy"), GAUSSLOBATTO, 1.0,
                   [6.2838924111412731E-01, 1.8880402299553984E+00, 2.5257816906017094E+00, 4.4137338696617938E+00, 2.5153751482782658E+00],
                   [6.7174769052792914E-06, 7.8936751849804373E-06, 2.8203611177044557E-06, 6.6884144365250127E-06, 1.4116624697295942E-05],
                   [1.2370219075207763E-04, 1.1813999720732922E-04, 4.0653170754839378E-05, 1.2358578587070169E-04, 2.1723315094097264E-04],
                   3.4973252274750376E-07),
        "entropy": (dict(inviscid="split-form", averaging="chandrasekar", riemann="matrix dissipation", gradient_variables="Entropy"), GAUSSLOBATTO, 1.0,
                    [6.2929844029491377E-01, 1.8894710750845625E+00, 2.5264755519384234E+00, 4.4146672499485859E+00, 2.5157533917446182E+00],
                    [3.3102799903292017E-05, 3.6405407003671022E-05, 1.5957629488942332E-05, 3.4132613560424396E-05, 6.6811198249671421E-05],
                    [9.2273153574772720E-04, 9.6259031804949867E-04, 3.8474183309141608E-04, 8.7717012428247660E-04, 1.6750805236722724E-03],
                    3.6207469616966779E-07)}[case]
    rec, err, qerr, sem = convergence_case(gpu_api_cls(), nodes=nodes, cfl=cfl, generated=True, **kw)
    assert abs(rec["t"] - 1.0) < 1e-13
    assert np.abs(rec["residuals"] - np.array(res)).max() < 1.0e-11
    assert np.abs(err - np.array(e0)).max() < 1.0e-11
    assert np.abs(qerr - np.array(q0)).max() < 1.0e-11
    assert abs(sem.volume_monitor("entropy rate") - er) < 1.0e-11


def test_1000_rk3_steps_traces_match_oracle(gpu_api_cls):
    """North-star acceptance: the L2 solution and kinetic-energy monitor traces after 1000 RK3 steps agree within 1e-10
    relative (here between the device and the restated reference; curved periodic mesh, P=3, CFL-limited steps)."""
    mesh = get_mesh(4, 3, GAUSS, 0.1, True)
    phys = make_physics(flow="NS", mach=0.08, reynolds=1600.0)
    recs = []
    for api in (OracleApi(), gpu_api_cls()):
        sem = DGSem(api, mesh, phys)
        sem.set_initial_condition(taylor_green_ic)
        recs.append((sem.integrate(1000, cfl=0.4, dcfl=0.4), sem.Q()))
    (ro, Qo), (rg, Qg) = recs
    assert len(ro) == len(rg) == 1001
    t = np.array([[a["t"], b["t"]] for a, b in zip(ro, rg)])
    ke = np.array([[a["kinetic energy"], b["kinetic energy"]] for a, b in zip(ro, rg)])
    ens = np.array([[a["enstrophy"], b["enstrophy"]] for a, b in zip(ro, rg)])
    assert np.abs(t[:, 0] - t[:, 1]).max() <= 1e-10 * t[-1, 0]
    assert np.abs(ke[:, 0] - ke[:, 1]).max() <= 1e-10 * np.abs(ke[:, 0]).max()
    assert np.abs(ens[:, 0] - ens[:, 1]).max() <= 1e-10 * np.abs(ens[:, 0]).max()
    l2 = np.sqrt(((Qg - Qo) ** 2).sum() / (Qo ** 2).sum())
    assert l2 <= 1e-10
    assert abs(ke[-1, 0] - ke[0, 0]) > 1e-6 * ke[0, 0]              # the flow evolved


def test_free_stream_preservation_full_size_p7(gpu_api_cls):
    """Size-independent property at the benchmark polynomial order: a uniform state has zero residual on a curved,
    randomly re-oriented mesh (metric identities + all eight face rotations)."""
    mesh = get_mesh(6, 7, GAUSS, 0.1, True)
    sem = DGSem(gpu_api_cls(), mesh, make_physics(flow="NS", mach=0.3, reynolds=100.0))
    Q = np.zeros((mesh.nElem, 8, 8, 8, 5)); Q[...] = [1.0, 0.3, -0.2, 0.5, 10.0]
    sem.set_Q(Q)
    sem.ComputeTimeDerivative(0.0)
    assert np.abs(sem.QDot()).max() < 2e-9
    sem.TakeRK3Step(0.0, 1e-3)
    assert np.abs(sem.Q() - Q).max() < 1e-11


def test_residual_is_idempotent_and_source_is_added(gpu_api_cls):
    mesh = get_mesh(3, 3, GAUSS, 0.1, True)
    sem = DGSem(gpu_api_cls(), mesh, make_physics(flow="NS", mach=0.08, reynolds=1600.0))
    sem.set_initial_condition(perturbed_tgv)
    sem.ComputeTimeDerivative(0.0); a = sem.QDot()
    sem.ComputeTimeDerivative(0.0); b = sem.QDot()
    assert np.array_equal(a, b)
    S = np.random.default_rng(0).standard_normal(a.shape)
    sem.set_source(S); sem.ComputeTimeDerivative(0.0); c = sem.QDot()
    assert np.array_equal(c, a + S)
    sem.set_source(None); sem.ComputeTimeDerivative(0.0)
    assert np.array_equal(sem.QDot(), a)


def test_nan_is_detected(gpu_api_cls):
    mesh = get_mesh(2, 2, GAUSS)
    sem = DGSem(gpu_api_cls(), mesh, make_physics())
    Q = taylor_green_ic(sem.node_coordinates())
    sem.set_Q(Q)
    assert not sem.checkForNan()
    Q[1, 0, 1, 2, 3] = np.nan
    sem.set_Q(Q)
    assert sem.checkForNan()


def test_errors_follow_the_reference_messages(gpu_api_cls):
    from horses3d_b200.capi import H3dError
    mesh = get_mesh(2, 2, GAUSS)
    bad = make_physics(); bad.riemann = 99
    with pytest.raises(H3dError, match="Riemann Solver not recognized"):
        DGSem(gpu_api_cls(), mesh, bad)
    with pytest.raises(H3dError, match="Gauss-Lobatto"):
        sem = DGSem(gpu_api_cls(), mesh, make_physics(flow="Euler", inviscid="split-form", averaging="pirozzoli"))
        sem.set_initial_condition(taylor_green_ic)
        sem.ComputeTimeDerivative(0.0)


def test_readiness_checks_refuse_incomplete_set_ups(gpu_api_cls):
    """a mesh with boundary faces needs its boundary table (a null table would be dereferenced on the device), the
    table must cover every zone of the mesh and hold known types, and LES needs the element volumes."""
    import ctypes as C
    from horses3d_b200.capi import H3dError, _ptr
    from horses3d_b200.hostmesh import NodalStorage
    phys = make_physics(flow="NS", mach=0.3, reynolds=200.0)
    mesh = get_mesh(2, 2, GAUSS, 0.1, True, bc="channel", phys=phys)
    api = gpu_api_cls()
    api.set_physics(phys); api.set_basis(NodalStorage(2, GAUSS)); api.set_mesh(mesh)
    Q = np.ascontiguousarray(channel_state(mesh.array("x").reshape(mesh.nElem, 3, 3, 3, 3), phys))
    api.call("upload_Q", _ptr(Q, np.float64))
    with pytest.raises(H3dError, match="boundary faces but h3d_set_boundary_conditions was not called"):
        api.call("compute_time_derivative", 0.0)
    with pytest.raises(H3dError, match="zone beyond this table"):
        api.set_boundary_conditions([P.BC_TYPES["inflow"]] * 2, np.zeros((2, 16)))
    with pytest.raises(H3dError, match="unknown boundary condition type"):
        api.set_boundary_conditions([77] * 6, np.zeros((6, 16)))
    sem = DGSem(gpu_api_cls(), mesh, phys)                        # the complete set-up works
    sem.set_initial_condition(lambda x: channel_state(x, phys))
    sem.ComputeTimeDerivative(0.0)
    assert np.isfinite(sem.QDot()).all()
