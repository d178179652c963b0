"""p-nonconforming path on the CPU: the kernel functors and the orchestration of horses3d_b200/csrc/h3d_mixed.cuh,
built for the host with a loop as the launcher (tests/emu), against the oracle.  Bit-exact: the functors keep the reference's
order of every sum and the host build disables FMA contraction as the device build does."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import mixed_cases as MC                                   # noqa: E402
from emu.emu_api import EmuApi                             # noqa: E402
from horses3d_b200.hostmesh import GAUSS, GAUSSLOBATTO     # noqa: E402
from horses3d_b200.physics import make_physics             # noqa: E402
from oracle import oracle_api                              # noqa: E402


def both(mesh_fn, phys, **kw):
    _, a = MC.run_case(oracle_api.OracleApi(), mesh_fn(), phys, **kw)
    _, b = MC.run_case(EmuApi(), mesh_fn(), phys, **kw)
    worst, bad = MC.compare(a, b)
    print(worst)
    assert not bad, bad


@pytest.mark.parametrize("riemann", ["roe", "lax-friedrichs", "standard roe", "central"])
def test_navier_stokes_periodic_box_random_anisotropic_orders(riemann):
    both(lambda: MC.periodic_box(3, 2, 5, seed=7), make_physics(flow="NS", mach=0.3, reynolds=200.0, riemann=riemann), source=True)


def test_larger_mesh_with_orders_up_to_nine():
    both(lambda: MC.periodic_box(6, 1, 9, seed=21), make_physics(flow="NS", mach=0.3, reynolds=400.0, riemann="roe"))


def test_edge_cases_self_periodic_element_and_orders_up_to_fifteen():
    """An element that is its own periodic neighbour on all six faces (one-element mesh), and the highest orders the path takes."""
    phys = make_physics(flow="NS", mach=0.3, reynolds=200.0, riemann="roe")
    both(lambda: MC.periodic_box(1, 2, 5, seed=3), phys)
    both(lambda: MC.periodic_box(2, 10, 15, seed=5), phys)


def test_euler_with_and_without_gradients():
    both(lambda: MC.periodic_box(3, 1, 4, seed=5), make_physics(flow="Euler", mach=0.3, riemann="roe"))
    both(lambda: MC.periodic_box(3, 1, 4, seed=5), make_physics(flow="Euler", mach=0.3, riemann="rusanov", compute_gradients=True))


def test_gauss_lobatto_nodes_and_isotropic_orders():
    both(lambda: MC.periodic_box(3, 2, 6, seed=9, nodes=GAUSSLOBATTO, anisotropic=False), make_physics(flow="NS", mach=0.3, reynolds=100.0, riemann="roe"))


@pytest.mark.parametrize("scheme", ["euler", "rk5", "lserk14-4", "ssprk33", "ssprk43"])
def test_runge_kutta_schemes(scheme):
    both(lambda: MC.periodic_box(2, 2, 4, seed=13), make_physics(flow="NS", mach=0.3, reynolds=200.0, riemann="roe"), scheme=scheme)


@pytest.mark.parametrize("gradvars", ["state", "entropy", "energy"])
def test_boundary_conditions_and_gradient_variables(gradvars):
    phys = make_physics(flow="NS", mach=0.3, reynolds=150.0, riemann="roe", gradient_variables=gradvars)
    for zone in (2, 4):      # no-slip wall, inflow
        both(lambda: MC.channel(phys), phys, zone=zone)


def test_k13_cylinder_different_orders_through_the_device_functors():
    """The reference's CylinderDifferentOrders regression (K13) through the emulated device path: 100 steps, residuals, drag, lift
    and the wake probe at the reference's tolerance -- and identical to the oracle."""
    from test_oracle_pins import K13, cylinder_different_orders
    api = EmuApi()
    _, res, cd, cl, wake_u = cylinder_different_orders(api)
    _, res0, cd0, cl0, wake0 = cylinder_different_orders()
    assert api.kernel_launches() >= 100 * 3 * 13
    assert np.abs(res - K13["residuals"]).max() < 1.0e-11 and abs(cd - K13["cd"]) < 1.2e-10 and abs(cl - K13["cl"]) < 1.0e-11 and abs(wake_u - K13["wake_u"]) < 1.0e-11
    assert np.array_equal(res, res0) and cd == cd0 and cl == cl0 and wake_u == wake0


SPLIT = [dict(flow="Euler", mach=0.3, inviscid="split-form", averaging="pirozzoli", riemann="roe"),
         dict(flow="Euler", mach=0.3, inviscid="split-form", averaging="standard", riemann="lax-friedrichs"),
         dict(flow="NS", mach=0.3, reynolds=200.0, inviscid="split-form", averaging="kennedy-gruber", riemann="roe"),
         dict(flow="NS", mach=0.3, reynolds=200.0, inviscid="split-form", averaging="chandrasekar", riemann="central", gradient_variables="entropy"),
         dict(flow="NS", mach=0.3, reynolds=200.0, inviscid="split-form", averaging="ducros", riemann="standard roe")]


@pytest.mark.parametrize("kw", SPLIT, ids=lambda k: k["flow"] + "-" + k["averaging"])
def test_split_form_on_random_orders(kw):
    """SplitDG on a p-nonconforming mesh: the two-point fluxes of every line with the element's own sharpD per direction."""
    both(lambda: MC.periodic_box(3, 2, 5, seed=17, nodes=GAUSSLOBATTO), make_physics(**kw))


def test_mixed_oracle_split_form_with_uniform_orders_equals_the_uniform_oracle():
    """Anchor of the split form on such meshes: with one order for every element the restatement is bit-identical to the uniform-order
    oracle, whose split form is pinned by the reference's TaylorGreenKEPEC, BoxAroundCirclePirozzoli and Convergence_* cases."""
    from horses3d_b200.dgsem import DGSem, taylor_green_ic
    from horses3d_b200.hostmesh import HostMesh
    for kw in SPLIT[::2]:
        phys, out = make_physics(**kw), []
        for mixed in (False, True):
            m = HostMesh.box(3, amp=0.15, shuffle=True).connect()
            m = m.geometry_p([4, 4, 4], GAUSSLOBATTO) if mixed else m.geometry(4, GAUSSLOBATTO, reference_order=True)
            sem = DGSem(oracle_api.OracleApi(), m, phys)
            sem.set_Q(taylor_green_ic(sem.node_coordinates().reshape(-1, 3), p0=1.0 / (1.4 * 0.3 ** 2)).reshape(sem._shape))
            sem.TakeRK3Step(0.0, 1.0e-3, ctd_after_step=True)
            out.append({k: v.reshape(-1, 5) for k, v in sem.download(Q=True, QDot=True).items()})
        assert all(np.array_equal(out[0][k], out[1][k]) for k in out[0]), kw


LES = [dict(les="smagorinsky", les_wall_model="linear"), dict(les="smagorinsky"), dict(les="wale"), dict(les="vreman")]


@pytest.mark.parametrize("kw", LES, ids=lambda k: k["les"] + ("-wall" if k.get("les_wall_model") else ""))
def test_les_models_on_random_orders(kw):
    """Filter widths from the element's / face's own orders (SpatialDiscretization.f90:420, 1378), wall distances at the packed nodes."""
    phys = make_physics(flow="NS", mach=0.3, reynolds=150.0, riemann="roe", **kw)
    both(lambda: MC.channel(phys), phys, zone=2)


def test_mixed_oracle_les_with_uniform_orders_equals_the_uniform_oracle():
    """Anchor of the LES models on such meshes: bit-identical to the uniform-order oracle (pinned by CylinderSmagorinsky / WALE / Vreman)."""
    from horses3d_b200.dgsem import DGSem
    for kw in LES[::2]:
        phys, out = make_physics(flow="NS", mach=0.3, reynolds=150.0, riemann="roe", **kw), []
        for uniform in (3, None):
            sem = DGSem(oracle_api.OracleApi(), MC.channel(phys, lo=3, hi=3, uniform=uniform), phys)
            sem.set_Q(MC.smooth_state(sem, 0.3).reshape(sem._shape))
            sem.TakeRK3Step(0.0, 1.0e-3, ctd_after_step=True)
            out.append({k: v.reshape(-1, 5) for k, v in sem.download(Q=True, QDot=True, gradients=True).items()})
        assert all(np.array_equal(out[0][k], out[1][k]) for k in out[0]), kw


@pytest.mark.parametrize("scheme", ["ssprk33", "ssprk43"])
def test_stage_limiter_and_statistics(scheme):
    _, a = MC.limiter_and_statistics_case(oracle_api.OracleApi(), scheme=scheme)
    _, b = MC.limiter_and_statistics_case(EmuApi(), scheme=scheme)
    _, c = MC.limiter_and_statistics_case(oracle_api.OracleApi(), limited=False, scheme=scheme)
    p = lambda A: 0.4 * (A[:, 4] - 0.5 * (A[:, 1:4] ** 2).sum(-1) / A[:, 0])
    assert c["Q_one_stage"][:, 0].min() < 0.01 and p(c["Q_one_stage"]).min() < 0.01            # without the limiter the pits survive the stage
    assert a["Q_one_stage"][:, 0].min() >= 0.05 * (1 - 1e-12) and p(a["Q_one_stage"]).min() >= 0.05 * (1 - 1e-9)
    assert a["samples"][0] == 3 and a["statistics"].shape[1] == 29
    worst, bad = MC.compare(a, b)
    assert not bad, bad


@pytest.mark.parametrize("kw", [dict(viscous="ip"), dict(viscous="ip", ip_variant="NIPG", penalty_parameter=3.0, gradient_variables="energy")],
                         ids=["sipg", "nipg-energy"])
def test_interior_penalty_on_random_orders(kw):
    """IP on a p-nonconforming mesh: local gradients prolonged before the lift, jumps 1/2 (UL - UR), penalty with max(Nf) and the faces' h
    (EllipticIP.f90:189-406, 678-761); anchored on the uniform oracle (CylinderIP, TaylorGreenKEPEC_IP) for uniform orders."""
    from horses3d_b200.dgsem import DGSem
    phys = make_physics(flow="NS", mach=0.3, reynolds=150.0, riemann="roe", **kw)
    both(lambda: MC.channel(phys), phys, zone=2)
    out = []
    for uniform in (3, None):
        sem = DGSem(oracle_api.OracleApi(), MC.channel(phys, lo=3, hi=3, uniform=uniform), phys)
        sem.set_Q(MC.smooth_state(sem, 0.3).reshape(sem._shape))
        sem.TakeRK3Step(0.0, 1.0e-3, ctd_after_step=True)
        out.append({k: v.reshape(-1, 5) for k, v in sem.download(Q=True, QDot=True, gradients=True).items()})
    assert all(np.array_equal(out[0][k], out[1][k]) for k in out[0])


def test_partitioned_interior_penalty_and_les():
    """Two more partitioned cases: the interior penalty (the MPI faces' h is the global face's) and Smagorinsky with the wall model (wall
    distances inherited from the global mesh)."""
    for kw in (dict(viscous="ip"), dict(les="smagorinsky", les_wall_model="linear")):
        phys = make_physics(flow="NS", mach=0.3, reynolds=150.0, riemann="roe", **kw)
        ref, got = _partitioned(2, lambda: MC.channel(phys, ne=4), phys, "metis", zone=2)
        worst, _ = MC.compare(ref, got)
        for k, v in worst.items():
            assert v <= (1e-13 if k in ("integrals", "surface") else 0.0), (kw, k, v)


def test_unsupported_configurations_are_refused():
    from horses3d_b200.capi import H3dError
    from horses3d_b200.dgsem import DGSem
    for kw in (dict(viscous="br2"),):
        with pytest.raises(H3dError):
            DGSem(EmuApi(), MC.periodic_box(2, 2, 3, seed=1, nodes=GAUSSLOBATTO), make_physics(flow="NS", mach=0.3, reynolds=100.0, riemann="roe", **kw))
    with pytest.raises(H3dError):      # the split form needs Gauss-Lobatto nodes
        DGSem(EmuApi(), MC.periodic_box(2, 2, 3, seed=1), make_physics(flow="Euler", mach=0.3, inviscid="split-form", averaging="pirozzoli"))


def _partitioned(world_size, mesh_fn, phys, method, zone=None, scheme="rk3"):
    """The same case on the single-domain oracle and on `world_size` emulated ranks (threads); returns both result sets with the
    partitioned element fields gathered into the global element order."""
    import threading
    from emu.emu_api import EmuWorld
    sem0, ref = MC.run_case(oracle_api.OracleApi(), mesh_fn(), phys, zone=zone, scheme=scheme)
    g = mesh_fn()
    part = g.partition(world_size, method)
    assert len(set(part)) == world_size
    world = EmuWorld(world_size)
    outs, errs = [None] * world_size, []

    def work(rank):
        try:
            m = g.extract(part, rank, inherit_geometry=True)
            sem, out = MC.run_case(EmuApi(world, rank), m, phys, zone=zone, scheme=scheme, state_from=(sem0, m.array("globalElem").copy()))
            outs[rank] = (sem, out, m.array("globalElem").copy(), int((m.array("faceType") == 3).sum()))
        except Exception as ex:      # a dead rank would leave the others at the barrier: open it
            errs.append(ex)
            world.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world_size)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    assert all(o[3] > 0 for o in outs)            # every rank has MPI faces
    got = {}
    for k, v in ref.items():
        if v.ndim == 2 and v.shape[0] == sem0.NDOF:        # element field: gather
            a = np.empty_like(v)
            for sem, out, ge, _ in outs:
                for l, e in enumerate(ge):
                    a[sem0.elem_offset[e]:sem0.elem_offset[e + 1]] = out[k][sem.elem_offset[l]:sem.elem_offset[l + 1]]
            got[k] = a
        else:                                               # reduced over the ranks: the same on all of them
            for _, out, _, _ in outs[1:]:
                assert np.array_equal(out[k], outs[0][1][k]), k
            got[k] = outs[0][1][k]
    return ref, got


@pytest.mark.parametrize("world_size,method", [(2, "metis"), (3, "block"), (4, "metis"), (8, "metis")])
def test_partitioned_mesh_reproduces_the_single_domain_oracle(world_size, method):
    """MPI faces of a p-nonconforming mesh: traces exchanged at the face order, mortar projection on each side's rank.  Element fields
    are bit-identical to the single-domain oracle (the partitions inherit the global geometry); the all-reduced scalars agree to
    the round-off of a different summation order."""
    phys = make_physics(flow="NS", mach=0.3, reynolds=200.0, riemann="roe")
    ref, got = _partitioned(world_size, lambda: MC.periodic_box(4, 2, 5, seed=7), phys, method)
    worst, _ = MC.compare(ref, got)
    print(worst)
    for k, v in worst.items():
        assert v <= (1e-13 if k in ("integrals", "surface") else 0.0), (k, v)


def test_partitioned_channel_with_boundaries_and_surface_integrals():
    phys = make_physics(flow="NS", mach=0.3, reynolds=150.0, riemann="roe", gradient_variables="energy")
    ref, got = _partitioned(2, lambda: MC.channel(phys, ne=4), phys, "metis", zone=2, scheme="ssprk33")
    worst, _ = MC.compare(ref, got)
    print(worst)
    for k, v in worst.items():
        assert v <= (1e-13 if k in ("integrals", "surface") else 0.0), (k, v)


def test_set_up_errors_are_reported():
    """The set-up checks of h3d_set_mesh_p / h3d_set_halo (shared by the device library and the host-loop backend)."""
    import ctypes as C
    from horses3d_b200.capi import H3dError, _ptr
    from horses3d_b200.dgsem import DGSem
    from horses3d_b200.hostmesh import NodalStorage
    phys = make_physics(flow="NS", mach=0.3, reynolds=100.0, riemann="roe")
    mesh = MC.periodic_box(2, 2, 4, seed=1)

    def attempt(skip_basis=None, skip_interp=False, face_order=None, msg=""):
        api = EmuApi()
        api.set_physics(phys)
        orders = np.array(mesh.array("elemOrder")).reshape(-1, 3); fo = np.array(mesh.array("faceOrder")).reshape(-1, 6)
        for N in sorted(set(orders.ravel()) | set(fo.ravel())):
            if N != skip_basis:
                api.set_basis(NodalStorage(int(N), mesh.nodes))
        if not skip_interp:
            from horses3d_b200.hostmesh import interpolation_matrix
            for a in range(2, 5):
                for b in range(2, 5):
                    if a != b:
                        api.call("set_interpolation", a, b, _ptr(interpolation_matrix(a, b, mesh.nodes), np.float64))
        if face_order is not None:
            saved = mesh.array("faceOrder").copy()
            mesh.array("faceOrder")[:] = face_order(saved)
        try:
            with pytest.raises(H3dError, match=msg):
                api.set_mesh_p(mesh)
        finally:
            if face_order is not None:
                mesh.array("faceOrder")[:] = saved

    attempt(skip_basis=3, msg="h3d_set_basis has not been called")
    attempt(skip_interp=True, msg="h3d_set_interpolation has not been called")
    attempt(face_order=lambda fo: np.where(np.arange(fo.size) == 0, fo + 1, fo), msg="faceOrder contradicts")
    # a partition on a single-rank context: MPI faces are refused
    part = mesh.partition(2, "block")
    with pytest.raises(H3dError, match="MPI faces on a single rank"):
        DGSem(EmuApi(), mesh.extract(part, 0, inherit_geometry=True), phys)
    # boundary faces without a boundary table
    ch = MC.channel(phys)
    ch.bcs = []
    sem = DGSem(EmuApi(), ch, phys)
    sem.set_Q(MC.smooth_state(sem, 0.3))
    with pytest.raises(H3dError, match="h3d_set_boundary_conditions was not called"):
        sem.ComputeTimeDerivative(0.0)


def test_k13_on_four_emulated_ranks():
    """The reference runs CylinderDifferentOrders under MPI in its parallel CI with the same expected values: the partitioned run
    (METIS weighted with the elements' degrees of freedom, four ranks as threads over the host-loop backend) meets them too."""
    import threading
    from emu.emu_api import EmuWorld
    from test_oracle_pins import K13, cylinder_different_orders
    world, out, errs = EmuWorld(4), [None] * 4, []
    part = cylinder_different_orders(steps=None).partition(4, "metis")      # once, before the threads: METIS is not re-entrant
    assert len(set(part)) == 4

    def work(rank):
        try:
            out[rank] = cylinder_different_orders(EmuApi(world, rank), partition=(part, rank))[1:]
        except Exception as ex:
            errs.append(ex)
            world.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    probes_found = [o[3] for o in out if o[3] is not None]
    assert len(probes_found) >= 1
    for res, cd, cl, _ in out:
        assert np.abs(res - K13["residuals"]).max() < 1.0e-11 and abs(cd - K13["cd"]) < 1.2e-10 and abs(cl - K13["cl"]) < 1.0e-11
    assert abs(probes_found[0] - K13["wake_u"]) < 1.0e-11
