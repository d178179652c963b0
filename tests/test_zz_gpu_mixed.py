"""p-nonconforming path on the device: libh3dgpu.so (h3d_set_mesh_p) against the oracle, through the C ABI.
The same cases run on the CPU through the emulated kernels in tests/test_mixed_emu.py.

The file name sorts last on purpose: this path was written in a session without GPU time (its kernels were checked through the
host-loop backend only), so under `pytest -x` a failure here must not hide the results of the device tests that were verified
on hardware."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import mixed_cases as MC                                   # noqa: E402
from horses3d_b200.hostmesh import GAUSS, GAUSSLOBATTO     # noqa: E402
from horses3d_b200.physics import make_physics             # noqa: E402
from oracle import oracle_api                              # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1.0e-13          # the build disables FMA contraction: the sums are the oracle's, term by term


def both(gpu_api_cls, mesh_fn, phys, tol=TOL, **kw):
    _, a = MC.run_case(oracle_api.OracleApi(), mesh_fn(), phys, **kw)
    api = gpu_api_cls()
    _, b = MC.run_case(api, mesh_fn(), phys, **kw)
    assert api.kernel_launches() > 0
    worst, bad = MC.compare(a, b, tol)
    print(worst)
    assert not bad, bad


def test_state_round_trip_and_euler_residual(gpu_api_cls):
    """The smallest steps first, so that a failure on a new device is easy to place: set-up, upload / download of the packed state
    (device-side transposition), then the Euler residual without gradients (trace, adaption, Riemann, projection, volume)."""
    from horses3d_b200.dgsem import DGSem
    phys = make_physics(flow="Euler", mach=0.3, riemann="roe")
    sem = DGSem(gpu_api_cls(), MC.periodic_box(2, 1, 4, seed=3), phys)
    Q = MC.smooth_state(sem, 0.3)
    sem.set_Q(Q)
    assert np.array_equal(sem.Q(), Q)
    ref = DGSem(oracle_api.OracleApi(), MC.periodic_box(2, 1, 4, seed=3), phys)
    ref.set_Q(Q)
    sem.ComputeTimeDerivative(0.0); ref.ComputeTimeDerivative(0.0)
    a, b = ref.QDot(), sem.QDot()
    assert np.abs(a - b).max() <= TOL * np.abs(a).max()


@pytest.mark.parametrize("riemann", ["roe", "lax-friedrichs", "standard roe", "central"])
def test_navier_stokes_periodic_box_random_anisotropic_orders(gpu_api_cls, riemann):
    both(gpu_api_cls, lambda: MC.periodic_box(3, 2, 5, seed=7), make_physics(flow="NS", mach=0.3, reynolds=200.0, riemann=riemann), source=True)


def test_larger_mesh_with_orders_up_to_nine(gpu_api_cls):
    both(gpu_api_cls, lambda: MC.periodic_box(6, 1, 9, seed=21), make_physics(flow="NS", mach=0.3, reynolds=400.0, riemann="roe"))


def test_edge_cases_self_periodic_element_and_orders_up_to_fifteen(gpu_api_cls):
    phys = make_physics(flow="NS", mach=0.3, reynolds=200.0, riemann="roe")
    both(gpu_api_cls, lambda: MC.periodic_box(1, 2, 5, seed=3), phys)
    both(gpu_api_cls, lambda: MC.periodic_box(2, 10, 15, seed=5), phys)      # beyond the orders the uniform kernels are instantiated for


def test_euler_with_and_without_gradients(gpu_api_cls):
    both(gpu_api_cls, lambda: MC.periodic_box(3, 1, 4, seed=5), make_physics(flow="Euler", mach=0.3, riemann="roe"))
    both(gpu_api_cls, lambda: MC.periodic_box(3, 1, 4, seed=5), make_physics(flow="Euler", mach=0.3, riemann="rusanov", compute_gradients=True))


def test_gauss_lobatto_nodes_and_isotropic_orders(gpu_api_cls):
    both(gpu_api_cls, lambda: MC.periodic_box(3, 2, 6, seed=9, nodes=GAUSSLOBATTO, anisotropic=False),
         make_physics(flow="NS", mach=0.3, reynolds=100.0, riemann="roe"))


@pytest.mark.parametrize("scheme", ["euler", "rk5", "lserk14-4", "ssprk33", "ssprk43"])
def test_runge_kutta_schemes(gpu_api_cls, scheme):
    both(gpu_api_cls, lambda: MC.periodic_box(2, 2, 4, seed=13), make_physics(flow="NS", mach=0.3, reynolds=200.0, riemann="roe"), scheme=scheme)


@pytest.mark.parametrize("gradvars", ["state", "entropy", "energy"])
def test_boundary_conditions_and_gradient_variables(gpu_api_cls, gradvars):
    phys = make_physics(flow="NS", mach=0.3, reynolds=150.0, riemann="roe", gradient_variables=gradvars)
    # CUDA's log differs from glibc's in the last bit for some arguments (entropy variables): the north-star bound itself
    tol = 1.0e-12 if gradvars == "entropy" else TOL
    for zone in (2, 4):      # no-slip wall, inflow
        both(gpu_api_cls, lambda: MC.channel(phys), phys, tol=tol, zone=zone)


def test_uniform_orders_through_set_mesh_p_equal_the_uniform_path(gpu_api_cls):
    """One order for every element: the p-nonconforming kernels and the tuned uniform-order kernels must agree (both reproduce the
    oracle's sums)."""
    from horses3d_b200.dgsem import DGSem
    from horses3d_b200.hostmesh import HostMesh
    phys = make_physics(flow="NS", mach=0.3, reynolds=200.0, riemann="roe")
    out = []
    for mixed in (False, True):
        m = HostMesh.box(3, amp=0.15, shuffle=True).connect()
        m = m.geometry_p([3, 3, 3], GAUSS) if mixed else m.geometry(3, GAUSS, reference_order=True)
        sem = DGSem(gpu_api_cls(), m, phys)
        sem.set_Q(MC.smooth_state(sem, 0.3).reshape(sem._shape))
        sem.TakeRK3Step(0.0, 1.0e-3, ctd_after_step=True)
        d = sem.download(Q=True, QDot=True, gradients=True)
        out.append({k: v.reshape(-1, 5) for k, v in d.items()})
    worst, bad = MC.compare(out[0], out[1], TOL)
    print(worst)
    assert not bad, bad


def test_k13_cylinder_different_orders_on_the_device(gpu_api_cls):
    """The reference's NavierStokes/CylinderDifferentOrders regression through libh3dgpu.so: expected values and the 1e-11
    tolerance of SETUP/ProblemFile.f90:553-614."""
    from test_oracle_pins import K13, cylinder_different_orders
    api = gpu_api_cls()
    _, res, cd, cl, wake_u = cylinder_different_orders(api)
    print("K13 device rel", (res - K13["residuals"]) / K13["residuals"], "cd", cd - K13["cd"], "cl", cl - K13["cl"], "wake_u", wake_u - K13["wake_u"])
    assert api.kernel_launches() >= 100 * 3 * 13
    assert np.abs(res - K13["residuals"]).max() < 1.0e-11
    assert abs(cd - K13["cd"]) < 1.0e-11 * 12.0
    assert abs(cl - K13["cl"]) < 1.0e-11
    assert abs(wake_u - K13["wake_u"]) < 1.0e-11


def test_unsupported_entry_points_are_refused(gpu_api_cls):
    from horses3d_b200.capi import H3dError
    from horses3d_b200.dgsem import DGSem
    phys = make_physics(flow="NS", mach=0.3, reynolds=100.0, riemann="roe")
    sem = DGSem(gpu_api_cls(), MC.periodic_box(2, 2, 3, seed=1), phys)
    with pytest.raises(H3dError):
        sem.ScalarVolumeIntegral(99)      # unknown kind
    with pytest.raises(H3dError):
        DGSem(gpu_api_cls(), MC.periodic_box(2, 2, 3, seed=1, nodes=GAUSSLOBATTO), make_physics(flow="NS", mach=0.3, reynolds=100.0, riemann="roe", viscous="br2"))
    with pytest.raises(H3dError):      # the split form needs Gauss-Lobatto nodes
        DGSem(gpu_api_cls(), MC.periodic_box(2, 2, 3, seed=1), make_physics(flow="Euler", mach=0.3, inviscid="split-form", averaging="pirozzoli"))


def test_cpp_driver_reproduces_the_different_orders_regression_on_the_device():
    """K13 through the native C++ driver (polynomial order file, h3d_set_mesh_p) and libh3dgpu.so."""
    from horses3d_b200 import build
    from test_cpp_driver import K13_ARGS, K13_RES, final_line, run_driver
    f = final_line(run_driver("--lib", build.build_gpu(), *K13_ARGS))
    assert f["iter"] == 100 and np.abs(f["residuals"] - K13_RES).max() < 1.0e-11


@pytest.mark.parametrize("name", ["box_ns_mixed_p2to4", "channel_ns_mixed_p2to4", "box_euler_mixed_p1to5", "box_euler_split_pirozzoli_mixed_p2to5"])
def test_device_reproduces_golden_mixed(gpu_api_cls, name):
    """The frozen oracle outputs of tests/golden (made by make_golden.py) through libh3dgpu.so."""
    from test_golden import check
    check(gpu_api_cls(), name, exact=False)


@pytest.mark.parametrize("kw", [dict(flow="Euler", mach=0.3, inviscid="split-form", averaging="pirozzoli", riemann="roe"),
                                dict(flow="NS", mach=0.3, reynolds=200.0, inviscid="split-form", averaging="kennedy-gruber", riemann="roe"),
                                dict(flow="NS", mach=0.3, reynolds=200.0, inviscid="split-form", averaging="chandrasekar", riemann="central", gradient_variables="entropy")],
                         ids=lambda k: k["flow"] + "-" + k["averaging"])
def test_split_form_on_random_orders(gpu_api_cls, kw):
    # logarithmic means and entropy variables take log(): last-bit differences between CUDA and glibc, hence the north-star bound
    both(gpu_api_cls, lambda: MC.periodic_box(3, 2, 5, seed=17, nodes=GAUSSLOBATTO), make_physics(**kw), tol=1.0e-12 if kw["averaging"] == "chandrasekar" else TOL)


@pytest.mark.parametrize("kw", [dict(les="smagorinsky", les_wall_model="linear"), dict(les="wale"), dict(les="vreman")], ids=lambda k: k["les"])
def test_les_models_on_random_orders(gpu_api_cls, kw):
    phys = make_physics(flow="NS", mach=0.3, reynolds=150.0, riemann="roe", **kw)
    both(gpu_api_cls, lambda: MC.channel(phys), phys, zone=2)


@pytest.mark.parametrize("scheme", ["ssprk33", "ssprk43"])
def test_stage_limiter_and_statistics(gpu_api_cls, scheme):
    _, a = MC.limiter_and_statistics_case(oracle_api.OracleApi(), scheme=scheme)
    _, b = MC.limiter_and_statistics_case(gpu_api_cls(), scheme=scheme)
    worst, bad = MC.compare(a, b, TOL)
    print(worst)
    assert not bad, bad


@pytest.mark.parametrize("kw", [dict(viscous="ip"), dict(viscous="ip", ip_variant="NIPG", penalty_parameter=3.0, gradient_variables="energy")], ids=["sipg", "nipg-energy"])
def test_interior_penalty_on_random_orders(gpu_api_cls, kw):
    phys = make_physics(flow="NS", mach=0.3, reynolds=150.0, riemann="roe", **kw)
    both(gpu_api_cls, lambda: MC.channel(phys), phys, zone=2)
