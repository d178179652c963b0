"""Per-DOF parity of the PERSISTENT kernel paths with the CPU oracle.  Needs a B200.

The element kernels are persistent (one or two CTAs per SM, each looping over tiles with the next tile's fields prefetched by
bulk copies into double-buffered face tables).  The small cases of test_gpu_parity.py give every CTA at most one tile, so the
prefetch branch, the mbarrier parity flips and the table double-buffering never run there.  These cases are sized so that every
staged instantiation runs at least three tiles per CTA, up to BASELINE configs[1] itself (32^3 elements, P=7, curved, randomly
re-oriented elements: 16.8 M DOF), and compare QDot, the gradients and Q after one RK3 step node by node with the oracle
(SpatialDiscretization.f90:227-320, ExplicitMethods.f90:667-788).

Tolerance as in test_gpu_parity.py: 1e-13 per equation relative to the field's max-norm (north star: 1e-12).
"""
import numpy as np
import pytest

from horses3d_b200.dgsem import DGSem
from horses3d_b200.hostmesh import GAUSS, GAUSSLOBATTO
from horses3d_b200.physics import make_physics
from oracle.oracle_api import OracleApi
from parity import channel_state, get_mesh, perturbed_tgv, rel_err

pytestmark = pytest.mark.gpu

TOL = 1.0e-13

NS = dict(flow="NS", mach=0.08, reynolds=1600.0)
SPLIT = dict(flow="Euler", mach=0.3, inviscid="split-form", averaging="pirozzoli")

# (ne, N, nodes, physics, tiles of the element kernels = elements / elements per CTA)
CASES = [
    (8, 7, GAUSS, NS),                  # 512 tiles: 3.5 per CTA, staged n=8 StandardDG + BR1 (the headline instantiation)
    (12, 5, GAUSS, NS),                 # 1728 tiles, staged n=6
    (16, 3, GAUSS, NS),                 # 4096 elements, 4 per CTA: 1024 tiles, staged n=4
    (24, 1, GAUSS, NS),                 # 13824 elements, 32 per CTA: 432 tiles, staged n=2
    (10, 4, GAUSS, NS),                 # odd n: 1000 tiles
    (8, 6, GAUSS, NS),                  # odd n = 7: 512 tiles
    (7, 8, GAUSS, NS),                  # odd n = 9
    (6, 9, GAUSS, NS),                  # n = 10
    (9, 7, GAUSSLOBATTO, SPLIT),        # 729 tiles, staged Euler SplitDG Pirozzoli n=8
    (12, 5, GAUSSLOBATTO, SPLIT),
    (16, 3, GAUSSLOBATTO, SPLIT),
    (8, 7, GAUSSLOBATTO, dict(flow="NS", mach=0.3, reynolds=200.0, inviscid="split-form", averaging="pirozzoli")),
    (8, 7, GAUSS, dict(flow="Euler", mach=0.3)),
    (8, 7, GAUSS, dict(flow="NS", mach=0.3, reynolds=200.0, viscous="BR2")),
    (8, 7, GAUSS, dict(flow="NS", mach=0.3, reynolds=200.0, viscous="IP")),
]


def compare(gpu_api_cls, mesh, phys, ic, dt=1.0e-3, tol=TOL):
    out = []
    for api in (OracleApi(), gpu_api_cls()):
        sem = DGSem(api, mesh, phys)
        sem.set_initial_condition(ic)
        sem.ComputeTimeDerivative(0.0)
        d = sem.download(Q=False, QDot=True, gradients=bool(phys.flowIsNavierStokes))
        sem.TakeRK3Step(0.0, dt)
        d["Q1"] = sem.Q()
        d["res"] = sem.ComputeMaxResiduals()
        out.append(d)
        del sem
    o, g = out
    assert np.abs(o["QDot"]).max() > 1e-3
    errs = {k: rel_err(g[k], o[k]) for k in o if k != "res"}
    for k, e in errs.items():
        assert e < tol, (k, errs)
    assert np.abs(o["res"] - g["res"]).max() <= 1e-12 * np.abs(o["res"]).max()
    return errs


@pytest.mark.parametrize("ne,N,nodes,kw", CASES)
def test_multi_tile_time_derivative_and_rk3_step_match_oracle(gpu_api_cls, ne, N, nodes, kw):
    mesh = get_mesh(ne, N, nodes, 0.1, True)
    compare(gpu_api_cls, mesh, make_physics(**kw), perturbed_tgv)


@pytest.mark.parametrize("ne,N,kw", [(8, 7, dict(flow="NS", mach=0.3, reynolds=200.0)),
                                    (12, 3, dict(flow="NS", mach=0.3, reynolds=200.0, les="smagorinsky", les_wall_model="linear"))])
def test_multi_tile_boundary_conditions_match_oracle(gpu_api_cls, ne, N, kw):
    phys = make_physics(**kw)
    mesh = get_mesh(ne, N, GAUSS, 0.1, True, bc="channel", phys=phys)
    if phys.les_wall_model:
        mesh.wall_distances()
    compare(gpu_api_cls, mesh, phys, lambda x: channel_state(x, phys))


def test_baseline_config1_tgv_32cubed_p7_matches_oracle(gpu_api_cls):
    """BASELINE configs[1] itself: Taylor-Green vortex, Re 1600, 32^3 curvilinear hex elements, P=7, StandardDG + BR1 + Roe,
    RK3: 16.8 M DOF, 221 tiles per CTA.  Every node of QDot, of the three gradients and of Q after one RK3 step."""
    from horses3d_b200.dgsem import taylor_green_ic
    mesh = get_mesh(32, 7, GAUSS, 0.1, True)
    errs = compare(gpu_api_cls, mesh, make_physics(flow="NS", mach=0.08, reynolds=1600.0, riemann="roe"), taylor_green_ic, dt=2.0e-4)
    print("configs[1] per-DOF max rel err vs oracle:", errs)


def _run_with_option(gpu_api_cls, mesh, phys, ic, *options):
    api = gpu_api_cls()
    sem = DGSem(api, mesh, phys)
    for option in options:
        if option:
            api.call("set_option", option.encode())
    sem.set_initial_condition(ic)
    sem.ComputeTimeDerivative(0.0)
    d = sem.download(QDot=True, gradients=True)
    sem.TakeRK3Step(0.0, 1.0e-3)
    d["Q1"] = sem.Q()
    return d


def test_second_generation_kernels_are_bit_identical(gpu_api_cls):
    """Option gen2=1 (h3d_kernels2.cuh: 256-thread CTAs, two per SM, fluxes in place over the staged gradients, swizzled
    prolongation buffers): same arithmetic, same summation order -- bit for bit the first-generation result, several tiles per CTA."""
    mesh = get_mesh(9, 7, GAUSS, 0.1, True)
    phys = make_physics(**NS)
    a = _run_with_option(gpu_api_cls, mesh, phys, perturbed_tgv, None)
    b = _run_with_option(gpu_api_cls, mesh, phys, perturbed_tgv, "gen2=1")
    for k in a:
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("gen2", [0, 1])
@pytest.mark.parametrize("mach,bound", [(0.3, 2.0e-12), (0.08, 1.0e-11)])
def test_fp64_tensor_core_contraction_stays_within_its_bound(gpu_api_cls, gen2, mach, bound):
    """Option mma=1 (h3d_mma.cuh): the derivative contractions on DMMA.8x8x4.  The summation is fused and re-ordered, so the
    result is not bit-identical: measured 3e-13..7e-13 of the field's max-norm at M 0.3 and 2e-12..4e-12 at M 0.08, where the
    background pressure 1 / (gamma M^2) = 112 makes every derivative a difference of numbers 10^3 times larger than the result
    (the oracle's own rounding error is of that size).  That is above the north star's 1e-12 for the headline state, which is
    why the path is an option and the bit-identical CUDA-core contraction the default."""
    mesh = get_mesh(8, 7, GAUSS, 0.1, True)
    phys = make_physics(flow="NS", mach=mach, reynolds=200.0 if mach > 0.1 else 1600.0)
    ic = (lambda x: perturbed_tgv(x, 0.1)) if mach > 0.1 else perturbed_tgv
    sem = DGSem(OracleApi(), mesh, phys)
    sem.set_initial_condition(ic)
    sem.ComputeTimeDerivative(0.0)
    o = sem.download(QDot=True, gradients=True)
    g = _run_with_option(gpu_api_cls, mesh, phys, ic, "mma=1", "gen2=1" if gen2 else None)
    errs = {k: rel_err(g[k], o[k]) for k in o}
    assert max(errs.values()) > 0.0                      # the tensor-core path really ran
    for k, e in errs.items():
        assert e < bound, errs
