"""Helpers shared by the parity tests: build a case, run it on a backend, compare with the oracle."""
import numpy as np

from horses3d_b200.dgsem import DGSem, taylor_green_ic
from horses3d_b200.hostmesh import GAUSS, GAUSSLOBATTO, HostMesh
from horses3d_b200.physics import make_physics

_mesh_cache = {}


def channel_bcs(phys):
    """Non-periodic box: inflow (left), outflow (right), moving and fixed no-slip walls (bottom/top), free-slip (front/back)."""
    from horses3d_b200.physics import bc_parameters
    bcs = [("left", "inflow", None), ("right", "outflow", None), ("bottom", "noslipwall", None), ("top", "noslipwall", None),
           ("front", "freeslipwall", None), ("back", "freeslipwall", None)]
    params = [bc_parameters("inflow", phys, aoa_theta=0.05, aoa_phi=0.02), bc_parameters("outflow", phys),
              bc_parameters("noslipwall", phys, wall_velocity=(0.1, 0.0, 0.0)), bc_parameters("noslipwall", phys),
              bc_parameters("freeslipwall", phys), bc_parameters("freeslipwall", phys)]
    return bcs, np.array(params)


def get_mesh(ne, N, nodes=GAUSS, amp=0.0, shuffle=False, bFaceOrder=2, seed=1234, bc=None, phys=None):
    key = (ne, N, nodes, amp, shuffle, bFaceOrder, seed, bc)
    if key not in _mesh_cache:
        m = HostMesh.box(ne, amp=amp, bFaceOrder=bFaceOrder, shuffle=shuffle, seed=seed)
        if bc == "channel":
            bcs, params = channel_bcs(phys)
            m.connect(bcs, params)
        else:
            m.connect()
        _mesh_cache[key] = m.geometry(N, nodes)
    return _mesh_cache[key]


def channel_state(x, phys):
    """Smooth non-uniform state around the inflow condition (rho ~ 1, |v| ~ 1, p ~ 1/(gamma M^2))."""
    X, Y, Z = x[..., 0], x[..., 1], x[..., 2]
    rho = 1.0 + 0.05 * np.sin(X) * np.cos(Y + 0.1) * np.cos(2 * Z)
    u = 1.0 + 0.1 * np.sin(Y) * np.cos(Z)
    v = 0.1 * np.sin(X + 0.3) * np.cos(Z)
    w = 0.05 * np.cos(X) * np.sin(2 * Y)
    p = (1.0 / phys.gammaM2) * (1.0 + 0.02 * np.cos(X) * np.sin(Y) * np.cos(Z))
    Q = np.empty(x.shape[:-1] + (5,))
    Q[..., 0], Q[..., 1], Q[..., 2], Q[..., 3] = rho, rho * u, rho * v, rho * w
    Q[..., 4] = p / phys.gammaMinus1 + 0.5 * rho * (u * u + v * v + w * w)
    return Q


def perturbed_tgv(x, mach_scale=1.0):
    """Taylor-Green state with a non-zero w and a density variation so that every flux term is exercised."""
    Q = taylor_green_ic(x, p0=100.0 * mach_scale)
    X, Y, Z = x[..., 0], x[..., 1], x[..., 2]
    rho = 1.0 + 0.1 * np.sin(X) * np.cos(2 * Y) * np.sin(Z + 0.3)
    w = 0.3 * np.cos(X + 0.2) * np.sin(Y) * np.sin(2 * Z)
    u, v = Q[..., 1] / Q[..., 0], Q[..., 2] / Q[..., 0]
    p = (Q[..., 4] - 0.5 * (u * u + v * v)) * 0.4
    Q[..., 0] = rho
    Q[..., 1], Q[..., 2], Q[..., 3] = rho * u, rho * v, rho * w
    Q[..., 4] = p / 0.4 + 0.5 * rho * (u * u + v * v + w * w)
    return Q


def rel_err(a, b):
    """max-abs difference per equation relative to the max-abs of the reference field of that equation."""
    a = a.reshape(-1, a.shape[-1])
    b = b.reshape(-1, b.shape[-1])
    scale = np.maximum(np.abs(b).max(axis=0), 1e-300)
    return (np.abs(a - b).max(axis=0) / scale).max()
