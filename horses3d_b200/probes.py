"""Host side of the probes (libs/monitors/Probe.f90): point location and the Lagrange vectors handed to h3d_probe.

`find_point` follows HexMesh_FindPointWithCoords / HexElement_FindPointWithCoords (HexMesh.f90:5483-5560,
HexElementClass.f90:616-739): Newton iterations on the nodal interpolant of the element geometry, tolerance 1e-12 on the
residual, at most 50 iterations, accepted when |xi| < 1 + 1e-8 in every direction; the first element (lowest id) that
accepts the point owns the probe."""
import numpy as np

from . import physics as P

VARIABLES = {"pressure": 0, "velocity": 1, "u": 2, "v": 3, "w": 4, "mach": 5, "k": 6}


def lagrange_vectors(nodes, xi):
    """lj(xi) and dlj(xi): the Lagrange basis on `nodes` and its derivative at xi (NodalStorageClass lj / dlj)."""
    n = len(nodes)
    l, dl = np.ones(n), np.zeros(n)
    for j in range(n):
        for m in range(n):
            if m != j:
                l[j] *= (xi - nodes[m]) / (nodes[j] - nodes[m])
        for m in range(n):
            if m == j:
                continue
            t = 1.0 / (nodes[j] - nodes[m])
            for k in range(n):
                if k != j and k != m:
                    t *= (xi - nodes[k]) / (nodes[j] - nodes[k])
            dl[j] += t
    return l, dl


def find_point_in_element(X, nodes, x, tol=1.0e-12, inside_tol=1.0e-8, max_iter=50):
    """X: node coordinates of one element [k][j][i][3].  Returns (inside, xi)."""
    xi = np.zeros(3)
    nodes3 = nodes if isinstance(nodes, (tuple, list)) else (nodes, nodes, nodes)     # anisotropic elements: one node set per direction
    for _ in range(max_iter):
        (lx, dlx), (ly, dly), (lz, dlz) = (lagrange_vectors(nodes3[d], xi[d]) for d in range(3))
        F = np.einsum("kjic,i,j,k->c", X, lx, ly, lz) - x
        if np.abs(F).max() < tol or np.abs(xi).max() >= 2.5:
            break
        J = np.stack([np.einsum("kjic,i,j,k->c", X, dlx, ly, lz), np.einsum("kjic,i,j,k->c", X, lx, dly, lz),
                      np.einsum("kjic,i,j,k->c", X, lx, ly, dlz)], axis=1)
        xi = xi + np.linalg.solve(J, -F)
    return bool((np.abs(xi) < 1.0 + inside_tol).all()), xi


def find_point(sem, x):
    """(element id, xi) of the first element containing x, or (None, None)."""
    x = np.asarray(x, dtype=np.float64)
    X = sem.node_coordinates()
    if getattr(sem, "mixed", False):
        for e in range(sem.nElem):
            Xe = sem.element_view(X, e)
            lo, hi = Xe.min(axis=(0, 1, 2)), Xe.max(axis=(0, 1, 2))
            pad = 0.5 * (hi - lo).max() + 1e-8
            if not ((x >= lo - pad) & (x <= hi + pad)).all():
                continue
            inside, xi = find_point_in_element(Xe, tuple(sem.sps[int(N)].x for N in sem.orders[e]), x)
            if inside:
                return int(e), xi
        return None, None
    lo, hi = X.min(axis=(1, 2, 3)), X.max(axis=(1, 2, 3))
    pad = 0.5 * (hi - lo).max(axis=1, keepdims=True) + 1e-8      # Gauss nodes do not reach the element boundary
    for e in np.nonzero(((x >= lo - pad) & (x <= hi + pad)).all(axis=1))[0]:
        inside, xi = find_point_in_element(X[e], sem.sp.x, x)
        if inside:
            return int(e), xi
    return None, None


class Probe:
    def __init__(self, sem, position, variable):
        self.variable = VARIABLES[variable.lower()]
        self.eID, self.xi = find_point(sem, position)
        self.active = self.eID is not None
        if self.active and getattr(sem, "mixed", False):
            # rows padded to the largest number of nodes per direction of the mesh (include/h3d_gpu.h, h3d_probe)
            ld = int(sem.orders.max()) + 1
            self.l = []
            for d in range(3):
                row = np.zeros(ld)
                lv = lagrange_vectors(sem.sps[int(sem.orders[self.eID][d])].x, self.xi[d])[0]
                row[:len(lv)] = lv
                self.l.append(row)
        elif self.active:
            self.l = [np.ascontiguousarray(lagrange_vectors(sem.sp.x, self.xi[d])[0]) for d in range(3)]


def evaluate(sem, probes):
    """Probe_Update for a list of probes living on this rank; returns their values."""
    from .capi import _ptr
    act = [p for p in probes if p.active]
    if not act:
        return np.zeros(0)
    elem = np.array([p.eID for p in act], dtype=np.int32)
    var = np.array([p.variable for p in act], dtype=np.int32)
    L = [np.ascontiguousarray(np.stack([p.l[d] for p in act])) for d in range(3)]
    out = np.zeros(len(act))
    sem.api.call("probe", len(act), _ptr(elem, np.int32), _ptr(var, np.int32), _ptr(L[0], np.float64), _ptr(L[1], np.float64),
                 _ptr(L[2], np.float64), _ptr(out, np.float64))
    return out
