"""ctypes front end of libh3dhost.so: the driver-side mesh/geometry the reference keeps in Fortran.

Mirrors what the reference driver does before the time loop (sem%construct, DGSEMClass.f90:268-468):
read or generate the mesh, build face connectivity with the boundary table of the control file,
construct nodal storage and metric terms.  The arrays exposed here are, one to one, the arguments of
the device C-ABI (include/h3d_gpu.h).
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

GAUSS = 1
GAUSSLOBATTO = 2

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.build_host()
        L = C.CDLL(path)
        L.h3dhost_last_error.restype = C.c_char_p
        L.h3dhost_mesh_box.restype = C.c_void_p
        L.h3dhost_mesh_box.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_uint]
        L.h3dhost_mesh_read.restype = C.c_void_p
        L.h3dhost_mesh_read.argtypes = [C.c_char_p]
        L.h3dhost_mesh_write.argtypes = [C.c_void_p, C.c_char_p]
        L.h3dhost_mesh_free.argtypes = [C.c_void_p]
        L.h3dhost_mesh_connect.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_void_p]
        L.h3dhost_mesh_geometry.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.h3dhost_mesh_geometry_p.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.h3dhost_interpolation_matrix.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.h3dhost_mesh_sizes.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
        L.h3dhost_get_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
        L.h3dhost_nodal.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 7
        L.h3dhost_partition.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.h3dhost_partition_weighted.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.h3dhost_extract_partition.restype = C.c_void_p
        L.h3dhost_extract_partition.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.h3dhost_inherit_geometry.argtypes = [C.c_void_p, C.c_void_p]
        L.h3dhost_wall_distance.argtypes = [C.c_void_p]
        L.h3dhost_set_num_threads.argtypes = [C.c_int]
        L.h3dhost_wall_points.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_longlong)]
        L.h3dhost_wall_distance_from.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
        _lib = L
    return _lib


def set_num_threads(n):
    """Threads of the host library's OpenMP loops; returns what a parallel region really gets.  Needed under launchers that pin
    OMP_NUM_THREADS=1 (torchrun): the OpenMP runtime has read the environment long before this module can change it."""
    return int(lib().h3dhost_set_num_threads(int(n)))


class HostError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise HostError(lib().h3dhost_last_error().decode())


PERIODIC_BOX_BCS = [
    ("front", "periodic", "back"), ("back", "periodic", "front"),
    ("bottom", "periodic", "top"), ("top", "periodic", "bottom"),
    ("left", "periodic", "right"), ("right", "periodic", "left"),
]


class NodalStorage:
    """1-D operators of one polynomial order (NodalStorageClass.f90:24-52)."""

    def __init__(self, N, nodes=GAUSS):
        n = N + 1
        self.N, self.nodes, self.n = N, nodes, n
        self.x, self.w = np.zeros(n), np.zeros(n)
        self.D, self.hatD, self.sharpD = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n))
        self.v, self.b = np.zeros((2, n)), np.zeros((2, n))
        _check(lib().h3dhost_nodal(N, nodes, *[a.ctypes.data for a in (self.x, self.w, self.D, self.hatD, self.sharpD, self.v, self.b)]))


def interpolation_matrix(Norigin, Ndest, nodes=GAUSS):
    """Tset(Norigin, Ndest) % T (InterpolationMatrices.f90:42-107): [Ndest+1, Norigin+1]; Lagrange interpolation when Norigin < Ndest,
    the L2 projection (its weighted transpose) otherwise."""
    T = np.zeros((Ndest + 1, Norigin + 1))
    _check(lib().h3dhost_interpolation_matrix(Norigin, Ndest, nodes, T.ctypes.data))
    return T


def read_order_file(path):
    """ReadOrderFile (libs/io/ReadInputFile.f90:132-153): number of elements, then Nx Ny Nz per element."""
    with open(path) as f:
        tok = f.read().replace(",", " ").split()
    n = int(tok[0])
    return np.array(tok[1:1 + 3 * n], dtype=np.int32).reshape(n, 3)


class HostMesh:
    """Driver-side HexMesh: raw mesh -> connectivity -> geometry (flat numpy views on C++ storage)."""

    def __init__(self, handle):
        if not handle:
            raise HostError(lib().h3dhost_last_error().decode())
        self._h = C.c_void_p(handle)
        self.halo = None
        self.mixed = False             # p-nonconforming: per-element polynomial orders (geometry_p)
        self.is_partition = False      # extracted from a global mesh: some faces are MPI faces
        self.wall_global = False       # wall distances measured against the wall nodes of the whole mesh

    # --- constructors
    @classmethod
    def box(cls, ne, L=2.0 * np.pi, amp=0.0, bFaceOrder=2, shuffle=False, seed=1234, ney=None, nez=None):
        return cls(lib().h3dhost_mesh_box(ne, ney or ne, nez or ne, L, amp, bFaceOrder, int(shuffle), seed))

    @classmethod
    def read(cls, path):
        return cls(lib().h3dhost_mesh_read(os.fsencode(path)))

    def write(self, path):
        _check(lib().h3dhost_mesh_write(self._h, os.fsencode(path)))

    def __del__(self):
        try:
            if self._h:
                lib().h3dhost_mesh_free(self._h)
                self._h = None
        except Exception:
            pass

    # --- pipeline
    def connect(self, bcs=PERIODIC_BOX_BCS, params=None):
        nb = len(bcs)
        names = (C.c_char_p * nb)(*[b[0].encode() for b in bcs])
        types = (C.c_char_p * nb)(*[b[1].encode() for b in bcs])
        coupled = (C.c_char_p * nb)(*[(b[2] or "").encode() for b in bcs])
        p = None
        if params is not None:
            p = np.ascontiguousarray(params, dtype=np.float64).reshape(nb, 16)
        _check(lib().h3dhost_mesh_connect(self._h, nb, names, types, coupled, p.ctypes.data if p is not None else None))
        self.bcs = list(bcs)
        self.bc_params = p
        return self

    def geometry(self, N, nodes=GAUSS, reference_order=False):
        """Metric terms (MappedGeometry.f90:174-384).  reference_order=True interpolates them from the Chebyshev-Lobatto grid to the
        nodes with the reference's full triple sum (n^6 operations per element) instead of the sum-factorised form: same values to
        round-off, the reference's rounding -- for the regression pins that run a thousand steps."""
        _check(lib().h3dhost_mesh_geometry(self._h, N, nodes + (16 if reference_order else 0)))
        self.N, self.nodes = N, nodes
        return self

    def geometry_p(self, orders, nodes=GAUSS):
        """Geometry of a p-nonconforming mesh: orders[e] = (Nx, Ny, Nz) of every element, as the reference's
        "polynomial order file".  Faces get the maximum order of their two sides per direction (FaceClass.f90:187-282)."""
        orders = np.ascontiguousarray(np.broadcast_to(np.asarray(orders, dtype=np.int32), (self.nElem, 3)))
        _check(lib().h3dhost_mesh_geometry_p(self._h, nodes, orders.ctypes.data))
        self.N, self.nodes, self.mixed = None, nodes, True
        self.orders = orders
        return self

    def wall_distances(self, gather=None):
        """e % geom % dWall / f % geom % dWall: distance to the nearest no-slip wall node (HexMesh.f90:5594-5692).

        On a partition the wall nodes of ALL ranks are needed (GatherAllWallCoordinates, HexMesh.f90:5696-5780): pass
        `gather`, a callable that takes this rank's [k,3] wall points and returns the concatenation over the ranks (an
        all-gather of the driver's communicator).  Without it a partition refuses: distances to the local wall nodes only
        would make the wall model depend on the partition."""
        if self.is_partition and gather is None:
            raise HostError("wall distances on a partition need the wall nodes of every rank: pass gather= (or extract the "
                            "partition with inherit_geometry=True from a global mesh that has its wall distances)")
        if gather is None:
            _check(lib().h3dhost_wall_distance(self._h))
        else:
            pts = np.ascontiguousarray(gather(self.wall_points()), dtype=np.float64).reshape(-1, 3)
            _check(lib().h3dhost_wall_distance_from(self._h, pts.ctypes.data, len(pts)))
        self.wall_global = True
        return self

    def wall_points(self):
        """Nodes of this mesh's no-slip wall faces, [k,3]."""
        cnt = C.c_longlong()
        _check(lib().h3dhost_wall_points(self._h, None, C.byref(cnt)))
        pts = np.zeros((cnt.value, 3))
        if cnt.value:
            _check(lib().h3dhost_wall_points(self._h, pts.ctypes.data, C.byref(cnt)))
        return pts

    def sizes(self):
        a = [C.c_int() for _ in range(4)]
        lib().h3dhost_mesh_sizes(self._h, *[C.byref(x) for x in a])
        return tuple(x.value for x in a)

    @property
    def nElem(self):
        return self.sizes()[0]

    @property
    def nFaces(self):
        return self.sizes()[1]

    def array(self, name):
        ptr, cnt, isint = C.c_void_p(), C.c_longlong(), C.c_int()
        _check(lib().h3dhost_get_array(self._h, name.encode(), C.byref(ptr), C.byref(cnt), C.byref(isint)))
        if cnt.value == 0:
            return np.zeros(0, dtype=np.int32 if isint.value else np.float64)
        ct = C.c_int32 if isint.value else C.c_double
        arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(cnt.value,))
        return arr   # a VIEW of the C++ storage: keep this HostMesh alive while it is in use

    # --- partitioning
    def partition(self, nparts, method="metis"):
        """Element -> rank.  A p-nonconforming mesh is partitioned with the elements' degrees of freedom as weights, as the reference
        does when the orders are not uniform (METISPartitioning.f90:125-151)."""
        part = np.zeros(self.nElem, dtype=np.int32)
        if self.mixed:
            w = np.ascontiguousarray(np.prod(self.orders + 1, axis=1), dtype=np.int32)
            _check(lib().h3dhost_partition_weighted(self._h, nparts, 0 if method == "metis" else 1, w.ctypes.data, part.ctypes.data))
        else:
            _check(lib().h3dhost_partition(self._h, nparts, 0 if method == "metis" else 1, part.ctypes.data))
        return part

    def extract(self, part, rank, inherit_geometry=False):
        """Local mesh of `rank`.  inherit_geometry=True copies this (global) mesh's metric terms instead of rebuilding them
        from the local elements: MPI faces then carry the global face geometry and results do not depend on the partition."""
        part = np.ascontiguousarray(part, dtype=np.int32)
        child = HostMesh(lib().h3dhost_extract_partition(self._h, part.ctypes.data, rank))
        child.bcs = getattr(self, "bcs", [])
        child.bc_params = getattr(self, "bc_params", None)
        child.is_partition = True
        if self.mixed and not inherit_geometry:
            raise HostError("a partition of a p-nonconforming mesh takes its geometry from the global mesh: extract(..., inherit_geometry=True)")
        if inherit_geometry:
            _check(lib().h3dhost_inherit_geometry(child._h, self._h))
            child.N, child.nodes = self.N, self.nodes
            child.wall_global = self.wall_global
            child.mixed = self.mixed
            if self.mixed:
                child.orders = np.array(child.array("elemOrder")).reshape(-1, 3)
        return child
