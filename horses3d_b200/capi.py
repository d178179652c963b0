"""ctypes binding of the C ABI declared in include/h3d_gpu.h.

`Binding` types the entry points of a library that exports that ABI under a given prefix.  The product
binds `h3d_*` from libh3dgpu.so (class GpuApi, below); there is NO CPU fallback: if the CUDA library is
missing or no B200 is visible, GpuApi raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build
from .physics import H3dPhysics

_D = C.c_void_p  # double* / int* passed as raw addresses


class H3dError(RuntimeError):
    pass


def _ptr(a, dtype):
    if a is None:
        return None
    if not (isinstance(a, np.ndarray) and a.dtype == dtype and a.flags.c_contiguous):
        raise TypeError("expected a C-contiguous %s array" % np.dtype(dtype).name)
    return a.ctypes.data


class Binding:
    """Typed entry points `<prefix>set_physics`, ... of one shared library."""

    COMMON = {
        "set_physics": [C.POINTER(H3dPhysics)],
        "set_basis": [C.c_int, C.c_int] + [_D] * 7,
        "set_mesh": [C.c_int, C.c_int] + [_D] * 19,
        "set_interpolation": [C.c_int, C.c_int, _D],
        "set_mesh_p": [C.c_int, C.c_int] + [_D] * 21,
        "set_boundary_conditions": [C.c_int, _D, _D],
        "set_wall_distance": [_D, _D],
        "set_face_h": [_D],
        "upload_Q": [_D],
        "download": [_D] * 5,
        "set_source": [_D],
        "compute_time_derivative": [C.c_double],
        "rk_step": [C.c_int, C.c_double, C.c_double, C.c_int],
        "rk_stage": [C.c_int, C.c_int, C.c_double, C.c_double],
        "enable_limiter": [C.c_int, C.c_double],
        "max_residuals": [_D],
        "max_timestep": [C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)],
        "volume_integral": [C.c_int, C.POINTER(C.c_double)],
        "has_nan": [C.POINTER(C.c_int)],
        "surface_integral": [C.c_int, C.c_int, _D],
        "statistics_update": [C.c_int],
        "snapshot_begin": [],
        "snapshot_end": [_D],
        "statistics_download": [_D, C.POINTER(C.c_int), C.POINTER(C.c_int)],
        "probe": [C.c_int, _D, _D, _D, _D, _D, _D],
    }

    def __init__(self, lib, prefix, extra=None, names=None):
        self.lib, self.prefix = lib, prefix
        table = dict(self.COMMON)
        table.update(extra or {})
        if names is not None:      # a library that exports a subset of the ABI (test doubles)
            table = {k: table[k] for k in names}
        for name, args in table.items():
            fn = getattr(lib, prefix + name)
            fn.argtypes = [C.c_void_p] + args
            fn.restype = C.c_int
            setattr(self, name, fn)
        self.last_error = getattr(lib, prefix + "last_error")
        self.last_error.argtypes = [C.c_void_p]
        self.last_error.restype = C.c_char_p


class Api:
    """Handle + checked calls; subclasses provide the handle and the Binding."""

    binding = None
    handle = None
    name = "?"

    def call(self, fname, *args):
        rc = getattr(self.binding, fname)(self.handle, *args)
        if rc != 0:
            msg = self.binding.last_error(self.handle)
            raise H3dError("%s%s failed (%d): %s" % (self.binding.prefix, fname, rc, msg.decode() if msg else ""))

    # -- typed conveniences shared by every backend
    def set_physics(self, p):
        self.call("set_physics", C.byref(p))

    def set_basis(self, sp):
        self.call("set_basis", sp.N, sp.nodes, *[_ptr(np.ascontiguousarray(a), np.float64) for a in (sp.x, sp.w, sp.D, sp.hatD, sp.sharpD, sp.v, sp.b)])

    def set_mesh(self, m):
        nE, nF = m.nElem, m.nFaces
        I = lambda k: _ptr(m.array(k), np.int32)
        Dp = lambda k: _ptr(m.array(k), np.float64)
        self.call("set_mesh", nE, nF, I("elemFace"), I("elemFaceSide"), I("faceElem"), I("faceElemSide"), I("faceRot"), I("faceType"),
                  I("faceZone"), Dp("jGradXi"), Dp("jGradEta"), Dp("jGradZeta"), Dp("jacobian"), Dp("x"), Dp("volume"),
                  Dp("faceNormal"), Dp("faceT1"), Dp("faceT2"), Dp("faceJacobian"), Dp("faceX"), Dp("faceSurface"))

    def set_mesh_p(self, m):
        """p-nonconforming mesh: as set_mesh with the elements' orders; arrays packed at the elements' / faces' own sizes."""
        nE, nF = m.nElem, m.nFaces
        I = lambda k: _ptr(m.array(k), np.int32)
        Dp = lambda k: _ptr(m.array(k), np.float64)
        self.call("set_mesh_p", nE, nF, I("elemOrder"), I("faceOrder"), I("elemFace"), I("elemFaceSide"), I("faceElem"), I("faceElemSide"), I("faceRot"),
                  I("faceType"), I("faceZone"), Dp("jGradXi"), Dp("jGradEta"), Dp("jGradZeta"), Dp("jacobian"), Dp("x"), Dp("volume"),
                  Dp("faceNormal"), Dp("faceT1"), Dp("faceT2"), Dp("faceJacobian"), Dp("faceX"), Dp("faceSurface"))

    def set_boundary_conditions(self, types, params):
        types = np.ascontiguousarray(types, dtype=np.int32)
        params = np.ascontiguousarray(params, dtype=np.float64)
        self.call("set_boundary_conditions", len(types), _ptr(types, np.int32), _ptr(params, np.float64))


_gpu_lib = None


def gpu_library():
    """Loads libh3dgpu.so (built in-tree by build.build_gpu).  Raises if it cannot be built/loaded."""
    global _gpu_lib
    if _gpu_lib is None:
        path = os.environ.get("H3D_GPU_LIB") or _build.GPU_LIB     # H3D_GPU_LIB: alternative build of the SAME sources
        if not os.path.exists(path):
            path = _build.build_gpu()
        # libh3dgpu.so needs libnccl.so.2.  PyTorch bundles a newer NCCL than the system one under the same soname; load
        # that one first so that this library and a later `import torch` share a single NCCL in the process.
        import sys
        for d in sys.path:
            cand = os.path.join(d, "nvidia", "nccl", "lib", "libnccl.so.2")
            if os.path.exists(cand):
                try:
                    C.CDLL(cand, mode=C.RTLD_GLOBAL)
                except OSError:
                    pass
                break
        _gpu_lib = C.CDLL(path)
    return _gpu_lib


class GpuApi(Api):
    """One context = one rank = one B200 (h3d_create).  No fallback: raises if the device path is unavailable."""

    name = "gpu"
    EXTRA = {
        "set_halo": [C.c_int, _D, _D, _D, _D],
        "synchronize": [],
        "timer_begin": [],
        "timer_end": [C.POINTER(C.c_double)],
        "set_option": [C.c_char_p],
    }

    def __init__(self, rank=0, nranks=1, device=0, nccl_id=None):
        lib = gpu_library()
        self.binding = Binding(lib, "h3d_", self.EXTRA)
        lib.h3d_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.h3d_create.restype = C.c_int
        lib.h3d_destroy.argtypes = [C.c_void_p]
        lib.h3d_kernel_launches.argtypes = [C.c_void_p]
        lib.h3d_kernel_launches.restype = C.c_longlong
        h = C.c_void_p()
        idbuf = None
        if nccl_id is not None:
            idbuf = (C.c_char * 128).from_buffer_copy(bytes(nccl_id))
        rc = lib.h3d_create(C.byref(h), rank, nranks, device, idbuf)
        if rc != 0:
            msg = self.binding.last_error(None)
            raise H3dError("h3d_create failed (%d): %s" % (rc, msg.decode() if msg else ""))
        self.handle = h
        self.rank, self.nranks = rank, nranks

    @staticmethod
    def nccl_unique_id():
        lib = gpu_library()
        buf = (C.c_char * 128)()
        lib.h3d_get_nccl_unique_id.argtypes = [C.c_void_p]
        if lib.h3d_get_nccl_unique_id(buf) != 0:
            raise H3dError("h3d_get_nccl_unique_id failed")
        return bytes(buf)

    def kernel_launches(self):
        return int(self.binding.lib.h3d_kernel_launches(self.handle))

    def set_halo(self, ranks, counts, faces, sides):
        a = [np.ascontiguousarray(x, dtype=np.int32) for x in (ranks, counts, faces, sides)]
        self.call("set_halo", len(a[0]), *[_ptr(x, np.int32) for x in a])

    def close(self):
        if self.handle:
            self.binding.lib.h3d_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
