"""TEST INFRASTRUCTURE -- host-loop backend of the p-nonconforming device path (tests/emu/h3d_mixed_emu.cpp).

Builds tests/emu/libh3dmixedemu.so from the product's own kernel source (horses3d_b200/csrc/h3d_mixed.cuh) with g++ and binds it
behind the `Api` interface, so that the functors and the orchestration libh3dgpu.so runs on the device can be compared with the
oracle on a machine without a GPU.  Only tests/ may import this module."""
import ctypes as C
import os
import subprocess

from horses3d_b200 import build as _build
from horses3d_b200.capi import Api, Binding

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libh3dmixedemu.so")
SRC = os.path.join(HERE, "h3d_mixed_emu.cpp")
NAMES = ["set_physics", "set_wall_distance", "set_face_h", "set_basis", "set_interpolation", "set_mesh_p", "set_boundary_conditions", "upload_Q", "download", "set_source",
         "compute_time_derivative", "rk_step", "rk_stage", "max_residuals", "max_timestep", "volume_integral", "has_nan",
         "surface_integral", "probe", "enable_limiter", "statistics_update", "statistics_download", "snapshot_begin", "snapshot_end"]


def build(force=False):
    deps = [SRC] + [os.path.join(_build.CSRC_DIR, f) for f in ("h3d_mixed.cuh", "h3d_physics.cuh")] + [os.path.join(_build.INCLUDE_DIR, "h3d_gpu.h")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        # -ffp-contract=off is the host image of the device build's -fmad=false
        cmd = ["g++", "-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-attributes", "-Wno-unknown-pragmas",
               "-I", os.path.join(_build.CUDA_HOME, "include"), "-I", _build.INCLUDE_DIR, "-I", _build.CSRC_DIR, SRC, "-o", LIB]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("emu build failed:\n" + r.stdout)
    return LIB


def _lib():
    lib = C.CDLL(build())
    lib.emu_create.restype = C.c_void_p
    lib.emu_create_rank.restype = C.c_void_p
    lib.emu_create_rank.argtypes = [C.c_void_p, C.c_int]
    lib.emu_world_create.restype = C.c_void_p
    lib.emu_world_create.argtypes = [C.c_int]
    lib.emu_world_destroy.argtypes = [C.c_void_p]
    lib.emu_world_abort.argtypes = [C.c_void_p]
    lib.emu_destroy.argtypes = [C.c_void_p]
    lib.emu_kernel_launches.argtypes = [C.c_void_p]
    lib.emu_kernel_launches.restype = C.c_longlong
    return lib


class EmuWorld:
    """The ranks of one emulated partitioned run: one EmuApi(world, rank) per rank, each driven from its own Python thread (the
    exchange and the all-reduce of the host-loop backend meet at a barrier inside the library)."""

    def __init__(self, nranks):
        self._lib = _lib()
        self.nranks = nranks
        self.handle = C.c_void_p(self._lib.emu_world_create(nranks))

    def abort(self):
        """A rank failed: open the barriers so that the other ranks finish and the error can be reported."""
        self._lib.emu_world_abort(self.handle)

    def __del__(self):
        try:
            self._lib.emu_world_destroy(self.handle)
        except Exception:
            pass


class EmuApi(Api):
    name = "emu"

    def __init__(self, world=None, rank=0):
        lib = _lib()
        self.binding = Binding(lib, "emu_", extra={"set_halo": [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]}, names=NAMES + ["set_halo"])
        self.world = world         # keeps the world alive
        self.handle = C.c_void_p(lib.emu_create_rank(world.handle, rank) if world is not None else lib.emu_create())
        self._lib = lib

    def set_halo(self, ranks, counts, faces, sides):
        import numpy as np
        a = [np.ascontiguousarray(x, dtype=np.int32) for x in (ranks, counts, faces, sides)]
        self.call("set_halo", len(a[0]), *[x.ctypes.data for x in a])

    def kernel_launches(self):
        return int(self._lib.emu_kernel_launches(self.handle))

    def close(self):
        if self.handle:
            self._lib.emu_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
