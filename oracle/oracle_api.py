"""TEST INFRASTRUCTURE -- ctypes binding of the CPU oracle (oracle/h3d_oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  It exposes the oracle behind the same `Api` interface as the product's GpuApi so that the very
same DGSem driver code runs against the restated reference algorithm.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from horses3d_b200 import build as _build          # noqa: E402
from horses3d_b200.capi import Api, Binding, _D, _ptr  # noqa: E402

_lib = None


def library():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build_oracle())
        _lib.orc_create.restype = C.c_void_p
        _lib.orc_destroy.argtypes = [C.c_void_p]
        _lib.orc_nodal.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 7
    return _lib


class OracleApi(Api):
    name = "oracle"
    EXTRA = {"download_faces": [_D] * 5}

    def __init__(self):
        lib = library()
        self.binding = Binding(lib, "orc_", self.EXTRA)
        self.handle = C.c_void_p(lib.orc_create())

    def download_faces(self, nFace, n):
        out = {k: np.empty((nFace, 2, n, n, 5)) for k in ("Q", "U_x", "U_y", "U_z", "fStar")}
        self.call("download_faces", *[_ptr(out[k], np.float64) for k in ("Q", "U_x", "U_y", "U_z", "fStar")])
        return out

    def close(self):
        if self.handle:
            library().orc_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nodal(N, nodes):
    """The oracle's own restatement of the 1-D operators (K6 pin)."""
    n = N + 1
    x, w = np.zeros(n), np.zeros(n)
    D, hatD, sharpD = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n))
    v, b = np.zeros((2, n)), np.zeros((2, n))
    library().orc_nodal(N, nodes, *[a.ctypes.data for a in (x, w, D, hatD, sharpD, v, b)])
    return dict(x=x, w=w, D=D, hatD=hatD, sharpD=sharpD, v=v, b=b)
