"""Prints the GPU-vs-oracle error of every parity case (tests/test_gpu_parity.py CASES) for the library selected
by H3D_GPU_LIB (default: production build).  Run on the GPU box; output is kept under profiles/."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from horses3d_b200.capi import GpuApi
from horses3d_b200.physics import make_physics
from parity import get_mesh, rel_err
from test_gpu_parity import CASES, run_pair

print("library:", os.environ.get("H3D_GPU_LIB", "libh3dgpu.so (production build, -fmad=false)"))
print("%-4s %-3s %-5s %-5s %-7s %-60s %10s %10s %12s" % ("ne", "N", "nodes", "amp", "shuffle", "physics", "err gradU", "err QDot", "bit-equal"))
for ne, N, nodes, amp, shuffle, kw in CASES:
    mesh = get_mesh(ne, N, nodes, amp, shuffle)
    (so, o), (sg, g) = run_pair(GpuApi, mesh, make_physics(**kw))
    eg = max(rel_err(g[k], o[k]) for k in ("U_x", "U_y", "U_z")) if kw.get("flow", "NS") != "Euler" else 0.0
    eq = rel_err(g["QDot"], o["QDot"])
    same = np.array_equal(g["QDot"], o["QDot"])
    print("%-4d %-3d %-5d %-5.2f %-7s %-60s %10.2e %10.2e %12s" % (ne, N, nodes, amp, shuffle, str(kw)[:60], eg, eq, same))
