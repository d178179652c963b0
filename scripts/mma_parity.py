"""Error of the DMMA (FP64 tensor-core) contraction path against the CPU oracle, per field, on n = 8 cases (GPU box).
usage: python scripts/mma_parity.py [ne ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from horses3d_b200.capi import GpuApi
from horses3d_b200.dgsem import DGSem, taylor_green_ic
from horses3d_b200.hostmesh import GAUSS
from horses3d_b200.physics import make_physics
from oracle.oracle_api import OracleApi
from parity import channel_state, get_mesh, perturbed_tgv, rel_err

def run(api, mesh, phys, ic, opt=None, dt=1e-3):
    sem = DGSem(api, mesh, phys)
    if opt: api.call("set_option", opt.encode())
    sem.set_initial_condition(ic)
    sem.ComputeTimeDerivative(0.0)
    d = sem.download(QDot=True, gradients=True)
    sem.TakeRK3Step(0.0, dt)
    d["Q1"] = sem.Q()
    return d

cases = [("tgv M0.08 perturbed ne=4", 4, dict(flow="NS", mach=0.08, reynolds=1600.0), perturbed_tgv, None),
         ("tgv M0.08 ne=8", 8, dict(flow="NS", mach=0.08, reynolds=1600.0), taylor_green_ic, None),
         ("M0.3 Re200 perturbed ne=6", 6, dict(flow="NS", mach=0.3, reynolds=200.0), lambda x: perturbed_tgv(x, 0.1), None),
         ("channel BC ne=4", 4, dict(flow="NS", mach=0.3, reynolds=200.0), None, "channel")]
for ne in [int(a) for a in sys.argv[1:]]:
    cases.append(("tgv M0.08 ne=%d" % ne, ne, dict(flow="NS", mach=0.08, reynolds=1600.0), taylor_green_ic, None))
for name, ne, kw, ic, bc in cases:
    phys = make_physics(**kw)
    mesh = get_mesh(ne, 7, GAUSS, 0.1, True, bc=bc, phys=phys)
    if ic is None: ic = lambda x: channel_state(x, phys)
    o = run(OracleApi(), mesh, phys, ic)
    g0 = run(GpuApi(), mesh, phys, ic, "mma=0")
    g1 = run(GpuApi(), mesh, phys, ic, "mma=1")
    print("%-28s" % name, " ".join("%s: cuda-core %.1e dmma %.1e |" % (k, rel_err(g0[k], o[k]), rel_err(g1[k], o[k])) for k in ("U_x", "U_y", "U_z", "QDot", "Q1")), flush=True)
