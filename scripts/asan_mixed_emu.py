"""Runs the emulated kernels of the p-nonconforming path (tests/emu) under AddressSanitizer:
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python scripts/asan_mixed_emu.py
No report = no out-of-bounds access in the functors or the orchestration of horses3d_b200/csrc/h3d_mixed.cuh."""
import subprocess
ROOT = __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))
subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=address", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-attributes",
                       "-Wno-unknown-pragmas", "-I", "/usr/local/cuda/include", "-I", ROOT + "/include", "-I", ROOT + "/horses3d_b200/csrc",
                       ROOT + "/tests/emu/h3d_mixed_emu.cpp", "-o", "/tmp/libh3dmixedemu_asan.so"])
import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from emu import emu_api
emu_api.LIB = '/tmp/libh3dmixedemu_asan.so'
emu_api.build = lambda force=False: emu_api.LIB
import mixed_cases as MC
from horses3d_b200.physics import make_physics
phys = make_physics(flow="NS", mach=0.3, reynolds=150.0, riemann="roe")
MC.run_case(emu_api.EmuApi(), MC.channel(phys), phys, zone=2)
MC.run_case(emu_api.EmuApi(), MC.periodic_box(3, 1, 6, seed=7), phys, source=True)
from horses3d_b200 import probes
from test_oracle_pins import cylinder_different_orders
print(cylinder_different_orders(emu_api.EmuApi(), steps=2)[1:])
print("asan run complete")
# the later additions: split form, interior penalty, LES with the wall model, limiter / statistics / snapshot, two emulated ranks
from horses3d_b200.hostmesh import GAUSSLOBATTO
MC.run_case(emu_api.EmuApi(), MC.periodic_box(3, 2, 5, seed=17, nodes=GAUSSLOBATTO), make_physics(flow="NS", mach=0.3, reynolds=200.0, inviscid="split-form", averaging="chandrasekar", riemann="central"))
for kw in (dict(viscous="ip"), dict(les="smagorinsky", les_wall_model="linear")):
    ph2 = make_physics(flow="NS", mach=0.3, reynolds=150.0, riemann="roe", **kw)
    MC.run_case(emu_api.EmuApi(), MC.channel(ph2), ph2, zone=2)
MC.limiter_and_statistics_case(emu_api.EmuApi())
import threading
g = MC.periodic_box(4, 2, 5, seed=7)
part = g.partition(2, "metis")
world = emu_api.EmuWorld(2)
ts = [threading.Thread(target=lambda r=r: MC.run_case(emu_api.EmuApi(world, r), g.extract(part, r, inherit_geometry=True), phys)) for r in range(2)]
[t.start() for t in ts]; [t.join() for t in ts]
print("asan run (second part) complete")
