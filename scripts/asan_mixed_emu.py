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
