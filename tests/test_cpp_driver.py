"""The native (C++) host-side driver above the C ABI (horses3d_b200/host/h3d_driver.cpp, dgsem.hpp): the reference's own
regressions run through it -- on the CPU against the oracle library (named by path and prefix from HERE: the product never
references the oracle), on the GPU against libh3dgpu.so."""
import os
import subprocess
import sys

import numpy as np
import pytest

from horses3d_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CYLINDER_MESH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "CylinderNSpol3.mesh")   # the reference's Solver/test/TestMeshes/CylinderNSpol3.mesh, copied
K1_RES = np.array([1.6417830052388520E-05, 1.2677577061211545E-01, 1.2677577048633804E-01, 2.4981129585617484E-01, 6.2174425106488129E-01])


def run_driver(*args, check=True):
    env = dict(os.environ)
    for d in sys.path:                      # the NCCL that PyTorch bundles, as horses3d_b200.capi.gpu_library prefers
        cand = os.path.join(d, "nvidia", "nccl", "lib")
        if os.path.exists(os.path.join(cand, "libnccl.so.2")):
            env["LD_LIBRARY_PATH"] = cand + os.pathsep + env.get("LD_LIBRARY_PATH", "")
            break
    r = subprocess.run([build.build_driver()] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=900)
    if check:
        assert r.returncode == 0, r.stderr[-2000:]
    return r


def final_line(r):
    v = [float(x) for x in [l for l in r.stdout.splitlines() if l.startswith("FINAL")][-1].split()[1:]]
    return dict(iter=int(v[0]), t=v[1], residuals=np.array(v[2:7]), ke=v[7], ke_rate=v[8], enstrophy=v[9])


def check_k1(f):
    assert f["iter"] == 5
    assert np.abs(f["residuals"] - K1_RES).max() < 1.0e-7          # test/NavierStokes/TaylorGreen/SETUP/ProblemFile.f90:317-366
    assert abs(f["ke"] - 1.2499879367819486E-01) < 1.0e-11
    assert abs(f["ke_rate"] - (-4.2807806718622574E-04)) < 1.0e-11
    assert abs(f["enstrophy"] - 3.7499683882517909E-01) < 1.0e-11


def test_cpp_driver_reproduces_k1_with_the_oracle_backend():
    lib = build.build_oracle()
    check_k1(final_line(run_driver("--lib", lib, "--prefix", "orc_", "--ne", 32, "--order", 3, "--steps", 5, "--cfl", 0.4, "--dcfl", 0.4)))


@pytest.mark.skipif(not os.path.exists(CYLINDER_MESH), reason="reference test mesh not available on this machine")
def test_cpp_driver_reproduces_the_cylinder_regression_with_the_oracle_backend():
    """test/NavierStokes/Cylinder (K5): SpecMesh file, boundary table, inflow / outflow / wall parameters and the uniform
    initial condition all built by the C++ driver."""
    r = run_driver("--lib", build.build_oracle(), "--prefix", "orc_", "--mesh", CYLINDER_MESH, "--order", 3, "--steps", 100, "--cfl", 0.3, "--dcfl", 0.3,
                   "--mach", 0.3, "--reynolds", 200, "--aoa-phi", 90, "--ic", "uniform", "--bc", "innercylinder:noslipwall", "--bc", "bottom:freeslipwall",
                   "--bc", "top:freeslipwall", "--bc", "back:inflow", "--bc", "left:inflow", "--bc", "front:inflow", "--bc", "right:outflow")
    f = final_line(r)
    res = np.array([8.8131248889811715E+00, 1.7608838068776613E+01, 1.9037533106262516E-01, 2.4301352846288605E+01, 2.4063786464536835E+02])
    assert f["iter"] == 100 and np.abs((f["residuals"] - res) / res).max() < 1.0e-11


@pytest.mark.skipif(not os.path.exists(CYLINDER_MESH), reason="reference test mesh not available on this machine")
def test_cpp_driver_reproduces_the_cylinder_wale_regression_with_the_oracle_backend():
    """test/NavierStokes/CylinderWALE through the C++ driver (LES model and its default intensity chosen in makePhysics)."""
    r = run_driver("--lib", build.build_oracle(), "--prefix", "orc_", "--mesh", CYLINDER_MESH, "--order", 3, "--steps", 100, "--cfl", 0.3, "--dcfl", 0.3,
                   "--mach", 0.3, "--reynolds", 200, "--aoa-phi", 90, "--ic", "uniform", "--les", "wale", "--bc", "innercylinder:noslipwall",
                   "--bc", "bottom:freeslipwall", "--bc", "top:freeslipwall", "--bc", "back:inflow", "--bc", "left:inflow", "--bc", "front:inflow", "--bc", "right:outflow")
    res = np.array([7.9687618041712476, 16.312135941662717, 0.2211855539938163, 21.313216389082029, 218.00956664214917])
    assert np.abs(final_line(r)["residuals"] - res).max() < 1.0e-7


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
K13_ARGS = ["--mesh", os.path.join(GOLDEN, "CylinderNSpol3_1elem_y.mesh"), "--order-file", os.path.join(GOLDEN, "OrdersN2N3N4N5_anisotropy.csv"), "--steps", 100,
            "--cfl", 0.2, "--dcfl", 0.2, "--mach", 0.3, "--reynolds", 45, "--aoa-phi", 90, "--ic", "uniform", "--bc", "innercylinder:noslipwall",
            "--bc", "bottom:freeslipwall", "--bc", "top:freeslipwall", "--bc", "back:inflow", "--bc", "left:inflow", "--bc", "front:inflow", "--bc", "right:outflow"]
K13_RES = np.array([9.5806856005342933E+00, 2.0804408993372231E+01, 3.7668665836122439E-01, 2.8294964263463125E+01, 2.6470704194989690E+02])


def test_cpp_driver_reproduces_the_different_orders_regression_with_the_oracle_backend():
    """test/NavierStokes/CylinderDifferentOrders (K13) through the C++ driver: polynomial order file, p-nonconforming geometry, nodal
    storages and interpolation matrices, h3d_set_mesh_p -- all built natively."""
    f = final_line(run_driver("--lib", build.build_oracle(), "--prefix", "orc_", *K13_ARGS))
    assert f["iter"] == 100 and np.abs(f["residuals"] - K13_RES).max() < 1.0e-11


def test_cpp_driver_fails_loudly_without_a_device_or_with_bad_options():
    import torch
    r = run_driver("--lib", "/nonexistent/libh3dgpu.so", check=False)
    assert r.returncode != 0 and "cannot load" in r.stderr
    r = run_driver("--lib", build.build_oracle(), "--prefix", "orc_", "--ne", 2, "--order", 2, "--riemann", "godunov", check=False)
    assert r.returncode != 0 and "Riemann solver not recognized" in r.stderr
    if not torch.cuda.is_available():
        r = run_driver("--lib", build.build_gpu(), "--ne", 2, "--order", 2, check=False)
        assert r.returncode != 0 and r.stderr.strip()            # no CPU fallback: creation fails with the library's message


@pytest.mark.gpu
def test_cpp_driver_reproduces_k1_on_the_device():
    check_k1(final_line(run_driver("--lib", build.build_gpu(), "--ne", 32, "--order", 3, "--steps", 5, "--cfl", 0.4, "--dcfl", 0.4)))


@pytest.mark.gpu
def test_cpp_driver_device_matches_oracle_on_a_curved_split_form_case():
    args = ["--ne", 4, "--amp", 0.1, "--order", 5, "--nodes", "gauss-lobatto", "--steps", 10, "--inviscid", "split-form", "--averaging", "pirozzoli",
            "--mach", 0.3, "--reynolds", 200, "--viscous", "BR2", "--scheme", "ssprk33"]
    a = final_line(run_driver("--lib", build.build_oracle(), "--prefix", "orc_", *args))
    b = final_line(run_driver("--lib", build.build_gpu(), *args))
    assert abs(a["t"] - b["t"]) <= 1e-13 * a["t"]
    assert np.abs(a["residuals"] - b["residuals"]).max() <= 1e-11 * np.abs(a["residuals"]).max()
    assert abs(a["ke"] - b["ke"]) <= 1e-12 * abs(a["ke"]) and abs(a["enstrophy"] - b["enstrophy"]) <= 1e-11 * abs(a["enstrophy"])
