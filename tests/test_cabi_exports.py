"""The C-ABI library must load and export every symbol that include/h3d_gpu.h declares (no compute calls here),
and must refuse to run without a device instead of falling back to anything."""
import ctypes as C
import os
import re

import pytest

from horses3d_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "h3d_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(h3d_[A-Za-z_0-9]+)\s*\(", src)))


def test_header_declares_the_documented_entry_points():
    names = declared_functions()
    for must in ("h3d_create", "h3d_set_physics", "h3d_set_basis", "h3d_set_mesh", "h3d_set_halo", "h3d_upload_Q", "h3d_download",
                 "h3d_compute_time_derivative", "h3d_rk_step", "h3d_max_residuals", "h3d_max_timestep", "h3d_volume_integral",
                 "h3d_has_nan", "h3d_destroy", "h3d_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    build.build_gpu()
    from horses3d_b200.capi import gpu_library
    lib = gpu_library()
    missing = [f for f in declared_functions() if not hasattr(lib, f)]
    assert not missing, missing


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from horses3d_b200.capi import GpuApi, H3dError
    with pytest.raises(H3dError, match="no CUDA device|no CPU fallback"):
        GpuApi()


def test_oracle_is_not_reachable_from_the_product_package():
    """The product must never import, link or call the oracle."""
    pkg = os.path.join(ROOT, "horses3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f == "build.py":      # holds the recipe that COMPILES the checker; building it is not using it
                continue
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_api" not in text and "h3d_oracle" not in text and "orc_" not in text, os.path.join(dirpath, f)
                # nor the host-loop backend of the p-nonconforming kernels (tests/emu): the device library has the CUDA backend only
                assert "emu_api" not in text and "h3d_mixed_emu" not in text and "emu_" not in text and "HostBackend" not in text, os.path.join(dirpath, f)
    import subprocess
    syms = subprocess.run(["nm", "-D", "--defined-only", os.path.join(pkg, "csrc", "libh3dgpu.so")], stdout=subprocess.PIPE, text=True).stdout
    assert "emu_" not in syms and "HostBackend" not in syms


def test_fortran_interfaces_are_in_step_with_the_header():
    """integration/h3d_gpu_interfaces.f90 (the ISO_C_BINDING interface block a maintainer of the reference compiles into the
    adapter) is generated from include/h3d_gpu.h: it must be current and cover every entry point and every field of H3dPhysics."""
    import subprocess
    import sys
    assert subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gen_fortran_interfaces.py"), "--check"]).returncode == 0
    text = open(os.path.join(ROOT, "integration", "h3d_gpu_interfaces.f90")).read()
    for f in declared_functions():
        assert 'bind(C, name="%s")' % f in text or 'name="%s")' % f in text, f
    from horses3d_b200.physics import H3dPhysics
    for name, _ in H3dPhysics._fields_:
        assert re.search(r":: %s\b" % name, text), name
    assert max(len(l) for l in text.splitlines()) <= 132


def test_fortran_adapter_calls_match_the_interface_block():
    """integration/H3DGpuAdapter.f90 (the module a maintainer adds to the reference; no Fortran compiler here): every h3d_*
    entry point it calls must be declared in the generated interface block with the same number of arguments."""
    import re
    iface = open(os.path.join(ROOT, "integration", "h3d_gpu_interfaces.f90")).read()
    decl = {}
    for m in re.finditer(r"function\s+(h3d_[a-zA-Z_0-9]+)\s*\((.*?)\)\s*bind", iface.replace("&\n", " "), flags=re.S):
        decl[m.group(1).lower()] = len([a for a in m.group(2).split(",") if a.strip()])
    src = open(os.path.join(ROOT, "integration", "H3DGpuAdapter.f90")).read()
    src = "\n".join(l.split("!")[0] for l in src.splitlines()).replace("&\n", " ")
    calls = 0
    for m in re.finditer(r"\b(h3d_[a-z_0-9]+)\s*\(", src, flags=re.I):
        name = m.group(1).lower()
        if name.startswith("h3d_gpu_") or name not in decl:
            assert name.startswith("h3d_gpu_") or name == "h3d_stepper", "unknown entry point %s" % name
            continue
        depth, i, args, cur = 1, m.end(), [], ""
        while depth:                                   # split the argument list at top-level commas
            c = src[i]
            depth += c in "([" ; depth -= c in ")]"
            if depth == 1 and c == ",":
                args.append(cur); cur = ""
            elif depth:
                cur += c
            i += 1
        args.append(cur)
        assert len([a for a in args if a.strip()]) == decl[name], (name, args, decl[name])
        calls += 1
    assert calls >= 18
    # the hooks the adapter exports are the ones INTEGRATION.md lists
    for hook in ("ComputeTimeDerivative_GPU", "TakeRK3Step_GPU", "TakeRK5Step_GPU", "ComputeMaxResiduals_GPU", "MaxTimeStep_GPU",
                 "ScalarVolumeIntegral_GPU", "checkForNan_GPU", "h3d_gpu_setup", "h3d_gpu_download_state"):
        assert hook in src and hook in open(os.path.join(ROOT, "INTEGRATION.md")).read(), hook
