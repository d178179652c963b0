"""In-tree builds of the three native libraries.

  libh3dhost.so    horses3d_b200/host/   g++  (driver-side mesh / geometry / partition, C++17 + OpenMP)
  libh3dgpu.so     horses3d_b200/csrc/   nvcc (sm_100a kernels + the C-ABI of include/h3d_gpu.h)
  libh3doracle.so  oracle/               g++  (CPU restatement of the reference algorithm; TEST INFRASTRUCTURE)

The product package never loads the oracle; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline do.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "horses3d_b200")
HOST_DIR = os.path.join(PKG, "host")
CSRC_DIR = os.path.join(PKG, "csrc")
ORACLE_DIR = os.path.join(ROOT, "oracle")
INCLUDE_DIR = os.path.join(ROOT, "include")

HOST_LIB = os.path.join(HOST_DIR, "libh3dhost.so")
DRIVER_BIN = os.path.join(HOST_DIR, "h3d_driver")
GPU_LIB = os.path.join(CSRC_DIR, "libh3dgpu.so")
GPU_LIB_FMA = os.path.join(CSRC_DIR, "libh3dgpu_fma.so")
ORACLE_LIB = os.path.join(ORACLE_DIR, "libh3doracle.so")

CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
METIS_A = os.path.join(CUDA_HOME, "targets/x86_64-linux/lib/libmetis_static.a")


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + "\n")
        raise RuntimeError("build failed: " + " ".join(cmd[:3]))
    return r.stdout


def _sources(d, exts):
    return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(exts)]


def build_host(force=False):
    srcs = _sources(HOST_DIR, (".cpp", ".hpp"))
    if force or _newer(HOST_LIB, srcs):
        cmd = ["g++", "-O2", "-std=c++17", "-fopenmp", "-fPIC", "-shared", os.path.join(HOST_DIR, "capi.cpp"), "-o", HOST_LIB]
        if os.path.exists(METIS_A):
            cmd[1:1] = ["-DH3D_HAS_METIS"]
            cmd.append(METIS_A)
        _run(cmd)
    return HOST_LIB


def build_driver(force=False):
    """h3d_driver: the native (C++) host-side driver above the C ABI (horses3d_b200/host/h3d_driver.cpp, dgsem.hpp)."""
    srcs = _sources(HOST_DIR, (".cpp", ".hpp")) + _sources(INCLUDE_DIR, (".h",))
    if force or _newer(DRIVER_BIN, srcs):
        _run(["g++", "-O2", "-std=c++17", "-fopenmp", os.path.join(HOST_DIR, "h3d_driver.cpp"), "-o", DRIVER_BIN, "-ldl"])
    return DRIVER_BIN


def nvcc_path():
    p = shutil.which("nvcc") or os.path.join(CUDA_HOME, "bin", "nvcc")
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


# -fmad=false: the reference's gfortran RELEASE build contracts no FMA; without contraction the device results are
# bit-identical to the oracle (profiles/r1_a_first_correct/parity_strict.txt) at a measured cost of ~2 % (HBM-bound kernels).
GPU_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
             "--expt-relaxed-constexpr", "-Xptxas", "-v", "-fmad=false"]


def build_gpu(force=False, extra=(), out=None):
    """libh3dgpu.so.  `extra`/`out` build a variant of the same sources (e.g. extra=("-fmad=false",) for the
    strict-IEEE parity build, out=GPU_LIB_STRICT)."""
    out = out or GPU_LIB
    srcs = _sources(CSRC_DIR, (".cu", ".cuh", ".h")) + _sources(INCLUDE_DIR, (".h",))
    if force or _newer(out, srcs):
        cus = [s for s in srcs if s.endswith(".cu")]
        cmd = [nvcc_path()] + GPU_FLAGS + list(extra) + ["-shared", "-I", INCLUDE_DIR, "-I", CSRC_DIR] + cus + ["-o", out, "-lnccl"]
        log = _run(cmd)
        with open(os.path.splitext(out)[0] + ".ptxas.txt", "w") as f:
            f.write(log)
    return out


def build_gpu_fma(force=False):
    """Variant WITH FMA contraction (not bit-comparable with the oracle; kept to measure what contraction buys)."""
    return build_gpu(force, extra=("-fmad=true",), out=GPU_LIB_FMA)


def build_oracle(force=False):
    srcs = _sources(ORACLE_DIR, (".cpp", ".hpp", ".inc"))
    if force or _newer(ORACLE_LIB, srcs):
        # -ffp-contract=off: the reference's gfortran RELEASE build (-O3, no -march, no -ffast-math) emits no FMA
        cmd = ["g++", "-O3", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared",
               os.path.join(ORACLE_DIR, "h3d_oracle.cpp"), "-o", ORACLE_LIB]
        _run(cmd)
    return ORACLE_LIB


def build_all(force=False):
    return build_host(force), build_gpu(force), build_oracle(force), build_driver(force)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    {"host": build_host, "gpu": build_gpu, "fma": build_gpu_fma, "oracle": build_oracle, "driver": build_driver, "all": build_all}[which](force=True)
    print("built", which)
