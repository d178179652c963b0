#!/usr/bin/env python
"""bench.py -- DOF-updates/s (and time per DOF per RK stage) of the explicit RK3 Navier-Stokes step.

Metric (BASELINE.json): Taylor-Green vortex, P=7, compressible NS (Re 1600, M 0.08), Standard DG + BR1 + Roe,
explicit RK3.  A "step" is one RK3 time step (three residual evaluations + updates).

  python bench.py --gpus N --steps K --warmup W            # this framework on N B200s
  python bench.py --impl reference --steps K --warmup W    # restated reference algorithm on the host cores

Workload (default): STRONG scaling of BASELINE configs[3], the north star's target: the 64^3-element P=7 curvilinear periodic
          box (134 M DOF, fits one B200) on N = 1, 2, 4, 8 GPUs, partitioned element-wise (METIS_PartMeshDual, as the
          reference) with NCCL face exchange.  The per-GPU rate of the N=1 run is that of configs[1] (32^3, 16.8 M DOF)
          within 2 % (profiles/r1_j_round_end): both are far larger than L2.
          --weak : 32^3 elements per GPU ((32,32,32) = configs[1], (64,32,32), (64,64,32), (64,64,64)), the round-1 behaviour
          --ne E : E^3 elements (strong) or E^3 elements per GPU (--weak); the configuration sweeps use it
One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B_ALG_NS_STAGE = lambda n: 600.0 + 3072.0 / n            # bytes per DOF per RK stage (NS/BR1/StandardDG)
B_ALG_VOLUME_KERNEL = lambda n: 360.0 + 480.0 / n        # k_volume: Q 40 + gradU 120 + metrics 80 + G 40 read, G 40 + Q 40 written,
                                                         # + fStar read (6 faces x 5 x n^2) + trace write (same) = 480/n  


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ne", type=int, default=0, help="elements per direction: of the whole mesh (strong scaling, default 64) or per GPU (--weak, default 32)")
    ap.add_argument("--weak", action="store_true", help="weak scaling: --ne^3 elements per GPU instead of a fixed global mesh")
    ap.add_argument("--no-self-check", action="store_true")
    ap.add_argument("--order", type=int, default=7)
    ap.add_argument("--amp", type=float, default=0.1, help="curvature amplitude of the mesh mapping")
    ap.add_argument("--partition", default="metis", choices=["metis", "block"])
    ap.add_argument("--dt", type=float, default=1.0e-4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--ref-ne", type=int, default=16, help="elements per direction of the bounded CPU sample (16^3 P=7 = 2.1 M DOF: out of the host's L3)")
    # the other configurations of SURVEY 8d (C1 P=3, C3 Euler split-form sweeps); the defaults are the headline C2 / C4
    ap.add_argument("--flow", default="NS", choices=["NS", "Euler"])
    ap.add_argument("--inviscid", default="standard", choices=["standard", "split-form"])
    ap.add_argument("--averaging", default="standard")
    ap.add_argument("--riemann", default="roe")
    ap.add_argument("--nodes", default="gauss", choices=["gauss", "gauss-lobatto"])
    ap.add_argument("--viscous", default="BR1", choices=["BR1", "BR2", "IP"])
    ap.add_argument("--les", default="none", choices=["none", "smagorinsky", "wale", "vreman"])
    ap.add_argument("--gradient-variables", default="State", choices=["State", "Entropy", "Energy"])
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks/throttle reasons while the timed region runs (B200_PROFILING.md, clocks line)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) >= 7 and s[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.samples)}


def host_threads():
    """All the cores this process may use (the affinity mask, not os.cpu_count(): containers and launchers restrict it)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_sample(args, steps, warmup):
    """The restated reference algorithm (oracle, g++ -O3 -fopenmp, no FMA) on a bounded sample of the same workload, on ALL host
    cores: torchrun exports OMP_NUM_THREADS=1 to its workers, so the thread count is set through the OpenMP runtime of the
    oracle library itself and the count REPORTED is the one its parallel regions really get (orc_num_threads)."""
    from horses3d_b200.dgsem import DGSem, taylor_green_ic
    from horses3d_b200.hostmesh import GAUSS, HostMesh
    from horses3d_b200.physics import make_physics
    from oracle.oracle_api import OracleApi, library
    os.environ["OMP_NUM_THREADS"] = str(host_threads())
    library().orc_set_num_threads(host_threads())
    from horses3d_b200 import hostmesh as _hm
    _hm.set_num_threads(host_threads())                      # the host-side geometry of the sample
    mesh = HostMesh.box(args.ref_ne, amp=args.amp, bFaceOrder=2).connect().geometry(args.order, GAUSS)
    sem = DGSem(OracleApi(), mesh, make_physics(flow="NS", mach=0.08, reynolds=1600.0, riemann="roe"))
    sem.set_initial_condition(taylor_green_ic)
    for _ in range(warmup):
        sem.TakeRK3Step(0.0, args.dt)
    t0 = time.perf_counter()
    for _ in range(steps):
        sem.TakeRK3Step(0.0, args.dt)
    dt = time.perf_counter() - t0
    value = sem.NDOF * 3 * steps / dt
    sample = "TGV P=%d, %d^3 curvilinear elements (%d DOF), %d RK3 steps" % (args.order, args.ref_ne, sem.NDOF, steps)
    return value, dt / steps * 1e3, int(library().orc_num_threads()), sample


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, ms, cores, sample = cpu_sample(args, max(args.steps, 1), max(args.warmup, 1))
    n = args.order + 1
    line = {
        "impl": "reference", "metric": "DOF-updates/s (TGV P=%d explicit RK3, NS/BR1/Roe)" % args.order, "value": value, "unit": "DOF-updates/s",
        "tpdof_s": 1.0 / value, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Taylor-Green vortex Re=1600 M=0.08, P=%d, StandardDG+BR1+Roe, RK3 (bounded CPU sample: %s)" % (args.order, sample), "nodes_per_element": n ** 3},
        "cpu_baseline": {"value": value, "unit": "DOF-updates/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "restated reference algorithm (oracle/h3d_oracle.cpp, g++ -O3 -fopenmp -ffp-contract=off); the Fortran reference cannot be built: no gfortran / MPI here or on the GPU box (profiles/r2_a_parity_dmma/probe.txt)"},
        "e2e": {"value": value, "unit": "DOF-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def grid_for(ngpus, ne):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(ngpus, (ngpus, 1, 1))


KERNEL_SOURCES = ("h3d_kernels.cuh", "h3d_kernels2.cuh", "h3d_physics.cuh", "h3d_tma.cuh", "h3d_mma.cuh")


def source_sha():
    """Fingerprint of what the timed kernels are made of: the headers that define them and the part of h3d_api.cu that launches them
    (launch configuration and the residual's orchestration, from makeOps to the C entry points).  The committed ncu traffic figure
    is only quoted for a library built from the sources it was captured on; files that neither define nor launch the timed
    kernels (the C entry points, the p-nonconforming path of h3d_mixed.cuh) do not enter."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "horses3d_b200", "csrc")
    for f in KERNEL_SOURCES:
        h.update(open(os.path.join(d, f), "rb").read())
    api = open(os.path.join(d, "h3d_api.cu")).read()
    h.update(api[api.index("template <int n> Ops<n> makeOps"):api.index('extern "C" {')].encode())
    return h.hexdigest()[:16]


# FP64 operations per node and residual evaluation, counted by ncu on the headline instantiations (thread-level DADD / DMUL /
# DFMA of the source page; a DFMA or DMMA lane-MAC counts 2): profiles/r2_*/flops.txt.  The peak is the measured DMUL+DADD issue
# rate of this build (no FMA contraction: 18.5 TFLOP/s, scripts/micro/dmma_rate.cu) -- the DFMA peak (33.9) is quoted beside it.
FP64_PEAK_NOFMA_TFLOPS = 18.5
FP64_PEAK_DFMA_TFLOPS = 33.9


def main_b200(args):
    import torch
    import torch.distributed as dist
    from horses3d_b200.capi import GpuApi, _ptr
    from horses3d_b200.dgsem import DGSem, taylor_green_ic
    from horses3d_b200.hostmesh import GAUSS, GAUSSLOBATTO, HostMesh
    from horses3d_b200.physics import make_physics
    import ctypes as C
    nodes = GAUSS if args.nodes == "gauss" else GAUSSLOBATTO
    euler = args.flow == "Euler"
    phys_kw = dict(flow=args.flow, mach=0.08, reynolds=1600.0, riemann=args.riemann, inviscid=args.inviscid, averaging=args.averaging,
                   viscous=args.viscous, gradient_variables=args.gradient_variables, les=args.les)
    headline = (not euler) and args.inviscid == "standard" and args.riemann == "roe" and args.nodes == "gauss" and args.viscous == "BR1" \
        and args.gradient_variables == "State" and args.les == "none"
    scheme = "%s, %s%s+%s" % (args.flow, "StandardDG" if args.inviscid == "standard" else "SplitDG-" + args.averaging,
                              "" if euler else "+" + args.viscous + ("" if args.gradient_variables == "State" else "(" + args.gradient_variables + " variables)")
                              + ("" if args.les == "none" else "+LES-" + args.les), args.riemann)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # torchrun pins OMP_NUM_THREADS=1; the host-side metric construction is OpenMP code: give each rank its share of the cores
    # (the environment variable alone does not do it: importing torch has already started the OpenMP runtime with torchrun's value)
    os.environ["OMP_NUM_THREADS"] = str(max(1, host_threads() // max(world, 1)))
    from horses3d_b200 import hostmesh as _hm
    host_omp_threads = _hm.set_num_threads(max(1, host_threads() // max(world, 1)))
    if world != args.gpus and world > 1:
        raise SystemExit("WORLD_SIZE (%d) != --gpus (%d)" % (world, args.gpus))
    torch.cuda.set_device(local)
    nccl_id = None
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        obj = [GpuApi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        nccl_id = obj[0]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- mesh: every rank builds the (cheap) global connectivity, partitions it identically, keeps its part
    N, n = args.order, args.order + 1
    px, py, pz = grid_for(world, args.ne)
    if args.weak:
        nel = args.ne or 32
        ex, ey, ez = nel * px, nel * py, nel * pz                  # nel^3 elements per GPU
    else:
        ex = ey = ez = args.ne or 64                               # the whole mesh, whatever N
    gmesh = HostMesh.box(ex, amp=args.amp, bFaceOrder=2, ney=ey, nez=ez).connect()
    nElemGlobal = gmesh.nElem
    if world > 1:
        if args.partition == "metis":
            part = gmesh.partition(world, "metis")
        else:
            idx = np.arange(nElemGlobal)
            bx, by, bz = ex // px, ey // py, ez // pz
            part = ((idx % ex) // bx + px * (((idx // ex) % ey) // by + py * ((idx // (ex * ey)) // bz))).astype(np.int32)
        mesh = gmesh.extract(part, rank)
        mesh.geometry(N, nodes)
    else:
        mesh = gmesh.geometry(N, nodes)
    api = GpuApi(rank=rank, nranks=world, device=local, nccl_id=nccl_id)
    sem = DGSem(api, mesh, make_physics(**phys_kw))
    Q0 = taylor_green_ic(sem.node_coordinates())
    sem.set_Q(Q0)
    ndof_global = nElemGlobal * n ** 3
    dt = args.dt

    # ---- device-resident timing (value)
    for _ in range(args.warmup):
        sem.TakeRK3Step(0.0, dt)
    api.call("synchronize")
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    l0 = api.kernel_launches()
    api.call("timer_begin")
    for _ in range(args.steps):
        sem.TakeRK3Step(0.0, dt)
    ms = C.c_double()
    api.call("timer_end", C.byref(ms))
    api.call("synchronize")
    launches = api.kernel_launches() - l0
    barrier()
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    t_ms = torch.tensor([ms.value], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    total_ms = float(t_ms.item())
    value = ndof_global * 3 * args.steps / (total_ms * 1e-3)
    assert not sem.checkForNan(), "solution diverged during the benchmark"

    # ---- kernel-level timing of the dominant kernel (roofline), a short profiled pass
    roof = None
    api.call("set_option", b"profile_kernels=1")        # every rank steps (the halo exchange is collective); rank 0 reports
    for _ in range(3):
        sem.TakeRK3Step(0.0, dt)
    prof = (C.c_double * 16)()
    api.binding.lib.h3d_kernel_profile.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
    api.binding.lib.h3d_kernel_profile(api.handle, prof, 16)
    api.call("set_option", b"profile_kernels=0")
    barrier()
    if rank == 0:
        try:
            # prof: [ms_gradient, n_gradient, ms_riemann, n_riemann, ms_volume, n_volume, ms_prolong, n_prolong]
            vol_ms = prof[4] / max(prof[5], 1.0)
            peaks = {}
            pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
            peak, src = 6650.0, "fallback"
            if os.path.exists(pk):
                peaks = json.load(open(pk))
                peak, src = float(peaks.get("hbm_gbs", 6650.0)), "measured"
            ndof_local = sem.NDOF
            # Euler without gradients: stage 240 + 936/n, volume kernel Q 40 + metrics 80 + G 40 read, G 40 + Q 40 written
            b_vol = (240.0 + 480.0 / n) if euler else B_ALG_VOLUME_KERNEL(n)
            b_stage = (240.0 + 936.0 / n) if euler else B_ALG_NS_STAGE(n)
            achieved = b_vol * ndof_local / (vol_ms * 1e-3) / 1e9
            stage_gbs = b_stage * value / world / 1e9
            # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this very workload
            traffic, tsrc = None, None
            tf = os.path.join(ROOT, "profiles", "ncu_traffic.json")
            if os.path.exists(tf) and world == 1 and headline:
                # per-launch DRAM bytes scale with the element count (every tile moves the same bytes): the capture is
                # quoted per element and only for the kernel sources it was taken on
                rec = json.load(open(tf)).get("P%d" % args.order)
                if rec and rec.get("source_sha") == source_sha():
                    traffic, tsrc = rec["k_volume_bytes_per_element"] * sem.nElem, rec["source"]
                elif rec:
                    tsrc = "stale: the committed ncu capture (%s) was taken on other kernel sources" % rec["source"]
            roof = {"bound": "hbm", "kernel": "k_volume<%d>" % n, "achieved": achieved, "peak": peak, "peak_source": src, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": tsrc, "avg_launch_ms": vol_ms,
                    "alg_bytes_per_dof": b_vol,
                    "per_kernel_ms": {"gradient": prof[0] / max(prof[1], 1), "riemann": prof[2] / max(prof[3], 1), "volume": vol_ms},
                    "stage": {"alg_bytes_per_dof_stage": b_stage, "achieved": stage_gbs, "frac": stage_gbs / peak}}
            ff = os.path.join(ROOT, "profiles", "fp64_flops.json")
            key = "%s_%s_P%d" % (args.flow, args.inviscid, args.order)
            if os.path.exists(ff) and key in json.load(open(ff)):
                rec = json.load(open(ff))[key]
                tf64 = rec["volume_flop_per_dof"] * ndof_local / (vol_ms * 1e-3) / 1e12
                roof["fp64"] = {"kernel_flop_per_dof": rec["volume_flop_per_dof"], "achieved_tflops": tf64, "peak_tflops": FP64_PEAK_NOFMA_TFLOPS,
                                "frac": tf64 / FP64_PEAK_NOFMA_TFLOPS, "peak_dfma_tflops": FP64_PEAK_DFMA_TFLOPS, "source": rec["source"],
                                "stage_flop_per_dof": rec.get("stage_flop_per_dof"),
                                "stage_achieved_tflops": rec["stage_flop_per_dof"] * value / world / 1e12 if rec.get("stage_flop_per_dof") else None}
                if rec.get("bound") == "fp64":
                    roof["bound"] = "fp64"
        except Exception as ex:  # profile hooks are optional
            roof = {"bound": "hbm", "error": str(ex)}

    # ---- timeline of one RK stage on both streams of every rank (h3d_stage_timeline): where the halo exchange sits
    timeline = None
    if world > 1:
        api.call("set_option", b"timeline=1")
        sem.TakeRK3Step(0.0, dt)
        marks = (C.c_double * 11)()
        api.binding.lib.h3d_stage_timeline.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        api.binding.lib.h3d_stage_timeline(api.handle, marks, 11)
        api.call("set_option", b"timeline=0")
        allm = [None] * world
        dist.all_gather_object(allm, [round(float(x), 4) for x in marks])
        timeline = {"marks": ["start", "q_halo_begin", "q_halo_end", "gradient_interior_end", "gradient_mpi_end", "grad_halo_begin", "grad_halo_end",
                              "riemann_local_end", "volume_interior_end", "riemann_mpi_end", "volume_mpi_end"],
                    "ms_per_rank": allm}
        barrier()

    # ---- end-to-end through the C ABI with HOST buffers (strict drop-in: state crosses PCIe every step)
    e2e = None
    if not args.no_e2e:
        hostQ = torch.empty(Q0.shape, dtype=torch.float64, pin_memory=True)
        hostQ.copy_(torch.from_numpy(sem.Q()))
        qnp = hostQ.numpy()
        nsteps = max(3, min(args.steps, 5))
        barrier()
        t0 = time.perf_counter()
        for _ in range(nsteps):
            api.call("upload_Q", _ptr(qnp, np.float64))
            sem.TakeRK3Step(0.0, dt)
            api.call("download", _ptr(qnp, np.float64), None, None, None, None)
            sem.ComputeMaxResiduals()
        barrier()
        el = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        e2e_val = ndof_global * 3 * nsteps / float(el.item())
        # resident mode: what a time loop of the reference does per step once the state lives on the device
        barrier()
        t0 = time.perf_counter()
        for _ in range(nsteps):
            sem.TakeRK3Step(0.0, dt)
            sem.ComputeMaxResiduals()
            sem.volume_monitors()
            sem.checkForNan()
        barrier()
        el2 = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(el2, op=dist.ReduceOp.MAX)
        e2e = {"value": e2e_val, "unit": "DOF-updates/s", "h2d_bytes_per_step": int(qnp.nbytes), "d2h_bytes_per_step": int(qnp.nbytes + 48),
               "mode": "strict drop-in: h3d_upload_Q + h3d_rk_step + h3d_download + h3d_max_residuals per step, pinned host buffers",
               "resident": {"value": ndof_global * 3 * nsteps / float(el2.item()), "unit": "DOF-updates/s", "d2h_bytes_per_step": 8 * (6 + 4 * 4 + 6),
                            "mode": "state resident on the device; per step h3d_rk_step + residuals + KE/KE-rate/enstrophy monitors + NaN check"}}

    # ---- self-check of the run that was just timed: restart from the initial condition, two RK3 steps, and compare the max
    # residuals, kinetic energy and enstrophy (globally reduced) with the single-GPU values of the same library committed under
    # tests/golden/ (the single-GPU path is bit-identical to the oracle: tests/test_gpu_parity_large.py).  The library gives both
    # ranks of an MPI face the geometry of its left-side owner (option sync_mpi_face_geometry), so a partitioned run differs from
    # the single-GPU one only by the order of the cross-rank reductions: 1e-11.  (Without that option each rank builds the geometry
    # of its MPI faces from its own element, as the reference does, and the residuals differ by up to 1.4e-4 of the largest one at
    # 8 GPUs: round-off of the metric terms amplified by the pressure and the lift weights, profiles/r2_k_scaling8.)  Every rank
    # computes the same verdict from globally reduced values: nobody is left waiting in a collective.
    check = None
    if headline and not args.no_self_check:
        sem.set_Q(Q0)
        for _ in range(2):
            sem.TakeRK3Step(0.0, 1.0e-4)
        got = [float(x) for x in sem.ComputeMaxResiduals()] + [float(sem.volume_monitor("kinetic energy")), float(sem.volume_monitor("enstrophy"))]
        nan = bool(sem.checkForNan())
        gf = os.path.join(ROOT, "tests", "golden", "scale_check.json")
        key = "ne%dx%dx%d_P%d_amp%g" % (ex, ey, ez, N, args.amp)
        ref = json.load(open(gf)).get(key) if os.path.exists(gf) else None
        if ref:
            rv = ref["values"]
            e_res = max(abs(a - b) for a, b in zip(got[:5], rv[:5])) / max(abs(b) for b in rv[:5])
            e_int = max(abs(a - b) / abs(b) for a, b in zip(got[5:], rv[5:]))
            tol_res = tol_int = 1e-11
            check = {"key": key, "residual_err_vs_single_gpu": e_res, "integral_err_vs_single_gpu": e_int, "tolerances": [tol_res, tol_int],
                     "ok": bool(e_res < tol_res and e_int < tol_int and not nan), "values": got}
        else:
            check = {"key": key, "values": got, "ok": not nan, "note": "no committed single-GPU values for this mesh"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cms, cores, sample = cpu_sample(args, 10, 2)
        cpu = {"value": v, "unit": "DOF-updates/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": "DOF-updates/s (TGV P=%d explicit RK3, %s)" % (N, "NS/BR1/Roe" if headline else scheme), "value": value, "unit": "DOF-updates/s", "tpdof_s": 1.0 / value,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Taylor-Green vortex Re=1600 M=0.08, %dx%dx%d curvilinear hex elements, P=%d %s, %s, RK3, fixed dt"
                                   % (ex, ey, ez, N, "Gauss" if args.nodes == "gauss" else "Gauss-Lobatto",
                                      "StandardDG+BR1+Roe" if headline else scheme),
                       "ndof": ndof_global, "elements_per_gpu": nElemGlobal // world, "partition": args.partition if world > 1 else "none",
                       "host_threads_per_rank": host_omp_threads,
                       "contraction": "DMMA" if os.environ.get("H3D_USE_MMA") == "1" else "CUDA cores, bit-identical to the oracle",
                       "l2": "inputs larger than L2 (state + gradients + metrics = %.1f GB per GPU)" % (sem.NDOF * 8 * (30 + 10) / 1e9)},
            "gpu_launches": int(launches), "clocks": sampler.summary() if sampler else None,
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "self_check": check, "timeline": timeline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if check is not None and not check["ok"] and rank == 0:
        # reported in the JSON line ("self_check": {"ok": false}); the exit code stays 0 so that the line is not lost
        sys.stderr.write("WARNING: self-check outside its tolerance: %s\n" % check)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_b200(a)
