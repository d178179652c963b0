"""BASELINE configs[4] (tutorials/Cylinder on MESH/cyl_circ.msh, Smagorinsky, explicit RK3): the GMSH reader stand-in, the
oracle, the device path and the partitioned multi-GPU path on the tutorial's own mesh."""
import math
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cylinder_tutorial as ct  # noqa: E402
from horses3d_b200.dgsem import DGSem  # noqa: E402
from horses3d_b200.hostmesh import GAUSS  # noqa: E402


def test_gmsh_reader_builds_the_tutorial_mesh():
    """Read_GMSH.f90:203-842: 2184 27-node hexahedra, boundary faces from the physical surfaces, all six faces of every element
    curved (order 2); the metric terms must be positive, the volume that of the annulus and the cylinder's area pi D H."""
    phys = ct.physics()
    m = ct.mesh(phys)
    assert m.sizes()[:2] == (2184, 8792)
    ft, fz = m.array("faceType"), m.array("faceZone")
    assert [int(((ft == 2) & (fz == z)).sum()) for z in range(5)] == [56, 2184, 2184, 28, 28] and int((ft == 0).sum()) == 0
    m.geometry(3, GAUSS)
    assert m.array("jacobian").min() > 0.0
    assert abs(m.array("volume").sum() - math.pi * (45.0 ** 2 - 0.5 ** 2)) < 3e-3         # order-2 approximation of the two circles
    assert abs(m.array("faceSurface")[(ft == 2) & (fz == 0)].sum() - math.pi) < 1e-6
    if os.path.exists("/root/reference/tutorials/Cylinder/MESH/cyl_circ.msh"):
        assert open("/root/reference/tutorials/Cylinder/MESH/cyl_circ.msh", "rb").read() == open(ct.MESH, "rb").read()


@pytest.mark.gpu
def test_tutorial_cylinder_smagorinsky_on_gpu_matches_oracle(gpu_api_cls):
    """20 RK3 steps of the tutorial configuration with Smagorinsky: residuals, drag and lift of the device path against the oracle."""
    from oracle.oracle_api import OracleApi
    phys = ct.physics()
    m = ct.mesh(phys).geometry(3, GAUSS)
    (ro, cdo, clo), (rg, cdg, clg) = [ct.run(DGSem(api, m, phys), 20) for api in (OracleApi(), gpu_api_cls())]
    assert np.abs(ro).max() > 1e-3 and abs(cdo) > 1e-3
    assert np.abs((rg - ro) / ro).max() < 1e-11
    assert abs(cdg - cdo) < 1e-11 * abs(cdo) and abs(clg - clo) < 1e-11 * max(abs(cdo), abs(clo))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from horses3d_b200.capi import GpuApi
        obj = [GpuApi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        phys = ct.physics()
        g = ct.mesh(phys)
        part = g.partition(world, "metis")                       # the reference's METIS_PartMeshDual (METISPartitioning.f90:151)
        m = g.extract(part, rank).geometry(3, GAUSS)
        res, cd, cl = ct.run(DGSem(GpuApi(rank=rank, nranks=world, device=rank, nccl_id=obj[0]), m, phys), 20)
        # the surface integrals come back summed over the ranks (h3d_surface_integral; SurfaceIntegrals.f90 does the MPI sum)
        tot = [None] * world
        dist.all_gather_object(tot, (m.nElem, len(m.array("haloCount"))))
        q.put((rank, res, cd, cl, [t[0] for t in tot], [t[1] for t in tot]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_tutorial_cylinder_smagorinsky_partitioned_matches_single_gpu(gpu_api_cls, world):
    """configs[4] as stated: METIS partition over `world` B200s, NCCL face exchange; residuals, drag and lift of the partitioned
    run against the single-GPU run (bounds: see tests/test_gpu_multirank.py on locally rebuilt MPI-face geometry)."""
    import queue
    import time

    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs, this box has %d (run under gpurun --gpus %d)" % (world, torch.cuda.device_count(), world))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29300 + os.getpid() % 500, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, t0 = [], time.time()
    while len(got) < world:
        try:
            got.append(q.get(timeout=2))
        except queue.Empty:
            dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
            if dead or time.time() - t0 > 400:
                for p in procs:
                    p.kill()
                pytest.fail("a rank exited with %s / timed out" % dead)
    for p in procs:
        p.join(timeout=60)
    phys = ct.physics()
    r1, cd1, cl1 = ct.run(DGSem(gpu_api_cls(), ct.mesh(phys).geometry(3, GAUSS), phys), 20)
    for rank, res, cd, cl, nel, nnb in got:
        assert sum(nel) == 2184 and min(nnb) >= 1
        assert np.abs((res - r1) / r1).max() < 1e-7, (res, r1)
        assert abs(cd - cd1) < 1e-8 * abs(cd1) and abs(cl - cl1) < 1e-8 * max(abs(cd1), abs(cl1)), (cd, cd1, cl, cl1)
    print("configs[4] on %d GPUs: elements per rank %s, cd %.12f (1 GPU %.12f), cl %.3e (1 GPU %.3e)" % (world, got[0][4], got[0][2], cd1, got[0][3], cl1))
