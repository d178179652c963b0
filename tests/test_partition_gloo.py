"""Host-side logic of the multi-rank path, world_size 2 and 3 over gloo (CPU only).

Each rank extracts its partition of a curved, randomly re-oriented periodic box, evaluates a smooth function at
the nodes of its MPI faces (face frame, from its own element's geometry) and exchanges it with the neighbour in the
halo order given to h3d_set_halo.  Both ranks must see the same physical points node by node, normals must be
identical (the face frame belongs to the global face) and the surface Jacobians must agree.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from horses3d_b200.hostmesh import GAUSS, HostMesh  # noqa: E402


def _worker(rank, world, port, method, N, result):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = HostMesh.box(4, amp=0.1, bFaceOrder=2, shuffle=True, seed=7).connect()
        part = g.partition(world, method)
        m = g.extract(part, rank).geometry(N, GAUSS)
        n2 = (N + 1) ** 2
        ranks, counts = m.array("haloRank").copy(), m.array("haloCount").copy()
        faces, sides = m.array("haloFace").copy(), m.array("haloSide").copy()
        ftype = m.array("faceType")
        assert counts.sum() == (ftype == 3).sum() and len(faces) == counts.sum()
        fx = m.array("faceX").reshape(-1, n2, 3)
        fn = m.array("faceNormal").reshape(-1, n2, 3)
        fj = m.array("faceJacobian").reshape(-1, n2)
        off = 0
        ok = True
        for nb, cnt in zip(ranks, counts):
            ids = faces[off:off + cnt]
            # positions are compared through sin/cos (periodic faces differ by the box length 2 pi)
            mine = torch.from_numpy(np.concatenate([fx[ids].reshape(cnt, -1), fn[ids].reshape(cnt, -1), fj[ids]], axis=1).copy())
            theirs = torch.empty_like(mine)
            reqs = [dist.isend(mine, int(nb)), dist.irecv(theirs, int(nb))]
            for r in reqs:
                r.wait()
            a, b = mine.numpy(), theirs.numpy()
            px = max(np.abs(np.sin(a[:, :3 * n2]) - np.sin(b[:, :3 * n2])).max(), np.abs(np.cos(a[:, :3 * n2]) - np.cos(b[:, :3 * n2])).max())  # periodic-safe
            pn = np.abs(a[:, 3 * n2:6 * n2] - b[:, 3 * n2:6 * n2]).max()
            pj = np.abs(a[:, 6 * n2:] - b[:, 6 * n2:]).max()
            ok = ok and px < 1e-11 and pn < 1e-11 and pj < 1e-11
            # the local side must be the complement of the neighbour's side on every shared face
            s_mine = torch.from_numpy(sides[off:off + cnt].astype(np.int64).copy()); s_theirs = torch.empty_like(s_mine)
            reqs = [dist.isend(s_mine, int(nb)), dist.irecv(s_theirs, int(nb))]
            for r in reqs:
                r.wait()
            ok = ok and bool(((s_mine + s_theirs) == 1).all())
            off += cnt
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        nel = torch.tensor([m.nElem])
        dist.all_reduce(nel)
        if rank == 0:
            result.put((int(flag.item()), int(nel.item()), g.nElem))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,method", [(2, "metis"), (2, "block"), (3, "metis")])
def test_halo_lists_and_mpi_face_geometry_agree_across_ranks(world, method):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, method, 3, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    ok, nel_sum, nel_global = q.get(timeout=10)
    assert ok == 1
    assert nel_sum == nel_global


def test_partition_is_balanced_and_complete():
    g = HostMesh.box(6).connect()
    for method in ("metis", "block"):
        part = g.partition(4, method)
        cnt = np.bincount(part, minlength=4)
        assert cnt.sum() == g.nElem and cnt.min() > 0.8 * g.nElem / 4


def _worker_face_h(rank, world, port, result):
    """CommunicateMPIFaceMinimumDistance (HexMesh.f90:3059-3145) over gloo: each rank sends the h of its MPI faces (its own
    element only), takes the minimum with the neighbour's and must land on the single-domain value of that face."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = HostMesh.box(4, amp=0.1, bFaceOrder=2, shuffle=True, seed=3).connect()
        part = g.partition(world, "metis")
        g.geometry(3, GAUSS)
        h_global = g.array("faceH").copy()
        m = g.extract(part, rank).geometry(3, GAUSS)                  # geometry rebuilt from the local elements
        inh = g.extract(part, rank, inherit_geometry=True)            # geometry copied from the global mesh
        gf = m.array("globalFace").copy()
        h_loc = m.array("faceH").copy()
        ranks, counts, faces = m.array("haloRank").copy(), m.array("haloCount").copy(), m.array("haloFace").copy()
        ok = np.allclose(inh.array("faceH"), h_global[inh.array("globalFace")], rtol=0, atol=0)
        interior = m.array("faceType") != 3
        ok = ok and np.allclose(h_loc[interior], h_global[gf[interior]], rtol=1e-12)
        off = 0
        for nb, cnt in zip(ranks, counts):
            ids = faces[off:off + cnt]
            mine = torch.from_numpy(h_loc[ids].copy()); theirs = torch.empty_like(mine)
            for r in [dist.isend(mine, int(nb)), dist.irecv(theirs, int(nb))]:
                r.wait()
            h_min = np.minimum(mine.numpy(), theirs.numpy())
            ok = ok and bool((h_loc[ids] >= h_global[gf[ids]] * (1 - 1e-12)).all())          # one element only: never smaller
            ok = ok and np.allclose(h_min, h_global[gf[ids]], rtol=1e-11)
            off += cnt
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            result.put(int(flag.item()))
    finally:
        dist.destroy_process_group()


def test_mpi_face_minimum_distance_matches_the_single_domain_value():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 17
    procs = [ctx.Process(target=_worker_face_h, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert q.get(timeout=10) == 1


def test_wall_distances_of_a_partition_use_the_wall_nodes_of_every_rank():
    """HexMesh_ComputeWallDistances (HexMesh.f90:5594-5780): a partition measures its nodes against the no-slip wall nodes gathered
    from ALL ranks; measuring against its own wall faces only is refused (the LES wall model would depend on the partition)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from horses3d_b200.dgsem import DGSem
    from horses3d_b200.hostmesh import HostError
    from horses3d_b200.physics import make_physics
    from parity import channel_bcs

    class NullApi:      # accepts the set-up calls of DGSem (the check under test is host-side)
        def __getattr__(self, name):
            return lambda *a, **k: None

    phys = make_physics(flow="NS", mach=0.3, reynolds=200.0, les="smagorinsky", les_wall_model="linear")
    bcs, params = channel_bcs(phys)
    g = HostMesh.box(4, amp=0.1, bFaceOrder=2, shuffle=True, seed=11).connect(bcs, params).geometry(3, GAUSS).wall_distances()
    world = 3
    part = g.partition(world, "metis")
    parts = [g.extract(part, r).geometry(3, GAUSS) for r in range(world)]
    pts = [m.wall_points() for m in parts]
    assert sum(len(p) for p in pts) == len(g.wall_points()) > 0
    assert min(len(p) for p in pts) < max(len(p) for p in pts) or world == 1      # the wall is not spread evenly
    for r, m in enumerate(parts):
        with pytest.raises(HostError):
            m.wall_distances()
        with pytest.raises(ValueError, match="wall distances"):
            DGSem(NullApi(), m, phys)
        m.wall_distances(gather=lambda p: np.concatenate(pts))
        ge = m.array("globalElem")
        assert np.array_equal(m.array("dWall").reshape(len(ge), -1), g.array("dWall").reshape(g.nElem, -1)[ge])
        local = np.sqrt(((m.array("x").reshape(-1, 1, 3) - pts[r][None]) ** 2).sum(-1).min(1)) if len(pts[r]) else np.full(m.array("dWall").shape, np.inf)
        assert (local >= m.array("dWall") - 1e-14).all() and (local > m.array("dWall") + 1e-6).any()   # the local-only answer differs
        mi = g.extract(part, r, inherit_geometry=True)
        assert mi.wall_global and np.array_equal(mi.array("dWall"), m.array("dWall"))
