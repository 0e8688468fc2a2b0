"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: slab partition + gather == full volume bit for bit,
frame broadcast, sharded normal equations summed by all-reduce == single-rank normal equations."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, _free_port_cached(), fn, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    return dict(ret)


_PORT = None


def _free_port_cached():
    global _PORT
    if _PORT is None:
        _PORT = _free_port()
    return _PORT


def test_slab_partition_covers_grid():
    from dynamicfusion_body_b200.dist import slab_partition
    for rx, w in ((512, 8), (513, 8), (7, 2), (10, 3), (1024, 8)):
        parts = slab_partition(rx, w)
        assert parts[0][0] == 0 and parts[-1][1] == rx and len(parts) == w
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1


def _slab_job(rank, world):
    """Each rank fuses its x-slab with the host build of the kernel logic; rank 0 gathers and compares with the
    single-volume result."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import hostshim_api as hs
    import scenes
    from dynamicfusion_body_b200 import dist as ddist
    from dynamicfusion_body_b200 import synth
    sc = synth.make_scene(res=24, k=4, n_nodes=80, seed=2, rows=48, cols=64)
    R = sc.res
    vox, idx, tie = scenes.oracle_knn((R, R, R), sc.node_pos, sc.k)
    t0, w0 = scenes.initial_state(R ** 3, tdist=sc.tdist)
    # frame broadcast: only rank 0 holds the real sensor data / transforms
    depths = torch.from_numpy(sc.depths.copy()) if rank == 0 else torch.zeros(sc.depths.shape)
    dq = torch.from_numpy(sc.node_dq.copy()) if rank == 0 else torch.zeros(sc.node_dq.shape)
    lw = torch.from_numpy(sc.lw.copy()) if rank == 0 else torch.zeros(8, dtype=torch.float64)
    ddist.broadcast_frame(depths, dq, lw)
    assert np.array_equal(depths.numpy(), sc.depths) and np.array_equal(dq.numpy(), sc.node_dq)
    x0, x1 = ddist.slab_partition(R, world)[rank]
    plane = R * R
    wf = hs.HostWarpField(sc.node_pos, dq.numpy(), np.float32(sc.node_w), sc.k, knn=idx[x0 * plane:x1 * plane], lw=lw.numpy())
    tv = t0[x0 * plane:x1 * plane].copy(); tw = w0[x0 * plane:x1 * plane].copy()
    hs.update_projective(tv, tw, (R, R, R), wf, depths.numpy(), sc.K, sc.Kinv, sc.tdist, x0=x0, x1=x1)
    full_v = ddist.gather_slabs(torch.from_numpy(tv).reshape(x1 - x0, R, R), R)
    full_w = ddist.gather_slabs(torch.from_numpy(tw).reshape(x1 - x0, R, R), R)
    if rank != 0:
        return True
    wf_all = hs.HostWarpField(sc.node_pos, sc.node_dq, np.float32(sc.node_w), sc.k, knn=idx, lw=sc.lw)
    rv, rw_ = t0.copy(), w0.copy()
    hs.update_projective(rv, rw_, (R, R, R), wf_all, sc.depths, sc.K, sc.Kinv, sc.tdist)
    return bool(np.array_equal(full_v.numpy().ravel(), rv) and np.array_equal(full_w.numpy().ravel(), rw_))


def test_slab_sharded_update_equals_full_volume():
    out = _run(_slab_job)
    assert out[0] is True and out[1] is True


def _gn_job(rank, world):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import hostshim_api as hs
    from scipy.spatial import cKDTree
    from dynamicfusion_body_b200 import dist as ddist
    from dynamicfusion_body_b200 import synth
    sc = synth.make_scene(res=32, k=4, n_nodes=40, seed=1, rows=48, cols=64)
    rng = np.random.default_rng(0)
    sel = rng.choice(len(sc.vertices), 400, replace=False)
    verts, norms, knn = sc.vertices[sel], sc.normals[sel], sc.vert_knn[sel]
    corr = sc.warped_vertices[sel] + rng.normal(size=(400, 3)) * 0.2
    _, nvi = cKDTree(verts.astype(np.float64)).query(sc.node_pos.astype(np.float64))
    x = sc.node_dq.reshape(-1).astype(np.float64)
    v0, v1 = ddist.residual_partition(len(verts), world)[rank]
    # shard: data residuals [v0,v1); the regularisation rows belong to rank 0 (rw = 0 elsewhere)
    nbr_src = hs.HostGN(verts, norms, corr, knn, sc.node_pos, np.float32(sc.node_w), nvi, sc.lw, 0.5)
    G = hs.HostGN(verts[v0:v1], norms[v0:v1], corr[v0:v1], knn[v0:v1], sc.node_pos, np.float32(sc.node_w), np.zeros(sc.n_nodes, int), sc.lw,
                  0.5 if rank == 0 else 0.0, huber=True, f_scale=0.3)
    G.node_nbr[...] = nbr_src.node_nbr
    H, g, cost = G.normal_eq_dense(x)
    Ht, gt, ct = torch.from_numpy(H), torch.from_numpy(g), torch.from_numpy(cost)
    ddist.allreduce_normal_equations(Ht, gt, ct)
    full = hs.HostGN(verts, norms, corr, knn, sc.node_pos, np.float32(sc.node_w), nvi, sc.lw, 0.5, huber=True, f_scale=0.3)
    Hf, gf, cf = full.normal_eq_dense(x)
    ok = (np.abs(Ht.numpy() - Hf).max() <= 1e-12 * np.abs(Hf).max() and np.abs(gt.numpy() - gf).max() <= 1e-12 * np.abs(gf).max()
          and abs(ct.numpy()[0] - cf[0]) <= 1e-12 * cf[0])
    return bool(ok)


def test_sharded_normal_equations_allreduce():
    out = _run(_gn_job)
    assert out[0] is True and out[1] is True


def _packet_job(rank, world):
    """FramePacket: one collective carries lw (float64), the depth views and the node transforms."""
    from dynamicfusion_body_b200 import dist as ddist
    rng = np.random.default_rng(7)
    lw = rng.normal(size=8)
    depths = rng.normal(size=(3, 12, 16)).astype(np.float32)
    dq = rng.normal(size=(37, 8)).astype(np.float32)
    pk = ddist.FramePacket(3, 12, 16, 37, torch.device("cpu"))
    assert pk.depths.is_contiguous() and pk.node_dq.is_contiguous() and pk.lw.dtype == torch.float64
    if rank == 0:
        pk.lw.copy_(torch.from_numpy(lw)); pk.depths.copy_(torch.from_numpy(depths)); pk.node_dq.copy_(torch.from_numpy(dq))
    pk.broadcast()
    ok = bool(np.array_equal(pk.lw.numpy(), lw) and np.array_equal(pk.depths.numpy(), depths) and np.array_equal(pk.node_dq.numpy(), dq))
    # the two halves separately (depth ahead of time on its own group, transforms inside the step)
    pk2 = ddist.FramePacket(3, 12, 16, 37, torch.device("cpu"))
    g2 = dist.new_group()
    if rank == 0:
        pk2.lw.copy_(torch.from_numpy(lw)); pk2.depths.copy_(torch.from_numpy(depths)); pk2.node_dq.copy_(torch.from_numpy(dq))
    pk2.broadcast_depths(group=g2)
    ok = ok and bool(np.array_equal(pk2.depths.numpy(), depths)) and (rank == 0 or not pk2.node_dq.numpy().any())
    pk2.broadcast_transforms()
    return ok and bool(np.array_equal(pk2.lw.numpy(), lw) and np.array_equal(pk2.node_dq.numpy(), dq))


def test_frame_packet_broadcast():
    out = _run(_packet_job)
    assert out[0] and out[1]


def _surface_job(rank, world):
    """Each rank extracts the surface of its x-slab (halo planes exchanged with its neighbours, host build of the extractor); the
    gathered mesh must be the single-volume mesh bit for bit -- with an explicit level and with the all-reduced default level."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import hostshim_api as hs
    from dynamicfusion_body_b200 import dist as ddist
    rng = np.random.default_rng(4)
    x, y, z = np.meshgrid(np.arange(21), np.arange(18), np.arange(40), indexing="ij")
    vol = (np.sqrt((x - 10.3) ** 2 + (y - 8.6) ** 2 + (z - 19.2) ** 2) - 7.7 + 0.3 * rng.normal(size=x.shape)).astype(np.float32)
    ext = lambda sub, s, lv, **kw: hs.marching_cubes(sub.numpy(), s, lv, **kw)
    ok = True
    for step, level in ((1, 0.0), (2, None), (1, None)):
        x0, x1 = ddist.slab_partition(vol.shape[0], world)[rank]
        part = ddist.extract_surface_slab(torch.from_numpy(vol[x0:x1].copy()), x0, x1, vol.shape[0], step, level, extractor=ext)
        got = ddist.allgather_mesh(part)
        want = hs.marching_cubes(vol, step, level)
        ok = ok and len(want[1]) > 100 and all(g.shape == w.shape and np.array_equal(g, w) for g, w in zip(got, want))
    return bool(ok)


def test_slab_sharded_surface_extraction_equals_full_volume():
    for world in (2, 3):
        out = _run(_surface_job, world=world)
        assert all(out[r] is True for r in range(world))
