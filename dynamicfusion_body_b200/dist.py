"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on the GPUs, gloo in the CPU tests).

The TSDF update shards with no halo and no data-path collective: GPU g owns the contiguous x-slab
[x0_g, x1_g) of the C-ordered volume (index = x*ry*rz + y*rz + z), every rank keeps a replica of the (tiny) node
table, and per frame rank 0 broadcasts the sensor data + node transforms (~1.3 MB).  The Gauss-Newton solve has one
real exchange: each rank assembles J^T W J / J^T W f over its range of data residuals into the SAME block pattern and
the blocks are summed with one all-reduce before every rank runs the same node-space solve (replicated; its atomically
accumulated inner products agree to rounding only, so the iterate is broadcast from rank 0 after every solve --
gn.Problem.gauss_newton(broadcast=...) -- and all ranks warp with bit-identical transforms).
"""
import numpy as np
import torch
import torch.distributed as dist


def slab_partition(rx, world):
    """Contiguous x-slabs, sizes differing by at most one: [(x0, x1)] * world."""
    base, rem = divmod(int(rx), int(world))
    out, x = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((x, x + n))
        x += n
    return out


def balanced_slab_partition(unit_cost, world, unit_planes=16, rx=None):
    """Contiguous x-slabs whose boundaries are multiples of `unit_planes`, chosen so that the LARGEST slab cost is as small as
    possible (the step time of a sharded volume is the maximum over ranks).  unit_cost[i] = estimated cost of planes
    [i * unit_planes, (i + 1) * unit_planes) -- e.g. from the brick classes of a frame (engine.slab_cost_profile).  Every slab gets
    at least one unit.  Returns [(x0, x1)] * world like slab_partition; rx clips the last boundary."""
    c = np.asarray(unit_cost, dtype=np.float64)
    n, world = len(c), int(world)
    if n < world:
        raise ValueError("fewer units (%d) than ranks (%d)" % (n, world))

    def cuts_for(cap):
        """greedy fill under capacity `cap`; None if more than `world` slabs would be needed"""
        cuts, acc, used = [], 0.0, 1
        for i in range(n):
            remaining_units, remaining_slabs = n - i, world - used
            if acc > 0 and (acc + c[i] > cap or remaining_units == remaining_slabs):
                cuts.append(i); acc = 0.0; used += 1
                if used > world:
                    return None
            acc += c[i]
        return cuts

    lo, hi = float(c.max()), float(c.sum())
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if cuts_for(mid) is None:
            lo = mid
        else:
            hi = mid
    cuts = cuts_for(hi) or []
    while len(cuts) < world - 1:                     # fewer slabs than ranks: split the widest slab
        b = [0] + cuts + [n]
        w = int(np.argmax(np.diff(b)))
        cuts = sorted(cuts + [b[w] + (b[w + 1] - b[w]) // 2])
    b = [0] + cuts + [n]
    total = n * unit_planes if rx is None else int(rx)
    return [(b[r] * unit_planes, min(b[r + 1] * unit_planes, total) if r < world - 1 else total) for r in range(world)]


def residual_partition(n_vert, world):
    """Contiguous ranges of data residuals per rank."""
    return slab_partition(n_vert, world)


def is_dist():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def broadcast_frame(depths, node_dq, lw=None, src=0, group=None):
    """Per-frame broadcast root -> all of the depth view(s) (V,rows,cols) f32, the node transforms (N,8) f32 and the
    global rigid dq (8,) f64 tensor.  In place; returns the tensors."""
    if is_dist():
        dist.broadcast(depths, src, group=group)
        dist.broadcast(node_dq, src, group=group)
        if lw is not None:
            dist.broadcast(lw, src, group=group)
    return depths, node_dq, lw


class FramePacket:
    """One frame in ONE contiguous float32 device buffer: [depth views | global rigid dq | node transforms].  `depths`
    (V, rows, cols) float32, `lw` (8,) float64 and `node_dq` (N, 8) float32 are views.  broadcast() sends everything in
    a single collective (two or three small NCCL launches cost more than the 1.3 MB they move); broadcast_depths() /
    broadcast_transforms() send the two halves separately, so that a streaming caller can ship the sensor data of frame
    t+1 on a side stream (and its own process group) while frame t is being fused -- depth does not depend on the fusion
    result, the transforms do."""

    def __init__(self, n_views, rows, cols, n_nodes, device):
        nd = n_views * rows * cols
        nd_pad = (nd + 3) // 4 * 4                      # keeps the float64 view 16-byte aligned
        self.flat = torch.zeros(nd_pad + 16 + 8 * n_nodes, dtype=torch.float32, device=device)
        self.depths = self.flat[:nd].view(n_views, rows, cols)
        self.transforms = self.flat[nd_pad:]
        self.lw = self.flat[nd_pad:nd_pad + 16].view(torch.float64)
        self.node_dq = self.flat[nd_pad + 16:].view(n_nodes, 8)

    def broadcast(self, src=0, group=None):
        if is_dist():
            dist.broadcast(self.flat, src, group=group)
        return self

    def broadcast_depths(self, src=0, group=None):
        if is_dist():
            dist.broadcast(self.depths, src, group=group)
        return self

    def broadcast_transforms(self, src=0, group=None):
        if is_dist():
            dist.broadcast(self.transforms, src, group=group)
        return self


def flat_normal_equations(H, g, cost):
    """The flat [H | g | cost] tensor when the three are adjacent views of one buffer (how gn.Problem.normal_equations
    allocates them), else None."""
    try:
        same = H.untyped_storage().data_ptr() == g.untyped_storage().data_ptr() == cost.untyped_storage().data_ptr()
    except Exception:
        return None
    if (same and H.is_contiguous() and g.is_contiguous() and cost.is_contiguous() and g.storage_offset() == H.storage_offset() + H.numel()
            and cost.storage_offset() == g.storage_offset() + g.numel()):
        return torch.as_strided(H, (H.numel() + g.numel() + cost.numel(),), (1,), H.storage_offset())
    return None


def allreduce_normal_equations(H, g, cost, group=None, comm=None):
    """Sum the block-sparse normal equations over ranks: ONE collective on the flat [H | g | cost] buffer, in place.
    comm: an engine.Comm (raw NCCL through the C ABI, dfb_comm_allreduce_f64); default: torch.distributed."""
    if comm is None and not is_dist():
        return H, g, cost
    flat = flat_normal_equations(H, g, cost)
    packed = flat is None
    if packed:                                            # separate tensors (callers outside gn.Problem): pack once
        flat = torch.cat([H.reshape(-1), g.reshape(-1), cost.reshape(-1)])
    if comm is not None:
        comm.allreduce_f64(flat)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if packed:
        nH, ng = H.numel(), g.numel()
        H.copy_(flat[:nH].view_as(H))
        g.copy_(flat[nH:nH + ng].view_as(g))
        cost.copy_(flat[nH + ng:].view_as(cost))
    return H, g, cost


def gather_slabs(slab, rx, dst=0, group=None):
    """Concatenate the x-slabs of every rank on `dst` (parity tests; SURVEY 8e: must equal the single-GPU volume
    bit for bit).  Returns the full tensor on dst, None elsewhere."""
    if not is_dist():
        return slab
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    parts = slab_partition(rx, world)
    shapes = [(x1 - x0,) + tuple(slab.shape[1:]) for x0, x1 in parts]
    if rank == dst:
        bufs = [torch.empty(s, dtype=slab.dtype, device=slab.device) for s in shapes]
        bufs[dst].copy_(slab)
        for r in range(world):
            if r != dst:
                dist.recv(bufs[r], src=r, group=group)
        return torch.cat(bufs, 0)
    dist.send(slab.contiguous(), dst=dst, group=group)
    return None


def max_over_ranks(value_ms, device):
    """Device-timed durations are reported as the max over ranks."""
    if not is_dist():
        return float(value_ms)
    t = torch.tensor([float(value_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- surface extraction over x-slabs (SURVEY 8e x 8f rank 3) ---------------------------------------------------------
def slab_sample_planes(x0, x1, rx, step):
    """Sample x-planes (voxel x, multiples of `step`) inside the slab [x0, x1): (first, last, last plane of the whole grid)."""
    s = int(step)
    last = (int(rx) - 1) // s * s
    a = -(-int(x0) // s) * s
    b = min((int(x1) - 1) // s * s, last)
    if b - a < s:
        raise ValueError("surface extraction over slabs needs at least two sample planes per slab (slab [%d, %d), step %d)" % (x0, x1, s))
    return a, b, last


def global_level(slab, group=None):
    """0.5 * (min + max) of the whole volume, float32 of the float64 mean like dfb_mc_level (the level skimage takes when the
    reference passes none, core/fusion.py:559)."""
    mm = torch.stack([slab.min(), -slab.max()]).to(torch.float32)
    if is_dist():
        dist.all_reduce(mm, op=dist.ReduceOp.MIN, group=group)
    mn, mx = float(mm[0].item()), -float(mm[1].item())
    return float(np.float32(0.5 * (np.float64(np.float32(mn)) + np.float64(np.float32(mx)))))


def slab_halo_volume(slab, x0, x1, rx, step, prev_plane=None, next_planes=None):
    """The slab extended by the sample planes the extractor needs from its neighbours: one before (gradient of the first owned plane)
    and two after (far face of the last owned cells; gradient on it).  Planes that are not sample planes stay zero (never read).
    Returns (sub volume, voxel x of sub[0], first owned plane, last owned plane)."""
    s = int(step)
    a, b, last = slab_sample_planes(x0, x1, rx, s)
    lo = a - s if prev_plane is not None else min(a, int(x0))
    hi = b + 2 * s if next_planes is not None else int(x1) - 1
    if prev_plane is None and a != 0:
        raise ValueError("slab does not start the grid: the sample plane before it is required")
    if next_planes is None and b != last:
        raise ValueError("slab does not end the grid: the two sample planes after it are required")
    sub = torch.zeros((hi - lo + 1,) + tuple(slab.shape[1:]), dtype=torch.float32, device=slab.device)
    sub[int(x0) - lo:int(x1) - lo] = slab
    if prev_plane is not None:
        sub[0] = prev_plane
    if next_planes is not None:
        sub[b + s - lo] = next_planes[0]
        sub[b + 2 * s - lo] = next_planes[1]
    return sub, lo, a, b


def cut_owned_mesh(mesh, lo, a, b, step):
    """From the mesh of a halo volume (with plane offsets) keep what the slab owns: the vertices of its sample planes [a, b] and the
    triangles of the cells that start on them.  Face ids are made relative to the first owned vertex; ids >= the number of owned
    vertices point into the next slab's first plane, whose vertices come in the same order there."""
    v, f, n, val, pv, pt = mesh
    s = int(step)
    j0, j1 = (a - lo) // s, (b - lo) // s
    v_lo, v_hi, t_lo, t_hi = int(pv[j0]), int(pv[j1 + 1]), int(pt[j0]), int(pt[j1 + 1])
    return v[v_lo:v_hi], f[t_lo:t_hi].astype(np.int64) - v_lo, n[v_lo:v_hi], val[v_lo:v_hi]


def extract_surface_slab(slab, x0, x1, rx, step=1, level=None, extractor=None, group=None):
    """This rank's part of the whole volume's surface mesh: (verts, faces with GLOBAL vertex ids int32, normals, values).  The parts
    of all ranks concatenated in rank order are the single-volume mesh of engine.marching_cubes bit for bit (vertices are ordered by
    owning sample and triangles by cell, x slowest, so a slab's share is a contiguous range of both).  Exchanges three (ry, rz)
    planes per slab boundary and the vertex counts; extractor(vol, step, level, x_origin=, plane_offsets=True) defaults to the
    device extractor."""
    if extractor is None:
        from . import engine
        extractor = engine.marching_cubes
    s = int(step)
    slab = slab.to(torch.float32)
    if level is None:
        level = global_level(slab, group)
    prev_plane = next_planes = None
    rank, world = 0, 1
    if is_dist():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        a, b, _ = slab_sample_planes(x0, x1, rx, s)
        plane = tuple(slab.shape[1:])
        ops, keep = [], []
        if rank > 0:
            prev_plane = torch.empty(plane, dtype=torch.float32, device=slab.device)
            first_two = torch.stack([slab[a - int(x0)], slab[a + s - int(x0)]]).contiguous()
            keep.append(first_two)
            ops += [dist.P2POp(dist.isend, first_two, rank - 1, group), dist.P2POp(dist.irecv, prev_plane, rank - 1, group)]
        if rank < world - 1:
            next_planes = torch.empty((2,) + plane, dtype=torch.float32, device=slab.device)
            mine = slab[b - int(x0)].contiguous()
            keep.append(mine)
            ops += [dist.P2POp(dist.isend, mine, rank + 1, group), dist.P2POp(dist.irecv, next_planes, rank + 1, group)]
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()
    sub, lo, a, b = slab_halo_volume(slab, x0, x1, rx, s, prev_plane, next_planes)
    v, f, n, val = cut_owned_mesh(extractor(sub, s, level, x_origin=lo // s, plane_offsets=True), lo, a, b, s)
    offset = 0
    if is_dist():
        counts = torch.zeros(world, dtype=torch.int64, device=slab.device)
        counts[rank] = len(v)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
        offset = int(counts[:rank].sum().item())
    return v, (f + offset).astype(np.int32), n, val


def allgather_mesh(part, group=None):
    """Concatenate the ranks' mesh parts (rank order) on every rank: the whole-volume mesh the reference keeps in
    `_vertices / _faces / _normals` (core/fusion.py:565-567)."""
    if not is_dist():
        return part
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, tuple(np.ascontiguousarray(x) for x in part), group=group)
    return tuple(np.concatenate([p[i] for p in parts]) for i in range(4))
