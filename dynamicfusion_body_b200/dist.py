"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on the GPUs, gloo in the CPU tests).

The TSDF update shards with no halo and no data-path collective: GPU g owns the contiguous x-slab
[x0_g, x1_g) of the C-ordered volume (index = x*ry*rz + y*rz + z), every rank keeps a replica of the (tiny) node
table, and per frame rank 0 broadcasts the sensor data + node transforms (~1.3 MB).  The Gauss-Newton solve has one
real exchange: each rank assembles J^T W J / J^T W f over its range of data residuals into the SAME block pattern and
the blocks are summed with one all-reduce before every rank runs the identical node-space solve.
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


def allreduce_normal_equations(H, g, cost, group=None):
    """Sum the block-sparse normal equations over ranks (one flat buffer -> one collective)."""
    if not is_dist():
        return H, g, cost
    flat = torch.cat([H.reshape(-1), g.reshape(-1), cost.reshape(-1)])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
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
