"""Multi-GPU check of the sharded surface extraction (SURVEY 8e x 8f rank 3), one process per GPU over NCCL:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_mc_slabs_nccl.py [R]
Every rank holds an x-slab of a truncated-sphere TSDF, extracts its share (dist.extract_surface_slab: halo planes by NCCL p2p) and
gathers the whole mesh; rank 0 compares it with the single-GPU extraction of the whole volume (must be identical) and prints timings."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamicfusion_body_b200 import dist as ddist  # noqa: E402
from dynamicfusion_body_b200 import engine  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    ax = torch.arange(R, device="cuda", dtype=torch.float32)
    c = 0.5 * R + 0.3
    vol = torch.clamp(torch.sqrt((ax[:, None, None] - c) ** 2 + (ax[None, :, None] - c + 1) ** 2 + (ax[None, None, :] - c - 1) ** 2) - 0.35 * R, -3, 3)
    x0, x1 = ddist.slab_partition(R, world)[rank]
    slab = vol[x0:x1].contiguous()
    ok = True
    for step, level in ((1, None), (2, 0.0)):
        for rep in range(3):                                               # first repetition warms NCCL up
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            part = ddist.extract_surface_slab(slab, x0, x1, R, step, level)
            t1 = time.perf_counter()
            mesh = ddist.allgather_mesh(part)
            t2 = time.perf_counter()
        if rank == 0:
            full = engine.marching_cubes(vol, step, level)
            same = all(a.shape == b.shape and np.array_equal(a, b) for a, b in zip(mesh, full))
            ok = ok and same
            print("R=%d step=%d world=%d: %d vertices, %d faces, identical=%s; slab extraction %.2f ms, gather %.2f ms"
                  % (R, step, world, len(full[0]), len(full[1]), same, (t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
