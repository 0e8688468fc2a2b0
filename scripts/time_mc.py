"""Times the device surface extractor (engine.marching_cubes = dfb_mc_level + dfb_mc_count + dfb_mc_emit + read-back) on a
truncated sphere TSDF.  python scripts/time_mc.py [R ...]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from dynamicfusion_body_b200 import engine  # noqa: E402

for R in [int(a) for a in sys.argv[1:]] or [256, 512]:
    ax = torch.arange(R, device="cuda", dtype=torch.float32)
    c = 0.5 * R + 0.3
    vol = torch.clamp(torch.sqrt((ax[:, None, None] - c) ** 2 + (ax[None, :, None] - c + 1) ** 2 + (ax[None, None, :] - c - 1) ** 2) - 0.35 * R, -3, 3)
    for step in (1, 2):
        engine.marching_cubes(vol, step)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            v, f, n, _ = engine.marching_cubes(vol, step)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / reps * 1e3
        print("R=%d step=%d: %d vertices, %d faces, %.2f ms per extraction incl. read-back (%.0f GB/s of volume bytes)"
              % (R, step, len(v), len(f), ms, vol.numel() * 4 / ms / 1e6), flush=True)
