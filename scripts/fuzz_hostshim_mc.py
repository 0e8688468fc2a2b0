"""CPU fuzzer of the surface extractor: the host build of csrc/dfb_mc.h (tests/hostshim) against oracle/mc.py on random volumes
(noise, blobs, quantised values that put samples exactly on the level, ragged shapes, steps 1-3, random x origins), and the slab
composition of dist.py against the whole-volume mesh.   python scripts/fuzz_hostshim_mc.py [cases] [seed] [--device]
--device: the CUDA path (engine.marching_cubes through the C ABI) takes the place of the host build (needs a GPU)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import hostshim_api as hs  # noqa: E402
from dynamicfusion_body_b200 import dist as ddist  # noqa: E402
from oracle import mc as omc  # noqa: E402


def random_volume(rng):
    shape = tuple(int(n) for n in rng.integers(2, 28, size=3))
    kind = rng.integers(0, 4)
    if kind == 0:
        vol = rng.normal(size=shape)
    elif kind == 1:                                                         # a few blobs
        g = np.stack(np.meshgrid(*[np.arange(n) for n in shape], indexing="ij"), -1).astype(np.float64)
        vol = np.full(shape, 1e9)
        for _ in range(int(rng.integers(1, 4))):
            c = rng.uniform(0, 1, 3) * (np.array(shape) - 1)
            vol = np.minimum(vol, np.linalg.norm(g - c, axis=-1) - rng.uniform(1.0, 0.5 * max(shape)))
    elif kind == 2:                                                         # quantised: many samples exactly at the level
        vol = np.round(rng.normal(size=shape) * 2) / 2
    else:                                                                   # truncated like a TSDF
        vol = np.clip(rng.normal(size=shape) * 4, -3, 3)
    return vol.astype(np.float32)


DEVICE = "--device" in sys.argv
if DEVICE:
    sys.argv.remove("--device")
    from dynamicfusion_body_b200 import engine  # noqa: E402


def extract(vol, step, level, **kw):
    if DEVICE:
        return engine.marching_cubes(torch.from_numpy(np.ascontiguousarray(vol)).cuda(), step, level, **kw)
    return hs.marching_cubes(vol, step, level, **kw)


def same(a, b):
    return all(x.shape == y.shape for x, y in zip(a, b)) and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) \
        and np.array_equal(a[3], b[3]) and (a[2].size == 0 or np.abs(a[2] - b[2]).max() <= 1e-6)


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    bad = nslab = nverts = 0
    for c in range(cases):
        vol = random_volume(rng)
        step = int(rng.integers(1, 4))
        level = None if rng.integers(0, 3) == 0 else float(np.float32(rng.choice([0.0, 0.5, 0.1, -0.25])))
        xo = int(rng.choice([0, 0, 3, 1000, 100000]))
        got = extract(vol, step, level, x_origin=xo)
        want = omc.marching_cubes(vol, step, level, x_origin=xo)
        nverts += len(want[0])
        if not same(got, want):
            bad += 1
            print("MISMATCH extractor vs oracle: case %d shape %s step %d level %s x_origin %d" % (c, vol.shape, step, level, xo))
        world = int(rng.integers(2, 5))
        try:
            parts_x = ddist.slab_partition(vol.shape[0], world)
            [ddist.slab_sample_planes(x0, x1, vol.shape[0], step) for x0, x1 in parts_x]
        except ValueError:
            continue                                                        # slabs thinner than two sample planes
        lv = float(omc.default_level(vol)) if level is None else level
        full = extract(vol, step, lv)
        parts, offset = [], 0
        for r, (x0, x1) in enumerate(parts_x):
            a, b, _ = ddist.slab_sample_planes(x0, x1, vol.shape[0], step)
            prev = torch.from_numpy(vol[a - step]) if r > 0 else None
            nxt = torch.from_numpy(np.stack([vol[b + step], vol[b + 2 * step]])) if r < world - 1 else None
            sub, lo, a, b = ddist.slab_halo_volume(torch.from_numpy(vol[x0:x1]), x0, x1, vol.shape[0], step, prev, nxt)
            v, f, n, val = ddist.cut_owned_mesh(extract(sub.numpy(), step, lv, x_origin=lo // step, plane_offsets=True), lo, a, b, step)
            parts.append((v, (f + offset).astype(np.int32), n, val))
            offset += len(v)
        cat = tuple(np.concatenate([p[i] for p in parts]) for i in range(4))
        nslab += 1
        if not all(x.shape == y.shape and np.array_equal(x, y) for x, y in zip(cat, full)):
            bad += 1
            print("MISMATCH slabs vs whole volume: case %d shape %s step %d world %d" % (c, vol.shape, step, world))
    print("%d cases (%d vertices), %d of them also through the slab composition: %d mismatches" % (cases, nverts, nslab, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
