"""Per-kernel timings of the a3 step via the profiling modes (scratch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynamicfusion_body_b200 import synth, engine, _capi
R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 4
V = int(sys.argv[4]) if len(sys.argv) > 4 else 1
sc = synth.make_scene(res=R, k=k, n_nodes=N, seed=0, background=True, n_views=V)
wf = engine.DeviceWarpField(k); wf.set_nodes(sc.node_pos, sc.node_dq, np.float32(sc.node_w))
x0, x1 = (int(v) for v in os.environ.get("SLAB", "0,%d" % R).split(","))     # SLAB=112,168: one rank's x-slab of the strong-scaling run
vol = engine.DeviceVolume((R, R, R), x0, x1, fill=sc.tdist)
depths = torch.from_numpy(sc.depths).cuda()
views = engine.make_views(depths, sc.K, sc.Kinv, sc.extrinsics)
wf.brick_nodes(vol.res, x0, x1); torch.cuda.synchronize()
modes = [("classify", 4), ("update", 7), ("exact", 3), ("full", 0)]
acc = {n: [] for n, _ in modes}
for i in range(8):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(modes) + 1)]
    ev[0].record()
    for j, (n, m) in enumerate(modes):
        engine.update_projective(vol, wf, sc.lw, depths, sc.K, sc.Kinv, sc.extrinsics, sc.tdist, views=views, mode=m)
        ev[j + 1].record()
    torch.cuda.synchronize()
    for j, (n, _) in enumerate(modes): acc[n].append(ev[j].elapsed_time(ev[j + 1]))
print("R=%d N=%d k=%d V=%d slab=[%d,%d)" % (R, sc.n_nodes, k, V, x0, x1), {n: round(float(np.mean(v[2:])), 4) for n, v in acc.items()}, vol.workspace.stats())
