"""Short single-config run of the a3 update for ncu (one scene, a few launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynamicfusion_body_b200 import synth, engine

R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 4
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
sc = synth.make_scene(res=R, k=k, n_nodes=N, seed=0, background=True)
wf = engine.DeviceWarpField(k); wf.set_nodes(sc.node_pos, sc.node_dq, np.float32(sc.node_w))
x0, x1 = (int(v) for v in os.environ.get("SLAB", "0,%d" % R).split(","))     # SLAB=192,256: one rank's x-slab of the strong-scaling run
vol = engine.DeviceVolume((R, R, R), x0, x1, fill=sc.tdist)
depths = torch.from_numpy(sc.depths).cuda()
views = engine.make_views(depths, sc.K, sc.Kinv, sc.extrinsics)
wf.knn_table(vol.res, x0, x1)
torch.cuda.synchronize()
for i in range(reps):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    engine.update_projective(vol, wf, sc.lw, depths, sc.K, sc.Kinv, sc.extrinsics, sc.tdist, views=views)
    b.record(); torch.cuda.synchronize()
    print("update %d: %.3f ms" % (i, a.elapsed_time(b)), vol.workspace.stats())
