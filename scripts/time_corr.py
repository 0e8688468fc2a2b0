"""Scratch: breakdown of the closest-point correspondence step at BASELINE config-3 size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynamicfusion_body_b200 import synth, engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
sc = synth.make_scene(res=256, k=4, n_nodes=1000, seed=0, background=True)
pd = synth.make_gn_problem(sc, n, seed=0)
wf = engine.DeviceWarpField(4); wf.set_nodes(sc.node_pos, sc.node_dq, np.float32(sc.node_w))
live = torch.from_numpy(pd.corr.astype(np.float32)).cuda()
verts = torch.from_numpy(pd.vertices).cuda(); norms = torch.from_numpy(pd.normals).cuda()
loc = torch.from_numpy(pd.vert_knn.astype(np.int32)).cuda()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): r = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, r


t_warp, (wv, wn) = timed(lambda: engine.warp_points(wf, sc.lw, verts, norms, idx=loc, k=4))
for cell in (None, 0.5, 1.0, 2.0, 4.0):
    t_grid, grid = timed(lambda: engine.PointGrid(live, cell=cell))
    t_knn, nn = timed(lambda: grid.knn(wv, 4))
    t_sel, _ = timed(lambda: engine.corr_select(wv, wn, live, nn))
    print("n=%d cell=%s (%.3f, dims %s): warp %.3f  grid build %.3f  knn %.3f  select %.3f ms" % (n, cell, grid.cell, grid.dims, t_warp, t_grid, t_knn, t_sel))
t_us, (sv, si) = timed(lambda: engine.uniform_sample(verts, 3.0 * 255 / 64), reps=2)
print("uniform_sample of %d points -> %d nodes: %.3f ms" % (n, len(si), t_us))
