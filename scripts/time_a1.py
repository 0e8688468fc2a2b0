"""Scratch: timing of the a1 path (Fusion.updateTSDF, volume-sampling warped update)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynamicfusion_body_b200 import synth, engine
R = int(sys.argv[1]) if len(sys.argv) > 1 else 128
sc = synth.make_scene(res=R, k=4, n_nodes=1000, seed=0)
nw = np.full(sc.n_nodes, sc.node_w)
lw = np.array([1, 0, 0, 0, 0, 0.1, 0, 0], np.float32)
wv = synth.blend_warp(sc.vertices, sc.node_pos, sc.node_dq, nw, sc.vert_knn, lw=lw.astype(np.float64))
live = synth.mesh_sdf_volume((R, R, R), wv, sc.warped_normals)
wf = engine.DeviceWarpField(4); wf.set_nodes(sc.node_pos, sc.node_dq, np.float32(sc.node_w))
for name, lv, tdist in (("untruncated SDF, tdist=max (reference usage test.py:110)", live, float(live.max())),
                        ("truncated TSDF +-%g" % sc.tdist, np.clip(live, -1.5 * sc.tdist, 1.5 * sc.tdist), sc.tdist)):
    vol = engine.DeviceVolume((R, R, R), fill=tdist)
    cur = torch.from_numpy(np.ascontiguousarray(lv, dtype=np.float32)).cuda()
    wf.knn_table(vol.res, 0, R); torch.cuda.synchronize()
    ts = []
    for i in range(5):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); engine.update_volume(vol, wf, lw, cur, tdist); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print("a1 R=%d %s: %.3f ms -> %.2f Gvox/s  %s" % (R, name, min(ts), R ** 3 / min(ts) / 1e6, vol.workspace.stats()))
