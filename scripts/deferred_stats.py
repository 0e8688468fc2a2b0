"""What the exact pass is spent on (scratch): after the FIRST frame on a fresh volume (v = tdist, w = 0) an updated voxel
holds v' = min(tdist, tl), so `v' < tdist` marks the voxels that are genuinely inside the truncation band."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynamicfusion_body_b200 import synth, engine
R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
sc = synth.make_scene(res=R, k=4, n_nodes=N, seed=0, background=True, n_views=1)
wf = engine.DeviceWarpField(4); wf.set_nodes(sc.node_pos, sc.node_dq, np.float32(sc.node_w))
vol = engine.DeviceVolume((R, R, R), fill=sc.tdist)
depths = torch.from_numpy(sc.depths).cuda()
mask, frus = engine.update_projective(vol, wf, sc.lw, depths, sc.K, sc.Kinv, sc.extrinsics, sc.tdist, want_masks=True)
torch.cuda.synchronize()
st = vol.workspace.stats()
n = st["deferred"]
lst = vol.workspace.list[:n].long()
v = vol.tsdf.reshape(-1)[lst]
m = mask.reshape(-1)[lst] & 1
f = frus.reshape(-1)[lst] & 1
inband = (m == 1) & (v < sc.tdist)
print("voxels", R ** 3, "deferred", n, "(%.3f %%)" % (100.0 * n / R ** 3))
print("  in band (needs an exact value)      ", int(inband.sum()), "%.1f %%" % (100.0 * inband.float().mean()))
print("  updated with the clamp value         ", int(((m == 1) & ~inband).sum()))
print("  not updated, inside the frustum      ", int(((m == 0) & (f == 1)).sum()))
print("  not updated, outside the frustum     ", int(((m == 0) & (f == 0)).sum()))
allv = vol.tsdf.reshape(-1)
print("all voxels in band", int(((mask.reshape(-1) & 1 == 1) & (allv < sc.tdist)).sum()))
