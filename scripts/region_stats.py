"""Region / brick classification statistics of one a3 step (scratch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynamicfusion_body_b200 import synth, engine
R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 4
V = int(sys.argv[4]) if len(sys.argv) > 4 else 1
sc = synth.make_scene(res=R, k=k, n_nodes=N, seed=0, background=True, n_views=V)
wf = engine.DeviceWarpField(k); wf.set_nodes(sc.node_pos, sc.node_dq, np.float32(sc.node_w))
vol = engine.DeviceVolume((R, R, R), fill=sc.tdist)
depths = torch.from_numpy(sc.depths).cuda()
br = wf.brick_nodes(vol.res, 0, R)
engine.update_projective(vol, wf, sc.lw, depths, sc.K, sc.Kinv, sc.extrinsics, sc.tdist, mode=4)
torch.cuda.synchronize()
rec = br[6].cpu().numpy().reshape(-1, 16)
code = rec[:, 15]
print("regions", len(code), "invalid", (code == 0).mean(), "valid-unresolved", (code == 1).mean(), "resolved", (code >= 2).mean(),
      "skip", (code == 2).mean(), "dev max/median", rec[code >= 1, 12:15].max(), np.median(rec[code >= 1, 12:15].max(axis=1)))
cls = vol.workspace.brick_cls.cpu().numpy()[: vol.workspace.n_bricks]
print("bricks", len(cls), "skip", (cls == 0).mean(), "mixed", (cls == 255).mean(), "clamp", ((cls != 0) & (cls != 255)).mean())
