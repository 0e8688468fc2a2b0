"""Graph revision cost at 512^3: full rebuild of the voxel kNN table + brick / region sets against the incremental update after
+1 % appended nodes (DeviceWarpField.append_nodes), with the bit-equality check against a full rebuild."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynamicfusion_body_b200 import synth, engine
R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
sc = synth.make_scene(res=R, k=4, n_nodes=N, seed=0, background=True)
dev = torch.device("cuda", 0)
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for frac in (0.01, 0.05):
    wf = engine.DeviceWarpField(4, dev); wf.set_nodes(sc.node_pos, sc.node_dq, np.float32(sc.node_w))
    torch.cuda.synchronize(); a = ev(); wf.knn_table((R, R, R), 0, R); b = ev(); wf.brick_nodes((R, R, R), 0, R); c = ev(); torch.cuda.synchronize()
    rng = np.random.default_rng(7)
    m = max(1, int(sc.n_nodes * frac))
    new_pos = sc.vertices[rng.choice(len(sc.vertices), m, replace=False)].astype(np.float32)
    new_dq = np.tile(np.array([1, 0, 0, 0, 0, 0, 0, 0], np.float32), (m, 1))
    torch.cuda.synchronize(); d = ev(); wf.append_nodes(new_pos, new_dq, np.float32(sc.node_w)); e = ev(); torch.cuda.synchronize()
    ref = engine.DeviceWarpField(4, dev); ref.set_nodes(wf.node_pos, wf.node_dq, wf.node_w)
    same = torch.equal(ref.knn_table((R, R, R), 0, R), wf.knn_table((R, R, R), 0, R))
    dirty = wf.last_dirty
    print("R=%d N=%d +%d nodes: full build knn %.2f ms + brick/region sets %.2f ms; incremental %.3f ms; 8^3 bricks with changed rows %.3f, "
          "rebuilt from scratch %.4f; table == full rebuild: %s" % (R, sc.n_nodes, m, a.elapsed_time(b), b.elapsed_time(c), d.elapsed_time(e),
                                                                       float((dirty == 1).float().mean()), float((dirty == 2).float().mean()), same))
    del wf, ref
