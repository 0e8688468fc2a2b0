"""Scratch: time one PCG solve of the bench GN problem."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynamicfusion_body_b200 import synth, engine, gn
sc = synth.make_scene(res=256, k=4, n_nodes=int(sys.argv[1]) if len(sys.argv) > 1 else 1000, seed=0, background=True)
pd = synth.make_gn_problem(sc, 300000, seed=0)
wf = engine.DeviceWarpField(4); wf.set_nodes(sc.node_pos, sc.node_dq, np.float32(sc.node_w))
prob = gn.Problem(wf, pd.vertices, pd.normals, pd.corr, pd.vert_knn, pd.node_vertex_idx)
x = torch.from_numpy(pd.x0).cuda()
H, g, c = prob.normal_equations(x, sc.lw, 0.05, huber=True)
for rep in range(3):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); c2 = torch.cuda.Event(enable_timing=True)
    a.record(); H, g, c = prob.normal_equations(x, sc.lw, 0.05, huber=True); b.record()
    xn, d, info = prob.solve_step(H, g, x, 1e-4, 400, 1e-9); c2.record(); torch.cuda.synchronize()
print("blocks", os.environ.get("DFB_PCG_BLOCKS"), "normal_eq %.3f ms  pcg %.3f ms  iters %d  nnzb %d" % (a.elapsed_time(b), b.elapsed_time(c2), int(info[6].item()), prob.pattern()[2]))
