"""Scratch: where a Gauss-Newton iteration spends its time (host wall clock with syncs vs device time)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynamicfusion_body_b200 import synth, engine, gn
sc = synth.make_scene(res=256, k=4, n_nodes=int(sys.argv[1]) if len(sys.argv) > 1 else 1000, seed=0, background=True)
pd = synth.make_gn_problem(sc, 300000, seed=0)
wf = engine.DeviceWarpField(4); wf.set_nodes(sc.node_pos, sc.node_dq, np.float32(sc.node_w))
prob = gn.Problem(wf, pd.vertices, pd.normals, pd.corr, pd.vert_knn, pd.node_vertex_idx)
x = torch.from_numpy(pd.x0).cuda()
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = prob.gauss_newton(x, sc.lw, 0.05, max_iter=15, huber=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("gauss_newton: %.3f ms / iteration (%d iterations, pcg %s)" % (1e3 * (t1 - t0) / res.iterations, res.iterations, [h["pcg_iterations"] for h in res.history]))
H, g, c = prob.normal_equations(x, sc.lw, 0.05, huber=True)
def T(fn, n=20):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t) / n
print("normal_equations        %.3f ms" % T(lambda: prob.normal_equations(x, sc.lw, 0.05, huber=True)))
print("solve_step (1 pcg iter) %.3f ms" % T(lambda: prob.solve_step(H, g, x, 1e-3, 1, 1e-9)))
print("solve_step (400)        %.3f ms" % T(lambda: prob.solve_step(H, g, x, 1e-3, 400, 1e-7)))
print("cost .item()            %.3f ms" % T(lambda: c[0].item()))
