"""Scratch: Gauss-Newton loop with different PCG tolerances / damping floors (inexact Newton): final cost, PCG iterations, time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynamicfusion_body_b200 import synth, engine, gn
sc = synth.make_scene(res=256, k=4, n_nodes=1000, seed=0, background=True)
pd = synth.make_gn_problem(sc, 300000, seed=0)
wf = engine.DeviceWarpField(4); wf.set_nodes(sc.node_pos, sc.node_dq, np.float32(sc.node_w))
prob = gn.Problem(wf, pd.vertices, pd.normals, pd.corr, pd.vert_knn, pd.node_vertex_idx)
x = torch.from_numpy(pd.x0).cuda()
prob.gauss_newton(x, sc.lw, 0.05, max_iter=2, huber=True)
for tol, lam_min in ((1e-9, 1e-5), (1e-7, 1e-5), (1e-5, 1e-5), (1e-3, 1e-5), (1e-2, 1e-5), (1e-3, 1e-4), (1e-3, 1e-3), (1e-3, 1e-2), (1e-5, 1e-3)):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = prob.gauss_newton(x, sc.lw, 0.05, max_iter=15, huber=True, ftol=0.0, pcg_tol=tol, lam_min=lam_min)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    its = [h["pcg_iterations"] for h in res.history]
    costs = [h["cost"] for h in res.history]
    xerr = float(torch.linalg.norm(res.x - torch.from_numpy(pd.x_true).cuda()) / np.linalg.norm(pd.x0 - pd.x_true))
    print("pcg_tol %.0e lam_min %.0e: %.3f ms/iter, accepted %d/%d, final cost %.9g, cost@3 %.6g cost@5 %.6g, |x-x*|/|x0-x*| %.4f, pcg its %s"
          % (tol, lam_min, 1e3 * (t1 - t0) / res.iterations, res.accepted, res.iterations, res.cost, costs[2], costs[4], xerr, its))
