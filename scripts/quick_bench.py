"""Scratch timing of the TSDF passes (not the contract bench; see bench.py)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynamicfusion_body_b200 import synth, engine

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))

for R, N, k, bg in ((256, 1000, 4, True), (256, 1000, 4, False), (512, 4000, 4, True), (512, 4000, 8, True)):
    t = time.time()
    sc = synth.make_scene(res=R, k=k, n_nodes=N, seed=0, background=bg)
    print(f"--- R={R} N={sc.n_nodes} k={k} bg={bg} scene {time.time()-t:.1f}s", flush=True)
    wf = engine.DeviceWarpField(k); wf.set_nodes(sc.node_pos, sc.node_dq, np.float32(sc.node_w))
    vol = engine.DeviceVolume((R, R, R), fill=sc.tdist)
    depths = torch.from_numpy(sc.depths).cuda()
    torch.cuda.synchronize(); t = time.time()
    wf.knn_table(vol.res, 0, R); torch.cuda.synchronize()
    print(f"knn build {1e3*(time.time()-t):.1f} ms")
    views = engine.make_views(depths, sc.K, sc.Kinv, sc.extrinsics)
    f = lambda: engine.update_projective(vol, wf, sc.lw, depths, sc.K, sc.Kinv, sc.extrinsics, sc.tdist, views=views)
    best, med = timeit(f)
    st = vol.workspace.stats()
    nv = R**3
    print(f"update: best {best:.3f} ms med {med:.3f} ms  -> {nv/best/1e6:.1f} Gvox/s  16B-roofline {16*nv/best/1e6/6534.1:.3f}  deferred {st['deferred']/nv:.4f} exact {st['exact_processed']/nv:.4f}")
    m, fr = engine.update_projective(vol, wf, sc.lw, depths, sc.K, sc.Kinv, sc.extrinsics, sc.tdist, views=views, want_masks=True)
    print(f"updated frac {(m!=0).float().mean().item():.3f} frustum frac {(fr!=0).float().mean().item():.3f}")
    del vol, wf; torch.cuda.empty_cache()
