"""Per-slab step time of the strong-scaling configuration on ONE GPU (no communication): for world = 1, 2, 4, 8 every rank's x-slab
of the 512^3 benchmark grid is run through the one-launch frame step.  Shows how much of the N-GPU step is compute imbalance /
fixed cost as opposed to NCCL.  usage: python scripts/slab_times.py [res] [nodes] [balanced]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dynamicfusion_body_b200 import synth, engine, dist as ddist

R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
sc = synth.make_scene(res=R, k=4, n_nodes=N, seed=0, background=True)
dev = torch.device("cuda", 0)
depth = torch.from_numpy(sc.depths).to(dev)
rng = np.random.default_rng(1)
dqs = [torch.from_numpy(sc.node_dq + (rng.normal(size=sc.node_dq.shape) * 1e-4).astype(np.float32)).to(dev) for _ in range(15)]
stream = torch.cuda.Stream(device=dev)
worlds = [int(w) for w in os.environ.get("WORLDS", "1,2,4,8").split(",")]
for world in worlds:
    parts = ddist.slab_partition(R, world)
    times = []
    for (x0, x1) in parts:
        wf = engine.DeviceWarpField(4, dev)
        wf.set_nodes(sc.node_pos, sc.node_dq, np.float32(sc.node_w))
        vol = engine.DeviceVolume((R, R, R), x0, x1, device=dev, fill=sc.tdist)
        views = engine.make_views(depth, sc.K, sc.Kinv, sc.extrinsics)
        step = engine.FrameStep(vol, wf, sc.lw, views, sc.tdist)
        io = step.io()
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            for i in range(5):
                wf.node_dq.copy_(dqs[i % 15], non_blocking=True)
                step.run(io)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(40):
                wf.node_dq.copy_(dqs[i % 15], non_blocking=True)
                step.run(io)
            e1.record(stream)
        torch.cuda.synchronize()
        st = vol.workspace.stats()
        times.append((e0.elapsed_time(e1) / 40, st["bricks_streamed"], st["bricks_mixed"], st["deferred"]))
        del step, vol, wf
        torch.cuda.empty_cache()
    ms = [t[0] for t in times]
    print("world %d: max %.4f ms  mean %.4f ms  per slab: %s" % (world, max(ms), float(np.mean(ms)), " ".join("%.4f" % m for m in ms)))
    print("   (streamed, mixed, deferred) per slab:", [(t[1], t[2], t[3]) for t in times])
