#!/usr/bin/env python
"""bench.py -- headline benchmark of the warped-TSDF hot path (BASELINE.json metric:
"warped-TSDF voxels/sec ... (% HBM roofline); GN solve ms/iter").

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...          (CPU: the reference path restated in oracle/, all host cores)

A step = one frame of the a3 path over one GPU's slab: new depth frame + new node transforms ->
node records packed -> warped projective TSDF update (fast pass + reference-exact pass).
Workload at N=1: 512^3 voxels, ~4k nodes, k=4 DQB, one 640x480 depth view (north_star target config;
BASELINE configs[4] at one GPU).  N>1: weak scaling -- the grid is ~(512*N^(1/3))^3 (y/z a multiple of 32: 640 / 800 /
1024 at N = 2 / 4 / 8) so every rank owns an x-slab of >= 512^3 voxels of it; rank 0 broadcasts the depth views of the
next frame on a side stream and the node transforms inside the step, over NCCL.

`value`   : voxels/s with the frame already resident in HBM.
`e2e`     : voxels/s through the reference-facing class call (Fusion.fuseFrame) with HOST numpy buffers:
            H2D of depth + node transforms and D2H of the per-frame statistics are inside the timed region.
            The TSDF volume itself is persistent device state (Fusion._tsdf), as in the reference where it
            is an attribute that lives across frames.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_VOXEL = 16.0  # read v, read w, write v, write w (fp32) -- SURVEY 8d
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture (profiles/), bytes
TRAFFIC_PER_LAUNCH = {"brick_update_kernel": 971359744, "proj_exact_kernel": 198060288, "brick_classify_kernel": 1899008,
                      "region_bounds_kernel": 5570304}  # profiles/r1_ncu_full_summary.md


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
            return
        t0 = time.time()                       # nvidia-smi needs a moment before its first sample
        while time.time() - t0 < 5.0:
            try:
                if os.path.getsize(self.path) > 0:
                    break
            except OSError:
                pass
            time.sleep(0.02)
        self.skip = sum(1 for _ in open(self.path))   # samples taken before the timed region starts

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for ln, line in enumerate(open(self.path)):
                if ln < getattr(self, "skip", 0):
                    continue
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def build_scene(res, n_nodes, k, n_views, seed=0):
    from dynamicfusion_body_b200 import synth
    return synth.make_scene(res=res, k=k, n_nodes=n_nodes, seed=seed, n_views=n_views, background=True)


def frame_dqs(sc, n_frames, seed=1):
    """Per-frame node transforms: the scene's field plus a small per-frame perturbation (15-frame sequence shape)."""
    rng = np.random.default_rng(seed)
    return [(sc.node_dq + (rng.normal(size=sc.node_dq.shape) * 1e-4).astype(np.float32)) for _ in range(n_frames)]


# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from dynamicfusion_body_b200 import _capi, engine
    from dynamicfusion_body_b200 import dist as ddist
    from dynamicfusion_body_b200.fusion import Fusion

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    per_gpu_res = args.res
    # y/z extent: the multiple of 32 nearest to per_gpu_res * N^(1/3), so that every z-row starts on a 128-byte line like the
    # power-of-two grids of the reference's configurations (a 645-voxel row at N=2 put the streaming pass on its unaligned
    # scalar path: 0.27 instead of 0.11 ms); x-thickness of a slab: the multiple of 4 giving >= per_gpu_res^3 voxels per rank
    res = max(32, int(round(per_gpu_res * world ** (1.0 / 3.0) / 32.0)) * 32)
    slab = -(-int(math.ceil(per_gpu_res ** 3 / float(res * res))) // 4) * 4
    res_x = slab * world
    grid = (res_x, res, res)
    x0, x1 = rank * slab, (rank + 1) * slab
    sc = build_scene(res, args.nodes, args.k, args.views)
    # the scene is generated for a cubic `res` grid; with world>1 res_x can differ from res by rounding
    fus = Fusion(sc.tdist, knn=args.k, device=dev, use_cnn=False, write_warpfield=False)
    fus.InitializeCanonicalSpace(tsdf_shape=grid, slab=(x0, x1), K=sc.K, vertices=sc.vertices, normals=sc.normals,
                                 nodes=sc.nodes_as_reference_tuples())
    fus._lw = sc.lw
    nvox_rank = (x1 - x0) * res * res
    n_frames = 15
    dqs = frame_dqs(sc, n_frames)
    dq_dev = [torch.from_numpy(d).to(dev) for d in dqs]
    depth_dev = torch.from_numpy(sc.depths.copy()).to(dev)
    depth_host = torch.from_numpy(sc.depths.copy()).pin_memory()
    dq_host = [torch.from_numpy(d).pin_memory() for d in dqs]
    fus.build_knn()                                                  # once per graph revision, outside the timed region
    torch.cuda.synchronize()

    # Frames travel in double-buffered packets.  The depth views of step t+1 (sensor data: independent of the fusion result)
    # are put into the other packet -- and, with several GPUs, broadcast from rank 0 over NVLink on a side stream with its own
    # process group -- while step t computes; the node transforms (which in the application come out of the solve against the
    # model updated by step t) are uploaded / broadcast inside their own step.
    shape = sc.depths.shape
    packets = [ddist.FramePacket(shape[0], shape[1], shape[2], sc.n_nodes, dev) for _ in range(2)]
    side = torch.cuda.Stream(device=dev)
    side_group = dist.new_group() if world > 1 else None
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    pending = {"h": None}

    def stage_depth(slot, src):
        side.wait_event(done[slot])                                   # the step that last read this packet has finished
        with torch.cuda.stream(side):
            if rank == 0:
                packets[slot].depths.copy_(src, non_blocking=True)    # pinned host (e2e) or device-resident frame
            packets[slot].broadcast_depths(group=side_group)
            ready[slot].record(side)

    def run_step(i, depth_src, dq_src):
        slot = i & 1
        pk = packets[slot]
        torch.cuda.current_stream().wait_event(ready[slot])
        if rank == 0:
            pk.node_dq.copy_(dq_src[i % n_frames], non_blocking=True)
        pk.broadcast_transforms()
        fus.set_node_dqs(pk.node_dq)
        fus.fuseFrame(pk.depths, extrinsics=sc.extrinsics)
        done[slot].record()
        stage_depth(slot ^ 1, depth_src)

    def begin(depth_src):
        def f():
            done[0].record(); done[1].record()
            stage_depth(0, depth_src)
        return f

    def step_resident(i):
        run_step(i, depth_dev, dq_dev)

    def step_e2e(i):
        run_step(i, depth_host, dq_host)
        h = fus.frame_stats_async()                                   # D2H of the per-frame counters (32 B) ...
        if pending["h"] is not None:
            pending["h"].result()                                     # ... read on the host while the next step is queued
        pending["h"] = h

    def e2e_end():
        if pending["h"] is not None:
            pending["h"].result()
            pending["h"] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, warmup, begin=None, end=None):
        if begin:
            begin()
        for i in range(warmup):
            step_fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(warmup, warmup + steps):
            step_fn(i)
        if end:
            end()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total = timed(step_resident, args.steps, args.warmup, begin(depth_dev))
    # the K timed steps last only a few tens of ms; keep the same step running (untimed, identical count on every rank so
    # that the collectives match) until the 20 ms sampler has seen ~0.4 s of this load
    n_extra = max(0, int(400.0 / max(ms_total / args.steps, 1e-3)) - args.steps)
    for i in range(min(n_extra, 2000)):
        step_resident(i)
        if i % 32 == 31:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    ms_e2e = timed(step_e2e, args.steps, max(3, args.warmup // 2), begin(depth_host), e2e_end)
    stats = fus.frame_stats()

    # ---- roofline leg: every kernel of the step timed alone with CUDA events on its stream ----
    kk = 4 if args.k <= 4 else 8
    # production launches per step: nodes_pack, region_bounds_kernel, brick_classify_kernel, brick_update_kernel (CLAMP + MIXED
    # bricks fused), proj_exact_kernel.  The two halves of the fused pass are also timed alone (profiling modes) for the breakdown.
    prod = [("region_bounds_kernel+brick_classify_kernel", _capi.MODE_BRICK_CLASSIFY), ("brick_update_kernel<%d>" % kk, _capi.MODE_BRICK_UPDATE),
            ("proj_exact_kernel<%d>" % kk, _capi.MODE_LIST_ONLY)]
    parts = [("region_bounds_kernel+brick_classify_kernel", _capi.MODE_BRICK_CLASSIFY), ("brick_stream_kernel", _capi.MODE_BRICK_STREAM),
             ("brick_mixed_kernel<%d>" % kk, _capi.MODE_BRICK_MIXED)]

    def time_modes(seq, reps):
        acc = {n: [] for n, _ in seq}
        for i in range(reps):
            fus.set_node_dqs(dq_dev[i % n_frames])
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(seq) + 1)]
            ev[0].record()
            for j, (n, mode) in enumerate(seq):
                fus.fuseFrame(depth_dev, extrinsics=sc.extrinsics, mode=mode)
                ev[j + 1].record()
            torch.cuda.synchronize()
            for j, (n, _) in enumerate(seq):
                acc[n].append(ev[j].elapsed_time(ev[j + 1]))
        return {n: float(np.mean(v[2:])) for n, v in acc.items()}

    reps = max(6, min(args.steps, 20))
    kms = time_modes(prod, reps)
    stats = fus.frame_stats()
    pms = time_modes(parts, reps)
    vox_per_brick = 4 * 4 * 32
    n_stream, n_mixed = stats["bricks_streamed"] * vox_per_brick, stats["bricks_mixed"] * vox_per_brick
    units = {prod[0][0]: 0, prod[1][0]: n_stream + n_mixed, prod[2][0]: stats["deferred"]}
    punits = {parts[1][0]: n_stream, parts[2][0]: n_mixed}
    names = prod
    # dominant = the slowest kernel that moves volume data (the classifier reads 48 B per brick, no voxels)
    dominant = max([n for n, _ in prod[1:]], key=kms.get)
    total_vox = nvox_rank * world
    value = total_vox * args.steps / (ms_total * 1e-3)
    e2e_value = total_vox * args.steps / (ms_e2e * 1e-3)
    peak, peak_src = measured_peaks()
    step_ms = ms_total / args.steps
    kernels = [{"kernel": n, "ms": kms[n], "voxels": int(units[n]),
                "achieved_GBps": ALG_BYTES_PER_VOXEL * units[n] / (kms[n] * 1e-3) / 1e9,
                "frac": ALG_BYTES_PER_VOXEL * units[n] / (kms[n] * 1e-3) / 1e9 / peak} for n, _ in names]
    kernels += [{"kernel": n + " (half of the fused pass, timed alone)", "ms": pms[n], "voxels": int(punits[n]),
                 "achieved_GBps": ALG_BYTES_PER_VOXEL * punits[n] / (pms[n] * 1e-3) / 1e9,
                 "frac": ALG_BYTES_PER_VOXEL * punits[n] / (pms[n] * 1e-3) / 1e9 / peak} for n, _ in parts[1:]]
    achieved = ALG_BYTES_PER_VOXEL * units[dominant] / (kms[dominant] * 1e-3) / 1e9
    step_achieved = ALG_BYTES_PER_VOXEL * nvox_rank / (step_ms * 1e-3) / 1e9
    out = {
        "metric": "warped_tsdf_voxels_per_sec", "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%d^3 voxels per GPU (grid %dx%dx%d), warped projective TSDF update (a3), k=%d DQB, %d nodes, "
                               "%d view(s) 640x480, 15-frame dq sequence" % (per_gpu_res, res_x, res, res, args.k, sc.n_nodes, args.views),
                   "l2": "inputs larger than L2 (%.0f MB of v,w,kNN per GPU)" % (nvox_rank * (8 + 2 * args.k) / 1e6),
                   "parallelism": "x-slab per GPU, frame broadcast over NCCL" if world > 1 else "single GPU",
                   "deferred_voxel_fraction": stats["deferred"] / nvox_rank},
        "e2e": {"value": e2e_value, "unit": "voxels/s", "h2d_bytes_per_step": int(sc.depths.nbytes + dqs[0].nbytes),
                "d2h_bytes_per_step": 32, "ms_per_step": ms_e2e / args.steps,
                "pipeline": "pinned-host depth frame of step t+1 uploaded on a copy stream while step t computes (double-buffered "
                            "frame packet); node transforms uploaded inside their own step; counters of step t read back while step "
                            "t+1 is queued; every step's H2D and D2H lie inside the timed region"},
        "gpu_launches": 5 * args.steps,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": TRAFFIC_PER_LAUNCH.get(dominant.split("<")[0]),
                     "note": "dominant kernel by time; its units = the voxels that launch processes x 16 B. The step as a whole "
                             "(all voxels of the slab x 16 B / step time) is in `step`.",
                     "algorithmic_bytes_per_voxel": ALG_BYTES_PER_VOXEL,
                     "step": {"achieved": step_achieved, "frac": step_achieved / peak, "ms": step_ms},
                     "kernels": kernels,
                     "bricks": {"total": stats["bricks"], "streamed": stats["bricks_streamed"], "mixed": stats["bricks_mixed"]}},
    }
    if not args.no_gn:
        out["gn"] = bench_gn(args, dev, rank, world)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(sc, (res_x, res, res), budget_s=args.cpu_seconds)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def bench_gn(args, dev, rank, world):
    """Second half of the BASELINE metric: Gauss-Newton ms/iteration (config 3: ~1k nodes, k=4, ~300k data residuals of
    one 640x480 frame, 15 iterations).  N>1: data residuals sharded over ranks, normal equations all-reduced (NCCL)."""
    import torch
    from dynamicfusion_body_b200 import dist as ddist
    from dynamicfusion_body_b200 import engine, gn, synth
    sc = synth.make_scene(res=256, k=4, n_nodes=args.gn_nodes, seed=0, background=True)
    pd = synth.make_gn_problem(sc, args.gn_points, seed=0)
    wf = engine.DeviceWarpField(4, dev)
    wf.set_nodes(sc.node_pos, sc.node_dq, np.float32(sc.node_w))
    shard = ddist.residual_partition(len(pd.vertices), world)[rank]
    prob = gn.Problem(wf, pd.vertices, pd.normals, pd.corr, pd.vert_knn, pd.node_vertex_idx, shard=shard, reg_owner=(rank == 0))
    x = torch.from_numpy(pd.x0).to(dev)
    allreduce = (lambda H, g, c: ddist.allreduce_normal_equations(H, g, c)) if world > 1 else None
    prob.pattern()
    res = prob.gauss_newton(x, sc.lw, 0.05, max_iter=2, huber=True, allreduce=allreduce)         # warm-up
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record()
    res = prob.gauss_newton(x, sc.lw, 0.05, max_iter=args.gn_iters, huber=True, ftol=0.0, allreduce=allreduce)
    e[1].record()
    torch.cuda.synchronize()
    total_ms = ddist.max_over_ranks(e[0].elapsed_time(e[1]), dev)
    # breakdown of one iteration
    H, g, c = prob.normal_equations(x, sc.lw, 0.05, huber=True)
    torch.cuda.synchronize()
    t = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t[0].record()
    H, g, c = prob.normal_equations(x, sc.lw, 0.05, huber=True)
    t[1].record()
    if allreduce:
        allreduce(H, g, c)
    t[2].record()
    x_new, delta, info = prob.solve_step(H, g, x, 1e-3, 400, 1e-9)
    t[3].record()
    torch.cuda.synchronize()
    row_ptr, col_idx, nnzb = prob.pattern()
    # the step in front of every solve: closest-point correspondences of all canonical points against a
    # live surface of the same size -- warp, search-grid build, exact 4-NN, best point-to-plane candidate
    live = torch.from_numpy(pd.corr.astype(np.float32)).to(dev)
    loc = torch.from_numpy(pd.vert_knn.astype(np.int32)).to(dev)

    def corr_step():
        wv, wn = engine.warp_points(wf, sc.lw, pd.vertices, pd.normals, idx=loc, k=4)
        nn = engine.PointGrid(live, device=dev).knn(wv, 4)
        return engine.corr_select(wv, wn, live, nn)

    corr_step()
    torch.cuda.synchronize()
    c = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    c[0].record()
    for _ in range(3):
        best, cost = corr_step()
    c[1].record()
    torch.cuda.synchronize()
    corr_ms = c[0].elapsed_time(c[1]) / 3
    return {"metric": "gn_solve_ms_per_iter", "value": total_ms / max(1, res.iterations), "unit": "ms", "iterations": res.iterations,
            "accepted": res.accepted, "cost0": res.cost0, "cost": res.cost,
            "config": {"workload": "warp-field Gauss-Newton, %d nodes, k=4, %d data + %d regularisation residuals, %d unknowns, %d 8x8 blocks"
                                   % (sc.n_nodes, len(pd.vertices), 3 * 4 * sc.n_nodes, 8 * sc.n_nodes, nnzb)},
            "breakdown_ms": {"normal_equations": t[0].elapsed_time(t[1]), "allreduce": t[1].elapsed_time(t[2]),
                             "pcg_solve_and_update": t[2].elapsed_time(t[3]), "pcg_iterations": int(info[6].item())},
            "correspondences": {"ms": corr_ms, "canonical_points": len(pd.vertices), "live_points": int(live.shape[0]), "k": 4,
                                "matched_within_0.2": float((cost <= 0.2).float().mean().item()),
                                "note": "setupCorrespondences body (core/fusion.py:258-276) incl. host->device upload of the "
                                        "canonical points and the search-grid build"},
            "reference_published_ms_per_iter": 70100.0}


def cpu_baseline(sc, res, budget_s=12.0):
    from oracle import driver
    sd = driver.scene_dict(sc)
    vps, n, dt = driver.time_projective(sd, res, 100_000)
    n_sample = int(min(4_000_000, max(100_000, vps * budget_s)))
    vps, n, dt = driver.time_projective(sd, res, n_sample, seed=1)
    return {"value": vps, "unit": "voxels/s", "cores": 1, "kind": "port",
            "sample": "%d random voxels of the same %dx%dx%d workload, numpy oracle (oracle/tsdf.py) incl. KD-tree kNN, %.1f s" % (n, res[0], res[1], res[2], dt)}


def run_reference(args):
    """The reference path on the host cores.  The reference is pure Python and does not exist on the GPU box, so
    this times its numpy restatement (oracle/, kind="port") on a bounded voxel sample per step, all cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import driver
    res = args.res
    sc = build_scene(res, args.nodes, args.k, args.views)
    sd = driver.scene_dict(sc)
    cores = os.cpu_count() or 1
    per_step = args.ref_sample
    for _ in range(min(1, args.warmup)):
        driver.time_projective_parallel(sd, (res, res, res), max(cores * 20000, per_step // 4), cores)
    t0 = time.perf_counter()
    nvox = 0
    rates = []
    for _ in range(args.steps):
        r, n, wall = driver.time_projective_parallel(sd, (res, res, res), per_step, cores)
        nvox += n
        rates.append(n / wall)
    wall = time.perf_counter() - t0
    value = nvox / wall
    out = {"impl": "reference", "metric": "warped_tsdf_voxels_per_sec", "value": value, "unit": "voxels/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "%d^3 voxels, warped projective TSDF update (a3), k=%d DQB, %d nodes, %d view(s) 640x480; each step = a "
                                  "bounded sample of %d voxels" % (res, args.k, sc.n_nodes, args.views, per_step)},
           "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": cores, "kind": "port",
                            "sample": "%d random voxels per step over %d worker processes (numpy oracle incl. KD-tree kNN)" % (per_step, cores)},
           "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--res", type=int, default=512, help="per-GPU cube edge (voxels per GPU = res^3)")
    ap.add_argument("--nodes", type=int, default=4000)
    ap.add_argument("--k", type=int, default=4)
    ap.add_argument("--views", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gn", action="store_true")
    ap.add_argument("--gn-nodes", type=int, default=1000)
    ap.add_argument("--gn-points", type=int, default=300000)
    ap.add_argument("--gn-iters", type=int, default=15)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-sample", type=int, default=1_600_000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
