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

    # STRONG scaling (BASELINE configs[4]): ONE res^3 grid, x-slab [x0, x1) per rank
    res = args.res
    grid = (res, res, res)
    x0, x1 = ddist.slab_partition(res, world)[rank]
    sc = build_scene(res, args.nodes, args.k, args.views)
    fus = Fusion(sc.tdist, knn=args.k, device=dev, use_cnn=False, write_warpfield=False)
    fus.InitializeCanonicalSpace(tsdf_shape=grid, slab=(x0, x1), K=sc.K, vertices=sc.vertices, normals=sc.normals,
                                 nodes=sc.nodes_as_reference_tuples())
    fus._lw = sc.lw
    nvox_rank = (x1 - x0) * res * res
    nvox_total = res ** 3
    n_frames = 15
    dqs = frame_dqs(sc, n_frames)
    dq_dev = [torch.from_numpy(d).to(dev) for d in dqs]
    depth_dev = torch.from_numpy(sc.depths.copy()).to(dev)
    depth_host = torch.from_numpy(sc.depths.copy()).pin_memory()

    # ---- graph revision: voxel kNN table + brick / region candidate sets (once per update_graph, core/fusion.py:229) ----
    torch.cuda.synchronize()
    rev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    rev[0].record()
    fus.build_knn()
    rev[1].record()
    torch.cuda.synchronize()
    graph_revision_ms = ddist.max_over_ranks(rev[0].elapsed_time(rev[1]), dev)

    # ---- communicators (raw NCCL through the C ABI): one for the in-step transform broadcast, one for the sensor prefetch ----
    comm = comm_pre = None
    if world > 1:
        comm = engine.Comm.from_torch(dev)
        comm_pre = engine.Comm.from_torch(dev)

    stream = torch.cuda.Stream(device=dev)           # a capturable (non-default) stream: the step is one CUDA-graph launch
    shape = sc.depths.shape
    packets = [torch.zeros(shape, dtype=torch.float32, device=dev) for _ in range(2)]
    dq_stage = [torch.zeros((sc.n_nodes, 8), dtype=torch.float32).pin_memory() for _ in range(2)]
    counters_host = [torch.zeros(8, dtype=torch.int32).pin_memory() for _ in range(2)]
    slot_done = [torch.cuda.Event(), torch.cuda.Event()]
    wf, vol = fus._wf, fus._vol
    steps_obj, ios_res, ios_e2e = [], [], []
    for slot in range(2):
        views = engine.make_views(packets[slot], sc.K, sc.Kinv, sc.extrinsics)
        st = engine.FrameStep(vol, wf, sc.lw, views, sc.tdist)
        steps_obj.append(st)
        # resident frame: the transforms are already in wf.node_dq on the root, the next depth frame is a device buffer
        ios_res.append(st.io(comm=comm, comm_prefetch=comm_pre, root=0, prefetch_dst=packets[slot ^ 1], prefetch_src=depth_dev))
        # end to end: transforms and depth come from pinned host memory, the counters go back to pinned host memory
        ios_e2e.append(st.io(comm=comm, comm_prefetch=comm_pre, root=0, dq_src=dq_stage[slot], prefetch_dst=packets[slot ^ 1],
                             prefetch_src=depth_host, counters_host=counters_host[slot]))

    def prime(src):
        """frame 0's sensor data into packet 0 (every later frame arrives through the prefetch branch of the step before it)"""
        with torch.cuda.stream(stream):
            if rank == 0:
                packets[0].copy_(src, non_blocking=True)
            if comm_pre is not None:
                comm_pre.broadcast(packets[0])
        stream.synchronize()

    def step_resident(i):
        slot = i & 1
        with torch.cuda.stream(stream):
            if rank == 0:
                wf.node_dq.copy_(dq_dev[i % n_frames], non_blocking=True)
            steps_obj[slot].run(ios_res[slot])

    def step_e2e(i):
        slot = i & 1
        slot_done[slot].synchronize()                                  # step i-2 has consumed this slot's staging buffers
        stats = counters_host[slot].numpy().copy()                     # D2H result of step i-2, read on the host
        if rank == 0:
            dq_stage[slot].numpy()[...] = dqs[i % n_frames]            # host -> pinned staging (the solver's output)
        with torch.cuda.stream(stream):
            steps_obj[slot].run(ios_e2e[slot])
            slot_done[slot].record()
        return stats

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, warmup, src):
        prime(src)
        for i in range(warmup):
            step_fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(warmup, warmup + steps):
            step_fn(i)
        e1.record(stream)
        barrier()
        return ddist.max_over_ranks(e0.elapsed_time(e1), dev)

    # ---- untimed parity check on the hardware (N > 1): slabs gathered over NCCL == a single-volume run on rank 0 ----
    parity = None
    if world > 1:
        parity = parity_check(args, sc, fus, dq_dev, depth_dev, steps_obj, ios_res, prime, stream, comm, rank, world, dev)
        # fresh state for the timed runs
        vol.tsdf.fill_(sc.tdist); vol.weight.zero_()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total = timed(step_resident, args.steps, args.warmup, depth_dev)
    # the K timed steps last only a few tens of ms; keep the same step running (untimed, identical count on every rank so
    # that the collectives match) until the 20 ms sampler has seen ~0.4 s of this load
    n_extra = max(0, int(400.0 / max(ms_total / args.steps, 1e-3)) - args.steps)
    for i in range(min(n_extra, 4000)):
        step_resident(i)
        if i % 64 == 63:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    for e in slot_done:
        e.record(stream)
    ms_e2e = timed(step_e2e, args.steps, max(3, args.warmup // 2), depth_host)
    stats = fus.frame_stats()
    gstats = steps_obj[0].stats()

    # ---- roofline leg: every kernel of the step timed alone with CUDA events on its stream ----
    kk = 4 if args.k <= 4 else 8
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
    value = nvox_total * args.steps / (ms_total * 1e-3)
    e2e_value = nvox_total * args.steps / (ms_e2e * 1e-3)
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
    traffic = measured_traffic(dominant) if world == 1 else None
    out = {
        "metric": "warped_tsdf_voxels_per_sec", "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%d^3 voxels total (one grid, x-slab of %d planes per GPU), warped projective TSDF update (a3), k=%d DQB, "
                               "%d nodes, %d view(s) 640x480, 15-frame dq sequence" % (res, x1 - x0, args.k, sc.n_nodes, args.views),
                   "l2": "inputs larger than L2 (%.0f MB of v,w,kNN per GPU)" % (nvox_rank * (8 + 2 * args.k) / 1e6)
                         if nvox_rank * (8 + 2 * args.k) > 2.5e8 else
                         "slab of %.0f MB (v,w,kNN) per GPU: at N>=4 a 512^3/N slab approaches the 126 MB L2; every step streams the whole slab "
                         "once, so the working set is re-read from HBM unless it fits" % (nvox_rank * (8 + 2 * args.k) / 1e6),
                   "parallelism": ("x-slab per GPU (strong scaling of ONE grid); per step ONE CUDA-graph launch per rank: NCCL broadcast of the "
                                   "node transforms + update kernels, next depth frame broadcast as a concurrent graph branch") if world > 1
                                  else "single GPU, one CUDA-graph launch per step",
                   "deferred_voxel_fraction": stats["deferred"] / nvox_rank},
        "e2e": {"value": e2e_value, "unit": "voxels/s", "h2d_bytes_per_step": int(sc.depths.nbytes + dqs[0].nbytes),
                "d2h_bytes_per_step": 32, "ms_per_step": ms_e2e / args.steps,
                "pipeline": "per step: host copies the frame's node transforms into pinned staging, ONE graph launch = [H2D transforms -> "
                            "(NCCL bcast) -> node records -> classify/update/exact kernels -> D2H of the 8 frame counters] with the H2D (+bcast) "
                            "of the next frame's pinned depth image as a concurrent branch; the host reads step t-2's counters before "
                            "re-using its slot; every step's H2D and D2H lie inside the timed region"},
        "gpu_launches": 5 * args.steps,
        "step_graph": gstats,
        "clocks": clocks,
        "graph_revision_ms": graph_revision_ms,
        "value_amortised": {"revision_every_frame": nvox_total / ((step_ms + graph_revision_ms) * 1e-3),
                            "revision_every_15_frames": nvox_total / ((step_ms + graph_revision_ms / 15.0) * 1e-3),
                            "note": "the reference rebuilds its KD-tree in update_graph (core/fusion.py:229) and queries it per voxel per frame "
                                    "(:175); here the voxel kNN table + brick/region sets are rebuilt once per graph revision "
                                    "(`graph_revision_ms`, full rebuild) and `value` excludes it"},
        "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "note": "dominant kernel by time; its units = the voxels that launch processes x 16 B. The step as a whole "
                             "(all voxels of the slab x 16 B / step time) is in `step`.",
                     "algorithmic_bytes_per_voxel": ALG_BYTES_PER_VOXEL,
                     "step": {"achieved": step_achieved, "frac": step_achieved / peak, "ms": step_ms},
                     "kernels": kernels,
                     "bricks": {"total": stats["bricks"], "streamed": stats["bricks_streamed"], "mixed": stats["bricks_mixed"]}},
    }
    if parity is not None:
        out["parity_check"] = parity
    if not args.no_gn:
        out["gn"] = bench_gn(args, dev, rank, world, comm)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(sc, grid, budget_s=args.cpu_seconds)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measured_traffic(kernel_name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed ncu --set full summary of
    THIS round's kernels (profiles/r2_ncu_full_summary.md, same 512^3 N=1 workload); None when no capture of that kernel is committed."""
    import re
    path = os.path.join(ROOT, "profiles", "r2_ncu_full_summary.md")
    if not os.path.isfile(path):
        return None
    base = kernel_name.split("<")[0]
    rd = wr = None
    active = False
    for line in open(path):
        if line.startswith("## "):
            active = base in line
        elif active:
            m = re.match(r"- dram__bytes_(read|write)\.sum = ([0-9.]+) (\w+)", line)
            if m:
                mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(m.group(3), 1.0)
                if m.group(1) == "read":
                    rd = float(m.group(2)) * mult
                else:
                    wr = float(m.group(2)) * mult
    return None if rd is None or wr is None else int(rd + wr)


def parity_check(args, sc, fus, dq_dev, depth_dev, steps_obj, ios_res, prime, stream, comm, rank, world, dev):
    """Three frames on the sharded volume (through the timed path: graph launches + NCCL broadcasts) against the same three
    frames on a single res^3 volume held by rank 0; slabs are gathered over NCCL.  SURVEY 8e: weights and values must be equal
    bit for bit."""
    import torch
    from dynamicfusion_body_b200 import engine
    res = args.res
    vol, wf = fus._vol, fus._wf
    vol.tsdf.fill_(sc.tdist); vol.weight.zero_()
    prime(depth_dev)
    n_chk = 3
    for i in range(n_chk):
        with torch.cuda.stream(stream):
            if rank == 0:
                wf.node_dq.copy_(dq_dev[i], non_blocking=True)
            steps_obj[i & 1].run(ios_res[i & 1])
    stream.synchronize()
    torch.cuda.synchronize()
    # gather the slabs on rank 0 (dfb_comm_sendrecv)
    from dynamicfusion_body_b200 import dist as ddist
    parts = ddist.slab_partition(res, world)
    result = None
    if rank == 0:
        full_t = torch.empty((res, res, res), dtype=torch.float32, device=dev)
        full_w = torch.empty_like(full_t)
        full_t[parts[0][0]:parts[0][1]] = vol.tsdf
        full_w[parts[0][0]:parts[0][1]] = vol.weight
        for r in range(1, world):
            a, b = parts[r]
            comm.sendrecv(recv=full_t[a:b], recv_peer=r)
            comm.sendrecv(recv=full_w[a:b], recv_peer=r)
        torch.cuda.synchronize()
        # the same frames on one volume
        wf1 = engine.DeviceWarpField(args.k, dev)
        wf1.set_nodes(sc.node_pos, sc.node_dq, np.float32(sc.node_w))
        vol1 = engine.DeviceVolume((res, res, res), device=dev, fill=sc.tdist)
        for i in range(n_chk):
            wf1.set_dq(dq_dev[i])
            engine.update_projective(vol1, wf1, sc.lw, depth_dev, sc.K, sc.Kinv, sc.extrinsics, sc.tdist)
        torch.cuda.synchronize()
        w_equal = bool(torch.equal(full_w, vol1.weight))
        v_equal = bool(torch.equal(full_t, vol1.tsdf))
        max_diff = float((full_t - vol1.tsdf).abs().max().item())
        updated = int((vol1.weight > 0).sum().item())
        result = {"tsdf": "ok" if (w_equal and v_equal) else ("ok_within_1e-6_tdist" if (w_equal and max_diff <= 1e-6 * sc.tdist) else "MISMATCH"),
                  "frames": n_chk, "weights_bit_equal": w_equal, "values_bit_equal": v_equal, "max_abs_value_diff": max_diff,
                  "updated_voxels": updated, "how": "slabs gathered on rank 0 over NCCL vs the same frames on one %d^3 volume" % res}
        del full_t, full_w, vol1, wf1
        torch.cuda.empty_cache()
    else:
        comm.sendrecv(send=vol.tsdf, send_peer=0)
        comm.sendrecv(send=vol.weight, send_peer=0)
        torch.cuda.synchronize()
    return result


def bench_gn(args, dev, rank, world, comm=None):
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
    allreduce = (lambda H, g, c: ddist.allreduce_normal_equations(H, g, c, comm=comm)) if world > 1 else None
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
