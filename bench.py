#!/usr/bin/env python
"""bench.py -- throughput of the B200 cuboid-proposal hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c4|c5]

A "step" is one pass of the hot path (detect_cuboid over every box of every frame) over one batch of
synthetic frames.  N=1 workload: BASELINE config 3 -- 256 synthetic 640x480 frames, ~3 boxes per frame.
N>1 (torchrun, one rank per GPU): every rank processes its own shard of that size (weak scaling), then one
NCCL all-gather of the top-K records.  Prints ONE JSON line on rank 0.

  value     scored (valid) cuboid proposals / s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       the same through cs_detect_cuboids_batch with pinned HOST buffers (H2D + kernels + D2H inside)
  roofline  dominant kernel: algorithmic bytes / CUDA-event time vs the measured HBM peak
  cpu_baseline  the CPU oracle (a port of the reference's algorithm) on the host cores, bounded sample
--impl reference times that CPU path alone (the reference itself cannot be compiled here: needs Eigen,
OpenCV C++ and ROS).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (frames per GPU, width, height, boxes/frame, kind, poisson, param overrides, description)
    "c2": (1, 640, 480, 5, "indoor", False, {}, "BASELINE config 2: single 640x480 frame, 5 boxes"),
    "c3": (256, 640, 480, 3, "indoor", True, {}, "BASELINE config 3: batch 256 synthetic 640x480 frames, avg 3 boxes/frame"),
    "c4": (128, 1242, 375, 8, "kitti", False, {}, "BASELINE config 4 shard: 128 KITTI-shape 1242x375 frames, 8 boxes/frame per GPU"),
    "c5": (64, 1280, 960, 8, "indoor", False, dict(yaw_step_deg=0.5, top_sample_count_override=30),
           "BASELINE config 5 shard: 64 frames 1280x960, 8 boxes, dense sweep 181 yaw x 30 top-x per GPU"),
}

# algorithmic bytes per unit of each stage
STAGE_BYTES = {
    "lsd": lambda s, shp: shp["frame_px"] * s["n_frames"] * (3 + 8 * 2 + 0.64 * 8 * 4),  # frame in, two f64 blur planes, scaled/modgrad/angle/list
    "gray": lambda s, shp: shp["frame_px"] * s["n_frames"] * 4,          # 3 B in + 1 B out per pixel
    "canny": lambda s, shp: s["roi_pixels"] * 2,                           # gray ROI read + edge map written
    "hyst": lambda s, shp: s["roi_pixels"] * 1,                            # edge map, in place
    "dt": lambda s, shp: s["roi_pixels"] * 5,                              # edge map read + f32 dist written
    "lines": lambda s, shp: s["n_lines_in"] * 32 * max(s["n_roi_jobs"], 1) / max(s["n_frames"], 1) + s["n_roi_jobs"] * 56 * 40,
    "sweep": lambda s, shp: s["n_valid"] * 200 + s["n_candidates"] * 1,    # 72 B error row + 128 B corners per scored proposal
    "fuse": lambda s, shp: s["n_valid"] * 16 + s["n_objects"] * 512,
}


def path_alg_bytes(stats, shp, topk):
    """W*H*3 + 16*M + sum 6*ROI_px + 200*P_valid + 512*N*topk, per batch."""
    return (shp["frame_px"] * stats["n_frames"] * 3 + 16 * stats["n_lines_in"] + 6 * stats["roi_pixels"] + 200 * stats["n_valid"] +
            512 * stats["n_objects"] * topk)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = False
        self.rows = []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = max((int(r[1]) for r in self.rows if r[1].isdigit()), default=None)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(self.rows)}


def make_workload(name, rank, seed_base=20260922):
    from cube_slam_b200 import synthetic as S
    F, w, h, nb, kind, poisson, over, desc = WORKLOADS[name]
    imgs, Ts, boxes, lines, K = S.make_batch(seed_base + 1000 * rank + sum(map(ord, name)), F, w, h, nb, kind=kind, poisson=poisson,
                                             distinct=min(F, 32))
    return dict(imgs=imgs, Ts=Ts, boxes=boxes, lines=lines, K=K, over=over, desc=desc, w=w, h=h, F=F)


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_run(wl, frame_ids, n_threads):
    """The reference's algorithm on the host (oracle port), one frame per task, n_threads workers."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle as O
    p = O.default_params(**wl["over"])
    O.lib()

    def one(f):
        r = O.detect_cuboid(wl["imgs"][f], wl["K"], wl["Ts"][f], wl["boxes"][f], wl["lines"][f], p)
        return r["n_valid"], r["n_candidates"]

    t0 = time.perf_counter()
    with ThreadPoolExecutor(n_threads) as ex:
        res = list(ex.map(one, frame_ids))
    dt = time.perf_counter() - t0
    return dt, sum(r[0] for r in res), sum(r[1] for r in res)


def cpu_sample_size(wl, budget_s, n_threads):
    dt, _, _ = cpu_run(wl, [0], 1)
    n = int(max(1, min(wl["F"], budget_s * n_threads / max(dt, 1e-4))))
    return n


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    wl = make_workload(args.workload, 0)
    cores = os.cpu_count() or 1
    n = cpu_sample_size(wl, 4.0, cores)
    ids = list(range(n))
    for _ in range(args.warmup):
        cpu_run(wl, ids[:max(1, n // 4)], cores)
    tot_t = tot_v = tot_c = 0.0
    for _ in range(args.steps):
        dt, v, c = cpu_run(wl, ids, cores)
        tot_t += dt
        tot_v += v
        tot_c += c
    val = tot_v / tot_t
    fps = n * args.steps / tot_t
    line = {
        "impl": "reference", "metric": "scored cuboid proposals/s", "value": val, "unit": "proposals/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "frames_per_s": fps, "candidates_per_s": tot_c / tot_t,
        "config": {"workload": wl["desc"], "sample_frames_per_step": n, "note": "CPU oracle port of the reference algorithm; the reference "
                   "itself needs Eigen/OpenCV C++/ROS and cannot be compiled on this image"},
        "cpu_baseline": {"value": val, "unit": "proposals/s", "cores": cores, "kind": "port",
                         "sample": "%d of %d frames per step, one frame per thread task, %d threads" % (n, wl["F"], cores)},
        "e2e": {"value": val, "unit": "proposals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, world, local_rank):
    import torch
    import cube_slam_b200 as cs
    from cube_slam_b200 import _lib
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    wl = make_workload(args.workload, rank)
    F, w, h = wl["F"], wl["w"], wl["h"]
    params = cs.default_params(**wl["over"])
    topk = int(params.max_cuboid_num)
    ctx = cs.Context(local_rank, w, h, F, 16, 8192)
    ctx.set_calibration(wl["K"])
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))

    recs_per_rank = 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gathered = C.c_void_p()

    # ---- resident-input throughput ("value").  `--inflight` contexts hold the same batch and are driven round-robin, so
    # the latency-bound kernels of one batch (distance transform, hysteresis, selection: one warp / CTA per ROI) overlap
    # the issue-bound kernels of the next; every timed step is still one full pass over one batch of F frames.
    ctxs, streams = [ctx], [stream]
    for _ in range(max(args.inflight, 1) - 1):
        c2 = cs.Context(local_rank, w, h, F, 16, 8192)
        c2.set_calibration(wl["K"])
        ctxs.append(c2)
        streams.append(torch.cuda.ExternalStream(c2.stream(), device=torch.device("cuda", local_rank)))
    for cx in ctxs:
        cx.upload(wl["imgs"], wl["Ts"], wl["boxes"], wl["lines"], params)
    # NCCL communicators owned by the library, one per context (the top-K all-gather is issued on that context's stream)
    if world > 1:
        nccl_path = None
        for d in sys.path:
            cand = os.path.join(d, "nvidia", "nccl", "lib", "libnccl.so.2")
            if os.path.exists(cand):
                nccl_path = cand
                break
        for cx in ctxs:
            uid = np.zeros(128, np.uint8)
            if rank == 0:
                cx.check(cx.L.cs_comm_unique_id(cx.h, (nccl_path or "").encode(), _lib.ptr(uid, C.c_uint8)))
            t = torch.from_numpy(uid).cuda()
            dist.broadcast(t, 0)
            uid = t.cpu().numpy()
            cx.check(cx.L.cs_comm_init(cx.h, (nccl_path or "").encode(), _lib.ptr(uid, C.c_uint8), world, rank))
        n_obj = torch.tensor([sum(len(b) for b in wl["boxes"])], device="cuda")
        dist.all_reduce(n_obj, op=dist.ReduceOp.MAX)
        recs_per_rank = int(n_obj.item()) * topk
    dbg_flags = (16 if args.no_prio else 0) | (32 if args.raster_dt else 0)
    for cx in ctxs:
        cx.set_profiling(dbg_flags)

    def step_i(i):
        cx = ctxs[i % len(ctxs)]
        cx.run_async()
        if world > 1:
            cx.check(cx.L.cs_allgather_topk(cx.h, recs_per_rank, C.byref(gathered)))

    for i in range(max(args.warmup, 3) * len(ctxs)):
        step_i(i)
    torch.cuda.synchronize()
    stats = ctx.stats()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev_end = [torch.cuda.Event(enable_timing=True) for _ in ctxs]
    stage_acc = {}
    barrier()
    ev0.record(streams[0])
    for st_ in streams[1:]:
        st_.wait_event(ev0)
    for i in range(args.steps):
        step_i(i)
    for e, st_ in zip(ev_end, streams):
        e.record(st_)
    barrier()
    ms_total = max(ev0.elapsed_time(e) for e in ev_end)
    # per-stage CUDA-event times of one more (profiled, alone) step
    ctx.set_profiling(1 | dbg_flags)
    ctx.run()
    stage_acc = ctx.stage_ms()
    sampler.stop_flag = True
    ms_t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
    cnt = torch.tensor([stats["n_valid"], stats["n_candidates"], stats["n_frames"], stats["n_objects"]], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_total = float(ms_t.item())
    n_valid_all, n_cand_all, n_frames_all, n_obj_all = [float(x) for x in cnt.tolist()]
    ms_per_step = ms_total / args.steps
    value = n_valid_all / (ms_per_step * 1e-3)

    # ---- online mode (object_slam main_obj.cpp:424-450): lines detected on the resident frames by the LSD kernels, N = 1 only
    online = None
    if world == 1 and not args.no_online:
        online = {}
        for flavour, use_lsd in (("lsd", True), ("edlines", False)):
            det = cs.line_lbd_detect(context=ctx)
            det.use_LSD = use_lsd            # object_slam sets true (main_obj.cpp:365); the class default is EDLines (line_lbd_allclass.cpp:121)
            det.line_length_thres = 15
            lp = det.params()
            # one batch alone, profiled: latency of a batch and its stage times
            ctx.set_profiling(1 | dbg_flags)
            ctx.upload_online(wl["imgs"], wl["Ts"], wl["boxes"], lp, params)
            for _ in range(2):
                ctx.run()
            st_on = ctx.stats()
            stage_on = ctx.stage_ms()
            # throughput: the same `--inflight` contexts as above, round-robin (the sequential half of a detector is one warp per frame)
            for cx in ctxs[1:]:
                cx.upload_online(wl["imgs"], wl["Ts"], wl["boxes"], lp, params)
            for cx in ctxs:
                cx.set_profiling(dbg_flags)
                cx.run_async()
            torch.cuda.synchronize()
            n_on = 2 * len(ctxs)
            o0 = torch.cuda.Event(enable_timing=True)
            o_end = [torch.cuda.Event(enable_timing=True) for _ in ctxs]
            o0.record(streams[0])
            for st_ in streams[1:]:
                st_.wait_event(o0)
            for i in range(n_on):
                ctxs[i % len(ctxs)].run_async()
            for e, st_ in zip(o_end, streams):
                e.record(st_)
            torch.cuda.synchronize()
            on_ms = max(o0.elapsed_time(e) for e in o_end) / n_on
            online[flavour] = {"workload": "same frames, lines from cs_detect_lines (%s, length > 15) on the device" % flavour, "ms_per_step": on_ms,
                               "frames_per_s": F / (on_ms * 1e-3), "value": st_on["n_valid"] / (on_ms * 1e-3), "unit": "proposals/s",
                               "n_valid": st_on["n_valid"], "batches_in_flight": len(ctxs), "one_batch_alone_ms": stage_on.get("total"),
                               "stage_ms": stage_on}
        online["note"] = "the sequential half of either detector (LSD seed loop / EDLines routing + fitting) is one warp per frame; at 256 frames it is latency-bound"
        for cx in ctxs:
            cx.upload(wl["imgs"], wl["Ts"], wl["boxes"], wl["lines"], params)

    # ---- end to end through the host-buffer ABI call ("e2e")
    pinned = torch.from_numpy(wl["imgs"]).pin_memory()
    imgs_pinned = pinned.numpy()
    out = np.zeros((max(int(stats["n_objects"]), 1), topk), cs.CUBOID_DTYPE)
    counts = np.zeros(max(int(stats["n_objects"]), 1), np.int32)

    # two host threads, one context each, call the synchronous ABI entry point: the copies of one batch overlap the kernels of the other
    # (single GPU only: two threads issuing NCCL calls on two communicators in an unordered way could deadlock across ranks)
    e2e_ctxs = ctxs[:2] if (len(ctxs) >= 2 and world == 1) else ctxs[:1]
    e2e_out = [(out, counts)] + [(np.zeros_like(out), np.zeros_like(counts)) for _ in e2e_ctxs[1:]]

    def step_e2e(k=0):
        cx = e2e_ctxs[k]
        cx.detect_batch_host(imgs_pinned, wl["Ts"], wl["boxes"], wl["lines"], params, e2e_out[k][0], e2e_out[k][1])
        if world > 1:
            cx.check(cx.L.cs_allgather_topk(cx.h, recs_per_rank, C.byref(C.c_void_p())))

    for k in range(len(e2e_ctxs)):
        for _ in range(2):
            step_e2e(k)
    barrier()
    e2e_steps = max(4, min(args.steps, 50))
    e2e_steps -= e2e_steps % len(e2e_ctxs)
    import threading

    def e2e_worker(k):
        torch.cuda.set_device(local_rank)
        for _ in range(e2e_steps // len(e2e_ctxs)):
            step_e2e(k)

    t0 = time.perf_counter()
    workers = [threading.Thread(target=e2e_worker, args=(k,)) for k in range(len(e2e_ctxs))]
    for t in workers:
        t.start()
    for t in workers:
        t.join()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3  # host-synchronous calls: the wall clock around all of them is the honest bound
    e2e_t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_ms_step = float(e2e_t.item()) / e2e_steps
    h2d = wl["imgs"].nbytes + sum(l.nbytes for l in wl["lines"]) + wl["Ts"].nbytes + sum(b.nbytes for b in wl["boxes"])
    d2h = out.nbytes + counts.nbytes

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    hbm_peak, peak_src = measured_peaks()
    shp = {"frame_px": w * h}
    kernel_ms = {k: v for k, v in stage_acc.items() if k not in ("total", "lsd")}
    dom = max(kernel_ms, key=kernel_ms.get)
    dom_bytes = float(STAGE_BYTES[dom](stats, shp))
    achieved = dom_bytes / (kernel_ms[dom] * 1e-3) / 1e9 if kernel_ms[dom] > 0 else 0.0
    path_bytes = float(path_alg_bytes(stats, shp, topk))
    traffic = None  # DRAM bytes of the stage's kernels per launch, from the committed ncu capture of this workload (profiles/)
    try:
        tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_traffic.json")))
        traffic = tj.get(args.workload, {}).get(dom)
    except (OSError, ValueError):
        pass
    LIMITERS = {"canny": "integer ALU pipe 70 % active, issue slots 73 % busy (ncu, profiles/r1_g_canny_nms_ncu_full.md): not an HBM-bound kernel",
                "dt": "dependency chain of H row steps per ROI (latency), DRAM traffic below the algorithmic bytes"}
    STAGE_KERNELS = {"dt": "k_dt_bi<N> (one launch per ROI width class)", "canny": "k_canny_nms", "hyst": "k_canny_hyst", "gray": "k_bgr2gray_flat",
                     "sweep": "k_sweep_warp", "fuse": "k_fuse_warp", "lines": "k_roi_lines", "lsd": "line detector kernels"}
    roofline = {"bound": "hbm", "kernel": "%s: %s" % (dom, STAGE_KERNELS.get(dom, dom)), "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "limiter": LIMITERS.get(dom), "peak_source": peak_src, "alg_bytes_per_launch": dom_bytes,
                "path": {"alg_bytes_per_step": path_bytes, "achieved": path_bytes / (ms_per_step * 1e-3) / 1e9,
                         "frac": path_bytes / (ms_per_step * 1e-3) / 1e9 / hbm_peak},
                "stage_ms": stage_acc}

    # ---- CPU baseline (rank 0, N=1 only), bounded sample
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        n = cpu_sample_size(wl, 3.0, cores)
        dt, v, c = cpu_run(wl, list(range(n)), cores)
        dt1, v1, _ = cpu_run(wl, list(range(min(n, 8))), 1)
        cpu = {"value": v / dt, "unit": "proposals/s", "cores": cores, "kind": "port", "frames_per_s": n / dt,
               "one_core_value": v1 / dt1, "sample": "%d of %d frames of this workload, one frame per task, %d threads" % (n, F, cores)}

    line = {
        "metric": "scored cuboid proposals/s", "value": value, "unit": "proposals/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "frames_per_s": n_frames_all / (ms_per_step * 1e-3), "candidates_per_s": n_cand_all / (ms_per_step * 1e-3),
        "config": {"workload": wl["desc"], "frames_per_gpu": F, "boxes_total": n_obj_all, "valid_fraction": n_valid_all / max(n_cand_all, 1),
                   "lines": "synthetic segments given as input (detect_cuboid entry point)", "l2": "inputs larger than L2 (%.0f MB of frames per GPU)" % (wl["imgs"].nbytes / 1e6),
                   "parallelism": "frames sharded x%d, one NCCL all-gather of top-K" % world if world > 1 else "single GPU",
                   "batches_in_flight": len(ctxs)},
        "e2e": {"value": n_valid_all / (e2e_ms_step * 1e-3), "unit": "proposals/s", "frames_per_s": n_frames_all / (e2e_ms_step * 1e-3),
                "ms_per_step": e2e_ms_step, "steps": e2e_steps, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "host_threads": len(e2e_ctxs), "timer": "wall clock around all calls (cs_detect_cuboids_batch is synchronous)"},
        "gpu_launches": int(stats["n_kernel_launches"]) * args.steps,
        "roofline": roofline, "cpu_baseline": cpu, "online": online, "clocks": sampler.summary(),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-online", action="store_true")
    ap.add_argument("--no-prio", action="store_true", help="A/B: keep each batch's whole chain on one stream (no high-priority tail)")
    ap.add_argument("--raster-dt", action="store_true", help="A/B: two-pass raster-scan distance transform kernel instead of the cone form")
    ap.add_argument("--inflight", type=int, default=4, help="batches in flight on one GPU (contexts driven round-robin)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 1000 if args.impl == "ours" else 5
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # convenience: python bench.py --gpus N re-launches itself under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
