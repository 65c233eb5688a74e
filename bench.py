#!/usr/bin/env python
"""bench.py -- headline benchmark of the SVGF + 1-spp path-trace hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3|c4|c5]

A "step" is one frame: pathtrace(pbo, frame) = 1-spp path trace -> temporal accumulation -> N-level a-trous ->
PBO pack, on synthetic input (the reference's own scene description, random-free; RNG seeded by pixel/frame).
N = 1 runs C2 (cornell 1920x1080, 5 a-trous levels), the configuration BASELINE.json's metric is quoted on;
N > 1 runs C4 (cornell 3840x2160) with the frame sharded by row strips over the N GPUs (strong scaling). `value` is
Mpixels/sec (BASELINE.json: "frames/sec & Mpixels/sec") so that the two resolutions share a unit; `fps` is alongside.
Prints ONE JSON line (rank 0). Keys follow the driver's contract; see DESIGN.md "Measurement".

  value     frames/s with everything resident on the device (no host image requested), CUDA events on the
            library's own stream around exactly K frames, after W warm-up frames.
  e2e       frames/s through the public C-ABI calls a user makes with HOST buffers: per frame the camera+parameter structs
            go in (kernel arguments) and the W*H*3 float image comes back to (page-locked) host memory. `value` uses the
            pipelined pair svgf_render_async / svgf_wait_image (frame N's image is consumed while frame N+1 renders: one
            frame in flight, every image waited for inside the timed region); `blocking` is the reference-shaped call
            svgf_render(..., host_image), which returns with the image in place (pathtrace.cu:450).
  roofline  the a-trous level kernel(s): algorithmic bytes (56 B/pixel, 68 B/pixel on the last level) / CUDA-event
            duration of each level launch (library event ring on the library's stream, a second pass over K frames so that
            the event records do not sit inside the `value` region), against the measured HBM copy peak.
  cpu_baseline  the CPU oracle (port of the reference path, OpenMP) timed on this box's host cores, bounded sample.

--impl reference runs the reference's OWN src/pathtrace.cu + src/denoise.cu (compiled from /root/reference into
oracle/_ref/libref_gpu.so, unmodified apart from the portability patch) on the same GPU, same scene, same settings.
The reference has no CPU implementation of this path (it is a CUDA program); if its binary did not travel, the arm
falls back to the oracle port on the host cores and says so.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "cuda-path-tracer-denoising_b200"

WORKLOADS = {   # BASELINE.json configs
    "c1": dict(scene="cornell", W=256, H=256, nlevel=3, moving=False, name="C1 cornell.txt 256x256, 1spp, 3 a-trous iters"),
    "c2": dict(scene="cornell", W=1920, H=1080, nlevel=5, moving=False, name="C2 cornell.txt 1920x1080, 1spp, 5 a-trous iters"),
    "c3": dict(scene="room", W=1920, H=1080, nlevel=5, moving=False, name="C3 room.txt 1920x1080, 1spp, 5 a-trous iters"),
    "c4": dict(scene="cornell", W=3840, H=2160, nlevel=5, moving=False, name="C4 cornell.txt 3840x2160, 1spp, 5 a-trous iters"),
    "c5": dict(scene="bunny", W=1920, H=1080, nlevel=5, moving=True, name="C5 bunny.txt moving camera 1920x1080, 1spp, 5 a-trous iters"),
    # north_star: "cornell.txt and room.txt at 720p/1080p/4K"
    "cornell720": dict(scene="cornell", W=1280, H=720, nlevel=5, moving=False, name="cornell.txt 1280x720, 1spp, 5 a-trous iters"),
    "room720": dict(scene="room", W=1280, H=720, nlevel=5, moving=False, name="room.txt 1280x720, 1spp, 5 a-trous iters"),
    "room4k": dict(scene="room", W=3840, H=2160, nlevel=5, moving=False, name="room.txt 3840x2160, 1spp, 5 a-trous iters"),
}
C5_SPEEDS = dict(camera_speed_x=0.05, camera_speed_y=0.02, camera_speed_z=0.02, camera_speed_theta=0.02, camera_speed_phi=0.05)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(wl, frames=2):
    """Oracle port on the host cores, bounded: `frames` whole frames from a reset, then the a-trous levels on their own
    (north_star: "a CPU a-trous baseline timed on the box's host cores (core count stated)") on the planes those frames left."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    sc = orc.Scene(wl["scene"]); o = orc.Oracle(sc, wl["W"], wl["H"]); P = orc.default_params(atrous_nlevel=wl["nlevel"])
    drv = orc.CameraDriver(sc, wl["W"], wl["H"], automate=wl["moving"])
    threads = orc.lib().orc_max_threads()
    t0 = time.perf_counter()
    for f in range(frames):
        o.frame(drv.step(), P, f, orc.VAR_JACOBI, threads)
    dt = time.perf_counter() - t0
    color, var, g = o.fetch("color_acc"), o.fetch("variance"), o.fetch("gbuffer")
    lv_ms = []
    for level in range(1, wl["nlevel"] + 1):
        t1 = time.perf_counter()
        color, var = orc.atrous_level(color, var, g, level, level == wl["nlevel"], P, orc.VAR_JACOBI, threads)
        lv_ms.append((time.perf_counter() - t1) * 1e3)
    px = wl["W"] * wl["H"]
    return {"value": frames / dt * px / 1e6, "unit": "Mpixels/sec", "fps": frames / dt, "cores": threads, "kind": "port",
            "atrous_ms_per_level": lv_ms,
            "atrous_gbs_per_level": [px * (68 if l == wl["nlevel"] - 1 else 56) / (t * 1e-3) / 1e9 for l, t in enumerate(lv_ms)],
            "sample": "%d frames of %s from reset, oracle/svgf_oracle.cpp (OpenMP, %d threads), %.1f s; then each a-trous level once on "
                      "the accumulated planes of the last frame (ATrousFilter port, same threads, %.0f ms)" % (frames, wl["name"], threads, dt, sum(lv_ms))}


def kernel_source_hash():
    """Hash of the a-trous kernel sources: profiles/atrous_traffic.json records the hash it was captured with."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, PKG, "csrc")
    for f in sorted(os.listdir(d)):
        if f.startswith("atrous") and (f.endswith(".cu") or f.endswith(".h") or f.endswith(".cuh")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refh
    px = wl["W"] * wl["H"]
    line = {"impl": "reference", "metric": "Mpixels/sec", "unit": "Mpixels/sec", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "scene": wl["scene"], "width": wl["W"], "height": wl["H"], "atrous_levels": wl["nlevel"]}}
    try:        # the reference's own error handling is print + exit(EXIT_FAILURE) (pathtrace.cu:25-43): never enter it without a device
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if refh.available("gpu") and have_gpu:
        try:
            h = refh.RefHarness("gpu")
            h.load_blob(wl["scene"], wl["W"], wl["H"])
            h.set_params(**refh.ALL_ON); h.set_params(atrous_nlevel=wl["nlevel"])
            if wl["moving"]:
                h.set_params(automate_camera=1, **C5_SPEEDS)
            h.time_frames(max(args.warmup, 1))
            ms = h.time_frames(args.steps)
            fps = 1000.0 * args.steps / ms
            line.update({"value": fps * px / 1e6, "fps": fps, "ms_per_step": ms / args.steps,
                         "cpu_baseline": {"value": fps * px / 1e6, "unit": "Mpixels/sec", "cores": 0, "kind": "reference",
                                          "sample": "the reference's own CUDA path (src/pathtrace.cu + src/denoise.cu built for sm_100, "
                                                    "oracle/_ref/libref_gpu.so) on 1 GPU: it has no CPU implementation; %d frames, host clock "
                                                    "(every reference frame ends in a blocking D2H)" % args.steps},
                         "e2e": {"value": fps * px / 1e6, "unit": "Mpixels/sec", "fps": fps, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": px * 12},
                         "config": dict(line["config"], parallelism="1 GPU (the reference is single-GPU)", device="gpu")})
            print(json.dumps(line), flush=True)
            return
        except Exception as e:      # fall through to the CPU port
            line["note"] = "reference GPU binary failed: %s" % e
    frames = max(1, min(args.steps, 2))
    cb = cpu_baseline(wl, frames)
    line.update({"value": cb["value"], "fps": cb["fps"], "ms_per_step": 1000.0 / cb["fps"], "cpu_baseline": cb,
                 "e2e": {"value": cb["value"], "unit": "Mpixels/sec", "fps": cb["fps"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "config": dict(line["config"], parallelism="host cores", device="cpu (oracle port: no CUDA device, or the reference binary did not travel)")})
    print(json.dumps(line), flush=True)


def run_ours(args, wl, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    m = importlib.import_module(PKG)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W, H, nl = wl["W"], wl["H"], wl["nlevel"]
    blob, R = m.open_scene(wl["scene"], W, H, device=local_rank)
    rows = [0, H]
    P = m.default_params(atrous_nlevel=nl)
    drv = blob.camera_driver(W, H, automate=wl["moving"])
    if world > 1:       # shard the frame by row strips; other strips' rows are read in place over NVLink (CUDA IPC)
        rows = m.connect_ranks(R, dist, rank, world)
        # Equal-height strips are not equal-time strips (path-tracing cost follows the geometry a row sees): measure each
        # rank's own device time per frame (waits excluded), move the boundaries, restart the history. Untimed set-up.
        for _ in range(3):
            R.set_profiling(True)
            for f in range(4):
                R.pathtrace(drv.cam if not drv.first else drv.step(), P, f)
            st = R.stage_times(); R.set_profiling(False)
            # per-level event intervals include the cross-rank waits, so the (uniform per pixel) denoise cost per row is
            # taken from the rank that waited least; path trace + temporal intervals contain no waits
            my_rows = max(1, rows[rank + 1] - rows[rank])
            mine = torch.tensor([float(st[0] + st[1]), float(sum(st[2:9]) + st[9]) / my_rows], device="cuda", dtype=torch.float64)
            allc = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allc, mine)
            kappa = min(float(t[1].item()) for t in allc)
            cost = [float(allc[r][0].item()) + kappa * (rows[r + 1] - rows[r]) for r in range(world)]
            rows = m.balanced_partition(rows, cost)
            dist.barrier(); R.sync(); R.reset()
            rows = m.connect_ranks(R, dist, rank, world, rows)
    stream = torch.cuda.ExternalStream(R.stream(), device=torch.device("cuda", local_rank))
    host = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
    host_np = host.numpy()
    host2 = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
    bufs = [host_np, host2.numpy()]
    frame = 0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank); clocks.start()       # from before the warm-up, so that short runs get samples too
    for _ in range(max(args.warmup, 3)):
        R.pathtrace(drv.step(), P, frame, host_image=host_np); frame += 1
    for i in range(3):      # sets up the copy stream and the second output buffer of the pipelined path
        R.pathtrace_async(drv.step(), P, frame, bufs[i & 1]); frame += 1
    R.wait_image(None)

    # ---- device-resident throughput (value): K frames back to back, CUDA events on the library's stream ----
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        R.pathtrace(drv.step(), P, frame); frame += 1
    e1.record(stream)
    R.sync(); barrier()
    ms_dev = e0.elapsed_time(e1)
    # ---- per-stage events (roofline): the same K frames again with the library's event ring on (12 records per frame) ----
    R.set_profiling(True)
    barrier()
    for _ in range(args.steps):
        R.pathtrace(drv.step(), P, frame); frame += 1
    R.sync(); barrier()
    stage = R.stage_times()
    R.set_profiling(False)
    # ---- end to end through the C ABI with host buffers ----
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    for _ in range(args.steps):
        R.pathtrace(drv.step(), P, frame, host_image=host_np); frame += 1
    e3.record(stream)
    R.sync(); barrier()
    ms_blk = max(e2.elapsed_time(e3), (time.perf_counter() - t0) * 1e3)
    # ---- end to end, pipelined: the image of frame N is waited for (= consumed) while frame N+1 renders ----
    barrier()
    t0 = time.perf_counter()
    t_cam = t_submit = t_wait = 0.0                 # where the host's share of the loop goes (reported, not subtracted)
    for i in range(args.steps):
        ta = time.perf_counter()
        cam_i = drv.step()
        tb = time.perf_counter()
        R.pathtrace_async(cam_i, P, frame, bufs[i & 1]); frame += 1
        tc = time.perf_counter()
        if i >= 1:
            R.wait_image(bufs[(i - 1) & 1])
        td = time.perf_counter()
        t_cam += tb - ta; t_submit += tc - tb; t_wait += td - tc
    R.wait_image(bufs[(args.steps - 1) & 1])
    R.sync()
    ms_e2e = (time.perf_counter() - t0) * 1e3      # host clock: the last image has landed
    host_split = {"camera_ms": t_cam * 1e3 / args.steps, "submit_ms": t_submit * 1e3 / args.steps, "wait_ms": t_wait * 1e3 / args.steps}
    barrier()
    clk = clocks.stop()
    if world > 1 and R.peer_error():
        raise SystemExit("bench.py: a cross-rank wait timed out (a peer stopped making progress)")
    # ---- N > 1: the same workload UNSHARDED on one GPU in the same run (rank 0), so that the strong-scaling factor is a
    # same-run, same-resolution number ----
    anchor = None
    if world > 1:
        if rank == 0:
            _, R1 = m.open_scene(wl["scene"], W, H, device=local_rank)
            d1 = blob.camera_driver(W, H, automate=wl["moving"])
            s1 = torch.cuda.ExternalStream(R1.stream(), device=torch.device("cuda", local_rank))
            f1 = 0
            for _ in range(max(args.warmup, 3)):
                R1.pathtrace(d1.step(), P, f1); f1 += 1
            R1.sync()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(s1)
            for _ in range(args.steps):
                R1.pathtrace(d1.step(), P, f1); f1 += 1
            a1.record(s1)
            R1.sync()
            ms1 = a0.elapsed_time(a1)
            anchor = {"value": args.steps * 1000.0 / ms1 * W * H / 1e6, "fps": args.steps * 1000.0 / ms1, "ms_per_step": ms1 / args.steps,
                      "what": "the same workload unsharded on ONE GPU (rank 0's), same run, device-resident"}
            R1.close()
        barrier()
    stages_per_rank = None
    if world > 1:
        t = torch.tensor([ms_dev, ms_e2e, ms_blk], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e, ms_blk = t.tolist()
        # every rank's own stage intervals (the per-level ones include its waits for the neighbours): where the frame goes
        mine = torch.tensor([float(v) for v in stage[:11]], device="cuda", dtype=torch.float64)
        alls = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(alls, mine)
        stages_per_rank = [{"rows": rows[r + 1] - rows[r], "pathtrace": round(a[0], 4), "temporal": round(a[1], 4), "atrous": [round(v, 4) for v in a[2:2 + nl]],
                            "pbo_pack": round(a[9], 4), "frame": round(a[10], 4)} for r, a in enumerate(x.tolist() for x in alls)]
    if rank != 0:
        return
    px = W * H
    # strong scaling: all ranks together render ONE frame per step (row strips)
    fps_dev = args.steps * 1000.0 / ms_dev
    fps_e2e = args.steps * 1000.0 / ms_e2e
    fps_blk = args.steps * 1000.0 / ms_blk
    peak, peak_src = measured_peak()
    lv_ms = [float(stage[2 + l]) for l in range(nl)]
    strip_px = W * (rows[rank + 1] - rows[rank]) if world > 1 else px       # rank 0's launches cover its strip
    lv_bytes = [strip_px * (68 if l == nl - 1 else 56) for l in range(nl)]
    lv_gbs = [b / (t * 1e-3) / 1e9 if t > 0 else 0.0 for b, t in zip(lv_bytes, lv_ms)]
    tot_ms = sum(lv_ms)
    # the a-trous stage as ONE launch (atrous_stage_kernel, the default): the levels overlap inside it and have no intervals of
    # their own; the library records every level's event after the launch, so the first interval is the whole stage
    one_launch = nl > 1 and lv_ms[0] > 0 and all(t < 1e-3 for t in lv_ms[1:])
    achieved = sum(lv_bytes) / (tot_ms * 1e-3) / 1e9 if tot_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "atrous_traffic.json")
    if os.path.exists(tp) and world == 1 and (W, H) == (1920, 1080):     # the capture is of C2; other workloads report null
        tj = json.load(open(tp))
        if tj.get("kernel_source_hash") == kernel_source_hash():            # a capture of other kernel code is stale: null
            traffic = tj.get("dram_bytes_per_launch")
    cb = cpu_baseline(wl, 2) if world == 1 and not args.no_cpu_baseline else None
    launches_per_frame = 1 + 1 + (1 if one_launch else 2 * nl) + 1 + (0 if world == 1 else 3)     # rt, temporal, a-trous stage (or (kl + tiled) x levels), pack [+ G-buffer halo push, frame wait/signal]
    line = {
        "metric": "Mpixels/sec", "value": fps_dev * px / 1e6, "unit": "Mpixels/sec", "fps": fps_dev, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "scene": wl["scene"], "width": W, "height": H, "atrous_levels": nl,
                   "parallelism": "1 GPU" if world == 1 else "%d row strips (cost-balanced, rows %s), peer reads over NVLink (CUDA IPC), no collective on the data path" % (world, rows),
                   "l2": "per-frame working set %.0f MB > 126 MB L2 (no flush needed)" % (px * 196 / 1e6)},
        "e2e": {"value": fps_e2e * px / 1e6, "unit": "Mpixels/sec", "fps": fps_e2e, "h2d_bytes_per_step": (84 + 80) * world,
                "d2h_bytes_per_step": px * 12, "ms_per_step": ms_e2e / args.steps,
                "api": "svgf_render_async + svgf_wait_image (one frame in flight, every image waited for; host clock)",
                "host_ms_per_step": host_split,
                "blocking": {"value": fps_blk * px / 1e6, "fps": fps_blk, "ms_per_step": ms_blk / args.steps,
                             "api": "svgf_render(..., host_image): returns with the image in place, like the reference's pathtrace()"},
                "note": "every rank copies its own strip of the image to its host buffer each frame" if world > 1 else "whole image to host each frame"},
        "e2e_blocking": {"value": fps_blk * px / 1e6, "unit": "Mpixels/sec", "fps": fps_blk,
                         "note": "like for like with the reference arm, whose pathtrace() blocks on its D2H every frame (e2e.value is the pipelined API)"},
        "gpu_launches": launches_per_frame * args.steps * 4 * world,
        "roofline": {"bound": "hbm", "kernel": "atrous level (all %d levels)" % nl, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "per_level_us": None if one_launch else [t * 1e3 for t in lv_ms], "per_level_gbs": None if one_launch else lv_gbs,
                     "per_level_frac": None if one_launch else [g / peak for g in lv_gbs],
                     "stage_us": tot_ms * 1e3, "launches_per_stage": 1 if one_launch else 2 * nl,
                     "algorithmic_bytes_per_pixel": [68 if l == nl - 1 else 56 for l in range(nl)]},
        "stages_ms": {"pathtrace": float(stage[0]), "temporal": float(stage[1]), "atrous": lv_ms, "pbo_pack": float(stage[9]), "frame": float(stage[10])},
        "clocks": clk,
    }
    if cb:
        line["cpu_baseline"] = cb
    if stages_per_rank:
        line["stages_ms_per_rank"] = stages_per_rank
    if anchor:
        line["anchor_1gpu"] = anchor
        line["speedup_vs_1gpu_same_run"] = line["value"] / anchor["value"]
    print(json.dumps(line), flush=True)


def run_verify(args, wl, rank, world, local_rank):
    """--verify: parity of the REAL multi-GPU path (one process per GPU, CUDA-IPC mappings, cross-device flags, dual stores
    over NVLink, an uneven cost-style partition): every rank renders the frames sharded AND, on its own GPU, the same frames
    unsharded, and compares its strip of every buffer bit for bit. Prints one JSON line; exit code 1 on any mismatch."""
    import numpy as np
    import torch
    import torch.distributed as dist
    m = importlib.import_module(PKG)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W, H, nl = wl["W"], wl["H"], wl["nlevel"]
    blob, R = m.open_scene(wl["scene"], W, H, device=local_rank)
    _, F = m.open_scene(wl["scene"], W, H, device=local_rank)
    P = m.default_params(atrous_nlevel=nl)
    rows = [0, H]
    if world > 1:       # deliberately uneven strips, like the cost-balanced partitions of the bench
        cost = [1.0 + 0.35 * ((r * 5) % 3) for r in range(world)]
        rows = m.balanced_partition(m.row_partition(H, world), cost)
        m.connect_ranks(R, dist, rank, world, rows)
    drv = blob.camera_driver(W, H, automate=wl["moving"])
    frames = max(2, min(args.steps, 8))
    keys = ["image", "gbuffer", "history_length", "moment_acc", "color_history", "variance", "denoised", "pbo"]
    bad = []
    for f in range(frames):
        cam = m.Camera.from_array(drv.step().as_array())
        R.pathtrace(cam, P, f); F.pathtrace(cam, P, f)
        if f in (0, frames // 2, frames - 1):
            R.sync(); F.sync()
            for k in keys:
                a, b = R.fetch(k)[rows[rank]:rows[rank + 1]], F.fetch(k)[rows[rank]:rows[rank + 1]]
                if not np.array_equal(a.view(np.uint8), b.view(np.uint8)):
                    bad.append("frame %d %s: %d differing bytes in rank %d's rows [%d, %d)" % (f, k, int((a.view(np.uint8) != b.view(np.uint8)).sum()), rank, rows[rank], rows[rank + 1]))
            if world > 1:
                dist.barrier()      # nobody races ahead into frames whose inputs a fetching rank still reads
    R.sync()
    perr = R.peer_error() if world > 1 else 0
    nbad = torch.tensor([len(bad) + (1 if perr else 0)], device="cuda", dtype=torch.int64)
    if world > 1:
        dist.all_reduce(nbad)
    for b in bad[:8]:
        print("rank %d: %s" % (rank, b), file=sys.stderr, flush=True)
    if rank == 0:
        print(json.dumps({"verify": "ok" if int(nbad.item()) == 0 else "MISMATCH", "n_gpus": world, "workload": wl["name"], "frames": frames,
                          "rows": rows, "buffers": keys, "mismatching_comparisons": int(nbad.item()),
                          "what": "every rank's strip of every buffer, sharded vs unsharded on the same GPU, bit for bit"}), flush=True)
    if world > 1:
        dist.barrier()
    sys.exit(0 if int(nbad.item()) == 0 else 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verify", action="store_true", help="multi-GPU parity: sharded == unsharded, bit for bit, on the real transport")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload or ("c2" if args.gpus == 1 else "c4")]
    if args.verify:
        run_verify(args, wl, rank, world, local_rank)
    elif args.impl == "reference":
        run_reference(args, wl, rank, world)
    else:
        run_ours(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
