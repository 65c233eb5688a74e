#!/usr/bin/env python
"""how long svgf_rebuild_bvh takes (the whole call on the host clock, and the device part between two CUDA
events on the library's stream), for the two mesh scenes. One JSON line per scene."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
m = importlib.import_module("cuda-path-tracer-denoising_b200")
for scene in ("bunny", "room"):
    blob, R = m.open_scene(scene, 640, 360)
    n = int(blob.counts["tris"])
    st = torch.cuda.ExternalStream(R.stream())
    R.rebuild_bvh(); R.sync()        # first call: module load, cub temp sizing
    host, dev = [], []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(st)
        R.rebuild_bvh()
        e1.record(st); R.sync(); t1 = time.perf_counter()
        host.append((t1 - t0) * 1e3); dev.append(e0.elapsed_time(e1))
    # refit in place (svgf_refit_bvh: upload of all triangles + boxes bottom-up), stream-ordered: timed to the end of the stream
    tris = np.ascontiguousarray(blob.triangles).view(np.uint8).reshape(n, 136)
    R.refit_bvh(tris); R.sync()
    rf = []
    for _ in range(10):
        t0 = time.perf_counter(); R.refit_bvh(tris); R.sync(); rf.append((time.perf_counter() - t0) * 1e3)
    rf.sort()
    # a frame through the rebuilt tree, so that the number belongs to a tree that renders
    P = m.default_params(); drv = blob.camera_driver(640, 360)
    R.pathtrace(drv.step(), P, 0)
    host.sort(); dev.sort()
    line = {"scene": scene, "triangles": n, "rebuild_ms_host_median": round(host[5], 3), "rebuild_ms_host_min": round(host[0], 3),
            "rebuild_ms_device_span_median": round(dev[5], 3), "what": "svgf_rebuild_bvh: Morton keys, cub radix sort, Karras tree, bottom-up boxes, pre-order emission, triangle reorder; includes its cudaMalloc/cudaFree and the final stream synchronisation"}
    if n: line["mtris_per_s"] = round(n / host[5] / 1e3, 2)
    line["refit_ms_host_median"] = round(rf[5], 3)
    print(json.dumps(line), flush=True)
    R.close()
