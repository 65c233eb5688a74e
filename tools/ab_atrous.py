#!/usr/bin/env python
"""A/B harness for the a-trous tile shapes (csrc/atrous.cu, g_at_shapes) and TMA descriptor options: one process, one
context per configuration (the SVGF_* switches are read at svgf_create), per-level CUDA-event times from the library's own
profiling ring over `--frames` frames of a BASELINE.json workload. Prints one JSON line per configuration.

    python tools/ab_atrous.py [--workload c2] [--frames 30] [--shapes 0,1,2,...] [--promo 0,2]
"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS, C5_SPEEDS  # noqa: E402


STRIP = None


def run(m, wl, frames, env):
    for k in ("SVGF_ATROUS_SHAPE", "SVGF_ATROUS_SHAPES", "SVGF_TMA_L2PROMO", "SVGF_ATROUS_VARIANT", "SVGF_ATROUS_BANDS", "SVGF_RT_ANYHIT", "SVGF_RT_MINBLOCKS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    W, H, nl = wl["W"], wl["H"], wl["nlevel"]
    blob, R = m.open_scene(wl["scene"], W, H)
    if STRIP:       # one rank's share of a sharded frame (taps outside the strip read this context's own planes)
        R.set_shard(0, 1, STRIP[0], STRIP[1])
    P = m.default_params(atrous_nlevel=nl)
    drv = blob.camera_driver(W, H, automate=wl["moving"])
    f = 0
    for _ in range(5):
        R.pathtrace(drv.step(), P, f); f += 1
    R.set_profiling(True)
    for _ in range(frames):
        R.pathtrace(drv.step(), P, f); f += 1
    st = R.stage_times()
    R.set_profiling(False)
    R.close()
    lv = [round(float(st[2 + l]) * 1e3, 2) for l in range(nl)]
    return {"env": env, "workload": wl["name"], "strip": STRIP, "rt_us": round(float(st[0]) * 1e3, 1), "temporal_us": round(float(st[1]) * 1e3, 1),
            "level_us": lv, "atrous_us": round(sum(lv), 1), "frame_us": round(float(st[10]) * 1e3, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--shapes", default="0,1,2,3,4,5,6,7")
    ap.add_argument("--promo", default="")
    ap.add_argument("--extra", default="", help="semicolon-separated KEY=VAL+KEY=VAL configurations")
    ap.add_argument("--strip", default="", help="row_begin,row_end: render only this strip (emulates one rank of a sharded frame)")
    a = ap.parse_args()
    global STRIP
    if a.strip:
        STRIP = tuple(int(v) for v in a.strip.split(","))
    m = importlib.import_module("cuda-path-tracer-denoising_b200")
    wl = WORKLOADS[a.workload]
    cfgs = [{}]
    cfgs += [{"SVGF_ATROUS_SHAPE": s} for s in a.shapes.split(",") if s != ""]
    for q in a.promo.split(","):
        if q != "":
            cfgs.append({"SVGF_TMA_L2PROMO": q})
            cfgs += [{"SVGF_TMA_L2PROMO": q, "SVGF_ATROUS_SHAPE": s} for s in a.shapes.split(",") if s != ""]
    for e in a.extra.split(";"):
        if e:
            cfgs.append(dict(kv.split("=") for kv in e.split("+")))
    for env in cfgs:
        try:
            print(json.dumps(run(m, wl, a.frames, env)), flush=True)
        except Exception as ex:      # a configuration that fails must not hide the others
            print(json.dumps({"env": env, "error": str(ex)}), flush=True)


if __name__ == "__main__":
    main()
