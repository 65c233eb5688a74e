#!/usr/bin/env python
"""Where the persistent blocks of the one-launch a-trous stage spend their time (diagnostic build, tools/build_stage_timers.sh;
run with SVGF_LIB_PATH=.../ab/libsvgf_timers.so): clock64 sums of thread 0 of every block, per category, per frame."""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS
m = importlib.import_module("cuda-path-tracer-denoising_b200")
wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
W, H, nl = wl["W"], wl["H"], wl["nlevel"]
blob, R = m.open_scene(wl["scene"], W, H)
P = m.default_params(atrous_nlevel=nl); drv = blob.camera_driver(W, H)
for f in range(5): R.pathtrace(drv.step(), P, f)
t0 = R.fetch_raw("stage_timers", 8, np.uint64).astype(np.float64)
N = 20
for f in range(N): R.pathtrace(drv.step(), P, 5 + f)
t1 = R.fetch_raw("stage_timers", 8, np.uint64).astype(np.float64)
d = (t1 - t0) / N
names = ["claim+decode", "dependency wait", "K item body", "tile load (issue..landed+barrier)", "tile compute+stores", "end barrier", "signal (fence+atomics)", "-"]
tot = d.sum()
print(json.dumps({"workload": wl["name"], "clock_sums_per_frame_over_740_blocks": {n: round(v) for n, v in zip(names, d)},
                  "share": {n: round(v / tot, 3) for n, v in zip(names, d)}, "us_per_block_at_1.965GHz": round(tot / 740 / 1965, 1)}))
