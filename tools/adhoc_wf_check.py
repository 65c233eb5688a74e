import os, sys, importlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
m = importlib.import_module("cuda-path-tracer-denoising_b200")
scene, W, H = "cornell", 96, 64
outs = []
for variant in ("mega", "wavefront"):
    os.environ["SVGF_RT_VARIANT"] = variant
    blob, R = m.open_scene(scene, W, H)
    P = m.default_params()
    drv = blob.camera_driver(W, H, automate=False)
    per = []
    for f in range(3):
        R.pathtrace(drv.step(), P, f)
        per.append({k: R.fetch(k) for k in ["image", "gbuffer", "denoised", "history_length"]})
    outs.append(per); R.close()
for f in range(3):
    for k in outs[0][f]:
        a, b = outs[0][f][k], outs[1][f][k]
        ne = (a.view(np.uint32) != b.view(np.uint32))
        print("frame", f, k, "mismatching words:", int(ne.sum()), "max abs diff", float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()))
        if k == "image" and ne.any():
            ys, xs = np.nonzero(ne.any(axis=-1)); print("   first pixels:", list(zip(ys[:6].tolist(), xs[:6].tolist())), a[ys[0], xs[0]], b[ys[0], xs[0]])
