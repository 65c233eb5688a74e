"""Ad-hoc GPU check (not a pytest file): product vs oracle vs reference GPU build, prints error statistics."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
svgf = importlib.import_module("cuda-path-tracer-denoising_b200")
import orc, refh

def stats(name, a, b, tol=1e-4):
    a = a.astype(np.float64); b = b.astype(np.float64)
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-2)
    print("   %-16s max_abs %.3g  max_rel %.3g  frac(rel>%g) %.5f" % (name, np.abs(a - b).max(), rel.max(), tol, (rel > tol).mean()), flush=True)

scene, W, H, nl = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
nframes = int(sys.argv[5]) if len(sys.argv) > 5 else 4
moving = len(sys.argv) > 6
blob, R = svgf.open_scene(scene, W, H)
P = svgf.default_params(atrous_nlevel=nl)
drv = blob.camera_driver(W, H, automate=moving)
osc = orc.Scene(scene); O = orc.Oracle(osc, W, H); OP = orc.default_params(atrous_nlevel=nl)
ref = None
if refh.available("gpu_jacobi"):
    ref = refh.RefHarness("gpu_jacobi"); ref.load_blob(scene, W, H); ref.set_params(**refh.ALL_ON); ref.set_params(atrous_nlevel=nl)
    if moving: ref.set_params(automate_camera=1, camera_speed_x=0.05, camera_speed_y=0.02, camera_speed_z=0.02, camera_speed_theta=0.02, camera_speed_phi=0.05)
host = np.zeros((H, W, 3), np.float32)
for f in range(nframes):
    cam = drv.step()
    R.pathtrace(cam, P, f, host_image=host)
    O.frame(orc.Camera.from_array(cam.as_array()), OP, f, orc.VAR_JACOBI, 0)
    print("frame", f)
    for other, nm in ((O, "oracle"), (ref, "ref_gpu")):
        if other is None: continue
        if nm == "ref_gpu":
            other.frame()
            print("  camera equal:", np.array_equal(other.fetch("camera").view(np.uint32), cam.as_array().view(np.uint32)))
        print("  vs", nm)
        g = R.fetch("gbuffer"); go = other.fetch("gbuffer")
        print("   geomId mismatches:", int((g[..., 12].view(np.int32) != go[..., 12].view(np.int32)).sum()), "of", W * H)
        stats("gbuf.normal", g[..., 0:3], go[..., 0:3]); stats("gbuf.position", g[..., 3:6], go[..., 3:6]); stats("gbuf.albedo", g[..., 6:9], go[..., 6:9])
        stats("image", R.fetch("image"), other.fetch("image"))
        print("   history_length mismatches:", int((R.fetch("history_length") != other.fetch("history_length")).sum()))
        stats("moment_acc", R.fetch("moment_acc"), other.fetch("moment_acc"))
        stats("color_history", R.fetch("color_history"), other.fetch("color_history"))
        stats("variance", R.fetch("variance"), other.fetch("variance"))
        stats("denoised", R.fetch("denoised"), other.fetch("denoised"))
        stats("host_image", host, other.fetch("host_image"))
        print("   pbo mismatches:", int((R.fetch("pbo") != other.fetch("pbo")).sum()))
R.set_profiling(True)
for f in range(nframes, nframes + 3):
    R.pathtrace(drv.step(), P, f, host_image=host)
print("stage ms [rt, temporal, L1..L7, pack, total]:", np.round(R.stage_times(), 4).tolist())
R.set_profiling(False)
t = time.time(); n = 50
for f in range(n): R.pathtrace(drv.step(), P, nframes + 3 + f, host_image=host)
print("e2e ms/frame (with D2H):", (time.time() - t) / n * 1e3)
t = time.time()
for f in range(n): R.pathtrace(drv.step(), P, nframes + 3 + n + f)
R.sync()
print("ms/frame (no D2H):", (time.time() - t) / n * 1e3)
