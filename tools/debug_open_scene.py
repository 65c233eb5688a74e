"""One-shot diagnosis: where does the first non-finite value appear when an OPEN scene (rays that miss) is rendered?"""
import importlib, os, struct, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
m = importlib.import_module("cuda-path-tracer-denoising_b200")
import orc

sc = m.SceneFile(os.path.join(ROOT, "tests", "golden", "scenes_txt", "two_meshes.txt"))
tex = (np.indices((8, 8)).sum(0) % 2 * 255).astype(np.uint8)[..., None].repeat(3, 2)
sc.set_texture(0, tex)
a = sc.arrays()
hdr = b"SVGFSCN1" + struct.pack("<6i", a["geoms"].size // 248, a["materials"].size // 56, a["triangles"].size // 136, a["bvh"].size // 40, a["boxes"].shape[0], 1) + struct.pack("<f", sc.fovy) + b"\0" * 4
cam = m.Camera(); cam.position[:] = list(sc.eye); cam.lookAt[:] = list(sc.lookat); cam.up[:] = list(sc.up)
blob = hdr + bytes(cam) + a["geoms"].tobytes() + a["materials"].tobytes() + a["triangles"].tobytes() + a["bvh"].tobytes() + a["boxes"].tobytes() + struct.pack("<3i", 8, 8, 3) + tex.tobytes()
open("/tmp/own.scene", "wb").write(blob)
W, H = 160, 100
osc = orc.Scene("/tmp/own.scene")
for over in ({"denoise_enable": 0}, {"atrous_nlevel": 0}, {"atrous_nlevel": 1}, {"atrous_nlevel": 3}, {"atrous_nlevel": 3, "temporal_enable": 0}):
    R = m.Renderer(sc.desc(W, H), W, H)
    P = m.default_params(**over); OP = orc.default_params(**over)
    O = orc.Oracle(osc, W, H)
    drv = sc.camera_driver(W, H)
    for f in range(3):
        c = drv.step()
        R.pathtrace(c, P, f)
        O.frame(orc.Camera.from_array(c.as_array()), OP, f, orc.VAR_JACOBI, 0)
        for name in ("image", "gbuffer", "variance", "color_history", "moment_history", "denoised"):
            try:
                g = R.fetch(name).astype(np.float64); o = O.fetch(name).astype(np.float64)
            except Exception as e:
                continue
            if name == "gbuffer":
                g = g[..., :12]; o = o[..., :12]
            bad = ~np.isfinite(g.reshape(H, W, -1)).all(axis=2)
            if bad.any():
                ys, xs = np.nonzero(bad)
                y, x = int(ys[0]), int(xs[0])
                print(over, "frame", f, name, "non-finite px:", int(bad.sum()), "first", (y, x), "gpu", g[y, x], "oracle", o[y, x])
                gb = R.fetch("gbuffer")[y, x]; ob = O.fetch("gbuffer")[y, x]
                print("   gbuffer gpu", gb[:12], gb[12:].view(np.int32), " oracle", ob[:12], ob[12:].view(np.int32))
                print("   image gpu", R.fetch("image")[y, x], "oracle", O.fetch("image")[y, x], " variance gpu", R.fetch("variance")[y, x], "oracle", O.fetch("variance")[y, x])
                im = R.fetch("image"); print("   image non-finite anywhere:", int((~np.isfinite(im)).sum()), " max", np.nanmax(im))
                break
        else:
            continue
        break
    else:
        d = R.fetch("denoised"); od = O.fetch("denoised")
        print(over, "all finite; max |gpu-oracle| denoised", float(np.abs(d - od).max()), "geomId agreement", float((R.fetch("gbuffer")[..., 12].view(np.int32) == O.fetch("gbuffer")[..., 12].view(np.int32)).mean()))
    R.close()
