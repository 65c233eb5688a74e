"""Generate golden vectors from the reference's own CUDA build (oracle/_ref/libref_gpu*.so) -- TEST INFRASTRUCTURE.

Run ON THE GPU BOX (gpurun):   python oracle/gen_golden_gpu.py gpurun_out/golden_gpu
then copy the .npz files into tests/golden/ and commit them together with this script.
Each invocation of a variant happens in a fresh subprocess because the reference keeps its state in
file-static variables (one instance per process).

Outputs:
  ref_gpu_jacobi_<scene>_<W>x<H>.npz   per-frame dumps from the race-free (Jacobi-variance) reference build
  ref_gpu_racy_<scene>_<W>x<H>.npz     final-frame denoised/variance from the unmodified reference, two runs
                                       (run-to-run spread of the reference's in-place variance race)
  ref_gpu_timing.json                  reference frame times (ms) per config, for DESIGN.md/BASELINE notes
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

WORKER = r'''
import sys, json
sys.path.insert(0, %(here)r)
import numpy as np, refh
variant, scene, W, H, nlevel, frames, fields, out, moving = json.loads(sys.argv[1])
h = refh.RefHarness(variant)
h.load_blob(scene, W, H)
h.set_params(**refh.ALL_ON)
h.set_params(atrous_nlevel=nlevel)
if moving:
    h.set_params(automate_camera=1, camera_speed_x=0.05, camera_speed_y=0.02, camera_speed_z=0.02,
                 camera_speed_theta=0.02, camera_speed_phi=0.05)
res = {}
for f in range(max(frames) + 1):
    h.frame()
    if f in frames:
        for k in fields:
            res["f%%d_%%s" %% (f, k)] = h.fetch(k)
        res["f%%d_camera" %% f] = h.fetch("camera")
np.savez_compressed(out, **res)
'''

TIMER = r'''
import sys, json
sys.path.insert(0, %(here)r)
import refh
variant, scene, W, H, nlevel, warm, timed, moving = json.loads(sys.argv[1])
h = refh.RefHarness(variant)
h.load_blob(scene, W, H)
h.set_params(**refh.ALL_ON)
h.set_params(atrous_nlevel=nlevel)
if moving:
    h.set_params(automate_camera=1, camera_speed_x=0.05, camera_speed_y=0.02, camera_speed_z=0.02,
                 camera_speed_theta=0.02, camera_speed_phi=0.05)
h.time_frames(warm)
ms = h.time_frames(timed)
print(json.dumps({"variant": variant, "scene": scene, "W": W, "H": H, "nlevel": nlevel, "moving": moving,
                  "ms_per_frame": ms / timed, "fps": 1000.0 * timed / ms}))
'''


def run(code, args):
    r = subprocess.run([sys.executable, "-c", code % {"here": HERE}, json.dumps(args)], capture_output=True, text=True)
    if r.returncode:
        print(r.stdout, r.stderr)
        raise SystemExit("worker failed: %r" % (args,))
    return r.stdout


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden_gpu"
    os.makedirs(out, exist_ok=True)
    full = ["image", "gbuffer", "denoised", "variance", "color_acc", "moment_acc", "history_length", "pbo"]
    jobs = [
        ("gpu_jacobi", "cornell", 64, 64, 3, [0, 1, 4], full, "ref_gpu_jacobi_cornell_64x64.npz", 0),
        ("gpu_jacobi", "cornell", 96, 64, 5, [0, 3], ["image", "denoised", "variance", "history_length"],
         "ref_gpu_jacobi_cornell_96x64.npz", 0),
        ("gpu_jacobi", "room", 64, 64, 5, [0, 2], ["image", "gbuffer", "denoised", "history_length"],
         "ref_gpu_jacobi_room_64x64.npz", 0),
        ("gpu_jacobi", "bunny", 64, 64, 5, [0, 3], ["image", "gbuffer", "denoised", "history_length"],
         "ref_gpu_jacobi_bunny_moving_64x64.npz", 1),
        ("gpu", "cornell", 64, 64, 3, [4], ["denoised", "variance"], "ref_gpu_racy_cornell_64x64_run0.npz", 0),
        ("gpu", "cornell", 64, 64, 3, [4], ["denoised", "variance"], "ref_gpu_racy_cornell_64x64_run1.npz", 0),
        ("gpu", "cornell", 256, 256, 3, [8], ["denoised", "variance"], "ref_gpu_racy_cornell_256x256_run0.npz", 0),
        ("gpu", "cornell", 256, 256, 3, [8], ["denoised", "variance"], "ref_gpu_racy_cornell_256x256_run1.npz", 0),
        ("gpu_jacobi", "cornell", 256, 256, 3, [8], ["denoised", "variance"], "ref_gpu_jacobi_cornell_256x256_f8.npz", 0),
    ]
    for v, s, W, H, nl, frames, fields, name, moving in jobs:
        run(WORKER, [v, s, W, H, nl, frames, fields, os.path.join(out, name), moving])
        print("wrote", name, flush=True)
    timing = []
    for v, s, W, H, nl, moving in [("gpu", "cornell", 256, 256, 3, 0), ("gpu", "cornell", 1920, 1080, 5, 0),
                                   ("gpu", "room", 1920, 1080, 5, 0), ("gpu", "cornell", 3840, 2160, 5, 0),
                                   ("gpu", "bunny", 1920, 1080, 5, 1)]:
        line = run(TIMER, [v, s, W, H, nl, 10, 50, moving]).strip().splitlines()[-1]
        print(line, flush=True)
        timing.append(json.loads(line))
    with open(os.path.join(out, "ref_gpu_timing.json"), "w") as f:
        json.dump(timing, f, indent=1)


if __name__ == "__main__":
    main()
