"""Route 2: runs the reference's own CUDA program (oracle/_ref/tracer_ref_cuda, built from the unmodified
reference sources for sm_100a with only Image.cpp replaced) on this box, records its own benchmark.txt time, and
compares its frame's channel means with the host-compiled reference (tests/golden/ref_meta.json) — risk R1: does the
GPU build's out-of-bounds read (Renderer.cu:216-219) tint the image?"""
import json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exe = os.path.join(ROOT, "oracle", "_ref", "tracer_ref_cuda")
out_dir = os.path.join(ROOT, "gpurun_out", "ref_cuda")
os.makedirs(out_dir, exist_ok=True)
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_meta.json")))["full_frame_means"]
res = {}
for scene, spp, tmo in (("spheres", 40, 600), ("mirrors", 16, 600), ("maze", 16, 900)):
    path = os.path.join(ROOT, "oracle", "_ref", "scenes", scene + ".json")
    t0 = time.time()
    try:
        r = subprocess.run([exe, "-d=10", f"-s={spp}", path], cwd=out_dir, capture_output=True, text=True, timeout=tmo)
        dt = time.time() - t0
        tail = r.stdout[-400:].replace("\r", "\n").splitlines()[-4:]
        f = os.path.join(out_dir, f"{scene}D10S{spp}.f64")
        entry = {"wall_s": dt, "returncode": r.returncode, "stdout_tail": tail}
        if os.path.isfile(f):
            raw = np.fromfile(f, dtype=np.uint8)
            w, h = np.frombuffer(raw[:8].tobytes(), dtype=np.uint32)
            img = np.frombuffer(raw[8:].tobytes(), dtype=np.float64).reshape(h, w, 3)
            entry["mean_rgb"] = img.mean(axis=(0, 1)).tolist()
            entry["host_reference_mean_rgb"] = gold[f"{scene}_d10_s{spp}"]["mean_rgb"]
            entry["nan_or_inf"] = int((~np.isfinite(img)).sum())
            entry["msamples_per_s_wall"] = w * h * spp / dt / 1e6
        res[scene] = entry
    except subprocess.TimeoutExpired:
        res[scene] = {"timeout_s": tmo}
    print(scene, json.dumps(res[scene]), flush=True)
bench = os.path.join(out_dir, "benchmark.txt")
if os.path.isfile(bench):
    res["benchmark_txt"] = open(bench).read()
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ref_cuda_results.json"), "w"), indent=1)
