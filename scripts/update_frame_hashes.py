"""Writes gpurun_out/frame_hashes.json: sha256 of the fp32 frame of every bench workload at the bench's own sizes, rendered
on ONE GPU through the C ABI (run on the GPU box; the result is committed as tests/golden/frame_hashes.json).
    python scripts/update_frame_hashes.py
The frames do not depend on batch size, tiling or GPU count (fixed-point accumulation), so bench.py must find the same hash
at N = 1, 2, 4, 8; tests/test_gpu_parity.py::test_committed_frame_hashes re-renders them."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "improved-path-tracer_b200"))
import bench
import pyipt

out = {}
for name, over, fp64 in [("spheres4k", {}, False)] + [(n, o, False) for n, o in bench.PER_CONFIG] + [("spheres4k", {"spp": 16}, True)]:
    wl = dict(bench.WORKLOADS[name]); wl.update(over)
    hs = pyipt.HostScene.load(bench.scene_file(wl["scene"]), width=wl["width"], height=wl["height"])
    c = pyipt.Context(0); c.set_scene(hs)
    c.render(wl["spp"], wl["depth"], seed=123456, flags=pyipt.FLAG_FP64 if fp64 else 0)
    frame = c.download(want64=fp64)
    c.close()
    key = bench.frame_key(name, hs.width, hs.height, wl["depth"], wl["spp"], 123456, fp64)
    out[key] = hashlib.sha256(np.ascontiguousarray(frame).tobytes()).hexdigest()
    print(key, out[key], float(frame.mean()), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "frame_hashes.json"), "w"), indent=1, sort_keys=True)
