"""Small renders through every kernel variant, for `compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py`."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "improved-path-tracer_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyipt
from scene_util import synthetic_scene, write_scene
scenes = os.path.join(ROOT, "oracle", "_ref", "scenes")
for name in ("spheres", "mirrors", "maze"):
    hs = pyipt.HostScene.load(os.path.join(scenes, name + ".json"), width=96, height=54)
    for flags in (0, pyipt.FLAG_FP64, pyipt.FLAG_RUSSIAN_ROULETTE):
        img, st = pyipt.render(hs, 3, 7, flags=flags, batch=2048)
        assert np.isfinite(img).all()
    img, st = pyipt.render(hs, 2, 131)                      # DEFER kernels
    hb = pyipt.HostScene.load(os.path.join(scenes, name + ".json"), width=96, height=54, brute_max=4)
    for flags in (0, pyipt.FLAG_FP64):
        img, st = pyipt.render(hb, 3, 7, flags=flags)        # BVH: split pipeline (fp32) and fused generic (fp64)
    img, st = pyipt.render(hb, 2, 131, flags=pyipt.FLAG_FP64)
    print(name, "ok", flush=True)
path = write_scene("/tmp/san_syn.json", synthetic_scene(900, 3, width=64, height=36))
hs = pyipt.HostScene.load(path)
c = pyipt.Context(0); c.set_scene(hs)
c.render(3, 8); c.download(); c.download_rgb8()
rays = np.random.default_rng(0).normal(size=(500, 6)); rays[:, :3] = rays[:, :3] * 300 + [640, 100, 360]
c.trace(rays, 0); c.trace(rays, pyipt.FLAG_FP64)
for rank in range(3):
    c.render(2, 6, rank=rank, world=3, tile=(16, 8))
c.close()
print("sanitize smoke ok")
