"""BASELINE config 5 on the GPU box: generate the synthetic 1M-primitive scene, time ingest (parse, BVH build, upload),
check nearest-hit parity against the oracle's linear scan on random rays, and measure render throughput."""
import os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "improved-path-tracer_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyipt, oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
path = f"/tmp/syn{n}.json"
t = time.time(); subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "make_synthetic_scene.py"), path, str(n)], check=True); print(f"generate {time.time()-t:.1f}s  {os.path.getsize(path)/1e6:.0f} MB", flush=True)
t = time.time(); hs = pyipt.HostScene.load(path, brute_max=10**9); t_parse = time.time() - t
t = time.time(); nodes = hs.build_bvh(int(os.environ.get("LEAF", "4")), 64); t_bvh = time.time() - t
print(f"parse {t_parse:.2f}s  bvh {nodes} nodes {t_bvh:.2f}s", flush=True)
ctx = pyipt.Context(0)
t = time.time(); ctx.set_scene(hs); print(f"upload {time.time()-t:.3f}s", flush=True)
# nearest-hit parity vs the oracle's linear scan (Renderer.cu:227-243)
sc = O.Scene.load(path)
rng = np.random.default_rng(1)
m = int(os.environ.get("NRAYS", "600"))
o = rng.uniform([30, -480, 30], [1250, 680, 690], size=(m, 3)); d = rng.normal(size=(m, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
rays = np.concatenate([o, d], axis=1)
t = time.time(); oi, ot = O.nearest_hit(sc, rays); print(f"oracle scan of {m} rays: {time.time()-t:.1f}s", flush=True)
for flags, tag in ((pyipt.FLAG_FP64, "fp64"), (0, "fp32")):
    gi, gt = ctx.trace(rays, flags)
    same = gi == oi; hit = same & (oi >= 0)
    print(f"trace {tag}: same object {same.mean():.5f}  hits {np.mean(oi>=0):.3f}  max rel dt {np.max(np.abs(gt[hit]-ot[hit])/ot[hit]):.2e}", flush=True)
for it in range(2):
    st = ctx.render(spp, 10)
print(f"render 1280x720 d10 s{spp}: {st['render_ms']:.1f} ms  {st['samples']/st['render_ms']*1e-3:.1f} Msamples/s  {st['traced_bounces']/st['render_ms']*1e-6:.3f} Gbounces/s  casts/sample {st['traced_bounces']/st['samples']:.2f}", flush=True)
img = ctx.download()
print("mean", img.mean(axis=(0, 1)), "finite", np.isfinite(img).all())
