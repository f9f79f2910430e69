"""First GPU contact: parity of the fp64 and fp32 paths against the oracle on the three shipped scenes, nearest-hit
parity, and raw throughput numbers."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "improved-path-tracer_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyipt, oracle as O

print("device:", pyipt.lib().ipt_device_name(0).decode(), "count", pyipt.lib().ipt_device_count(), flush=True)
for name in ("spheres", "mirrors", "maze"):
    path = O.scene_path(name)
    W, H, spp, depth, seed = 320, 180, 8, 10, 11
    sc = O.Scene.load(path, W, H)
    t = time.time(); ref, cnt = O.render(sc, spp, depth, rng=O.RNG_COUNTER, seed=seed); t_or = time.time() - t
    for brute_max in (64, 4):
        hs = pyipt.HostScene.load(path, width=W, height=H, brute_max=brute_max)
        for flags, tag in ((pyipt.FLAG_FP64, "fp64"), (0, "fp32")):
            img, st = pyipt.render(hs, spp, depth, seed=seed, flags=flags)
            d = np.abs(img - ref)
            tol = (1e-9 if flags else 1e-3) * np.maximum(1e-3, np.abs(ref))
            frac = np.mean(np.all(d <= tol, axis=2))
            print(f"{name:8s} bvh={'yes' if brute_max==4 else 'no '} {tag}: pixels within tol {frac:.5f}  mean gpu {img.mean(axis=(0,1))} oracle {ref.mean(axis=(0,1))}"
                  f"  traced {st['traced_bounces']} needed {cnt['casts_needed']}  render {st['render_ms']:.2f} ms", flush=True)
        # nearest-hit parity on random rays
        rng = np.random.default_rng(5)
        n = 20000
        o = rng.uniform([0, -500, 0], [1280, 700, 720], size=(n, 3)); dd = rng.normal(size=(n, 3)); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
        rays = np.concatenate([o, dd], axis=1)
        oi, ot = O.nearest_hit(sc, rays)
        ctx = pyipt.Context(0); ctx.set_scene(hs)
        for flags, tag in ((pyipt.FLAG_FP64, "fp64"), (0, "fp32")):
            gi, gt = ctx.trace(rays, flags)
            same = gi == oi
            hit = same & (oi >= 0)
            print(f"   trace {tag}: same object {same.mean():.5f}  max |dt|/t on agreeing hits {np.max(np.abs(gt[hit]-ot[hit])/ot[hit]):.3e}", flush=True)
        ctx.close()
# throughput at full size
for name in ("spheres", "mirrors", "maze"):
    hs = pyipt.HostScene.load(O.scene_path(name))
    ctx = pyipt.Context(0); ctx.set_scene(hs)
    for flags, tag in ((0, "fp32"), (pyipt.FLAG_FP64, "fp64")):
        for it in range(2):
            st = ctx.render(40, 10, flags=flags)
        print(f"{name:8s} {tag} 1280x720 d10 s40: {st['render_ms']:.2f} ms  {st['samples']/st['render_ms']*1e-3:.1f} Msamples/s  {st['traced_bounces']/st['render_ms']*1e-6:.2f} Gbounces/s  launches {st['kernel_launches']}", flush=True)
    img = ctx.download()
    print("   mean", img.mean(axis=(0,1)))
    ctx.close()
