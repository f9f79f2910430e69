"""Generates the golden fixtures in this directory from the reference's own code compiled for the host
(oracle/_ref/libref_host.so, built by `make -C oracle ref` where /root/reference is mounted).  Run once, here:
    python tests/golden/make_golden.py [--converged]
The reference ships no tests or golden vectors, so outputs of the reference itself are the anchor.
"""
import json, os, sys, time
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ctypes
import oracle as O

def small_images():
    out = {}
    for name, d, s, w, h in (("spheres", 10, 8, 128, 72), ("mirrors", 10, 8, 128, 72), ("maze", 10, 8, 128, 72),
                             ("spheres", 3, 4, 67, 41), ("spheres", 4, 4, 128, 72), ("mirrors", 32, 4, 96, 54),
                             ("spheres", 10, 4, 21, 22), ("spheres", 10, 4, 22, 22)):
        img = O.ref_render(name, s, d, width=w, height=h)
        out[f"{name}_d{d}_s{s}_{w}x{h}"] = img
        print(name, d, s, w, h, img.mean(axis=(0, 1)))
    np.savez_compressed(os.path.join(HERE, "ref_images_small.npz"), **out)

def function_kats():
    R = O.ref()
    rng = np.random.default_rng(20260101)
    dp = ctypes.POINTER(ctypes.c_double)
    def P(a): return np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(dp)
    n = 400
    sph = {"radius": rng.uniform(1, 600, n), "c": rng.uniform(-500, 1500, (n, 3)), "o": rng.uniform(-500, 1500, (n, 3)), "d": rng.normal(size=(n, 3))}
    sph["d"] /= np.linalg.norm(sph["d"], axis=1, keepdims=True)
    sph["d"][::7] *= rng.uniform(0.33, 1.0, (len(sph["d"][::7]), 1))        # non-unit directions (refracted rays)
    sph["o"][::3] = sph["c"][::3] + sph["radius"][::3, None] * sph["d"][::3] * rng.uniform(0, 2.5, (len(sph["o"][::3]), 1))
    sph["t"] = np.array([R.ref_sphere_intersect(sph["radius"][i], P(sph["c"][i]), P(sph["o"][i]), P(sph["d"][i])) for i in range(n)])
    pl = {"north": np.zeros((n, 3)), "east": np.zeros((n, 3)), "c": rng.uniform(-500, 1500, (n, 3)), "o": rng.uniform(-500, 1500, (n, 3)), "d": rng.normal(size=(n, 3))}
    for i in range(n):
        ax = rng.permutation(3)
        pl["north"][i, ax[0]] = rng.uniform(5, 600) * rng.choice([-1, 1]); pl["east"][i, ax[1]] = rng.uniform(5, 600) * rng.choice([-1, 1])
        if i % 5 == 0:   # general (non axis-aligned, still orthogonal) rectangles
            a = rng.normal(size=3); a /= np.linalg.norm(a); b = np.cross(a, rng.normal(size=3)); b /= np.linalg.norm(b)
            pl["north"][i] = a * rng.uniform(5, 600); pl["east"][i] = b * rng.uniform(5, 600)
        if i % 11 == 0:  # non-orthogonal north/east: the reference rejects (almost) every hit
            pl["east"][i] += 0.3 * pl["north"][i]
        # aim most rays at the rectangle's neighbourhood so that hits, near-edge hits and misses all occur
        tgt = pl["c"][i] + pl["north"][i] * rng.uniform(-1.3, 1.3) + pl["east"][i] * rng.uniform(-1.3, 1.3)
        if i % 4: pl["d"][i] = tgt - pl["o"][i]
    pl["d"] /= np.linalg.norm(pl["d"], axis=1, keepdims=True)
    pl["t"] = np.array([R.ref_plane_intersect(P(pl["north"][i]), P(pl["east"][i]), P(pl["c"][i]), P(pl["o"][i]), P(pl["d"][i])) for i in range(n)])
    sc = {"kind": rng.integers(0, 2, n), "reflection": rng.integers(0, 3, n), "depth": rng.integers(0, 6, n), "subseq": rng.integers(0, 484, n),
          "geom": np.zeros((n, 9)), "P": np.zeros((n, 3)), "in": rng.normal(size=(n, 3)), "out": np.zeros((n, 16))}
    sc["in"] /= np.linalg.norm(sc["in"], axis=1, keepdims=True)
    sc["in"][::5] *= 0.5
    for i in range(n):
        if sc["kind"][i] == 0:
            r = rng.uniform(10, 600); c = rng.uniform(-500, 1500, 3); nrm = rng.normal(size=3); nrm /= np.linalg.norm(nrm)
            sc["geom"][i, :4] = [r, *c]; sc["P"][i] = c + r * nrm
        else:
            ax = rng.permutation(3); north = np.zeros(3); east = np.zeros(3)
            north[ax[0]] = rng.uniform(5, 600); east[ax[1]] = rng.uniform(5, 600); c = rng.uniform(-500, 1500, 3)
            sc["geom"][i] = [*north, *east, *c]; sc["P"][i] = c + north * rng.uniform(-1, 1) + east * rng.uniform(-1, 1)
        out = np.zeros(16)
        R.ref_scatter(int(sc["kind"][i]), P(sc["geom"][i]), int(sc["reflection"][i]), P(sc["P"][i]), P(sc["in"][i]), int(sc["depth"][i]), int(sc["subseq"][i]), P(out))
        sc["out"][i] = out
    np.savez_compressed(os.path.join(HERE, "ref_function_kats.npz"), **{f"sphere_{k}": v for k, v in sph.items()},
                        **{f"plane_{k}": v for k, v in pl.items()}, **{f"scatter_{k}": v for k, v in sc.items()})
    raw = (ctypes.c_uint * 2)(); u = ctypes.c_double(); kat = {}
    for s in (0, 1, 483):
        R.ref_xorwow_kat(s, raw, ctypes.byref(u)); kat[str(s)] = [int(raw[0]), int(raw[1]), u.value]
    return kat

def full_means():
    res = {}
    for name, d, s in (("spheres", 10, 40), ("mirrors", 10, 16), ("maze", 10, 16)):
        t = time.time(); img = O.ref_render(name, s, d); dt = time.time() - t
        res[f"{name}_d{d}_s{s}"] = {"mean_rgb": img.mean(axis=(0, 1)).tolist(), "seconds_8_threads": dt, "samples": int(img.shape[0] * img.shape[1] * s)}
        print(name, res[f"{name}_d{d}_s{s}"])
    return res

def converged():
    out = {}
    for name, w, h in (("spheres", 96, 54), ("mirrors", 64, 36), ("maze", 64, 36)):
        t = time.time(); img = O.ref_render(name, 65535, 10, width=w, height=h)
        out[f"{name}_d10_s65535_{w}x{h}"] = img
        print(name, w, h, time.time() - t, img.mean(axis=(0, 1)), flush=True)
        np.savez_compressed(os.path.join(HERE, "ref_converged.npz"), **out)

if __name__ == "__main__":
    if "--converged" in sys.argv:
        converged()
    else:
        small_images()
        meta = {"xorwow_kat_seed123456": function_kats(), "full_frame_means": full_means(),
                "generated_by": "tests/golden/make_golden.py from oracle/_ref/libref_host.so (reference Renderer.cu compiled for the host, g++ -O2)"}
        json.dump(meta, open(os.path.join(HERE, "ref_meta.json"), "w"), indent=1)
