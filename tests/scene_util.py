"""Helpers shared by the tests: synthetic scenes in the reference's scenes/*.json schema."""
import json

import numpy as np


def vec(a):
    return {"xx": float(a[0]), "yy": float(a[1]), "zz": float(a[2])}


def room_objects():
    """The six walls and the light of spheres.json's layout (values written out here, not read from the reference)."""
    walls = [((640, 720, 360), (0, 0, 370), (650, 0, 0), (.75, .75, .75)), ((-10, 95, 360), (0, 0, 370), (0, 641, 0), (.75, .25, .25)),
             ((1290, 95, 360), (0, 0, 370), (0, 641, 0), (.25, .25, .75)), ((640, 95, -10), (0, 641, 0), (650, 0, 0), (.75, .75, .75)),
             ((640, 95, 730), (0, 641, 0), (650, 0, 0), (.75, .75, .75)), ((640, -546, 360), (0, 0, 370), (650, 0, 0), (.25, .75, .25))]
    objs = [{"type": "plane", "position": vec(p), "north": vec(n), "east": vec(e), "color": vec(c), "emission": vec((0, 0, 0)), "reflection": 0}
            for p, n, e, c in walls]
    objs.append({"type": "sphere", "radius": 600.0, "position": vec((640, 95, 1320)), "color": vec((0, 0, 0)), "emission": vec((20, 20, 20)), "reflection": 0})
    return objs


def synthetic_scene(n_small, seed, width=160, height=90, general_rects=True):
    """Room + n_small small spheres / rectangles of all three materials, a few emissive (BASELINE config 5 in miniature)."""
    rng = np.random.default_rng(seed)
    objs = room_objects()
    for i in range(n_small):
        pos = rng.uniform([60, -300, 40], [1220, 650, 680])
        refl = int(rng.choice([0, 0, 0, 1, 2]))
        col = rng.uniform(0.2, 0.9, 3)
        emi = rng.uniform(5, 20, 3) if rng.random() < 0.03 else np.zeros(3)
        if rng.random() < 0.7:
            objs.append({"type": "sphere", "radius": float(rng.uniform(8, 40)), "position": vec(pos), "color": vec(col), "emission": vec(emi), "reflection": refl})
        else:
            ax = rng.permutation(3)
            north, east = np.zeros(3), np.zeros(3)
            north[ax[0]] = rng.uniform(10, 50)
            east[ax[1]] = rng.uniform(10, 50)
            if general_rects and i % 4 == 0:   # rotated, still orthogonal
                a = rng.normal(size=3); a /= np.linalg.norm(a)
                b = np.cross(a, rng.normal(size=3)); b /= np.linalg.norm(b)
                north, east = a * rng.uniform(10, 50), b * rng.uniform(10, 50)
            objs.append({"type": "plane", "position": vec(pos), "north": vec(north), "east": vec(east), "color": vec(col), "emission": vec(emi), "reflection": refl})
    return {"width": width, "height": height,
            "camera": {"position": vec((640, 0, 360)), "direction": vec((0, 1, 0)), "orientation": vec((-1, 0, 0))},
            "objects": objs}


def scale_scene(scene, k):
    """Every length of the scene times k (positions, radii, edge vectors, camera position, frame size): SURVEY.md §8d's
    evenly loaded 4K variant of spheres.json is k = 3."""
    out = json.loads(json.dumps(scene))
    out["width"], out["height"] = int(out["width"] * k), int(out["height"] * k)
    for o in [out["camera"]] + out["objects"]:
        for key in ("position", "north", "east"):
            if key in o:
                o[key] = {a: float(k) * b for a, b in o[key].items()}
        if "radius" in o:
            o["radius"] = float(k) * o["radius"]
    return out


def write_scene(path, scene):
    with open(path, "w") as f:
        json.dump(scene, f)
    return str(path)
