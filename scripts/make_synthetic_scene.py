"""BASELINE config 5: a schema-conformant synthetic scene with ~1M primitives.
    python scripts/make_synthetic_scene.py OUT.json [N_TOTAL=1000000] [WIDTH=1280 HEIGHT=720]
The 6 room rectangles + light sphere of the spheres.json layout, plus N_TOTAL-7 small primitives jittered on a lattice
inside the room: 70 % spheres (r in [1,3]), 30 % axis-aligned rectangles (half-extents in [1,4], random axis pair);
materials 80 % diffuse / 10 % specular / 10 % refractive, colour U[0.2,0.9]^3, 0.1 % emissive (E = U[5,20]).
Seeded: numpy default_rng(20260101)."""
import sys
import numpy as np


def main():
    out = sys.argv[1]
    n_total = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    W = int(sys.argv[3]) if len(sys.argv) > 3 else 1280
    H = int(sys.argv[4]) if len(sys.argv) > 4 else 720
    rng = np.random.default_rng(20260101)
    n = n_total - 7
    side = int(np.ceil(n ** (1 / 3)))
    idx = rng.permutation(side ** 3)[:n]
    cell = np.stack([idx % side, (idx // side) % side, idx // (side * side)], axis=1).astype(np.float64)
    lo, hi = np.array([20.0, -500.0, 20.0]), np.array([1260.0, 690.0, 700.0])
    pos = lo + (cell + rng.uniform(0.2, 0.8, (n, 3))) / side * (hi - lo)
    is_sphere = rng.random(n) < 0.7
    radius = rng.uniform(1, 3, n)
    ext = rng.uniform(1, 4, (n, 2))
    axes = np.array([rng.permutation(3)[:2] for _ in range(64)])[rng.integers(0, 64, n)]
    refl = rng.choice([0, 1, 2], size=n, p=[0.8, 0.1, 0.1])
    col = rng.uniform(0.2, 0.9, (n, 3))
    emissive = rng.random(n) < 0.001
    emi = np.where(emissive[:, None], rng.uniform(5, 20, (n, 3)), 0.0)

    def v(a):
        return '{"xx":%.6g,"yy":%.6g,"zz":%.6g}' % (a[0], a[1], a[2])

    walls = [((640, 720, 360), (0, 0, 370), (650, 0, 0), (.75, .75, .75)), ((-10, 95, 360), (0, 0, 370), (0, 641, 0), (.75, .25, .25)),
             ((1290, 95, 360), (0, 0, 370), (0, 641, 0), (.25, .25, .75)), ((640, 95, -10), (0, 641, 0), (650, 0, 0), (.75, .75, .75)),
             ((640, 95, 730), (0, 641, 0), (650, 0, 0), (.75, .75, .75)), ((640, -546, 360), (0, 0, 370), (650, 0, 0), (.25, .75, .25))]
    with open(out, "w") as f:
        f.write('{"width":%d,"height":%d,"camera":{"position":%s,"direction":%s,"orientation":%s},"objects":[\n' %
                (W, H, v((640, 0, 360)), v((0, 1, 0)), v((-1, 0, 0))))
        for p, no, ea, c in walls:
            f.write('{"type":"plane","position":%s,"north":%s,"east":%s,"color":%s,"emission":%s,"reflection":0},\n' % (v(p), v(no), v(ea), v(c), v((0, 0, 0))))
        f.write('{"type":"sphere","radius":600,"position":%s,"color":%s,"emission":%s,"reflection":0}' % (v((640, 95, 1320)), v((0, 0, 0)), v((20, 20, 20))))
        chunk = []
        for i in range(n):
            if is_sphere[i]:
                chunk.append(',\n{"type":"sphere","radius":%.6g,"position":%s,"color":%s,"emission":%s,"reflection":%d}' %
                             (radius[i], v(pos[i]), v(col[i]), v(emi[i]), refl[i]))
            else:
                no, ea = [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]
                no[axes[i, 0]] = ext[i, 0]
                ea[axes[i, 1]] = ext[i, 1]
                chunk.append(',\n{"type":"plane","position":%s,"north":%s,"east":%s,"color":%s,"emission":%s,"reflection":%d}' %
                             (v(pos[i]), v(no), v(ea), v(col[i]), v(emi[i]), refl[i]))
            if len(chunk) >= 20000:
                f.write("".join(chunk)); chunk = []
        f.write("".join(chunk))
        f.write("\n]}\n")


if __name__ == "__main__":
    main()
