"""CPU tests of the C++ host layer and of the C-ABI library surface (no GPU, no compute calls)."""
import ctypes
import json
import os
import re
import subprocess

import numpy as np
import pytest

from scene_util import synthetic_scene, vec, write_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pyipt):
    declared = set()
    for h in ("ipt_abi.h", "ipt_host.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        declared |= set(re.findall(r"\b(ipt_[a-z0-9_]+)\s*\(", src))
    assert len(declared) >= 25
    assert declared == set(pyipt.ABI_SYMBOLS) | set(pyipt.HOST_SYMBOLS)
    L = pyipt.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.ipt_abi_version() == 3


def test_no_torch_and_no_oracle_in_the_product():
    """The shared library and its sources reference neither torch nor anything under oracle/."""
    pkg = os.path.join(ROOT, "improved-path-tracer_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".hpp", ".py", "Makefile")):
                text = open(os.path.join(d, f)).read()
                assert "oracle" not in text.replace("the oracle run on", "").replace("CPU oracle see", "").lower() or f in ("ipt_device.cuh",), f
                assert "torch" not in text, f
    out = subprocess.run(["ldd", os.path.join(pkg, "libipt_b200.so")], capture_output=True, text=True).stdout
    assert "torch" not in out and "oracle" not in out


def test_loader_matches_python_json(pyipt, oracle):
    for name in ("spheres", "mirrors", "maze"):
        hs = pyipt.HostScene.load(oracle.scene_path(name))
        sc = oracle.Scene.load(name)
        a = hs.arrays()
        v = hs.view.contents
        assert (v.width, v.height, v.n_objects) == (sc.width, sc.height, len(sc.objects))
        assert np.array_equal(list(v.cam_origin), sc.cam_pos) and np.array_equal(list(v.cam_dir), sc.cam_dir) and np.array_equal(list(v.cam_orient), sc.cam_orient)
        assert np.array_equal(a["mat_color"], [o["color"] for o in sc.objects])
        assert np.array_equal(a["mat_emission"], [o["emission"] for o in sc.objects])
        assert np.array_equal(a["mat_reflection"], [o["reflection"] for o in sc.objects])
        sph = [i for i, o in enumerate(sc.objects) if o["type"] == 0]
        rec = [i for i, o in enumerate(sc.objects) if o["type"] == 1]
        assert list(a["sphere_object"]) == sph and list(a["rect_object"]) == rec   # object (JSON) order kept
        for k, i in enumerate(sph):
            assert np.array_equal(a["sphere_cxyzr"][k], list(sc.objects[i]["position"]) + [sc.objects[i]["radius"]])
        for k, i in enumerate(rec):
            N, E, c = (np.array(sc.objects[i][q]) for q in ("north", "east", "position"))
            n = np.cross(N, E); n = n * (1 / np.sqrt(n @ n))
            assert np.allclose(a["rect_plane"][k], [*n, n @ c], rtol=0, atol=1e-12)
            u, vv, b = a["rect_u"][k], a["rect_v"][k], a["rect_bounds"][k]
            assert abs(u[:3] @ E) < 1e-9 and abs(u[:3] @ n) < 1e-12 and abs(np.linalg.norm(u[:3]) - 1) < 1e-12
            assert abs(vv[:3] @ N) < 1e-9 and abs(vv[:3] @ n) < 1e-12
            assert np.allclose(b, [0, np.linalg.norm(N) + 5e-5, 0, np.linalg.norm(E) + 5e-5], rtol=0, atol=1e-9)
        assert a["n_bvh_nodes"] == 0   # <= 64 primitives: brute force from shared memory


def test_rect_bounds_reproduce_reference_test(pyipt, oracle, tmp_path):
    """The flattened bounds test == Plane.cu:87-100 (through the oracle's literal restatement) on random hits,
    for orthogonal, rotated and NON-orthogonal rectangles."""
    rng = np.random.default_rng(1)
    scene = synthetic_scene(0, 0)
    rects = []
    for i in range(40):
        a = rng.normal(size=3); a /= np.linalg.norm(a)
        b = np.cross(a, rng.normal(size=3)); b /= np.linalg.norm(b)
        north, east = a * rng.uniform(5, 300), b * rng.uniform(5, 300)
        if i % 4 == 0:
            east = east + rng.uniform(0.01, 0.6) * north   # non-orthogonal
        rects.append((rng.uniform(-200, 1200, 3), north, east))
    scene["objects"] = [{"type": "plane", "position": vec(c), "north": vec(n), "east": vec(e), "color": vec((.5, .5, .5)), "emission": vec((0, 0, 0)), "reflection": 0} for c, n, e in rects]
    path = write_scene(tmp_path / "rects.json", scene)
    a = pyipt.HostScene.load(path).arrays()
    dp = ctypes.POINTER(ctypes.c_double)
    P = lambda x: np.ascontiguousarray(x, dtype=np.float64).ctypes.data_as(dp)
    agree = total = hits = 0
    for k, (c, north, east) in enumerate(rects):
        for _ in range(200):
            target = c + north * rng.uniform(-1.2, 1.2) + east * rng.uniform(-1.2, 1.2)
            o = target + rng.normal(size=3) * 300
            d = target - o; d /= np.linalg.norm(d)
            t_ref = oracle.lib().or_plane_intersect(P(north), P(east), P(c), P(o), P(d))
            pl, u, v, b = a["rect_plane"][k], a["rect_u"][k], a["rect_v"][k], a["rect_bounds"][k]
            den = pl[:3] @ d
            t = (pl[3] - pl[:3] @ o) / den if den != 0 else 0.0
            ok = t > 1e-4
            if ok:
                hit = o + d * t
                su, sv = abs(u[:3] @ hit - u[3]), abs(v[:3] @ hit - v[3])
                ok = b[0] <= su <= b[1] and b[2] <= sv <= b[3]
            total += 1
            agree += (ok == (t_ref != 0.0))
            hits += t_ref != 0.0
            if ok and t_ref != 0.0:
                assert abs(t - t_ref) <= 1e-9 * max(1.0, abs(t_ref))
    assert agree == total and hits > 0.2 * total


LOAD_ERRORS = [
    (lambda s: s.pop("width"), "Missing height or witdh data!"),
    (lambda s: s.pop("camera"), "No camera data!"),
    (lambda s: s["camera"].pop("orientation"), "Camera data could not be read!"),
    (lambda s: s["camera"]["position"].pop("yy"), "Camera data could not be parsed!"),
    (lambda s: s.pop("objects"), "No objects data!"),
    (lambda s: s["objects"][2].pop("color"), "Could not validate object data!"),
    (lambda s: s["objects"][1]["emission"].pop("xx"), "Could not validate object data!"),
    (lambda s: s["objects"][0].update(type="cube"), "Unknown object type"),
    (lambda s: s["objects"][6].pop("radius"), "Broken sphere object! "),
    (lambda s: s["objects"][0].pop("north"), "Broken plane object! "),
    (lambda s: s.update(objects=[]), "Object list empty! Cannot build scene"),
]


@pytest.mark.parametrize("case", range(len(LOAD_ERRORS)))
def test_loader_messages(pyipt, tmp_path, case):
    """SceneData.cpp's messages, in the reference's validation order."""
    mutate, message = LOAD_ERRORS[case]
    scene = synthetic_scene(3, 0)
    mutate(scene)
    path = write_scene(tmp_path / "bad.json", scene)
    with pytest.raises(pyipt.IptError) as e:
        pyipt.HostScene.load(path)
    assert str(e.value) == message


def test_loader_missing_and_malformed_file(pyipt, tmp_path):
    with pytest.raises(pyipt.IptError) as e:
        pyipt.HostScene.load(str(tmp_path / "nope.json"))
    assert str(e.value) == "Could not load provided json file!"
    p = tmp_path / "broken.json"
    p.write_text('{"width": 10, "height": ')
    with pytest.raises(pyipt.IptError):
        pyipt.HostScene.load(str(p))
    # not a file at all: a message, never an exception across the C ABI (a directory once ended the process in std::length_error)
    for special in (str(tmp_path), "/dev/null", "/proc/self/status"):
        with pytest.raises(pyipt.IptError) as e:
            pyipt.HostScene.load(special)
        assert str(e.value) == "Could not load provided json file!"
    (tmp_path / "empty.json").write_text("")
    with pytest.raises(pyipt.IptError):
        pyipt.HostScene.load(str(tmp_path / "empty.json"))
    # a scene that is not a regular file (read in pieces instead of mapped)
    import threading
    fifo = str(tmp_path / "scene.fifo")
    os.mkfifo(fifo)
    text = json.dumps(synthetic_scene(40, 1))
    w = threading.Thread(target=lambda: open(fifo, "w").write(text))
    w.start()
    hs = pyipt.HostScene.load(fifo)
    w.join()
    assert hs.view.contents.n_objects == len(json.loads(text)["objects"])


def test_loader_survives_mutated_scenes(pyipt, oracle, tmp_path):
    """SceneData.cpp:61-96 answers an unreadable scene with a message and no scene - where nlohmann does not throw: a parse or
    type error leaves main() uncaught upstream and the program dies.  The pull parser answers all of them with a message, on
    truncated, spliced and byte-flipped copies of spheres.json: every load
    either fails with a message or yields a scene whose BVH builds. (400 seeded mutations here; 21 000 were run once.)"""
    import random
    src = open(oracle.scene_path("spheres"), "rb").read()
    rnd = random.Random(20261019)
    toks = [b"{", b"}", b"[", b"]", b",", b":", b'"', b"-", b"e", b".", b"null", b"true", b"1e999", b"\\", b"\\u00",
            b"0", b" ", b'"objects"', b'"camera"', b"-0", b"1e-999", b"99999999999999999999999"]
    loaded = failed = 0
    p = tmp_path / "mutant.json"
    for _ in range(400):
        b = bytearray(src)
        for _ in range(rnd.randint(1, 2)):
            k, at = rnd.random(), rnd.randrange(max(1, len(b)))
            if k < 0.3:
                del b[at:at + rnd.randint(1, 40)]
            elif k < 0.6:
                b[at:at] = rnd.choice(toks)
            elif k < 0.8 and b:
                b[min(at, len(b) - 1)] = rnd.randrange(256)
            elif k < 0.9:
                b = b[:at]
            else:
                q = rnd.randrange(max(1, len(b)))
                b[at:at] = b[q:q + rnd.randint(1, 200)]
        p.write_bytes(bytes(b))
        try:
            pyipt.HostScene.load(str(p), brute_max=0)
            loaded += 1
        except pyipt.IptError as e:
            assert str(e)
            failed += 1
    assert loaded + failed == 400 and failed > 100 and loaded > 0, (loaded, failed)


def test_loader_matches_the_reference_loader(pyipt, oracle, tmp_path):
    """Differential test against the reference's OWN loader: SceneData.cpp, unmodified, inside oracle/_ref/libref_host.so
    (ref_scene_load_text / ref_scene_objects), run in child processes because it aborts on some inputs.  Seeded mutants of
    mirrors.json - keys removed, renamed or duplicated, values replaced by wrong types, empty containers, huge numbers - and
    byte-level mutants.  Where the reference accepts a file, the host layer must too, with the same objects and camera; where
    the reference prints a message, the same message; where the reference dies (uncaught nlohmann exception, or an
    assertion inside nlohmann) the host layer must refuse the file with a message."""
    import copy, random
    if not oracle.ref_available() or not hasattr(oracle.ref(), "ref_scene_load_text"):
        pytest.skip("oracle/_ref/libref_host.so without ref_scene_load_text (make -C oracle ref where the reference tree is)")
    rnd = random.Random(11)
    base = json.load(open(oracle.scene_path("mirrors")))
    junk = [None, True, "abc", -1, 0, 1e400, [], {}, [1, 2, 3], {"xx": 1}, "sphere", "plane", "cube", 3.5, -0.0, "", 2 ** 40, {"xx": 1, "yy": 2, "zz": 3}]

    def all_paths(node, path=()):
        out = [path]
        if isinstance(node, dict):
            for k, v in node.items():
                out += all_paths(v, path + (k,))
        elif isinstance(node, list):
            for i, v in enumerate(node):
                out += all_paths(v, path + (i,))
        return out

    files = []
    for i in range(260):
        sc = copy.deepcopy(base)
        for _ in range(rnd.randint(1, 2)):
            p = rnd.choice([q for q in all_paths(sc) if q])
            parent = sc
            for k in p[:-1]:
                parent = parent[k]
            k, r = p[-1], rnd.random()
            if r < 0.4:
                if isinstance(parent, dict):
                    parent.pop(k)
                else:
                    del parent[k]
            elif r < 0.8:
                parent[k] = copy.deepcopy(rnd.choice(junk))
            elif isinstance(parent, dict):
                parent[k + "_"] = parent.pop(k)
            else:
                parent.append(copy.deepcopy(parent[k]))
        f = tmp_path / f"m{i}.json"
        f.write_text(json.dumps(sc))
        files.append(str(f))
    src = open(oracle.scene_path("mirrors"), "rb").read()
    for i in range(80):
        b = bytearray(src)
        at = rnd.randrange(len(b))
        k = rnd.random()
        if k < 0.4:
            del b[at:at + rnd.randint(1, 30)]
        elif k < 0.8:
            b[at:at] = rnd.choice([b"{", b"}", b"[", b",", b'"', b"-", b"1e999", b"null", b" "])
        else:
            b[at] = rnd.randrange(32, 127)
        f = tmp_path / f"b{i}.json"
        f.write_bytes(bytes(b))
        files.append(str(f))
    verdicts = oracle.ref_scene_load_texts(files)
    both_ok = same_message = refused = 0
    for f, v in zip(files, verdicts):
        try:
            ours, message = pyipt.HostScene.load(f, brute_max=100000), None
        except pyipt.IptError as e:
            ours, message = None, str(e)
        lines = [] if v is None else [l for l in v[1].split("\n") if l.strip()]
        if v is not None and v[0] >= 0:                                  # the reference accepts the file
            assert ours is not None, (f, message)
            raw, cam, W, H, n = oracle.ref_scene_objects(f)
            theirs = pyipt.HostScene.from_objects(raw, n, W, H, list(np.frombuffer(cam, np.float64)), brute_max=100000)
            a, b = ours.arrays(), theirs.arrays()
            for key in ("sphere_cxyzr", "sphere_object", "rect_plane", "rect_u", "rect_v", "rect_bounds", "rect_object", "mat_color", "mat_emission", "mat_reflection"):
                assert np.array_equal(a[key], b[key], equal_nan=True), (f, key)
            vo, vt = ours.view.contents, theirs.view.contents
            assert (vo.width, vo.height) == (W, H) and list(vo.cam_origin) + list(vo.cam_dir) + list(vo.cam_orient) == list(vt.cam_origin) + list(vt.cam_dir) + list(vt.cam_orient)
            both_ok += 1
        elif v is not None and len(lines) >= 2:                          # the reference prints why it refuses it
            assert ours is None and message.strip() == lines[-1].strip(), (f, lines, message)
            same_message += 1
        else:                                                            # the reference dies: uncaught exception or abort
            assert ours is None and message, f
            refused += 1
    assert both_ok >= 10 and same_message >= 80 and refused >= 80, (both_ok, same_message, refused)


def test_loader_ignores_unknown_keys_and_key_order(pyipt, tmp_path):
    scene = synthetic_scene(5, 2)
    a = pyipt.HostScene.load(write_scene(tmp_path / "a.json", scene)).arrays()
    scene2 = {"objects": [dict(reversed(list(o.items())), extra={"a": [1, {"b": None}], "s": "x\\\"y"}) for o in scene["objects"]],
              "note": "ignored", "height": scene["height"], "camera": scene["camera"], "width": scene["width"]}
    b = pyipt.HostScene.load(write_scene(tmp_path / "b.json", scene2)).arrays()
    for k in ("sphere_cxyzr", "rect_plane", "rect_bounds", "mat_color", "mat_reflection"):
        assert np.array_equal(a[k], b[k])


def test_builders_survive_non_finite_objects(pyipt, oracle, tmp_path):
    """ipt_render_objects takes the caller's ObjectData bytes (ObjectData.hpp:15-31) as they are; the reference's scan
    (Renderer.cu:227-243) never faults on NaN or infinite coordinates, so the BVH and grid builders must not either
    (a NaN cost once left the exact-SAH split of a small range without an axis)."""
    rnd = np.random.default_rng(5)
    # random bytes: whatever object kinds and coordinates they happen to spell
    for i in range(60):
        n = int(rnd.integers(1, 300))
        raw = rnd.integers(0, 256, 144 * n, dtype=np.uint8).tobytes()
        if i % 3 == 0:
            raw = rnd.normal(0, 100, 18 * n).tobytes()
        pyipt.HostScene.from_objects(raw, n, 64, 48, list(rnd.normal(0, 1, 9)), brute_max=0)
    # a scene large enough for the uniform grid, with special values poked into it
    sc = oracle.Scene.load(write_scene(tmp_path / "s.json", synthetic_scene(3000, 3)))
    cs = sc.c_scene()
    raw0 = ctypes.string_at(cs.objects, 144 * cs.n_objects)
    special = [np.nan, np.inf, -np.inf, 1e300, -1e300, 3.5e38, -3.5e38, 1e-320, 0.0, -0.0, 1e20]
    for i in range(20):
        a = np.frombuffer(bytearray(raw0), dtype=np.float64).copy()
        k = int(rnd.integers(1, 50))
        a[rnd.integers(0, a.size, k)] = rnd.choice(special, k)
        h = pyipt.HostScene.from_objects(a.tobytes(), cs.n_objects, sc.width, sc.height, list(cs.camera))
        assert h.arrays()["mat_color"].shape[0] == cs.n_objects


def test_parallel_parse_equals_serial_parse(pyipt, tmp_path, monkeypatch):
    """Files above 8 MB (config 5's is 200 MB) are cut into parts that are scanned for the array's element boundaries on all host
    threads - without knowing, at a part's first byte, the nesting depth or whether it lies inside a string - and the elements
    are parsed in parallel.  Forced on small files here (IPT_PARSE_PARALLEL_MIN=0) with part counts from 4 to 256: the same
    arrays as the serial parse for a scene whose extra keys carry braces, brackets, commas, quotes and backslashes inside
    strings, nested arrays and objects; the same verdict on truncated copies."""
    scene = synthetic_scene(700, 5)
    nasty = ['}{', '],[', 'a,b', 'q\\"uote', 'back\\\\slash\\\\', '"', '\\', '{"xx":1}', '[[[', ']]]}}}', ',,,,', 'x' * 300, '']
    for i, o in enumerate(scene["objects"]):
        o["note" + nasty[i % len(nasty)]] = nasty[(i * 7) % len(nasty)]
        if i % 3 == 0:
            o["extra"] = {"list": [1, [2, {"k": "]}"}], "}"], "s": nasty[(i * 5) % len(nasty)]}
    path = write_scene(tmp_path / "nasty.json", scene)
    text = open(path).read()
    keys = ("sphere_cxyzr", "sphere_object", "rect_plane", "rect_u", "rect_v", "rect_bounds", "rect_object", "mat_color", "mat_emission", "mat_reflection")
    serial = pyipt.HostScene.load(path).arrays()
    cuts = [len(text) * k // 23 for k in range(1, 23)]
    verdict = {}
    for c in cuts:
        (tmp_path / "cut.json").write_text(text[:c])
        try:
            pyipt.HostScene.load(str(tmp_path / "cut.json"))
            verdict[c] = "ok"
        except pyipt.IptError as e:
            verdict[c] = str(e)
    monkeypatch.setenv("IPT_PARSE_PARALLEL_MIN", "0")
    for threads in (1, 2, 5, 16, 64):
        monkeypatch.setenv("IPT_HOST_THREADS", str(threads))
        par = pyipt.HostScene.load(path).arrays()
        for k in keys:
            assert np.array_equal(serial[k], par[k]), (threads, k)
        for c in cuts:
            (tmp_path / "cut.json").write_text(text[:c])
            try:
                pyipt.HostScene.load(str(tmp_path / "cut.json"))
                got = "ok"
            except pyipt.IptError as e:
                got = str(e)
            assert got == verdict[c], (threads, c)


def test_from_objects_equals_loader(pyipt, oracle):
    """ipt_host_from_objects (the reference's ObjectData[] AoS) flattens to the same arrays as the JSON loader."""
    sc = oracle.Scene.load("mirrors")
    cs = sc.c_scene()
    raw = ctypes.string_at(cs.objects, 144 * cs.n_objects)
    a = pyipt.HostScene.from_objects(raw, cs.n_objects, sc.width, sc.height, list(cs.camera)).arrays()
    b = pyipt.HostScene.load(oracle.scene_path("mirrors")).arrays()
    for k in ("sphere_cxyzr", "sphere_object", "rect_plane", "rect_u", "rect_v", "rect_bounds", "rect_object", "mat_color", "mat_emission", "mat_reflection"):
        assert np.array_equal(a[k], b[k]), k


def test_bvh_is_a_valid_partition(pyipt, tmp_path):
    scene = synthetic_scene(700, 3)
    hs = pyipt.HostScene.load(write_scene(tmp_path / "s.json", scene), leaf_size=4, brute_max=64)
    a = hs.arrays()
    n = len(scene["objects"])
    assert a["n_bvh_nodes"] > n // 8
    slots = a["bvh_slot_prim"]
    ns = len(a["sphere_object"])
    flat = np.where(slots & 0x80000000, (slots & 0x7fffffff) + ns, slots)
    assert sorted(flat.tolist()) == list(range(n))             # every primitive exactly once
    # primitive boxes
    lo, hi = np.zeros((n, 3)), np.zeros((n, 3))
    k_s = k_r = 0
    for o in scene["objects"]:
        c = np.array([o["position"][q] for q in ("xx", "yy", "zz")])
        if o["type"] == "sphere":
            lo[k_s], hi[k_s] = c - o["radius"], c + o["radius"]; k_s += 1
        else:
            N = np.array([o["north"][q] for q in ("xx", "yy", "zz")]); E = np.array([o["east"][q] for q in ("xx", "yy", "zz")])
            corners = np.array([c + sn * N + se * E for sn in (-1, 1) for se in (-1, 1)])
            lo[ns + k_r], hi[ns + k_r] = corners.min(0), corners.max(0); k_r += 1
    seen = np.zeros(n, bool)

    def walk(i, blo, bhi):
        nd = a["bvh_nodes"][i]
        for k, (clo, chi) in enumerate(((nd.lo0, nd.hi0), (nd.lo1, nd.hi1))):
            clo, chi = np.array(clo[:], dtype=np.float64), np.array(chi[:], dtype=np.float64)
            assert np.all(clo >= blo - 1e-3) and np.all(chi <= bhi + 1e-3) or blo is None
            ch = nd.child[k]
            if ch >= 0:
                assert ch > i                                   # breadth-first numbering
                walk(ch, clo, chi)
            else:
                first, cnt = ~ch, nd.count[k]
                assert 1 <= cnt <= 4
                for s in range(first, first + cnt):
                    p = flat[s]
                    assert not seen[p]
                    seen[p] = True
                    assert np.all(lo[p] >= clo) and np.all(hi[p] <= chi)   # conservative boxes

    walk(0, np.full(3, -np.inf), np.full(3, np.inf))
    assert seen.all()


def test_to_rgb_matches_reference_mapping(pyipt, oracle):
    """Image.cpp:19-22 through the oracle's restatement."""
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-0.5, 1.5, 2000), [0.0, 1.0, 1 / 255, 254.999 / 255, 20.0, 8.0e6, -8.0e6, 0.999999]])
    for x in xs:
        assert pyipt.lib().ipt_host_to_rgb(float(x)) == oracle.lib().or_to_rgb(float(x))
    # int(x*255) overflows (undefined behaviour upstream) beyond |x| ~ 8.4e6 and for NaN: defined here as saturation / 0
    assert pyipt.lib().ipt_host_to_rgb(float("nan")) == 0 and pyipt.lib().ipt_host_to_rgb(1e12) == 255 and pyipt.lib().ipt_host_to_rgb(-1e12) == 0


def test_png_roundtrip(pyipt, tmp_path):
    from PIL import Image
    rng = np.random.default_rng(4)
    img = rng.uniform(-0.2, 1.4, (37, 53, 3)).astype(np.float32)
    p = str(tmp_path / "x.png")
    assert pyipt.lib().ipt_host_write_png(p.encode(), img.ctypes.data, 53, 37) == 0
    got = np.asarray(Image.open(p))
    assert got.shape == (37, 53, 3) and got.dtype == np.uint8
    want = np.clip((img.astype(np.float64) * 255).astype(np.int64), 0, 255)
    assert np.array_equal(got, want)
    # a frame large enough for several deflate bands (they are compressed on separate threads and concatenated):
    # chunk CRCs, one zlib stream, its Adler-32 over all bands (zlib.decompress raises otherwise), the same pixels
    import struct
    import zlib
    for w, h in ((1280, 720), (1, 1), (5000, 3), (2, 1500)):
        img = rng.uniform(-0.2, 1.4, (h, w, 3)).astype(np.float32)
        assert pyipt.lib().ipt_host_write_png(p.encode(), img.ctypes.data, w, h) == 0
        d = open(p, "rb").read()
        pos, idat = 8, b""
        while pos < len(d):
            n, = struct.unpack(">I", d[pos:pos + 4])
            typ, data = d[pos + 4:pos + 8], d[pos + 8:pos + 8 + n]
            assert zlib.crc32(typ + data) & 0xffffffff == struct.unpack(">I", d[pos + 8 + n:pos + 12 + n])[0], typ
            idat += data if typ == b"IDAT" else b""
            pos += 12 + n
        raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 3 * w)
        assert not raw[:, 0].any()                                # filter type 0 on every row
        want = np.clip((img.astype(np.float64) * 255).astype(np.int64), 0, 255)
        assert np.array_equal(raw[:, 1:].reshape(h, w, 3), want)
        assert np.array_equal(np.asarray(Image.open(p)), want)


def test_time_string_and_benchmark_record(pyipt, tmp_path):
    buf = ctypes.create_string_buffer(64)
    cases = {0: "00:00:00.0", 5007: "00:00:05.7", 59999: "00:00:59.999", 3723456: "01:02:03.456", 36000000 + 61001: "10:01:01.1"}
    for ms, want in cases.items():
        pyipt.lib().ipt_host_time_string(ms, buf, 64)
        assert buf.value.decode() == want                       # Measurements.cpp:21-41: ms not zero-padded
    f = str(tmp_path / "benchmark.txt").encode()
    pyipt.lib().ipt_host_append_benchmark(f, b"spheresD10S40", b"00:00:05.7")
    pyipt.lib().ipt_host_append_benchmark(f, b"mazeD10S40", b"00:01:00.12")
    assert open(f).read() == "spheresD10S40;00:00:05.7;mazeD10S40;00:01:00.12;"   # no newline (Measurements.cpp:53)


def test_time_string_and_benchmark_record_match_the_reference(pyipt, tmp_path, monkeypatch):
    """Differential: getTimeString and saveBenchmark of the reference's own Measurements.cpp (unmodified, reached through
    oracle/_ref/libref_cli.so) against ipt_host_time_string / ipt_host_append_benchmark, on edge and random durations and a
    sequence of records appended to a fresh and to an existing file."""
    so = os.path.join(ROOT, "oracle", "_ref", "libref_cli.so")
    if not os.path.isfile(so) or not hasattr(ctypes.CDLL(so), "ref_time_string"):
        pytest.skip("oracle/_ref/libref_cli.so not built (needs the reference tree: make -C oracle ref)")
    R = ctypes.CDLL(so)
    R.ref_time_string.argtypes = [ctypes.c_ulonglong, ctypes.c_char_p, ctypes.c_int]
    R.ref_save_benchmark.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
    L = pyipt.lib()
    rng = np.random.default_rng(9)
    cases = [0, 1, 9, 10, 99, 100, 999, 1000, 1001, 9999, 59999, 60000, 60001, 599999, 3599999, 3600000, 3600001, 35999999,
             36000000, 86399999, 86400000, 359999999, 360000000, 2 ** 32 - 1, 2 ** 32, 2 ** 40]
    cases += [int(x) for x in rng.integers(0, 4 * 3600000, 2000)] + [int(x) for x in rng.integers(0, 2 ** 36, 500)]
    a, b = ctypes.create_string_buffer(64), ctypes.create_string_buffer(64)
    for ms in cases:
        R.ref_time_string(ms, a, 64)
        L.ipt_host_time_string(ms, b, 64)
        assert a.value == b.value, ms
    monkeypatch.chdir(tmp_path)                                        # the reference always writes ./benchmark.txt
    ours = str(tmp_path / "ours.txt").encode()
    for ident, t in ((b"spheresD10S40", b"00:00:05.7"), (b"mazeD10S40", b"00:01:00.12"), (b"my.scene.v2D255S65535", b"10:01:01.1"), (b"", b"")):
        R.ref_save_benchmark(ident, t)
        L.ipt_host_append_benchmark(ours, ident, t)
        assert open(tmp_path / "benchmark.txt", "rb").read() == open(ours, "rb").read()


def _cli(pyipt, args, capfd):
    argv = (ctypes.c_char_p * (len(args) + 1))(b"tracer", *[a.encode() for a in args])
    out = pyipt.Cli()
    ok = pyipt.lib().ipt_host_parse_cli(len(args) + 1, argv, ctypes.byref(out))
    import sys
    sys.stdout.flush()
    return ok, out, capfd.readouterr().out


def test_cli_grammar(pyipt, oracle, capfd, tmp_path):
    """InputParser.cpp:72-258 accept/reject table."""
    scene = oracle.scene_path("spheres")
    ok, o, _ = _cli(pyipt, [scene], capfd)
    assert ok == 1 and (o.samples, o.max_depth, o.scene_name) == (40, 10, b"spheres")
    ok, o, _ = _cli(pyipt, ["-d=32", "--samples=1024", scene], capfd)
    assert ok == 1 and (o.samples, o.max_depth) == (1024, 32)
    ok, o, _ = _cli(pyipt, ["s=-8", scene], capfd)                 # one dash ANYWHERE in the token counts (:136-142)
    assert ok == 1 and o.samples == 8
    ok, o, _ = _cli(pyipt, ["-s=4", "-s=9", scene], capfd)          # repeats allowed, last wins
    assert ok == 1 and o.samples == 9
    dotted = tmp_path / "my.scene.v2.json"
    dotted.write_text("{}")
    ok, o, _ = _cli(pyipt, [str(dotted)], capfd)
    assert ok == 1 and o.scene_name == b"my.scene.v2"               # up to the LAST dot (:41-55)
    rejects = [
        ([], "Got 0 arguments! Expected between 1 and 3 arguments"),
        (["-s=4", "-d=4", "-d=5", scene], "Got 4 arguments! Expected between 1 and 3 arguments"),
        (["/no/such/file.json"], "Path does not exist"),
        ([str(tmp_path)], "Not a file"),
        (["s=4", scene], "Arguments can have 1 or 2 (-)! Please check your input"),
        (["---s=4", scene], "Arguments can have 1 or 2 (-)! Please check your input"),
        (["-s", scene], "Cannot parse argument: s"),
        (["-s=4=5", scene], "Cannot parse argument: s=4=5"),
        (["-x=4", scene], "Unknown short argument: x=4"),
        (["--s=4", scene], "Unknown long argument: s=4"),
        (["-samples=4", scene], "Unknown short argument: samples=4"),
        (["-s=3", scene], "Number of samples out of range!"),
        (["-s=65536", scene], "Number of samples out of range!"),
        (["-s=99999999999", scene], "Number of samples out of range!"),
        (["-s=abc", scene], "Could not convert samples to number!"),
        (["-d=2", scene], "Depth out of range!"),
        (["-d=256", scene], "Depth out of range!"),
        (["--depth=x", scene], "Could not convert depth to number!"),
        ([scene, "-s=4"], "Path does not exist"),                   # the path must be the LAST argument (:93)
    ]
    for args, cause in rejects:
        ok, _, out = _cli(pyipt, args, capfd)
        assert ok == 0, args
        lines = out.splitlines()
        assert lines[0] == "Error parsing input!" and lines[1] == "Cause: " + cause and lines[2] == "Usage:", (args, out)
        assert lines[3] == "tracer [arguments] [path_to_scene]"
    ok, _, out = _cli(pyipt, ["--help"], capfd)
    assert ok == 0 and out.startswith("tracer [arguments] [path_to_scene]") and "between 4 and 65535" in out and "between 3 and 255" in out


def test_cli_matches_the_reference_parser(pyipt, tmp_path):
    """Differential test of the command-line grammar against the reference's OWN parser: InputParser.cpp, unmodified, compiled
    by oracle/Makefile into oracle/_ref/libref_cli.so (oracle/ref_cli_shim.cpp).  For seeded random argument lists - known and
    unknown keys, one to three dashes anywhere, missing / doubled '=', out-of-range, negative, hexadecimal, padded and
    non-numeric values, files with dots, dashes and '=' in their names, directories, missing paths, --help, 0 to 4 arguments -
    the two agree on validity, on the parsed samples / depth / path / scene name, and on every byte printed
    (22 000 lists were run once: no difference)."""
    import random
    so = os.path.join(ROOT, "oracle", "_ref", "libref_cli.so")
    if not os.path.isfile(so):
        pytest.skip("oracle/_ref/libref_cli.so not built (needs the reference tree: make -C oracle ref)")
    R = ctypes.CDLL(so)
    L = pyipt.lib()
    files = [str(tmp_path / n) for n in ("a.json", "my.scene.v2.json", "noext", "with-dash.json", "k=v.json", ".hidden", "x.")]
    for f in files:
        open(f, "w").write("{}")
    sub = tmp_path / "sub.dir"
    sub.mkdir()
    (sub / "inner").write_text("{}")
    paths = files + [str(sub / "inner"), str(tmp_path), str(sub), "/no/such/file.json", "--help", "", "-s=4", "relative.json"]
    keys = ["-s", "--samples", "-d", "--depth", "s", "d", "-x", "--foo", "samples", "-samples", "--s", "---s", "-s-", "--d", "-depth", "- s", ""]
    vals = ["4", "3", "65535", "65536", "-8", "abc", "", "1e3", "40.5", " 7", "99999999999999", "0x10", "+9", "255", "256", "2", "10", "--5", "4=5", "007"]
    seps = ["=", "=", "=", "==", "", ":", " = "]

    def run_ref(args):
        argv = (ctypes.c_char_p * (len(args) + 2))(b"tracer", *[a.encode() for a in args], None)
        sp, sn = ctypes.create_string_buffer(4096), ctypes.create_string_buffer(1024)
        ns, nd = ctypes.c_int(), ctypes.c_int()
        ok = R.ref_parse_cli(len(args) + 1, argv, sp, 4096, sn, 1024, ctypes.byref(ns), ctypes.byref(nd))
        return ok, sp.value, sn.value, ns.value, nd.value

    def run_ours(args):
        argv = (ctypes.c_char_p * (len(args) + 2))(b"tracer", *[a.encode() for a in args], None)
        o = pyipt.Cli()
        ok = L.ipt_host_parse_cli(len(args) + 1, argv, ctypes.byref(o))
        return ok, o.scene_path, o.scene_name, o.samples, o.max_depth

    def captured(fn, args):      # both libraries print through their own std::cout: catch file descriptor 1 itself
        import sys
        sys.stdout.flush()
        r, w = os.pipe()
        saved = os.dup(1)
        os.dup2(w, 1)
        os.close(w)
        try:
            res = fn(args)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        os.set_blocking(r, False)
        out = b""
        try:
            while True:
                chunk = os.read(r, 65536)
                if not chunk:
                    break
                out += chunk
        except BlockingIOError:
            pass
        os.close(r)
        return res, out

    rnd = random.Random(7)
    valid = invalid = 0
    for _ in range(600):
        n = rnd.choice([0, 1, 1, 2, 2, 2, 3, 3, 3, 4])
        args = [rnd.choice(keys) + rnd.choice(seps) + rnd.choice(vals) if rnd.random() < 0.9 else rnd.choice(paths) for _ in range(max(0, n - 1))]
        if n >= 1:
            args.append(rnd.choice(paths) if rnd.random() < 0.9 else rnd.choice(keys) + "=" + rnd.choice(vals))
        a, printed_a = captured(run_ref, args)
        b, printed_b = captured(run_ours, args)
        assert a[0] == b[0] and printed_a == printed_b, (args, a, b, printed_a, printed_b)
        if a[0]:
            assert a == b, (args, a, b)
            valid += 1
        else:
            invalid += 1
    assert valid > 20 and invalid > 200, (valid, invalid)


def test_tile_schedule_is_a_balanced_partition(pyipt):
    L = pyipt.lib()
    for world in (1, 2, 3, 4, 8):
        for tiles_x, tiles_y in ((20, 23), (60, 68), (7, 5)):
            owners = np.array([[L.ipt_tile_owner(x, y, tiles_x, world) for x in range(tiles_x)] for y in range(tiles_y)])
            assert owners.min() == 0 and owners.max() == world - 1
            counts = np.bincount(owners.ravel(), minlength=world)
            assert counts.max() - counts.min() <= max(tiles_x, tiles_y)
            if world > 1 and tiles_x >= world:   # every tile row is spread over all ranks
                assert all(len(set(r)) == world for r in owners)


def test_no_device_fails_loudly(pyipt, oracle):
    """Without a CUDA device nothing renders and nothing falls back: IPT_ERR_NO_DEVICE, and `tracer` prints the
    reference's message and exits 0 (CudaUtils.cu:13-17, main.cu:24-27)."""
    if pyipt.lib().ipt_device_count() > 0:
        pytest.skip("a GPU is present")
    hs = pyipt.HostScene.load(oracle.scene_path("spheres"))
    with pytest.raises(pyipt.IptError) as e:
        pyipt.render(hs, 4, 3)
    assert "(-1)" in str(e.value) and "CUDA capable device not found" in str(e.value)
    with pytest.raises(pyipt.IptError):
        pyipt.Context(0)
    r = subprocess.run([os.path.join(ROOT, "improved-path-tracer_b200", "tracer"), oracle.scene_path("spheres")], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == "CUDA capable device not found! Cannot continue"


def test_bvh_traversal_equals_linear_scan_on_cpu(tmp_path):
    """tools/bvh_stats.cpp in check mode: ordered traversal of the host-built BVH (the pruning k_extend_bvh uses) finds
    exactly the nearest distance of a scan over every primitive, for leaf sizes 1, 4 and 16, on 6000 primitives."""
    import subprocess, sys
    sys.path.insert(0, os.path.dirname(__file__))
    from scene_util import synthetic_scene, write_scene
    exe = str(tmp_path / "bvh_stats")
    pkg = os.path.join(ROOT, "improved-path-tracer_b200")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "bvh_stats.cpp"),
                    "-L" + pkg, "-lipt_b200", "-Wl,-rpath," + pkg, "-o", exe], check=True)
    path = write_scene(tmp_path / "syn6000.json", synthetic_scene(6000, seed=11, general_rects=True))
    for leaf in (1, 4, 16):
        r = subprocess.run([exe, path, str(leaf), "3000", "check"], capture_output=True, text=True)
        assert r.returncode == 0 and "check: 0 of 3000" in r.stdout, r.stdout + r.stderr


def test_wide_tree_equals_linear_scan_on_cpu(tmp_path):
    """tools/wide_stats.cpp in check mode: the 8-wide quantised tree of csrc/ipt_wide.h (SAH-optimal collapse and the first,
    greedy one), walked the way k_extend_cw walks it - fp32 arithmetic on the quantised planes, implicit child and primitive
    addressing, octant-ordered hit masks - finds exactly the nearest distance of a scan over every primitive; the explicit
    links of every visited node agree with the implicit addressing; a 2-wide tree with leaves above 4 primitives is refused."""
    import subprocess, sys
    sys.path.insert(0, os.path.dirname(__file__))
    from scene_util import synthetic_scene, write_scene
    exe = str(tmp_path / "wide_stats")
    pkg = os.path.join(ROOT, "improved-path-tracer_b200")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "wide_stats.cpp"),
                    "-L" + pkg, "-lipt_b200", "-Wl,-rpath," + pkg, "-o", exe], check=True)
    path = write_scene(tmp_path / "syn6000.json", synthetic_scene(6000, seed=11, general_rects=True))
    for leaf2, leaf_max, extra in ((4, 4, []), (1, 1, []), (2, 4, []), (4, 2, ["greedy"])):
        r = subprocess.run([exe, path, str(leaf2), str(leaf_max), "3000", "check"] + extra, capture_output=True, text=True)
        assert r.returncode == 0 and "check: 0 of 3000" in r.stdout and " 0 link mismatches" in r.stdout, r.stdout + r.stderr
    r = subprocess.run([exe, path, "8", "4", "10"], capture_output=True, text=True)
    assert r.returncode == 1 and "unsupported" in r.stderr


def test_bvh_depth_is_bounded_on_a_skewed_scene(pyipt, tmp_path):
    """A geometric progression of sizes and positions makes SAH peel one primitive per level; the builder halves ranges by
    index from level 40 on, so the tree stays below the 60 levels ipt_ctx_set_scene accepts (traversal stacks: 64 entries)."""
    import json, sys
    sys.path.insert(0, os.path.dirname(__file__))
    from scene_util import vec
    objs = []
    for i in range(200):
        x = 1.5 ** i
        objs.append({"type": "sphere", "radius": 0.01 * x, "position": vec((x, 0.0, 0.0)), "color": vec((.5, .5, .5)), "emission": vec((0, 0, 0)), "reflection": 0})
    scene = {"width": 64, "height": 36, "camera": {"position": vec((0, -10, 0)), "direction": vec((0, 1, 0)), "orientation": vec((-1, 0, 0))}, "objects": objs}
    path = tmp_path / "skew.json"
    path.write_text(json.dumps(scene))
    hs = pyipt.HostScene.load(str(path), leaf_size=1)
    a = hs.arrays()
    nodes = a["bvh_nodes"]
    assert len(nodes) > 100
    depth = [0] * len(nodes)
    depth[0] = 1
    for i, nd in enumerate(nodes):
        for k in range(2):
            if nd.child[k] >= 0:
                assert nd.child[k] > i                       # parents first: what ipt_ctx_set_scene requires
                depth[nd.child[k]] = depth[i] + 1
    assert 40 <= max(depth) <= 60, max(depth)


def test_uniform_grid_files_every_primitive_under_the_cells_it_touches(pyipt, tmp_path):
    """host/grid.cpp: a lattice of small primitives qualifies for the uniform grid; every small primitive is referenced from
    every cell it reaches (so a ray that walks the cells along its path meets every primitive it can hit), cell
    lists are sorted and start ascending, the walls and the large light are the 'big' list.  A room of large spheres is
    rejected (too many references per primitive) and keeps the tree only; IPT_NO_GRID switches the grid off."""
    import subprocess, sys
    path = str(tmp_path / "syn20k.json")
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "make_synthetic_scene.py"), path, "20000", "64", "36"], check=True)
    hs = pyipt.HostScene.load(path)
    v = hs.view.contents
    res = [v.grid_res[k] for k in range(3)]
    assert all(r > 1 for r in res) and v.n_bvh_nodes > 0 and v.n_grid_big == 7
    n_cells = res[0] * res[1] * res[2]
    start = np.ctypeslib.as_array(v.grid_cell_start, shape=(n_cells + 1,)).astype(np.int64)
    refs = np.ctypeslib.as_array(v.grid_refs, shape=(v.n_grid_refs,)).astype(np.int64)
    big = set(int(v.grid_big[i]) for i in range(v.n_grid_big))
    assert start[0] == 0 and start[-1] == v.n_grid_refs and np.all(np.diff(start) >= 0)
    assert refs.min() >= 0 and refs.max() < v.n_objects and not (set(refs.tolist()) & big)
    a = hs.arrays()
    lo = np.array([v.grid_lo[k] for k in range(3)], dtype=np.float64); cs = np.array([v.grid_cell[k] for k in range(3)], dtype=np.float64)
    slot_prim = a["bvh_slot_prim"]
    rng = np.random.default_rng(0)
    cell_sets = {}
    for slot in rng.choice(v.n_objects, 400, replace=False):
        if int(slot) in big:
            continue
        prim = int(slot_prim[slot])
        if prim & 0x80000000:
            continue                                     # spheres are enough to check the filing rule (boxes from centre and radius)
        c = a["sphere_cxyzr"][prim]
        blo, bhi = c[:3] - abs(c[3]), c[:3] + abs(c[3])
        i0 = np.clip(np.floor((blo - lo) / cs), 0, np.array(res) - 1).astype(int); i1 = np.clip(np.floor((bhi - lo) / cs), 0, np.array(res) - 1).astype(int)
        for z in range(i0[2], i1[2] + 1):
            for y in range(i0[1], i1[1] + 1):
                for x in range(i0[0], i1[0] + 1):
                    # filed under every cell the sphere reaches (centre-to-cell distance <= radius), not the corners of its box
                    ca = lo + np.array([x, y, z]) * cs
                    gap = np.maximum(np.maximum(ca - c[:3], c[:3] - (ca + cs)), 0.0)
                    if float(gap @ gap) > abs(c[3]) ** 2:
                        continue
                    ci = x + res[0] * (y + res[1] * z)
                    if ci not in cell_sets:
                        seg = refs[start[ci]:start[ci + 1]]
                        assert np.all(np.diff(seg) > 0)           # sorted, no duplicates
                        cell_sets[ci] = set(seg.tolist())
                    assert int(slot) in cell_sets[ci], (slot, x, y, z)
    # a room of large spheres: too many references per primitive -> no grid
    sys.path.insert(0, os.path.dirname(__file__))
    from scene_util import synthetic_scene, write_scene
    hs2 = pyipt.HostScene.load(write_scene(tmp_path / "room40k.json", synthetic_scene(40000, 3, width=64, height=36)))
    assert hs2.view.contents.grid_res[0] == 0 and hs2.view.contents.n_bvh_nodes > 0
    os.environ["IPT_NO_GRID"] = "1"
    try:
        hs3 = pyipt.HostScene.load(path)
        assert hs3.view.contents.grid_res[0] == 0 and hs3.view.contents.n_bvh_nodes > 0
    finally:
        del os.environ["IPT_NO_GRID"]
