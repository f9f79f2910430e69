"""TEST INFRASTRUCTURE ONLY — ctypes access to the CPU oracle (oracle/liboracle.so) and to the
reference's own code compiled for the host (oracle/_ref/libref_host.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Scene files are read with Python's json module here, independently of the product's C++ loader.
"""
import ctypes
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SCENES_DIR = os.path.join(REF_DIR, "scenes")

RNG_REFERENCE = 0
RNG_COUNTER = 1


class OrObject(ctypes.Structure):
    # the reference's ObjectData layout (ObjectData.hpp:15-31), 144 bytes
    _fields_ = [("type", ctypes.c_int32), ("pad0", ctypes.c_int32), ("radius", ctypes.c_double),
                ("north", ctypes.c_double * 3), ("east", ctypes.c_double * 3), ("position", ctypes.c_double * 3),
                ("emission", ctypes.c_double * 3), ("color", ctypes.c_double * 3),
                ("reflection", ctypes.c_int32), ("pad1", ctypes.c_int32)]


class OrScene(ctypes.Structure):
    _fields_ = [("width", ctypes.c_uint32), ("height", ctypes.c_uint32), ("camera", ctypes.c_double * 9),
                ("n_objects", ctypes.c_uint32), ("pad", ctypes.c_uint32), ("objects", ctypes.POINTER(OrObject))]


class OrCounts(ctypes.Structure):
    _fields_ = [("samples", ctypes.c_uint64), ("casts_reference", ctypes.c_uint64), ("casts_needed", ctypes.c_uint64)]


assert ctypes.sizeof(OrObject) == 144


def scene_path(name):
    """Path of one of the reference's benchmark scenes (copied to oracle/_ref/scenes by oracle/Makefile)."""
    p = name if os.path.isfile(name) else os.path.join(SCENES_DIR, name if name.endswith(".json") else name + ".json")
    if not os.path.isfile(p):
        raise FileNotFoundError(f"{p}: run `make -C oracle` where /root/reference is mounted")
    return p


def _vec(d):
    return (float(d["xx"]), float(d["yy"]), float(d["zz"]))


class Scene:
    """A scenes/*.json file as Python data (schema: SceneData.cpp:98-225)."""

    def __init__(self, width, height, cam_pos, cam_dir, cam_orient, objects):
        self.width, self.height = int(width), int(height)
        self.cam_pos = np.asarray(cam_pos, dtype=np.float64)
        d = np.asarray(cam_dir, dtype=np.float64)
        o = np.asarray(cam_orient, dtype=np.float64)
        # SceneData.cpp:143-145 normalises direction and orientation at load: v * (1/sqrt(v.v))
        self.cam_dir = d * (1.0 / np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))
        self.cam_orient = o * (1.0 / np.sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2]))
        self.objects = objects  # list of dicts: type, radius, north, east, position, emission, color, reflection

    @staticmethod
    def load(path, width=None, height=None):
        with open(scene_path(path)) as f:
            j = json.load(f)
        objs = []
        for o in j["objects"]:
            e = {"type": 0 if o["type"] == "sphere" else 1, "radius": float(o.get("radius", 0.0)),
                 "north": _vec(o["north"]) if "north" in o else (0.0, 0.0, 0.0),
                 "east": _vec(o["east"]) if "east" in o else (0.0, 0.0, 0.0),
                 "position": _vec(o["position"]), "emission": _vec(o["emission"]), "color": _vec(o["color"]),
                 "reflection": int(o["reflection"])}
            objs.append(e)
        c = j["camera"]
        return Scene(width or j["width"], height or j["height"], _vec(c["position"]), _vec(c["direction"]),
                     _vec(c["orientation"]), objs)

    def c_scene(self):
        arr = (OrObject * len(self.objects))()
        for i, o in enumerate(self.objects):
            a = arr[i]
            a.type, a.radius, a.reflection = o["type"], o["radius"], o["reflection"]
            for k in ("north", "east", "position", "emission", "color"):
                getattr(a, k)[:] = o[k]
        s = OrScene()
        s.width, s.height = self.width, self.height
        s.camera[:] = list(self.cam_pos) + list(self.cam_dir) + list(self.cam_orient)
        s.n_objects = len(self.objects)
        s.objects = arr
        s._keep = arr
        return s


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        p = os.path.join(HERE, "liboracle.so")
        if not os.path.isfile(p):
            raise FileNotFoundError(f"{p}: run `make -C oracle liboracle.so`")
        L = ctypes.CDLL(p)
        L.or_render.argtypes = [ctypes.POINTER(OrScene), ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_uint64,
                                ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(OrCounts)]
        L.or_render.restype = ctypes.c_int
        L.or_nearest_hit.argtypes = [ctypes.POINTER(OrScene), ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]
        L.or_nearest_hit.restype = None
        L.or_render_accel.argtypes = L.or_render.argtypes
        L.or_render_accel.restype = ctypes.c_int
        L.or_nearest_hit_accel.argtypes = L.or_nearest_hit.argtypes
        L.or_nearest_hit_accel.restype = None
        dp = ctypes.POINTER(ctypes.c_double)
        L.or_sphere_intersect.argtypes = [ctypes.c_double, dp, dp, dp]
        L.or_sphere_intersect.restype = ctypes.c_double
        L.or_plane_intersect.argtypes = [dp, dp, dp, dp, dp]
        L.or_plane_intersect.restype = ctypes.c_double
        L.or_scatter.argtypes = [ctypes.c_int, dp, ctypes.c_int, dp, dp, ctypes.c_int, ctypes.c_ulonglong, dp]
        L.or_scatter.restype = None
        L.or_philox4x32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        for f in (L.or_sym24, L.or_uniform23):
            f.argtypes = [ctypes.c_uint32]
            f.restype = ctypes.c_double
        L.or_to_rgb.argtypes = [ctypes.c_double]
        L.or_to_rgb.restype = ctypes.c_int
        _lib = L
    return _lib


def ref_available():
    return os.path.isfile(os.path.join(REF_DIR, "libref_host.so"))


def ref():
    """The reference's own Renderer.cu compiled for the host (Route 1)."""
    global _ref
    if _ref is None:
        L = ctypes.CDLL(os.path.join(REF_DIR, "libref_host.so"))
        L.ref_render_cells.argtypes = [ctypes.c_char_p] + [ctypes.c_int] * 7 + [ctypes.c_void_p]
        L.ref_render_cells.restype = ctypes.c_int
        L.ref_render_cell_list.argtypes = [ctypes.c_char_p] + [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.ref_render_cell_list.restype = ctypes.c_int
        L.ref_scene_dims.argtypes = [ctypes.c_char_p] + [ctypes.POINTER(ctypes.c_int)] * 3
        L.ref_num_cells.argtypes = [ctypes.c_int, ctypes.c_int]
        dp = ctypes.POINTER(ctypes.c_double)
        L.ref_sphere_intersect.argtypes = [ctypes.c_double, dp, dp, dp]
        L.ref_sphere_intersect.restype = ctypes.c_double
        L.ref_plane_intersect.argtypes = [dp, dp, dp, dp, dp]
        L.ref_plane_intersect.restype = ctypes.c_double
        L.ref_scatter.argtypes = [ctypes.c_int, dp, ctypes.c_int, dp, dp, ctypes.c_int, ctypes.c_ulonglong, dp]
        L.ref_scatter.restype = None
        L.ref_xorwow_kat.argtypes = [ctypes.c_ulonglong, ctypes.c_void_p, ctypes.c_void_p]
        if hasattr(L, "ref_scene_load_text"):
            L.ref_scene_load_text.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int]
            L.ref_scene_load_text.restype = ctypes.c_int
        if hasattr(L, "ref_scene_objects"):
            L.ref_scene_objects.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
            L.ref_scene_objects.restype = ctypes.c_int
        _ref = L
    return _ref


def render(scene, samples, depth, rng=RNG_COUNTER, seed=0, begin=0, end=-1, nthreads=None, out=None, accel=False):
    """Oracle render. Returns (image float64 [H,W,3], counts dict).  accel=True: nearest hits through the oracle's own box
    tree (SURVEY.md §8c Route 3, for scenes the reference's scan cannot finish) - bit-identical frames, see restate.cpp."""
    cs = scene.c_scene()
    if out is None:
        out = np.zeros((scene.height, scene.width, 3), dtype=np.float64)
    cnt = OrCounts()
    fn = lib().or_render_accel if accel else lib().or_render
    rc = fn(ctypes.byref(cs), samples, depth, rng, seed, begin, end, nthreads or os.cpu_count() or 1,
            out.ctypes.data, ctypes.byref(cnt))
    if rc != 0:
        raise RuntimeError("or_render failed")
    return out, {"samples": cnt.samples, "casts_reference": cnt.casts_reference, "casts_needed": cnt.casts_needed}


def nearest_hit(scene, rays, accel=False):
    """rays: float64 [n,6] (origin, direction) -> (index int32 [n], t float64 [n]); Renderer.cu:227-243.
    accel=True: the same through the oracle's own box tree (Route 3)."""
    cs = scene.c_scene()
    rays = np.ascontiguousarray(rays, dtype=np.float64)
    n = rays.shape[0]
    idx = np.zeros(n, dtype=np.int32)
    t = np.zeros(n, dtype=np.float64)
    (lib().or_nearest_hit_accel if accel else lib().or_nearest_hit)(ctypes.byref(cs), rays.ctypes.data, n, idx.ctypes.data, t.ctypes.data)
    return idx, t


def ref_scene_objects(path):
    """The reference's own in-memory scene, byte for byte: (objects bytes [n*144], camera bytes [72], W, H, n) as
    SceneData::getObjectsData() / getCamera() hold them (what RenderContoller::start() uploads)."""
    path = scene_path(path) if not os.path.isfile(path) else path
    W, H = ctypes.c_int(), ctypes.c_int()
    n = ref().ref_scene_objects(os.fsencode(path), None, 0, None, ctypes.byref(W), ctypes.byref(H))
    if n < 0:
        raise RuntimeError(f"reference SceneData could not load {path}")
    objs = ctypes.create_string_buffer(144 * n)
    cam = ctypes.create_string_buffer(72)
    ref().ref_scene_objects(os.fsencode(path), objs, n, cam, ctypes.byref(W), ctypes.byref(H))
    return objs.raw, cam.raw, W.value, H.value, n


def ref_scene_load_texts(paths):
    """For every file: (number of objects or -1, what SceneData::initScene printed), or None where the reference's own loader
    aborts the process (nlohmann's const operator[] asserts on a missing key: undefined behaviour upstream).  Runs in child
    processes for that reason; a child that dies is restarted behind the file that killed it."""
    import json, subprocess, sys
    child = ("import sys, json, ctypes\n"
             "sys.path.insert(0, %r)\n"
             "import oracle as O\n"
             "buf = ctypes.create_string_buffer(8192)\n"
             "for p in json.loads(sys.stdin.read()):\n"
             "    n = O.ref().ref_scene_load_text(p.encode(), buf, 8192)\n"
             "    print(json.dumps([n, buf.value.decode(errors='replace')]), flush=True)\n") % HERE
    out, at = [], 0
    paths = [os.fspath(p) for p in paths]
    while at < len(paths):
        r = subprocess.run([sys.executable, "-c", child], input=json.dumps(paths[at:]), capture_output=True, text=True)
        lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("[")]
        out += [tuple(l) for l in lines]
        at += len(lines)
        if at < len(paths) and r.returncode != 0:
            out.append(None)                      # the reference died on paths[at]
            at += 1
        elif at < len(paths):
            raise RuntimeError("reference loader child stopped early: " + r.stderr[-500:])
    return out


def ref_render(path, samples, depth, width=0, height=0, cell_begin=0, cell_end=-1, nthreads=None):
    """The reference's own routine (host-compiled). Returns image float64 [H,W,3]."""
    p = scene_path(path).encode()
    W, H, N = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    if ref().ref_scene_dims(p, ctypes.byref(W), ctypes.byref(H), ctypes.byref(N)) != 0:
        raise RuntimeError("reference could not load " + path)
    w, h = width or W.value, height or H.value
    out = np.zeros((h, w, 3), dtype=np.float64)
    rc = ref().ref_render_cells(p, samples, depth, width, height, cell_begin, cell_end, nthreads or os.cpu_count() or 1,
                                out.ctypes.data)
    if rc != 0:
        raise RuntimeError("ref_render_cells failed")
    return out


def ref_time_cells(path, samples, depth, width, height, cells, nthreads=None):
    """Wall time of the reference routine over the given thread cells (no image returned): bench.py's CPU baseline."""
    import time
    arr = (ctypes.c_int * len(cells))(*cells)
    t0 = time.perf_counter()
    rc = ref().ref_render_cell_list(scene_path(path).encode(), samples, depth, width or 0, height or 0, arr, len(cells),
                                    nthreads or os.cpu_count() or 1, None)
    if rc != 0:
        raise RuntimeError("ref_render_cell_list failed")
    return time.perf_counter() - t0


def philox(counter, key, rounds=0):
    """Philox4x32 with `rounds` rounds (0: the 7 rounds of the counter stream)."""
    c = (ctypes.c_uint32 * 4)(*counter)
    k = (ctypes.c_uint32 * 2)(*key)
    o = (ctypes.c_uint32 * 4)()
    lib().or_philox4x32(c, k, rounds, o)
    return tuple(o)
