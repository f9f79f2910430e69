"""ctypes binding of libipt_b200.so (include/ipt_abi.h, include/ipt_host.h).

Plumbing for tests/ and bench.py: every call goes through the C ABI, host buffers in and out, exactly as a C or
C++ caller (or the reference's RenderContoller::start) would use it.  There is no Python or
CPU implementation behind these functions: if the shared library is missing, or no CUDA device is present, the
calls raise.
"""
import ctypes
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(PKG_DIR, "libipt_b200.so")

FLAG_FP64 = 0x1
FLAG_FLOAT_ACCUM = 0x4
FLAG_RUSSIAN_ROULETTE = 0x8
FLAG_STRATIFIED = 0x10
FLAG_NEXT_EVENT = 0x20

ABI_SYMBOLS = [
    "ipt_abi_version", "ipt_device_count", "ipt_device_name", "ipt_last_error", "ipt_render", "ipt_render_rgb8", "ipt_render_objects",
    "ipt_ctx_create", "ipt_ctx_destroy", "ipt_ctx_set_scene", "ipt_ctx_render", "ipt_ctx_download", "ipt_ctx_download_rgb8",
    "ipt_ctx_export_frame", "ipt_ctx_set_gather_target_ipc", "ipt_ctx_set_gather_target", "ipt_tile_owner",
    "ipt_ctx_trace", "ipt_alloc_pinned", "ipt_free_pinned", "ipt_set_progress",
]
HOST_SYMBOLS = [
    "ipt_host_load_scene", "ipt_host_from_objects", "ipt_host_free_scene", "ipt_host_scene_view", "ipt_host_set_size",
    "ipt_host_build_bvh", "ipt_host_to_rgb", "ipt_host_write_png", "ipt_host_write_png_rgb8", "ipt_host_time_string", "ipt_host_append_benchmark",
    "ipt_host_parse_cli",
]


class BvhNode(ctypes.Structure):
    _fields_ = [("lo0", ctypes.c_float * 3), ("hi0", ctypes.c_float * 3), ("lo1", ctypes.c_float * 3),
                ("hi1", ctypes.c_float * 3), ("child", ctypes.c_int32 * 2), ("count", ctypes.c_uint32 * 2)]


class Scene(ctypes.Structure):
    _fields_ = [("width", ctypes.c_uint32), ("height", ctypes.c_uint32),
                ("cam_origin", ctypes.c_double * 3), ("cam_dir", ctypes.c_double * 3), ("cam_orient", ctypes.c_double * 3),
                ("n_objects", ctypes.c_uint32), ("n_spheres", ctypes.c_uint32), ("n_rects", ctypes.c_uint32),
                ("reserved0", ctypes.c_uint32),
                ("sphere_cxyzr", ctypes.POINTER(ctypes.c_double)), ("sphere_object", ctypes.POINTER(ctypes.c_uint32)),
                ("rect_plane", ctypes.POINTER(ctypes.c_double)), ("rect_u", ctypes.POINTER(ctypes.c_double)),
                ("rect_v", ctypes.POINTER(ctypes.c_double)), ("rect_bounds", ctypes.POINTER(ctypes.c_double)),
                ("rect_object", ctypes.POINTER(ctypes.c_uint32)),
                ("mat_color", ctypes.POINTER(ctypes.c_double)), ("mat_emission", ctypes.POINTER(ctypes.c_double)),
                ("mat_reflection", ctypes.POINTER(ctypes.c_int32)),
                ("n_bvh_nodes", ctypes.c_uint32), ("n_bvh_slots", ctypes.c_uint32),
                ("bvh_nodes", ctypes.POINTER(BvhNode)), ("bvh_slot_prim", ctypes.POINTER(ctypes.c_uint32)),
                ("grid_res", ctypes.c_uint32 * 3), ("n_grid_big", ctypes.c_uint32), ("grid_lo", ctypes.c_float * 3),
                ("grid_cell", ctypes.c_float * 3), ("n_grid_refs", ctypes.c_uint32), ("reserved1", ctypes.c_uint32),
                ("grid_cell_start", ctypes.POINTER(ctypes.c_uint32)), ("grid_refs", ctypes.POINTER(ctypes.c_uint32)),
                ("grid_big", ctypes.POINTER(ctypes.c_uint32))]


class Params(ctypes.Structure):
    _fields_ = [("samples", ctypes.c_uint32), ("max_depth", ctypes.c_uint32), ("seed", ctypes.c_uint64),
                ("flags", ctypes.c_uint32), ("tile_w", ctypes.c_uint32), ("tile_h", ctypes.c_uint32),
                ("rank", ctypes.c_uint32), ("world", ctypes.c_uint32), ("batch_samples", ctypes.c_uint32),
                ("reserved", ctypes.c_uint32 * 4)]


class Stats(ctypes.Structure):
    _fields_ = [("samples", ctypes.c_uint64), ("traced_bounces", ctypes.c_uint64), ("kernel_launches", ctypes.c_uint64),
                ("batches", ctypes.c_uint64), ("render_ms", ctypes.c_double), ("upload_ms", ctypes.c_double),
                ("download_ms", ctypes.c_double), ("h2d_bytes", ctypes.c_uint64), ("d2h_bytes", ctypes.c_uint64),
                ("per_gpu_render_ms", ctypes.c_double * 8), ("per_gpu_bounces", ctypes.c_uint64 * 8),
                ("active_pixels", ctypes.c_uint64), ("queue_bytes", ctypes.c_uint64),
                ("node_steps", ctypes.c_uint64), ("box_tests", ctypes.c_uint64), ("leaf_steps", ctypes.c_uint64),
                ("sphere_tests", ctypes.c_uint64), ("rect_tests", ctypes.c_uint64)]

    def as_dict(self):
        return {k: (list(getattr(self, k)) if k.startswith("per_gpu") else getattr(self, k)) for k, _ in self._fields_}


class Cli(ctypes.Structure):
    _fields_ = [("scene_path", ctypes.c_char * 4096), ("scene_name", ctypes.c_char * 1024),
                ("samples", ctypes.c_uint16), ("max_depth", ctypes.c_uint8)]


class IptError(RuntimeError):
    pass


_lib = None


def lib():
    """Loads libipt_b200.so (built by improved-path-tracer_b200/Makefile). Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise IptError(f"{LIB_PATH} is missing: run `make -C improved-path-tracer_b200` (there is no fallback path)")
    L = ctypes.CDLL(LIB_PATH)
    vp, u32, i32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int
    L.ipt_abi_version.restype = i32
    L.ipt_device_count.restype = i32
    L.ipt_device_name.argtypes = [i32]
    L.ipt_device_name.restype = ctypes.c_char_p
    L.ipt_last_error.restype = ctypes.c_char_p
    L.ipt_render.argtypes = [ctypes.POINTER(Scene), ctypes.POINTER(Params), i32, vp, vp, ctypes.POINTER(Stats)]
    L.ipt_render_rgb8.argtypes = [ctypes.POINTER(Scene), ctypes.POINTER(Params), i32, vp, ctypes.POINTER(Stats)]
    L.ipt_render_objects.argtypes = [vp, u32, u32, u32, vp, u32, u32, i32, vp]
    L.ipt_ctx_create.argtypes = [i32]
    L.ipt_ctx_create.restype = vp
    L.ipt_ctx_destroy.argtypes = [vp]
    L.ipt_ctx_destroy.restype = None
    L.ipt_ctx_set_scene.argtypes = [vp, ctypes.POINTER(Scene)]
    L.ipt_ctx_render.argtypes = [vp, ctypes.POINTER(Params), ctypes.POINTER(Stats)]
    L.ipt_alloc_pinned.restype = vp
    L.ipt_alloc_pinned.argtypes = [ctypes.c_size_t]
    L.ipt_free_pinned.restype = None
    L.ipt_free_pinned.argtypes = [vp]
    L.ipt_ctx_download.argtypes = [vp, vp, vp]
    L.ipt_ctx_download_rgb8.argtypes = [vp, vp]
    L.ipt_ctx_export_frame.argtypes = [vp, vp]
    L.ipt_ctx_set_gather_target_ipc.argtypes = [vp, vp]
    L.ipt_ctx_set_gather_target.argtypes = [vp, vp]
    L.ipt_tile_owner.argtypes = [u32, u32, u32, u32]
    L.ipt_tile_owner.restype = u32
    L.ipt_ctx_trace.argtypes = [vp, vp, u32, u32, vp, vp]
    L.ipt_set_progress.argtypes = [vp, vp]
    L.ipt_set_progress.restype = None
    L.ipt_host_load_scene.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
    L.ipt_host_load_scene.restype = vp
    L.ipt_host_from_objects.argtypes = [vp, u32, u32, u32, vp]
    L.ipt_host_from_objects.restype = vp
    L.ipt_host_free_scene.argtypes = [vp]
    L.ipt_host_free_scene.restype = None
    L.ipt_host_scene_view.argtypes = [vp]
    L.ipt_host_scene_view.restype = ctypes.POINTER(Scene)
    L.ipt_host_set_size.argtypes = [vp, u32, u32]
    L.ipt_host_set_size.restype = None
    L.ipt_host_build_bvh.argtypes = [vp, u32, u32]
    L.ipt_host_to_rgb.argtypes = [ctypes.c_double]
    L.ipt_host_write_png.argtypes = [ctypes.c_char_p, vp, u32, u32]
    L.ipt_host_write_png_rgb8.argtypes = [ctypes.c_char_p, vp, u32, u32]
    L.ipt_host_time_string.argtypes = [ctypes.c_uint64, ctypes.c_char_p, ctypes.c_size_t]
    L.ipt_host_time_string.restype = None
    L.ipt_host_append_benchmark.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
    L.ipt_host_parse_cli.argtypes = [i32, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(Cli)]
    _lib = L
    return L


def _check(rc, what):
    if rc != 0:
        raise IptError(f"{what} failed ({rc}): {lib().ipt_last_error().decode(errors='replace')}")


def make_params(samples, depth, seed=123456, flags=0, tile=(0, 0), rank=0, world=1, batch=0):
    p = Params()
    p.samples, p.max_depth, p.seed, p.flags = samples, depth, seed, flags
    p.tile_w, p.tile_h, p.rank, p.world, p.batch_samples = tile[0], tile[1], rank, world, batch
    return p


class HostScene:
    """A scene loaded and flattened by the C++ host layer (ipt_host_load_scene / ipt_host_from_objects)."""

    def __init__(self, handle):
        self.handle = handle

    @staticmethod
    def load(path, width=None, height=None, leaf_size=4, brute_max=192):
        msg = ctypes.create_string_buffer(256)
        h = lib().ipt_host_load_scene(os.fsencode(path), msg, 256)
        if not h:
            raise IptError(msg.value.decode())
        s = HostScene(h)
        if width and height:
            lib().ipt_host_set_size(h, width, height)
        s.build_bvh(leaf_size, brute_max)
        return s

    @staticmethod
    def from_objects(objects_bytes, n, width, height, camera9, leaf_size=4, brute_max=192):
        cam = (ctypes.c_double * 9)(*camera9)
        buf = ctypes.create_string_buffer(bytes(objects_bytes), len(objects_bytes))
        h = lib().ipt_host_from_objects(buf, n, width, height, cam)
        if not h:
            raise IptError("ipt_host_from_objects failed")
        s = HostScene(h)
        s.build_bvh(leaf_size, brute_max)
        return s

    def build_bvh(self, leaf_size=4, brute_max=192):
        n = lib().ipt_host_build_bvh(self.handle, leaf_size, brute_max)
        if n < 0:
            raise IptError("ipt_host_build_bvh failed")
        return n

    @property
    def view(self):
        return lib().ipt_host_scene_view(self.handle)

    @property
    def width(self):
        return self.view.contents.width

    @property
    def height(self):
        return self.view.contents.height

    def arrays(self):
        """The flattened arrays as numpy copies (for host-logic tests)."""
        v = self.view.contents

        def arr(ptr, n, dt):
            return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt).copy() if n else np.zeros(0, dt)

        return {
            "sphere_cxyzr": arr(v.sphere_cxyzr, v.n_spheres * 4, np.float64).reshape(-1, 4),
            "sphere_object": arr(v.sphere_object, v.n_spheres, np.uint32),
            "rect_plane": arr(v.rect_plane, v.n_rects * 4, np.float64).reshape(-1, 4),
            "rect_u": arr(v.rect_u, v.n_rects * 4, np.float64).reshape(-1, 4),
            "rect_v": arr(v.rect_v, v.n_rects * 4, np.float64).reshape(-1, 4),
            "rect_bounds": arr(v.rect_bounds, v.n_rects * 4, np.float64).reshape(-1, 4),
            "rect_object": arr(v.rect_object, v.n_rects, np.uint32),
            "mat_color": arr(v.mat_color, v.n_objects * 3, np.float64).reshape(-1, 3),
            "mat_emission": arr(v.mat_emission, v.n_objects * 3, np.float64).reshape(-1, 3),
            "mat_reflection": arr(v.mat_reflection, v.n_objects, np.int32),
            "n_bvh_nodes": v.n_bvh_nodes,
            "bvh_slot_prim": arr(v.bvh_slot_prim, v.n_bvh_slots, np.uint32),
            "bvh_nodes": [v.bvh_nodes[i] for i in range(min(v.n_bvh_nodes, 1 << 22))],
        }

    def close(self):
        if self.handle:
            lib().ipt_host_free_scene(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def render(scene, samples, depth, n_gpus=1, seed=123456, flags=0, tile=(0, 0), batch=0, want64=True):
    """One-shot ipt_render(): host scene in, host frame out. Returns (image [H,W,3], stats dict)."""
    v = scene.view.contents
    out32 = np.zeros((v.height, v.width, 3), dtype=np.float32)
    out64 = np.zeros((v.height, v.width, 3), dtype=np.float64) if want64 else None
    p = make_params(samples, depth, seed, flags, tile, 0, 1, batch)
    st = Stats()
    rc = lib().ipt_render(scene.view, ctypes.byref(p), n_gpus, out32.ctypes.data, out64.ctypes.data if want64 else None,
                          ctypes.byref(st))
    _check(rc, "ipt_render")
    return (out64 if want64 else out32), st.as_dict()


def render_rgb8(scene, samples, depth, n_gpus=1, seed=123456, flags=0):
    """One-shot ipt_render_rgb8(): toRgb applied on the device, uint8 frame [H,W,3] back."""
    v = scene.view.contents
    out = np.zeros((v.height, v.width, 3), dtype=np.uint8)
    p = make_params(samples, depth, seed, flags)
    st = Stats()
    _check(lib().ipt_render_rgb8(scene.view, ctypes.byref(p), n_gpus, out.ctypes.data, ctypes.byref(st)), "ipt_render_rgb8")
    return out, st.as_dict()


class PinnedArray:
    """numpy view of a page-locked host buffer from ipt_alloc_pinned (freed with the object)."""

    def __init__(self, shape, dtype=np.float32):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = lib().ipt_alloc_pinned(n)
        if not self.ptr:
            raise IptError("ipt_alloc_pinned: " + lib().ipt_last_error().decode(errors="replace"))
        buf = (ctypes.c_char * n).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def close(self):
        if self.ptr:
            self.array = None
            lib().ipt_free_pinned(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """Resident per-GPU context (ipt_ctx_*)."""

    def __init__(self, device=0):
        self.h = lib().ipt_ctx_create(device)
        if not self.h:
            raise IptError("ipt_ctx_create: " + lib().ipt_last_error().decode(errors="replace"))
        self.scene = None

    def set_scene(self, scene):
        _check(lib().ipt_ctx_set_scene(self.h, scene.view), "ipt_ctx_set_scene")
        self.scene = scene

    def render(self, samples, depth, seed=123456, flags=0, tile=(0, 0), rank=0, world=1, batch=0):
        p = make_params(samples, depth, seed, flags, tile, rank, world, batch)
        st = Stats()
        _check(lib().ipt_ctx_render(self.h, ctypes.byref(p), ctypes.byref(st)), "ipt_ctx_render")
        return st.as_dict()

    def download(self, want64=True, out=None):
        v = self.scene.view.contents
        if out is None:
            out = np.zeros((v.height, v.width, 3), dtype=np.float64 if want64 else np.float32)
        if out.dtype == np.float64:
            _check(lib().ipt_ctx_download(self.h, None, out.ctypes.data), "ipt_ctx_download")
        else:
            _check(lib().ipt_ctx_download(self.h, out.ctypes.data, None), "ipt_ctx_download")
        return out

    def download_rgb8(self):
        v = self.scene.view.contents
        out = np.zeros((v.height, v.width, 3), dtype=np.uint8)
        _check(lib().ipt_ctx_download_rgb8(self.h, out.ctypes.data), "ipt_ctx_download_rgb8")
        return out

    def trace(self, rays, flags=0):
        rays = np.ascontiguousarray(rays, dtype=np.float64)
        n = rays.shape[0]
        obj = np.zeros(n, dtype=np.int32)
        t = np.zeros(n, dtype=np.float64)
        _check(lib().ipt_ctx_trace(self.h, rays.ctypes.data, n, flags, obj.ctypes.data, t.ctypes.data), "ipt_ctx_trace")
        return obj, t

    def export_frame(self):
        buf = ctypes.create_string_buffer(64)
        _check(lib().ipt_ctx_export_frame(self.h, buf), "ipt_ctx_export_frame")
        return buf.raw

    def set_gather_target_ipc(self, handle):
        buf = ctypes.create_string_buffer(handle, 64)
        _check(lib().ipt_ctx_set_gather_target_ipc(self.h, buf), "ipt_ctx_set_gather_target_ipc")

    def set_gather_target(self, owner):
        _check(lib().ipt_ctx_set_gather_target(self.h, owner.h), "ipt_ctx_set_gather_target")

    def close(self):
        if self.h:
            lib().ipt_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
