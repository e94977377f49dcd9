"""ctypes binding of librtb200.so (include/rtb200.h) plus a thin object wrapper.

This is plumbing for tests and bench.py: the product is the shared library and the C++ facade in
include/igx_rt.hpp.  Nothing here computes; if the library is missing the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RTB_LIB", os.path.join(HERE, "librtb200.so"))   # RTB_LIB: load a tuning variant built by build.py

# enums of rtb200.h
(BUF_CAMERA, BUF_SEED, BUF_SCENE_INFO, BUF_SHADOW_PROPS, BUF_TRIANGLES, BUF_SPHERES, BUF_CUBES, BUF_PLANES, BUF_LIGHTS,
 BUF_MATERIALS, BUF_MATERIAL_INDICES) = range(11)
PASS_INIT, PASS_RAYGEN, PASS_SHADOW, PASS_LIGHTING, PASS_COMPOSITE, PASS_FRAME = range(6)
TGT_DIR_T, TGT_UV_NORMAL, TGT_SHADOW_BITS, TGT_LIGHTING, TGT_ACCUM, TGT_RGBA8, TGT_SEED, TGT_RGBA8_TILED, TGT_ACCEL_NODES, TGT_ACCEL_TRIANGLES = range(10)
ACCEL_BRUTE, ACCEL_BVH, ACCEL_BVH2 = 0, 1, 2
OPT_COUNTERS, OPT_TILE_RANK, OPT_TILE_COUNT, OPT_SHADER_BUILD, OPT_PRIMARY_PACKETS, OPT_FUSE_PRIMARY, OPT_SHADOW_ORDER, OPT_ACCEL_BUILDER, OPT_FRAME_LANES, OPT_FRAME_GRAPH, OPT_LIGHTS, OPT_HISTORY_ALPHA, OPT_PRIMITIVE_TREES, OPT_FRAME_OVERLAP, OPT_LIGHT_CACHE = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14
BUILDER_HOST, BUILDER_DEVICE = 0, 1
SHADER_DEBUG, SHADER_RELEASE = 0, 1
NO_RAY_HIT = 0xFFFFFFFF
NO_HIT = np.float32(3.4028235e38)

EXPORTS = [
    "rtb_create", "rtb_destroy", "rtb_last_error", "rtb_set_option", "rtb_set_stream", "rtb_resize", "rtb_upload",
    "rtb_upload_skybox", "rtb_build_accel", "rtb_refit_accel", "rtb_accel_info_get", "rtb_dispatch", "rtb_readback", "rtb_readback_async", "rtb_readback_wait", "rtb_device_ptr", "rtb_sync",
    "rtb_counters_get", "rtb_probe_l2_read_gbs", "rtb_last_frame_ms", "rtb_trace_rays", "rtb_occlusion_rays", "rtb_untile", "rtb_untile_on", "rtb_present_host", "rtb_path_frame", "rtb_path_stats_get", "rtb_pack_triangle",
    "rtb_pack_light_directional", "rtb_pack_light_point", "rtb_pack_material", "rtb_pack_camera", "rtb_load_hdr", "rtb_write_png",
    "rtb_gen_soup", "rtb_gen_heightfield",
]


class PathStats(C.Structure):
    _fields_ = [("depths", C.c_uint32), ("closest_launches", C.c_uint32), ("shadow_launches", C.c_uint32), ("kernel_launches", C.c_uint32),
                ("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("closest_rays_at_depth", C.c_uint64 * 16), ("shadow_rays_at_depth", C.c_uint64 * 16),
                ("closest_ms", C.c_float), ("shadow_ms", C.c_float), ("total_ms", C.c_float), ("closest_ms_at_depth", C.c_float * 16), ("shadow_ms_at_depth", C.c_float * 16)]

    def as_dict(self):
        d = int(self.depths)
        return dict(depths=d, kernel_launches=int(self.kernel_launches), closest_rays=int(self.closest_rays), shadow_rays=int(self.shadow_rays),
                    closest_rays_at_depth=[int(v) for v in self.closest_rays_at_depth[:d]], shadow_rays_at_depth=[int(v) for v in self.shadow_rays_at_depth[:d]],
                    closest_ms=float(self.closest_ms), shadow_ms=float(self.shadow_ms), total_ms=float(self.total_ms),
                    closest_ms_at_depth=[float(v) for v in self.closest_ms_at_depth[:d]], shadow_ms_at_depth=[float(v) for v in self.shadow_ms_at_depth[:d]])


class Limits(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("max_triangles", "max_spheres", "max_cubes", "max_planes", "max_lights", "max_materials")]


class AccelInfo(C.Structure):
    _fields_ = [("mode", C.c_uint32), ("node_count", C.c_uint32), ("node_bytes", C.c_uint32), ("leaf_count", C.c_uint32),
                ("max_depth", C.c_uint32), ("tri_record_bytes", C.c_uint32), ("sah_cost", C.c_float), ("build_ms", C.c_float),
                ("leaf_node_extent", C.c_float), ("refits", C.c_uint32), ("primary_packets", C.c_uint32), ("builder", C.c_uint32),
                ("sphere_tree_nodes", C.c_uint32), ("cube_tree_nodes", C.c_uint32)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("primary_rays", "shadow_rays", "primary_nodes", "primary_tris", "shadow_nodes",
                                          "shadow_tris", "primary_hits", "shadow_occluded")]


class RtbError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Loads librtb200.so.  Raises if it was not built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RtbError(f"{LIB_PATH} is missing: run `python -m igx_raytracing_b200.build` (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        vp, u32, u64, sz, f = C.c_void_p, C.c_uint32, C.c_uint64, C.c_size_t, C.c_float
        L.rtb_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(Limits)]
        L.rtb_destroy.argtypes = [vp]; L.rtb_destroy.restype = None
        L.rtb_last_error.argtypes = [vp]; L.rtb_last_error.restype = C.c_char_p
        L.rtb_set_option.argtypes = [vp, C.c_int, u32]
        L.rtb_set_stream.argtypes = [vp, vp]
        L.rtb_resize.argtypes = [vp, u32, u32, u32]
        L.rtb_upload.argtypes = [vp, C.c_int, sz, sz, vp]
        L.rtb_upload_skybox.argtypes = [vp, u32, u32, vp]
        L.rtb_build_accel.argtypes = [vp, C.c_int]
        L.rtb_refit_accel.argtypes = [vp]
        L.rtb_accel_info_get.argtypes = [vp, C.POINTER(AccelInfo)]
        L.rtb_dispatch.argtypes = [vp, C.c_int]
        L.rtb_readback.argtypes = [vp, C.c_int, vp, sz]
        L.rtb_readback_async.argtypes = [vp, C.c_int, vp, sz]
        L.rtb_readback_wait.argtypes = [vp]
        L.rtb_device_ptr.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(sz)]
        L.rtb_sync.argtypes = [vp]
        L.rtb_counters_get.argtypes = [vp, C.POINTER(Counters)]
        L.rtb_last_frame_ms.argtypes = [vp, C.POINTER(f * 8)]
        L.rtb_trace_rays.argtypes = [vp, vp, u64, vp, vp, vp, vp]
        L.rtb_occlusion_rays.argtypes = [vp, vp, u64, vp, vp, vp]
        L.rtb_untile.argtypes = [vp, vp, u32, u32, vp]
        L.rtb_untile_on.argtypes = [vp, vp, u32, u32, vp, vp]
        L.rtb_present_host.argtypes = [vp, vp, vp, vp]
        L.rtb_path_frame.argtypes = [vp, u32]
        L.rtb_path_stats_get.argtypes = [vp, vp]
        L.rtb_pack_triangle.argtypes = [vp, vp, vp]; L.rtb_pack_triangle.restype = None
        L.rtb_pack_light_directional.argtypes = [vp, vp, f, vp]; L.rtb_pack_light_directional.restype = None
        L.rtb_pack_light_point.argtypes = [vp, vp, f, f, f, vp]; L.rtb_pack_light_point.restype = None
        L.rtb_pack_material.argtypes = [vp, vp, vp, f, f, f, vp]; L.rtb_pack_material.restype = None
        L.rtb_pack_camera.argtypes = [vp, f, f, f, f, f, f, u32, u32, u32, u32, f, vp, vp]; L.rtb_pack_camera.restype = None
        L.rtb_load_hdr.argtypes = [C.c_char_p, vp, C.POINTER(u32), C.POINTER(u32)]
        L.rtb_gen_soup.argtypes = [u64, u64, vp]; L.rtb_gen_soup.restype = None
        L.rtb_gen_heightfield.argtypes = [u32, u64, vp]; L.rtb_gen_heightfield.restype = None
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---- host-side packing (no GPU needed) -------------------------------------------------------------------
def pack_triangle(p, normals=None) -> np.ndarray:
    p = np.ascontiguousarray(p, np.float32).reshape(9)
    n = None if normals is None else np.ascontiguousarray(normals, np.float32).reshape(9)
    out = np.zeros(48, np.uint8)
    lib().rtb_pack_triangle(_p(p), _p(n), _p(out))
    return out


def pack_light_directional(direction, color, angular_extent=None) -> np.ndarray:
    if angular_extent is None:
        angular_extent = np.float32(0.533 * (3.141592653589793 / 180))
    d, c = np.asarray(direction, np.float32), np.asarray(color, np.float32)
    out = np.zeros(32, np.uint8)
    lib().rtb_pack_light_directional(_p(d), _p(c), np.float32(angular_extent), _p(out))
    return out


def pack_light_point(pos, color, rad, origin, specularity=1.0) -> np.ndarray:
    p, c = np.asarray(pos, np.float32), np.asarray(color, np.float32)
    out = np.zeros(32, np.uint8)
    lib().rtb_pack_light_point(_p(p), _p(c), np.float32(rad), np.float32(origin), np.float32(specularity), _p(out))
    return out


def pack_material(albedo, ambient, emission, metallic, roughness, transparency=1.0) -> np.ndarray:
    a, b, e = (np.asarray(v, np.float32) for v in (albedo, ambient, emission))
    out = np.zeros(32, np.uint8)
    lib().rtb_pack_material(_p(a), _p(b), _p(e), np.float32(metallic), np.float32(roughness), np.float32(transparency), _p(out))
    return out


def pack_camera(width, height, eye=(4, 2, -2), pitch=0.0, yaw=0.0, roll=0.0, left_fov=70.0, right_fov=70.0, ipd=62.0,
                projection=0, flags=0, exposure=1.0, skybox_color=(0.25, 0.5, 1.0)) -> np.ndarray:
    e, sc = np.asarray(eye, np.float32), np.asarray(skybox_color, np.float32)
    out = np.zeros(144, np.uint8)
    lib().rtb_pack_camera(_p(e), pitch, yaw, roll, left_fov, right_fov, ipd, projection, width, height, flags, exposure, _p(sc), _p(out))
    return out


def make_seed(cpu_offset=(0.0, 0.0), sample_count=0, sample_offset=0, random=(0.0, 0.0)) -> np.ndarray:
    s = np.zeros(24, np.uint8)
    s[:16].view(np.float32)[:] = [random[0], random[1], cpu_offset[0], cpu_offset[1]]
    s[16:].view(np.uint32)[:] = [sample_count, sample_offset]
    return s


def write_png(path, rgba8, flip_vertically=True):
    """rgba8: (h, w) uint32 or (h, w, 4) uint8 frame as read back from TGT_RGBA8 (row 0 = bottom of the view)."""
    a = np.ascontiguousarray(rgba8)
    h, w = a.shape[0], a.shape[1]
    lib().rtb_write_png.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int]
    rc = lib().rtb_write_png(str(path).encode(), w, h, _p(a), 1 if flip_vertically else 0)
    if rc:
        raise IOError(f"rtb_write_png({path}): error {rc}")


def load_hdr(path: str) -> np.ndarray:
    w, h = C.c_uint32(0), C.c_uint32(0)
    rc = lib().rtb_load_hdr(path.encode(), None, C.byref(w), C.byref(h))
    if rc:
        raise IOError(f"rtb_load_hdr({path}): error {rc}")
    out = np.zeros((h.value, w.value, 4), np.uint16)
    rc = lib().rtb_load_hdr(path.encode(), _p(out), C.byref(w), C.byref(h))
    if rc:
        raise IOError(f"rtb_load_hdr({path}): error {rc}")
    return out


def gen_soup(n: int, seed: int = 0xB200) -> np.ndarray:
    out = np.zeros(n * 48, np.uint8)
    lib().rtb_gen_soup(n, seed, _p(out))
    return out


def gen_heightfield(grid: int, seed: int = 0xB200) -> np.ndarray:
    out = np.zeros(2 * grid * grid * 48, np.uint8)
    lib().rtb_gen_heightfield(grid, seed, _p(out))
    return out


def shadow_words(w, h, samples):
    return ((w + 15) // 16) * ((h + 1) // 2) * samples


def niels_scene(time: float = 0.0) -> dict:
    """The reference's default scene (test/scene/niels_scene.cpp:5-69) as raw GPU-layout buffers, in the order
    SceneGraph::update leaves them (lights directional < spot < point; object ids tri, sphere, cube, plane)."""
    import math
    mats = [((1, .5, 1), (.05, .01, .05), (0, 0, 0), 0, 1), ((0, 1, 0), (0, .05, 0), (0, 0, 0), 0, 1),
            ((0, 0, 1), (0, 0, .05), (0, 0, 0), 0, 1), ((1, 0, 1), (.05, 0, .05), (0, 0, 0), 0, 1),
            ((1, 1, 0), (.05, .05, 0), (0, 0, 0), 0, 1), ((0, 1, 1), (0, .05, .05), (0, 0, 0), 0, 1),
            ((0, 0, 0), (0, 0, 0), (0, 0, 0), 1, 0), ((0, 0, 0), (0, 0, 0), (0, 0, 0), .25, .5)]
    materials = np.concatenate([pack_material(a, b, e, m, r, 1.0) for a, b, e, m, r in mats])
    tris = np.concatenate([pack_triangle(p) for p in ([1, 1, 0, -1, 1, 0, 1, 0, 1], [-1, 4, 0, 1, 4, 0, 1, 3, 1], [-1, 7, 0, 1, 7, 0, 1, 5, 1])])
    t, c = np.float32(math.sin(time)), np.float32(math.cos(time))
    spheres = np.array([[0, 1, 5, 1], [0, 1, -5, 1], [3, 1, 0, 1], [0, 6, 0, 1],
                        [7, np.float32(2) + np.float32(math.sin(0.0)), 0, 1], [np.float32(-5) + t, np.float32(2) + c, 0, 1],
                        [t, np.float32(3) + c, 0, 1]], np.float32)
    cubes = np.array([[0, 0, 0, 1, 1, 1], [-2, 0, -2, -1, 1, -1]], np.float32)
    planes = np.array([[0, 1, 0, 0]], np.float32)
    d = np.array([-0.5, -2, -1], np.float32)
    s = np.float32(0)
    for v in d:   # core2 Vec::normalize: f32 accumulation, f64 sqrt, f32 divide
        s = np.float32(s + np.float32(v * v))
    d = (d / np.float32(np.sqrt(np.float64(s)))).astype(np.float32)
    lights = np.concatenate([pack_light_directional(d, (0.9, 0.9, 0.9)), pack_light_point((0, .1, 0), (1, 0, 0), 5, .3, 1),
                             pack_light_point((2, 2, 2), (0, 1, 1), 7, .6, 1)])
    return dict(triangles=tris, spheres=spheres.view(np.uint8).reshape(-1), cubes=cubes.view(np.uint8).reshape(-1),
                planes=planes.view(np.uint8).reshape(-1), lights=lights, materials=materials,
                material_indices=np.array([3, 4, 5, 0, 1, 2, 3, 4, 0, 7, 1, 2, 0], np.uint32),
                info=np.array([3, 8, 3, 7, 2, 1, 1, 0, 2], np.uint32))


class Context:
    """One rtb_ctx.  Methods map 1:1 onto the C ABI and raise RtbError with rtb_last_error on failure."""

    def __init__(self, device=0, max_triangles=65536, max_spheres=16384, max_cubes=32768, max_planes=256, max_lights=65536,
                 max_materials=65536):
        self.L = lib()
        self.h = C.c_void_p()
        lim = Limits(max_triangles, max_spheres, max_cubes, max_planes, max_lights, max_materials)
        rc = self.L.rtb_create(C.byref(self.h), device, C.byref(lim))
        if rc:
            raise RtbError(f"rtb_create failed ({rc}): {self.L.rtb_last_error(None).decode()}")
        self.width = self.height = self.samples = 0

    def close(self):
        if self.h:
            self.L.rtb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise RtbError(f"rtb error {rc}: {self.L.rtb_last_error(self.h).decode()}")

    def set_option(self, opt, value): self._ck(self.L.rtb_set_option(self.h, opt, value))
    def set_history_alpha(self, alpha):
        self.set_option(OPT_HISTORY_ALPHA, int(np.array([alpha], np.float32).view(np.uint32)[0]))

    def set_stream(self, stream_ptr): self._ck(self.L.rtb_set_stream(self.h, C.c_void_p(stream_ptr)))

    def resize(self, w, h, samples=1):
        self._ck(self.L.rtb_resize(self.h, w, h, samples))
        self.width, self.height, self.samples = w, h, samples

    def upload(self, buf, data, offset=0):
        a = np.ascontiguousarray(data)
        self._ck(self.L.rtb_upload(self.h, buf, offset, a.nbytes, _p(a)))

    def upload_raw(self, buf, ptr, nbytes, offset=0):
        self._ck(self.L.rtb_upload(self.h, buf, offset, nbytes, C.c_void_p(ptr)))

    def upload_skybox(self, sky):
        if sky is None:
            self._ck(self.L.rtb_upload_skybox(self.h, 0, 0, None))
        else:
            sky = np.ascontiguousarray(sky, np.uint16)
            self._ck(self.L.rtb_upload_skybox(self.h, sky.shape[1], sky.shape[0], _p(sky)))

    def upload_scene(self, scene: dict, skybox=None):
        """scene: dict of raw buffers as produced by niels_scene()."""
        for key, buf in (("triangles", BUF_TRIANGLES), ("spheres", BUF_SPHERES), ("cubes", BUF_CUBES), ("planes", BUF_PLANES),
                         ("lights", BUF_LIGHTS), ("materials", BUF_MATERIALS), ("material_indices", BUF_MATERIAL_INDICES)):
            if scene.get(key) is not None and np.asarray(scene[key]).size:
                self.upload(buf, scene[key])
        self.upload(BUF_SCENE_INFO, np.asarray(scene["info"], np.uint32))
        self.upload_skybox(skybox)

    def build_accel(self, mode=ACCEL_BVH): self._ck(self.L.rtb_build_accel(self.h, mode))

    def refit_accel(self): self._ck(self.L.rtb_refit_accel(self.h))

    def accel_bytes(self, target=None):
        """Raw node (or traversal-triangle) records of the acceleration structure."""
        target = TGT_ACCEL_NODES if target is None else target
        _, n = self.device_ptr(target)
        out = np.zeros(n, np.uint8)
        if n:
            self._ck(self.L.rtb_readback(self.h, target, _p(out), n))
        return out

    def accel_info(self) -> AccelInfo:
        a = AccelInfo()
        self._ck(self.L.rtb_accel_info_get(self.h, C.byref(a)))
        return a

    def dispatch(self, p): self._ck(self.L.rtb_dispatch(self.h, p))
    def sync(self): self._ck(self.L.rtb_sync(self.h))

    def readback(self, target):
        w, h, s = self.width, self.height, self.samples
        shape, dt = {TGT_DIR_T: ((h, w, 4), np.float32), TGT_UV_NORMAL: ((h, w, 4), np.float32),
                     TGT_SHADOW_BITS: ((self.device_ptr(TGT_SHADOW_BITS)[1] // 4 if target == TGT_SHADOW_BITS else 0,), np.uint32), TGT_LIGHTING: ((h, w, 4), np.uint16),
                     TGT_ACCUM: ((h, w, 4), np.float32), TGT_RGBA8: ((h, w), np.uint32), TGT_SEED: ((24,), np.uint8),
                     TGT_RGBA8_TILED: ((self.device_ptr(TGT_RGBA8_TILED)[1] // 4 if target == TGT_RGBA8_TILED else 0,), np.uint32)}[target]
        out = np.zeros(shape, dt)
        self._ck(self.L.rtb_readback(self.h, target, _p(out), out.nbytes))
        return out

    def readback_into(self, target, ptr, nbytes): self._ck(self.L.rtb_readback(self.h, target, C.c_void_p(ptr), nbytes))
    def readback_async_into(self, target, ptr, nbytes): self._ck(self.L.rtb_readback_async(self.h, target, C.c_void_p(ptr), nbytes))
    def readback_wait(self): self._ck(self.L.rtb_readback_wait(self.h))

    def device_ptr(self, target):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.L.rtb_device_ptr(self.h, target, C.byref(p), C.byref(n)))
        return p.value, n.value

    def probe_l2_read_gbs(self, nbytes=64 << 20) -> float:
        v = C.c_double()
        self.L.rtb_probe_l2_read_gbs.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_double)]
        self._ck(self.L.rtb_probe_l2_read_gbs(self.h, nbytes, C.byref(v)))
        return v.value

    def counters(self) -> Counters:
        c = Counters()
        self._ck(self.L.rtb_counters_get(self.h, C.byref(c)))
        return c

    def last_frame_ms(self):
        ms = (C.c_float * 8)()
        self._ck(self.L.rtb_last_frame_ms(self.h, C.byref(ms)))
        return list(ms)

    def trace_rays(self, rays, prev=None):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        obj, t, uv = np.zeros(n, np.uint32), np.zeros(n, np.float32), np.zeros((n, 2), np.float32)
        prev = None if prev is None else np.ascontiguousarray(prev, np.uint32)
        self._ck(self.L.rtb_trace_rays(self.h, _p(rays), n, _p(prev), _p(obj), _p(t), _p(uv)))
        return obj, t, uv

    def occlusion_rays(self, rays, max_dist=None, prev=None):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        occ = np.zeros(n, np.uint8)
        md = None if max_dist is None else np.ascontiguousarray(max_dist, np.float32)
        prev = None if prev is None else np.ascontiguousarray(prev, np.uint32)
        self._ck(self.L.rtb_occlusion_rays(self.h, _p(rays), n, _p(md), _p(prev), _p(occ)))
        return occ

    def path_frame(self, bounces): self._ck(self.L.rtb_path_frame(self.h, int(bounces)))

    def path_stats(self) -> PathStats:
        st = PathStats()
        self._ck(self.L.rtb_path_stats_get(self.h, C.byref(st)))
        return st

    def untile_on(self, tiled_all_ptr, nranks, slots_per_rank, out_ptr, stream_ptr):
        self._ck(self.L.rtb_untile_on(self.h, C.c_void_p(tiled_all_ptr), nranks, slots_per_rank, C.c_void_p(out_ptr), C.c_void_p(stream_ptr)))

    def present_host(self, host_ptr, tiled_src_ptr=None, stream_ptr=None):
        self._ck(self.L.rtb_present_host(self.h, C.c_void_p(host_ptr), C.c_void_p(tiled_src_ptr), C.c_void_p(stream_ptr)))

    def untile(self, tiled_all_ptr, nranks, slots_per_rank, out_ptr):
        self._ck(self.L.rtb_untile(self.h, C.c_void_p(tiled_all_ptr), nranks, slots_per_rank, C.c_void_p(out_ptr)))
