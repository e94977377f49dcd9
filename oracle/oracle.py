"""ctypes binding of the CPU oracle (oracle/oracle.h).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")

NO_RAY_HIT = 0xFFFFFFFF
NO_HIT = np.float32(3.4028235e38)

FLAG_EDGE, FLAG_TIE, FLAG_PARALLEL, FLAG_NAN = 1, 2, 4, 8


class OrcScene(C.Structure):
    _fields_ = [
        ("triangles", C.c_void_p), ("spheres", C.c_void_p), ("cubes", C.c_void_p), ("planes", C.c_void_p),
        ("lights", C.c_void_p), ("materials", C.c_void_p), ("material_indices", C.c_void_p),
        ("skybox", C.c_void_p), ("info", C.c_uint32 * 9), ("sky_w", C.c_uint32), ("sky_h", C.c_uint32),
        ("_pad", C.c_uint32),
    ]


def build(force: bool = False) -> None:
    """Compile oracle/_build/liboracle{,_fast}.so with the committed Makefile."""
    if force or not all(os.path.exists(os.path.join(_BUILD, n)) for n in ("liboracle.so", "liboracle_fast.so")):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Scene:
    """Raw GPU-layout scene buffers (numpy, host) + the ctypes view handed to the oracle."""

    def __init__(self, triangles=None, spheres=None, cubes=None, planes=None, lights=None, materials=None,
                 material_indices=None, info=None, skybox=None):
        z = lambda n: np.zeros(n, np.uint8)
        self.triangles = np.ascontiguousarray(triangles if triangles is not None else z(0)).view(np.uint8).reshape(-1)
        self.spheres = np.ascontiguousarray(spheres if spheres is not None else z(0)).view(np.uint8).reshape(-1)
        self.cubes = np.ascontiguousarray(cubes if cubes is not None else z(0)).view(np.uint8).reshape(-1)
        self.planes = np.ascontiguousarray(planes if planes is not None else z(0)).view(np.uint8).reshape(-1)
        self.lights = np.ascontiguousarray(lights if lights is not None else z(0)).view(np.uint8).reshape(-1)
        self.materials = np.ascontiguousarray(materials if materials is not None else z(0)).view(np.uint8).reshape(-1)
        self.material_indices = np.ascontiguousarray(
            material_indices if material_indices is not None else np.zeros(0, np.uint32), dtype=np.uint32)
        if info is None:
            nl = self.lights.size // 32
            ltypes = self.lights.view(np.uint16).reshape(-1, 16)[:, 15] if nl else np.zeros(0, np.uint16)
            info = [nl, self.materials.size // 32, self.triangles.size // 48, self.spheres.size // 16,
                    self.cubes.size // 24, self.planes.size // 16,
                    int((ltypes == 0).sum()), int((ltypes == 1).sum()), int((ltypes == 2).sum())]
        self.info = np.asarray(info, np.uint32)
        self.skybox = None if skybox is None else np.ascontiguousarray(skybox, dtype=np.uint16)  # (H, W, 4)

    @property
    def geometry_count(self):
        return int(self.info[2] + self.info[3] + self.info[4] + self.info[5])

    def c(self) -> OrcScene:
        s = OrcScene()
        s.triangles, s.spheres, s.cubes, s.planes = _ptr(self.triangles), _ptr(self.spheres), _ptr(self.cubes), _ptr(self.planes)
        s.lights, s.materials, s.material_indices = _ptr(self.lights), _ptr(self.materials), _ptr(self.material_indices)
        for i in range(9):
            s.info[i] = int(self.info[i])
        if self.skybox is not None:
            s.skybox = _ptr(self.skybox)
            s.sky_h, s.sky_w = self.skybox.shape[0], self.skybox.shape[1]
        return s


def shadow_words(w, h, samples):
    return ((w + 15) // 16) * ((h + 1) // 2) * samples


MODE_DEBUG, MODE_RELEASE = 0, 1   # which build of the reference shaders is modelled (oracle.cpp D10)


def build_native() -> str:
    """liboracle_native.so: the optimised build with -march=native, compiled ON THE MACHINE THAT RUNS IT (bench.py's CPU
    legs call this on the GPU box; a library built with the build container's -march=native must not travel)."""
    import hashlib
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    tag = hashlib.sha1(flags.encode()).hexdigest()[:10]
    out = os.path.join(_BUILD, f"liboracle_native_{tag}.so")
    src = os.path.join(_HERE, "oracle.cpp")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(_BUILD, exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O3", "-march=native", "-funroll-loops", "-ffp-contract=off", "-fno-fast-math",
                               "-fPIC", "-shared", "-pthread", "-o", out, src])
    return out


class Oracle:
    def __init__(self, fast: bool = False, native: bool = False):
        build()
        path = build_native() if native else os.path.join(_BUILD, "liboracle_fast.so" if fast else "liboracle.so")
        self.lib = C.CDLL(path)
        L = self.lib
        L.orc_f16_trunc.restype = C.c_uint16
        L.orc_f16_trunc.argtypes = [C.c_float]
        L.orc_f16_to_f32.restype = C.c_float
        L.orc_f16_to_f32.argtypes = [C.c_uint16]
        L.orc_f32_to_f16_rtne.restype = C.c_uint16
        L.orc_f32_to_f16_rtne.argtypes = [C.c_float]
        L.orc_get_threads.restype = C.c_int
        L.orc_load_hdr.restype = C.c_int
        L.orc_frame_pixels.restype = C.c_uint64
        L.orc_light_directional.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        L.orc_light_point.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.orc_material.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.orc_camera.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                 C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_void_p, C.c_void_p]
        L.orc_niels_scene.argtypes = [C.c_double] + [C.c_void_p] * 8
        L.orc_raygen.argtypes = [C.c_void_p] * 7
        L.orc_trace_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64] + [C.c_void_p] * 6
        L.orc_occlusion_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64] + [C.c_void_p] * 3
        L.orc_shadow.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_lighting.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32] + [C.c_void_p] * 5
        L.orc_composite.argtypes = [C.c_void_p] * 8
        L.orc_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32] + [C.c_void_p] * 6
        L.orc_frame_pixels.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_load_hdr.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]

    # ---- threads -------------------------------------------------------------------------------
    def set_threads(self, n):
        self.lib.orc_set_threads(C.c_int(n))

    def threads(self):
        return int(self.lib.orc_get_threads())

    def set_mode(self, mode):
        """MODE_DEBUG (default: the shipped .spv) or MODE_RELEASE; process-wide for this library."""
        self.lib.orc_set_mode(C.c_int(mode))

    def mode(self):
        return int(self.lib.orc_get_mode())

    def set_all_lights(self, on):
        """every light evaluated (layer = light * samples + sample) instead of the reference's light 0 x lightCount"""
        self.lib.orc_set_all_lights(C.c_int(1 if on else 0))
        self.all_lights = bool(on)

    def set_history(self, alpha):
        self.lib.orc_set_history(C.c_float(alpha))

    def history_blend(self, history, lighting, first):
        self.lib.orc_history_blend.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
        self.lib.orc_history_blend(_ptr(history), _ptr(lighting), history.size // 4, 1 if first else 0)

    # ---- packing -------------------------------------------------------------------------------
    def f16_trunc(self, v):
        return int(self.lib.orc_f16_trunc(np.float32(v)))

    def f16_to_f32(self, h):
        return float(self.lib.orc_f16_to_f32(int(h)))

    def f32_to_f16_rtne(self, v):
        return int(self.lib.orc_f32_to_f16_rtne(np.float32(v)))

    def spheremap(self, n):
        n = np.asarray(n, np.float32)
        out = np.zeros(2, np.uint16)
        self.lib.orc_spheremap(_ptr(n), _ptr(out))
        return out

    def encode_normal_cpu(self, n):
        n = np.asarray(n, np.float32)
        out = np.zeros(2, np.uint32)
        self.lib.orc_encode_normal_cpu(_ptr(n), _ptr(out))
        return out

    def triangle_flat(self, p):
        p = np.ascontiguousarray(p, np.float32).reshape(9)
        out = np.zeros(48, np.uint8)
        self.lib.orc_triangle_flat(_ptr(p), _ptr(out))
        return out

    def triangle_normals(self, p, n):
        p = np.ascontiguousarray(p, np.float32).reshape(9)
        n = np.ascontiguousarray(n, np.float32).reshape(9)
        out = np.zeros(48, np.uint8)
        self.lib.orc_triangle_normals(_ptr(p), _ptr(n), _ptr(out))
        return out

    def light_directional(self, direction, color, angular_extent=None):
        if angular_extent is None:
            angular_extent = np.float32(0.533 * (3.141592653589793 / 180))
        d, c = np.asarray(direction, np.float32), np.asarray(color, np.float32)
        out = np.zeros(32, np.uint8)
        self.lib.orc_light_directional(_ptr(d), _ptr(c), np.float32(angular_extent), _ptr(out))
        return out

    def light_point(self, pos, color, rad, origin, specularity=1.0):
        p, c = np.asarray(pos, np.float32), np.asarray(color, np.float32)
        out = np.zeros(32, np.uint8)
        self.lib.orc_light_point(_ptr(p), _ptr(c), np.float32(rad), np.float32(origin), np.float32(specularity), _ptr(out))
        return out

    def material(self, albedo, ambient, emission, metallic, roughness, transparency=1.0):
        a, b, e = (np.asarray(v, np.float32) for v in (albedo, ambient, emission))
        out = np.zeros(32, np.uint8)
        self.lib.orc_material(_ptr(a), _ptr(b), _ptr(e), np.float32(metallic), np.float32(roughness),
                              np.float32(transparency), _ptr(out))
        return out

    def camera(self, width, height, eye=(4, 2, -2), pitch=0.0, yaw=0.0, roll=0.0, left_fov=70.0, right_fov=70.0,
               ipd=62.0, projection=0, flags=0, exposure=1.0, skybox_color=(0.25, 0.5, 1.0)):
        e, sc = np.asarray(eye, np.float32), np.asarray(skybox_color, np.float32)
        out = np.zeros(144, np.uint8)
        self.lib.orc_camera(_ptr(e), pitch, yaw, roll, left_fov, right_fov, ipd, projection, width, height, flags,
                            exposure, _ptr(sc), _ptr(out))
        return out

    def niels_scene(self, time=0.0, skybox=None) -> Scene:
        tri, sph, cub, pla = np.zeros(3 * 48, np.uint8), np.zeros(7 * 16, np.uint8), np.zeros(2 * 24, np.uint8), np.zeros(16, np.uint8)
        lig, mat, idx, info = np.zeros(3 * 32, np.uint8), np.zeros(8 * 32, np.uint8), np.zeros(13, np.uint32), np.zeros(9, np.uint32)
        self.lib.orc_niels_scene(float(time), _ptr(tri), _ptr(sph), _ptr(cub), _ptr(pla), _ptr(lig), _ptr(mat), _ptr(idx), _ptr(info))
        return Scene(tri, sph, cub, pla, lig, mat, idx, info, skybox)

    def gen_soup(self, n, seed=0xB200):
        out = np.zeros(int(n) * 48, np.uint8)
        self.lib.orc_gen_soup(C.c_uint64(int(n)), C.c_uint64(seed), _ptr(out))
        return out

    def gen_heightfield(self, grid, seed=0xB200):
        out = np.zeros(2 * int(grid) * int(grid) * 48, np.uint8)
        self.lib.orc_gen_heightfield(C.c_uint32(int(grid)), C.c_uint64(seed), _ptr(out))
        return out

    def load_hdr(self, path):
        w, h = C.c_uint32(0), C.c_uint32(0)
        rc = self.lib.orc_load_hdr(path.encode(), None, C.byref(w), C.byref(h))
        if rc:
            raise IOError(f"orc_load_hdr({path}) header failed: {rc}")
        out = np.zeros((h.value, w.value, 4), np.uint16)
        rc = self.lib.orc_load_hdr(path.encode(), _ptr(out), C.byref(w), C.byref(h))
        if rc:
            raise IOError(f"orc_load_hdr({path}) failed: {rc}")
        return out

    # ---- passes --------------------------------------------------------------------------------
    @staticmethod
    def seed(cpu_offset=(0.0, 0.0), sample_count=0, sample_offset=0, random=(0.0, 0.0)):
        s = np.zeros(24, np.uint8)
        s[:16].view(np.float32)[:] = [random[0], random[1], cpu_offset[0], cpu_offset[1]]
        s[16:].view(np.uint32)[:] = [sample_count, sample_offset]
        return s

    def init_pass(self, seed):
        self.lib.orc_init_pass(_ptr(seed))
        return seed

    @staticmethod
    def _wh(cam):
        u = cam.view(np.uint32)
        return int(u[3]), int(u[7])

    def raygen(self, scene, cam, seed, want_rays=False, want_flags=False):
        w, h = self._wh(cam)
        dirT, uvN = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)
        rays = np.zeros((h, w, 6), np.float32) if want_rays else None
        flags = np.zeros((h, w), np.uint8) if want_flags else None
        cs = scene.c()
        self.lib.orc_raygen(C.byref(cs), _ptr(cam), _ptr(seed), _ptr(dirT), _ptr(uvN), _ptr(rays), _ptr(flags))
        return dirT, uvN, rays, flags

    def trace_rays(self, scene, rays, prev=None, want_flags=False):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        obj, t, uv, nrm = np.zeros(n, np.uint32), np.zeros(n, np.float32), np.zeros((n, 2), np.float32), np.zeros((n, 2), np.uint32)
        flags = np.zeros(n, np.uint8) if want_flags else None
        prev = None if prev is None else np.ascontiguousarray(prev, np.uint32)
        cs = scene.c()
        self.lib.orc_trace_rays(C.byref(cs), _ptr(rays), n, _ptr(prev), _ptr(obj), _ptr(t), _ptr(uv), _ptr(nrm), _ptr(flags))
        return obj, t, uv, nrm, flags

    def occlusion_rays(self, scene, rays, max_dist=None, prev=None):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        occ = np.zeros(n, np.uint8)
        md = None if max_dist is None else np.ascontiguousarray(max_dist, np.float32)
        prev = None if prev is None else np.ascontiguousarray(prev, np.uint32)
        cs = scene.c()
        self.lib.orc_occlusion_rays(C.byref(cs), _ptr(rays), n, _ptr(md), _ptr(prev), _ptr(occ))
        return occ

    def _layers(self, scene, samples):
        return samples * (int(scene.info[0]) if getattr(self, "all_lights", False) else 1)

    def shadow(self, scene, cam, seed, samples, dirT, want_rays=False):
        w, h = self._wh(cam)
        bits = np.zeros(shadow_words(w, h, self._layers(scene, samples)), np.uint32)
        rays = np.zeros((self._layers(scene, samples), h, w, 6), np.float32) if want_rays else None
        cs = scene.c()
        self.lib.orc_shadow(C.byref(cs), _ptr(cam), _ptr(seed), samples, _ptr(dirT), _ptr(bits), _ptr(rays))
        return (bits, rays) if want_rays else bits

    def lighting(self, scene, cam, samples, dirT, uvN, bits):
        w, h = self._wh(cam)
        l16, l32 = np.zeros((h, w, 4), np.uint16), np.zeros((h, w, 4), np.float32)
        cs = scene.c()
        self.lib.orc_lighting(C.byref(cs), _ptr(cam), samples, _ptr(dirT), _ptr(uvN), _ptr(bits), _ptr(l16), _ptr(l32))
        return l16, l32

    def composite(self, scene, cam, seed, dirT, uvN, l16, accum=None):
        w, h = self._wh(cam)
        rgba = np.zeros((h, w), np.uint32)
        cs = scene.c()
        self.lib.orc_composite(C.byref(cs), _ptr(cam), _ptr(seed), _ptr(dirT), _ptr(uvN), _ptr(l16), _ptr(accum), _ptr(rgba))
        return rgba

    def frame(self, scene, cam, seed, samples, accum=None, prefill=None):
        """K0..K4. Returns dict of every intermediate; `seed` is updated in place.  `prefill`: what the targets hold before
        the frame (a RELEASE build leaves the texels it does not store as they were)."""
        w, h = self._wh(cam)
        pre = prefill or {}
        out = dict(dirT=pre.get("dirT", np.zeros((h, w, 4), np.float32)).copy(), uvN=pre.get("uvN", np.zeros((h, w, 4), np.float32)).copy(),
                   bits=pre.get("bits", np.zeros(shadow_words(w, h, self._layers(scene, samples)), np.uint32)).copy(),
                   lighting=pre.get("lighting", np.zeros((h, w, 4), np.uint16)).copy(), rgba8=np.zeros((h, w), np.uint32))
        cs = scene.c()
        self.lib.orc_frame(C.byref(cs), _ptr(cam), _ptr(seed), samples, _ptr(out["dirT"]), _ptr(out["uvN"]),
                           _ptr(out["bits"]), _ptr(out["lighting"]), _ptr(accum), _ptr(out["rgba8"]))
        return out

    def path_frame(self, scene, cam, seed, bounces, accum=None):
        """K0 + one frame with diffuse bounces (our definition, see oracle.h); `seed` is updated in place."""
        w, h = self._wh(cam)
        out = dict(dirT=np.zeros((h, w, 4), np.float32), uvN=np.zeros((h, w, 4), np.float32), rgba8=np.zeros((h, w), np.uint32),
                   radiance=np.zeros((h, w, 3), np.float32))
        rays = C.c_uint64(0)
        cs = scene.c()
        self.lib.orc_path_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32] + [C.c_void_p] * 6
        self.lib.orc_path_frame(C.byref(cs), _ptr(cam), _ptr(seed), int(bounces), _ptr(out["dirT"]), _ptr(out["uvN"]), _ptr(accum),
                                _ptr(out["rgba8"]), _ptr(out["radiance"]), C.byref(rays))
        out["rays"] = int(rays.value)
        return out

    def frame_pixels_ex(self, scene, cam, seed, samples, xy):
        """frame_pixels plus the G-buffer texel, sample 0's shadow bit and the primary ray's edge / tie flags"""
        xy = np.ascontiguousarray(xy, np.uint32).reshape(-1, 2)
        n = xy.shape[0]
        out = dict(rgba8=np.zeros(n, np.uint32), object=np.zeros(n, np.uint32), t=np.zeros(n, np.float32), dirT=np.zeros((n, 4), np.float32),
                   shadowed=np.zeros(n, np.uint8), flags=np.zeros(n, np.uint8))
        cs = scene.c()
        self.lib.orc_frame_pixels_ex.restype = C.c_uint64
        self.lib.orc_frame_pixels_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64] + [C.c_void_p] * 6
        out["rays"] = int(self.lib.orc_frame_pixels_ex(C.byref(cs), _ptr(cam), _ptr(seed), samples, _ptr(xy), n, _ptr(out["rgba8"]), _ptr(out["object"]),
                                                       _ptr(out["t"]), _ptr(out["dirT"]), _ptr(out["shadowed"]), _ptr(out["flags"])))
        return out

    def frame_pixels(self, scene, cam, seed, samples, xy):
        xy = np.ascontiguousarray(xy, np.uint32).reshape(-1, 2)
        n = xy.shape[0]
        rgba, obj, t = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.float32)
        cs = scene.c()
        rays = self.lib.orc_frame_pixels(C.byref(cs), _ptr(cam), _ptr(seed), samples, _ptr(xy), n, _ptr(rgba), _ptr(obj), _ptr(t))
        return int(rays), rgba, obj, t
