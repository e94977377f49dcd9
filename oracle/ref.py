"""ctypes binding of oracle/_ref/libigxref_{debug,release}.so — the REFERENCE'S OWN shader code compiled for the host
(oracle/ref_shim/: Makefile, glsl_front.py, glsl_shim.h, ref_tu.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/ and scripts/make_golden.py to pin the hand-written oracle.  The libraries are
built from /root/reference where the sources lie (build container only); on the GPU box the prebuilt .so files travel with
the snapshot and `available()` says whether they are there.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .oracle import Scene, shadow_words, _ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
_OUT = os.path.join(_HERE, "_ref")
_SHIM = os.path.join(_HERE, "ref_shim")
REFERENCE = os.environ.get("IGX_REFERENCE", "/root/reference")


class RefBind(C.Structure):
    _fields_ = [
        ("camera144", C.c_void_p), ("scene_info9", C.c_void_p), ("seed24", C.c_void_p),
        ("triangles", C.c_void_p), ("spheres", C.c_void_p), ("cubes", C.c_void_p), ("planes", C.c_void_p),
        ("lights", C.c_void_p), ("materials", C.c_void_p), ("material_indices", C.c_void_p),
        ("skybox", C.c_void_p), ("sky_w", C.c_uint32), ("sky_h", C.c_uint32),
        ("width", C.c_uint32), ("height", C.c_uint32), ("samples", C.c_uint32),
        ("dirT", C.c_void_p), ("uvN", C.c_void_p), ("shadow_bits", C.c_void_p), ("lighting", C.c_void_p),
        ("accum", C.c_void_p), ("rgba8", C.c_void_p), ("debug_type", C.c_uint32), ("nan_only", C.c_uint32),
    ]


def _lib_path(debug: bool) -> str:
    return os.path.join(_OUT, "libigxref_debug.so" if debug else "libigxref_release.so")


def can_build() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "res", "shaders"))


def build(force: bool = False) -> bool:
    """Compile oracle/_ref/ from the reference sources with the committed recipe; False when the reference tree is absent."""
    if not can_build():
        return available()
    cmd = ["make", "-C", _SHIM, f"REF={REFERENCE}"] + (["-B"] if force else [])
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return True


def available() -> bool:
    return os.path.exists(_lib_path(True)) and os.path.exists(_lib_path(False))


class Ref:
    """The reference's shaders on the host.  debug=True is the -DDEBUG build (what the shipped .spv binaries are)."""

    def __init__(self, debug: bool = True):
        if not build() and not available():
            raise RuntimeError("oracle/_ref is not built and the reference tree is absent")
        self.lib = C.CDLL(_lib_path(debug))
        self.debug = bool(self.lib.ref_is_debug())
        assert self.debug == debug
        for name in ("ref_init", "ref_raygen", "ref_shadow", "ref_lighting", "ref_composite"):
            getattr(self.lib, name).argtypes = [C.c_void_p]
        self.lib.ref_trace_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64] + [C.c_void_p] * 5
        self.lib.ref_occlusion_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64] + [C.c_void_p] * 3
        self.lib.ref_primary_rays.argtypes = [C.c_void_p, C.c_void_p]

    def set_threads(self, n):
        self.lib.ref_set_threads(C.c_int(n))

    def _bind(self, scene: Scene, cam=None, seed=None, samples=1, keep=None, **targets) -> RefBind:
        b = RefBind()
        keep = keep if keep is not None else []
        keep += [scene, cam, seed]
        if cam is not None:
            b.camera144 = _ptr(cam)
            u = cam.view(np.uint32)
            b.width, b.height = int(u[3]), int(u[7])
        else:
            zero_cam = np.zeros(144, np.uint8)
            keep.append(zero_cam)
            b.camera144 = _ptr(zero_cam)
        info = np.ascontiguousarray(scene.info, np.uint32)
        keep.append(info)
        b.scene_info9 = _ptr(info)
        b.seed24 = _ptr(seed)
        b.triangles, b.spheres, b.cubes, b.planes = _ptr(scene.triangles), _ptr(scene.spheres), _ptr(scene.cubes), _ptr(scene.planes)
        b.lights, b.materials, b.material_indices = _ptr(scene.lights), _ptr(scene.materials), _ptr(scene.material_indices)
        if scene.skybox is not None:
            b.skybox = _ptr(scene.skybox)
            b.sky_h, b.sky_w = scene.skybox.shape[0], scene.skybox.shape[1]
        b.samples = samples
        for k, v in targets.items():
            setattr(b, k, _ptr(v))
            keep.append(v)
        b._keep = keep
        return b

    # ---- passes ------------------------------------------------------------------------------------
    def init_pass(self, scene, seed):
        b = self._bind(scene, None, seed)
        self.lib.ref_init(C.byref(b))
        return seed

    def raygen(self, scene, cam, seed, dirT=None, uvN=None):
        b = self._bind(scene, cam, seed)
        h, w = b.height, b.width
        dirT = np.zeros((h, w, 4), np.float32) if dirT is None else dirT
        uvN = np.zeros((h, w, 4), np.float32) if uvN is None else uvN
        b.dirT, b.uvN = _ptr(dirT), _ptr(uvN)
        self.lib.ref_raygen(C.byref(b))
        return dirT, uvN

    def primary_rays(self, scene, cam, seed):
        b = self._bind(scene, cam, seed)
        rays = np.zeros((b.height, b.width, 6), np.float32)
        self.lib.ref_primary_rays(C.byref(b), _ptr(rays))
        return rays

    def shadow(self, scene, cam, seed, samples, dirT, bits=None):
        b = self._bind(scene, cam, seed, samples, dirT=dirT)
        bits = np.zeros(shadow_words(b.width, b.height, samples), np.uint32) if bits is None else bits
        b.shadow_bits = _ptr(bits)
        self.lib.ref_shadow(C.byref(b))
        return bits

    def lighting(self, scene, cam, samples, dirT, uvN, bits, l16=None):
        b = self._bind(scene, cam, None, samples, dirT=dirT, uvN=uvN, shadow_bits=bits)
        l16 = np.zeros((b.height, b.width, 4), np.uint16) if l16 is None else l16
        b.lighting = _ptr(l16)
        self.lib.ref_lighting(C.byref(b))
        return l16

    def composite(self, scene, cam, seed, dirT, uvN, l16, accum=None, rgba8=None, debug_type=0, nan_only=0):
        b = self._bind(scene, cam, seed, 1, dirT=dirT, uvN=uvN, lighting=l16)
        rgba8 = np.zeros((b.height, b.width), np.uint32) if rgba8 is None else rgba8
        b.rgba8, b.accum = _ptr(rgba8), _ptr(accum)
        b.debug_type, b.nan_only = debug_type, nan_only
        self.lib.ref_composite(C.byref(b))
        return rgba8

    def frame(self, scene, cam, seed, samples, accum=None, prefill=None):
        """init -> raygen -> shadow -> lighting -> composite, the order the reference records (src/rt/raytracing_interface.cpp:
        144-179).  `prefill` = dict of arrays the targets start from (what a RELEASE build leaves untouched stays as given)."""
        self.init_pass(scene, seed)
        u = cam.view(np.uint32)
        w, h = int(u[3]), int(u[7])
        pre = prefill or {}
        out = dict(dirT=pre.get("dirT", np.zeros((h, w, 4), np.float32)).copy(), uvN=pre.get("uvN", np.zeros((h, w, 4), np.float32)).copy(),
                   bits=pre.get("bits", np.zeros(shadow_words(w, h, samples), np.uint32)).copy(),
                   lighting=pre.get("lighting", np.zeros((h, w, 4), np.uint16)).copy(), rgba8=np.zeros((h, w), np.uint32))
        self.raygen(scene, cam, seed, out["dirT"], out["uvN"])
        self.shadow(scene, cam, seed, samples, out["dirT"], out["bits"])
        self.lighting(scene, cam, samples, out["dirT"], out["uvN"], out["bits"], out["lighting"])
        self.composite(scene, cam, seed, out["dirT"], out["uvN"], out["lighting"], accum, out["rgba8"])
        return out

    # ---- function-level ----------------------------------------------------------------------------------
    def trace_rays(self, scene, rays, prev=None):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        obj, t, uv, nrm = np.zeros(n, np.uint32), np.zeros(n, np.float32), np.zeros((n, 2), np.float32), np.zeros((n, 2), np.uint32)
        prev = None if prev is None else np.ascontiguousarray(prev, np.uint32)
        b = self._bind(scene)
        self.lib.ref_trace_rays(C.byref(b), _ptr(rays), n, _ptr(prev), _ptr(obj), _ptr(t), _ptr(uv), _ptr(nrm))
        return obj, t, uv, nrm

    def occlusion_rays(self, scene, rays, max_dist=None, prev=None):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        occ = np.zeros(n, np.uint8)
        md = None if max_dist is None else np.ascontiguousarray(max_dist, np.float32)
        prev = None if prev is None else np.ascontiguousarray(prev, np.uint32)
        b = self._bind(scene)
        self.lib.ref_occlusion_rays(C.byref(b), _ptr(rays), n, _ptr(md), _ptr(prev), _ptr(occ))
        return occ
