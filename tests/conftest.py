import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (test infrastructure): strict-FP build of oracle/oracle.cpp."""
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def rtb():
    """The product's C ABI through ctypes; builds the library if the .so is stale or missing."""
    from igx_raytracing_b200 import build, rtb as m
    build.build_library()
    m.lib()
    return m


def synthetic_sky(w=64, h=32, seed=7):
    """Small deterministic rgba16f equirect (the reference's 4k .hdr does not travel to the GPU box)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = 0.3 + 0.7 * (yy / (h - 1))[..., None] * np.array([0.6, 0.8, 1.0]) + 0.2 * np.sin(xx / w * 6.283)[..., None]
    base = base + rng.random((h, w, 3)) * 0.25
    base[h // 4, w // 3] = [30.0, 28.0, 20.0]   # a "sun" texel: exercises HDR range
    sky = np.zeros((h, w, 4), np.float16)
    sky[..., :3] = base.astype(np.float16)
    return sky.view(np.uint16)


@pytest.fixture(scope="session")
def sky():
    return synthetic_sky()


def degenerate_triangles(tri_bytes):
    """fixture variant: triangle 1 gets p1 = p0 (a zero first edge: rejected by the DEBUG shaders, primitive.glsl:248-253)"""
    t = np.array(tri_bytes, copy=True).view(np.uint8).reshape(-1, 48)
    t[1, 16:28] = t[1, 0:12]
    return t.reshape(-1)


def case_scene(oracle, case):
    """the oracle-side Scene of a tests/golden case"""
    scene = oracle.niels_scene(case["time"], synthetic_sky() if case["sky"] else None)
    if case.get("degenerate"):
        scene.triangles = degenerate_triangles(scene.triangles)
    return scene
