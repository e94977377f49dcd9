"""Pins the hand-written oracle (oracle/oracle.cpp) against oracle/_ref: the reference's OWN shader sources compiled for the
host (oracle/ref_shim/).  Everything is compared BITWISE: whole frames of every committed fixture's inputs (init -> raygen ->
shadow -> lighting -> composite), explicit rays through the reference's traceGeometry / traceOcclusion / encodeNormal, both
shader builds (-DDEBUG = what the shipped .spv are, -DRELEASE), and the quirks SURVEY.md 8(a) lists (degenerate triangles,
inside-a-cube, grazing plane, point light first, every projection mode).

Runs wherever oracle/_ref exists: in the build container it is (re)built from /root/reference by the committed recipe; on the
GPU box the prebuilt libraries travel with the snapshot."""
import json
import os

import numpy as np
import pytest

from conftest import case_scene, synthetic_sky

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = json.load(open(os.path.join(GOLDEN, "cases.json")))["cases"]
KEYS = ("dirT", "uvN", "bits", "lighting", "rgba8", "accum")


@pytest.fixture(scope="module")
def refs():
    from oracle import ref
    if not ref.build() and not ref.available():
        pytest.skip("oracle/_ref is not built and /root/reference is absent")
    return {True: ref.Ref(debug=True), False: ref.Ref(debug=False)}


@pytest.fixture()
def modal_oracle(oracle):
    yield oracle
    oracle.set_mode(0)


def words_differing(a, b):
    a, b = np.asarray(a).view(np.uint8).reshape(-1, 4), np.asarray(b).view(np.uint8).reshape(-1, 4)
    return int((a != b).any(axis=-1).sum())


def garbage(w, h, samples, rng):
    """what the targets hold before the frame: a RELEASE build must leave unstored texels exactly like this"""
    from oracle.oracle import shadow_words
    return dict(dirT=rng.random((h, w, 4), np.float32), uvN=rng.random((h, w, 4), np.float32),
                bits=rng.integers(0, 2**32, shadow_words(w, h, samples), dtype=np.uint32),
                lighting=rng.integers(0, 0x3C00, (h, w, 4), dtype=np.uint16))


@pytest.mark.parametrize("debug", [True, False], ids=["DEBUG", "RELEASE"])
@pytest.mark.parametrize("name", [n for n in CASES if CASES[n].get("path_bounces") is None])
def test_whole_frames_equal_the_reference_shaders(modal_oracle, refs, name, debug):
    oracle, ref = modal_oracle, refs[debug]
    oracle.set_mode(0 if debug else 1)
    case = CASES[name]
    scene = case_scene(oracle, case)
    w, h = case["w"], case["h"]
    cam = oracle.camera(w, h, **case["cam"])
    s_o, s_r = oracle.seed(tuple(case["off"])), oracle.seed(tuple(case["off"]))
    a_o, a_r = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)
    pre = garbage(w, h, case["samples"], np.random.default_rng(5))
    for _ in range(case["frames"]):
        got = oracle.frame(scene, cam, s_o, case["samples"], accum=a_o, prefill=pre)
        want = ref.frame(scene, cam, s_r, case["samples"], accum=a_r, prefill=pre)
        pre = {k: want[k] for k in ("dirT", "uvN", "bits", "lighting")}
    got["accum"], want["accum"] = a_o, a_r
    assert np.array_equal(s_o, s_r), "Seed after init.comp differs"
    for k in KEYS:
        assert words_differing(got[k], want[k]) == 0, f"{name} [{'DEBUG' if debug else 'RELEASE'}]: oracle {k} != reference shaders"
    if not debug:   # the RELEASE build really skips stores: some prefilled texels survive where rays missed
        miss = want["dirT"][..., 3].view(np.uint32) == 0xFFFFFFFF
        if miss.any() and case["frames"] == 1:
            assert np.array_equal(want["uvN"][miss], garbage(w, h, case["samples"], np.random.default_rng(5))["uvN"][miss])


def random_scene(oracle, rng, n_tri=200, degenerate=True):
    """triangles (some with p1 == p0, some needle-thin), spheres, cubes, planes, one sun: every primitive loop of trace.glsl"""
    from oracle.oracle import Scene
    tris = []
    for i in range(n_tri):
        c = rng.uniform(-3, 3, 3)
        p = c + rng.uniform(-0.7, 0.7, (3, 3))
        if degenerate and i % 17 == 3:
            p[1] = p[0]                       # zero first edge: DEBUG rejects, RELEASE evaluates (NaN path)
        if degenerate and i % 23 == 5:
            p[2] = p[0] + (p[1] - p[0]) * 0.5  # collinear: a == 0
        tris.append(oracle.triangle_flat(p.astype(np.float32)))
    tris = np.concatenate(tris)
    spheres = np.concatenate([rng.uniform(-3, 3, (6, 3)), rng.uniform(0.2, 0.8, (6, 1))], axis=1).astype(np.float32)
    lo = rng.uniform(-3, 2, (5, 3))
    cubes = np.concatenate([lo, lo + rng.uniform(0.3, 1.2, (5, 3))], axis=1).astype(np.float32)
    planes = np.array([[0, 1, 0, 3.5], [0.3, 1, 0.1, 4.0]], np.float32)
    lights = np.concatenate([oracle.light_directional((-0.5, -2, -1), (0.9, 0.9, 0.9))])
    mats = np.concatenate([oracle.material((0.8, 0.7, 0.6), (0.05, 0.05, 0.05), (0, 0, 0), 0.2, 0.6),
                           oracle.material((0.1, 0.9, 0.3), (0.0, 0.02, 0.0), (0.1, 0, 0), 0.9, 0.1)])
    n_obj = n_tri + 6 + 5 + 2
    idx = (np.arange(n_obj) % 2).astype(np.uint32)
    return Scene(tris, spheres, cubes, planes, lights, mats, idx, None, synthetic_sky())


def random_rays(rng, n):
    o = rng.uniform(-5, 5, (n, 3))
    tgt = rng.uniform(-3, 3, (n, 3))
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d], axis=1).astype(np.float32)
    rays[::97, 3:] = [0.0, -1.0, 0.0]      # axis-parallel: 1/0 in the cube slabs
    rays[::89, 4] = 0.0                     # grazing the planes' family
    return rays


@pytest.mark.parametrize("debug", [True, False], ids=["DEBUG", "RELEASE"])
def test_explicit_rays_equal_the_reference_functions(modal_oracle, refs, debug):
    """traceGeometry + encodeNormal and traceOcclusion of the reference (trace.glsl:9-98, primitive.glsl:85-88,173-333) on
    20k random rays against a scene with every primitive type; with and without a previous-hit exclusion."""
    oracle, ref = modal_oracle, refs[debug]
    oracle.set_mode(0 if debug else 1)
    rng = np.random.default_rng(11)
    scene = random_scene(oracle, rng)
    rays = random_rays(rng, 20000)
    for prev in (None, rng.integers(0, scene.geometry_count, rays.shape[0]).astype(np.uint32)):
        o_obj, o_t, o_uv, o_n, _ = oracle.trace_rays(scene, rays, prev)
        r_obj, r_t, r_uv, r_n = ref.trace_rays(scene, rays, prev)
        assert np.array_equal(o_obj, r_obj)
        assert np.array_equal(o_t.view(np.uint32), r_t.view(np.uint32))
        assert np.array_equal(o_uv.view(np.uint32), r_uv.view(np.uint32))
        assert np.array_equal(o_n, r_n)
        for md in (None, rng.uniform(0.5, 6.0, rays.shape[0]).astype(np.float32)):
            assert np.array_equal(oracle.occlusion_rays(scene, rays, md, prev), ref.occlusion_rays(scene, rays, md, prev))
    assert (o_obj != 0xFFFFFFFF).mean() > 0.5


def test_the_two_shader_builds_differ_on_a_degenerate_triangle(modal_oracle, refs):
    """The DEBUG-only reject (primitive.glsl:248-253) is visible: a triangle with p1 == p0 can be 'hit' (NaN accepted) by the
    RELEASE maths and never by the DEBUG build.  Guards against a shim that ignores -DDEBUG."""
    oracle = modal_oracle
    from oracle.oracle import Scene
    p = np.array([[0, 0, 0], [0, 0, 0], [1, 0, 0]], np.float32)
    scene = Scene(oracle.triangle_flat(p), None, None, None, oracle.light_directional((0, -1, 0), (1, 1, 1)),
                  oracle.material((1, 1, 1), (0, 0, 0), (0, 0, 0), 0, 1), np.zeros(1, np.uint32), None, None)
    rays = np.array([[0.25, 1, 0, 0, -1, 0], [0.5, 2, 0.0, 0, -1, 0]], np.float32)
    d_obj = refs[True].trace_rays(scene, rays)[0]
    r_obj = refs[False].trace_rays(scene, rays)[0]
    assert (d_obj == 0xFFFFFFFF).all()
    for debug in (True, False):
        oracle.set_mode(0 if debug else 1)
        got = oracle.trace_rays(scene, rays)
        want = refs[debug].trace_rays(scene, rays)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32))
    # what RELEASE does with it is whatever IEEE gives (0 * inf = NaN passes every reject): pin that it is the same on both sides
    assert np.array_equal(oracle.trace_rays(scene, rays)[0], r_obj)


@pytest.mark.parametrize("projection", [0, 1, 2, 3, 4, 5])
def test_primary_rays_of_every_projection_mode(oracle, refs, projection):
    """calculatePrimary (camera.glsl:53-138) for all six projection types, jittered by the init.comp seed."""
    scene = oracle.niels_scene(0.0, None)
    cam = oracle.camera(80, 48, eye=(6, 5, 12), yaw=0.3, pitch=-0.1, projection=projection)
    seed = oracle.seed((17.0, -3.5))
    oracle.init_pass(seed)
    want = refs[True].primary_rays(scene, cam, seed)
    _, _, got, _ = oracle.raygen(scene, cam, seed, want_rays=True)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("debug", [True, False], ids=["DEBUG", "RELEASE"])
def test_point_light_first_frames(modal_oracle, refs, debug):
    """lights[0] a point light: the sphere-light branch of getDirToLight / shadow.comp (light.glsl:107-126,
    nv_all.shadow.comp:106-122), which NielsScene (sun first) never takes."""
    oracle, ref = modal_oracle, refs[debug]
    oracle.set_mode(0 if debug else 1)
    scene = oracle.niels_scene(0.0, synthetic_sky())
    lights = scene.lights.reshape(3, 32).copy()
    scene.lights = np.ascontiguousarray(lights[[2, 1, 0]]).reshape(-1)
    w, h, samples = 96, 54, 2
    cam = oracle.camera(w, h, eye=(6, 5, 12))
    s_o, s_r = oracle.seed((1.0, 2.0)), oracle.seed((1.0, 2.0))
    got = oracle.frame(scene, cam, s_o, samples)
    want = ref.frame(scene, cam, s_r, samples)
    for k in ("dirT", "uvN", "bits", "lighting", "rgba8"):
        assert words_differing(got[k], want[k]) == 0, k
    assert want["bits"].any()


def test_soup_subsample_equals_the_reference_loop(oracle, refs):
    """The brute-force loop over a random soup (the configs[2] generator at 60000 triangles): reference shaders == oracle on
    4096 camera-like rays and on their shadow rays."""
    from igx_raytracing_b200 import rtb
    from oracle.oracle import Scene
    tris = rtb.gen_soup(60000, 0xB200)
    scene = Scene(tris, None, None, None, oracle.light_directional((-0.5, -2, -1), (0.9, 0.9, 0.9)),
                  oracle.material((0.8, 0.8, 0.8), (0, 0, 0), (0, 0, 0), 0.0, 1.0), np.zeros(60000, np.uint32), None, None)
    rng = np.random.default_rng(3)
    o = np.tile(np.array([0, 0, 13.9], np.float32), (4096, 1))
    d = np.concatenate([rng.uniform(-1.2, 1.2, (4096, 2)), -np.ones((4096, 1))], axis=1)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d], axis=1).astype(np.float32)
    a = oracle.trace_rays(scene, rays)
    b = refs[True].trace_rays(scene, rays)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)) and np.array_equal(a[3], b[3])
    assert (a[0] != 0xFFFFFFFF).sum() > 50
