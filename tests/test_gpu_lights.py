"""Beyond the reference: every light evaluated instead of lights[0] x lightCount (RTB_OPT_LIGHTS = 1), the
same through per-tile light lists (= 2, the reference's LIGHTS_PER_TILE sketch), and the temporal History blend of the lighting
texture (RTB_OPT_HISTORY_ALPHA).  The reference implements none of these, so the semantics are ours (include/rtb200.h) and the
checker is the oracle's statement of the same definitions; the tile lists must not change a single bit of the all-lights result."""
import numpy as np
import pytest

from conftest import synthetic_sky


def many_lights(rtb, rng, n, radius, sun=True):
    out = []
    if sun:
        out.append(rtb.pack_light_directional((-0.5, -2.0, -1.0), (0.3, 0.3, 0.3)))
    for _ in range(n):
        pos = rng.uniform([-8, 0.2, -8], [8, 4, 8])
        out.append(rtb.pack_light_point(pos, rng.uniform(0.2, 1.0, 3), float(radius * rng.uniform(0.6, 1.4)), float(0.05 * radius), float(rng.choice([1.0, 2.0, 0.5]))))
    return np.concatenate([np.asarray(l).view(np.uint8).reshape(-1) for l in out])


def render(rtb, scene, sky, w, h, samples, mode, cam_kw, accel=1, frames=1, history=0.0, seeds=((3.0, 9.0),)):
    ctx = rtb.Context(max_lights=max(16, int(scene["info"][0])))
    ctx.set_option(rtb.OPT_LIGHTS, mode)
    if history:
        ctx.set_history_alpha(history)
    ctx.resize(w, h, samples)
    ctx.upload_scene(scene, sky)
    ctx.build_accel(accel)
    ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, **cam_kw))
    outs = []
    for f in range(frames):
        ctx.upload(rtb.BUF_SEED, rtb.make_seed(seeds[f % len(seeds)]))
        ctx.dispatch(rtb.PASS_FRAME)
        outs.append(dict(bits=ctx.readback(rtb.TGT_SHADOW_BITS), lighting=ctx.readback(rtb.TGT_LIGHTING), rgba8=ctx.readback(rtb.TGT_RGBA8)))
    ctx.close()
    return outs


@pytest.mark.gpu
@pytest.mark.parametrize("accel", [0, 1, 2])
def test_all_lights_equals_the_oracle_statement(rtb, oracle, accel):
    """NielsScene's three lights (a sun and two point lights), two samples: every target against the oracle with all lights on"""
    w, h, samples = 160, 90, 2
    cam_kw = dict(eye=(6, 5, 12))
    sky = synthetic_sky()
    oracle.set_all_lights(True)
    try:
        want = oracle.frame(oracle.niels_scene(0.0, sky), oracle.camera(w, h, **cam_kw), oracle.seed((3.0, 9.0)), samples)
    finally:
        oracle.set_all_lights(False)
    for mode in (1, 2):
        got = render(rtb, rtb.niels_scene(0.0), sky, w, h, samples, mode, cam_kw, accel)[0]
        assert got["bits"].size == want["bits"].size == 3 * samples * ((w + 15) // 16) * ((h + 1) // 2)
        assert np.array_equal(got["bits"], want["bits"]), f"mode {mode}: shadow layers differ"
        assert int((got["lighting"].reshape(-1, 4) != want["lighting"].reshape(-1, 4)).any(-1).sum()) <= 2
        assert int((got["rgba8"] != want["rgba8"]).sum()) <= 2
    ref_mode = render(rtb, rtb.niels_scene(0.0), sky, w, h, samples, 0, cam_kw, accel)[0]
    assert not np.array_equal(ref_mode["rgba8"], got["rgba8"]), "all lights must differ from light 0 x lightCount on this scene"


@pytest.mark.gpu
@pytest.mark.parametrize("radius,expect_overflow", [(1.5, False), (30.0, True)])
def test_tile_light_lists_do_not_change_a_bit(rtb, oracle, radius, expect_overflow):
    """300 point lights + a sun over NielsScene: the per-tile lists (mode 2) against the full loop (mode 1), small radii (lists of a
    few lights) and radii that cover the scene (every tile overflows its 32 entries and keeps the full loop); and against the oracle"""
    rng = np.random.default_rng(9)
    scene = rtb.niels_scene(0.0)
    scene["lights"] = many_lights(rtb, rng, 300, radius)
    info = np.asarray(scene["info"]).copy()
    info[0], info[6], info[8] = 301, 1, 300
    scene["info"] = info
    w, h, samples = 128, 72, 1
    cam_kw = dict(eye=(6, 5, 12))
    a = render(rtb, scene, synthetic_sky(), w, h, samples, 1, cam_kw)[0]
    b = render(rtb, scene, synthetic_sky(), w, h, samples, 2, cam_kw)[0]
    for k in ("bits", "lighting", "rgba8"):
        assert np.array_equal(a[k], b[k]), k
    assert a["bits"].any() and a["lighting"][..., :3].any()
    from oracle.oracle import Scene
    osc = oracle.niels_scene(0.0, synthetic_sky())
    osc = Scene(osc.triangles, osc.spheres, osc.cubes, osc.planes, scene["lights"], osc.materials, osc.material_indices, info, synthetic_sky())
    oracle.set_all_lights(True)
    try:
        want = oracle.frame(osc, oracle.camera(w, h, **cam_kw), oracle.seed((3.0, 9.0)), samples)
    finally:
        oracle.set_all_lights(False)
    assert np.array_equal(b["bits"], want["bits"])
    assert int((b["lighting"].reshape(-1, 4) != want["lighting"].reshape(-1, 4)).any(-1).sum()) <= 3
    assert int((b["rgba8"] != want["rgba8"]).sum()) <= 3


@pytest.mark.gpu
def test_history_blend_equals_the_oracle_statement(rtb, oracle):
    """four frames with different seeds, alpha 0.25: lighting and the frame against the oracle's pass-by-pass composition"""
    w, h, samples, alpha = 128, 72, 1, 0.25
    cam_kw = dict(eye=(6, 5, 12))
    sky = synthetic_sky()
    seeds = ((3.0, 9.0), (1.0, 2.0), (7.5, -3.0), (0.0, 4.0))
    got = render(rtb, rtb.niels_scene(0.0), sky, w, h, samples, 0, cam_kw, frames=4, history=alpha, seeds=seeds)
    osc, cam = oracle.niels_scene(0.0, sky), oracle.camera(w, h, **cam_kw)
    oracle.set_history(alpha)
    history = np.zeros((h, w, 4), np.uint16)
    for f in range(4):
        seed = oracle.init_pass(oracle.seed(seeds[f]))
        dirT, uvN, _, _ = oracle.raygen(osc, cam, seed)
        bits = oracle.shadow(osc, cam, seed, samples, dirT)
        l16, _ = oracle.lighting(osc, cam, samples, dirT, uvN, bits)
        oracle.history_blend(history, l16, first=(f == 0))
        rgba = oracle.composite(osc, cam, seed, dirT, uvN, l16)
        assert int((got[f]["lighting"].reshape(-1, 4) != l16.reshape(-1, 4)).any(-1).sum()) <= 2, f
        assert int((got[f]["rgba8"] != rgba).sum()) <= 2, f
    plain = render(rtb, rtb.niels_scene(0.0), sky, w, h, samples, 0, cam_kw, frames=4, seeds=seeds)
    assert not np.array_equal(plain[3]["lighting"], got[3]["lighting"])
