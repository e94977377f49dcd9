"""Committed golden fixtures (tests/golden/, written by scripts/make_golden.py FROM THE REFERENCE'S OWN SHADERS compiled for
the host: oracle/_ref, see oracle/ref_shim/).

CPU half: the oracle reproduces every fixture byte for byte (so the checker cannot drift from the reference, also where
/root/reference is absent), and the reference's own f16 known answers are read from the fixture file.  GPU half: the CUDA
path, through the C ABI, against the same bytes — DEBUG-build fixtures (what the shipped .spv are), RELEASE-build fixtures
(RTB_OPT_SHADER_BUILD = 1) and a scene with a degenerate triangle.
"""
import json
import os

import numpy as np
import pytest

from conftest import case_scene, degenerate_triangles, synthetic_sky

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = json.load(open(os.path.join(GOLDEN, "cases.json")))["cases"]
KEYS = ("dirT", "uvN", "bits", "lighting", "rgba8", "accum")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def test_f16_kat_fixture(oracle):
    kat = json.load(open(os.path.join(GOLDEN, "f16_kat.json")))
    assert len(kat["f32_bits_to_f16_bits"]) == 22
    for bits, want in kat["f32_bits_to_f16_bits"]:
        v = np.array([bits], np.uint32).view(np.float32)[0]
        assert oracle.f16_trunc(v) == want


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_golden(oracle, name):
    case, gold = CASES[name], load(name)
    scene = case_scene(oracle, case)
    w, h = case["w"], case["h"]
    cam = oracle.camera(w, h, **case["cam"])
    assert np.array_equal(cam, gold["camera"])
    seed = oracle.seed(tuple(case["off"]))
    accum = np.zeros((h, w, 4), np.float32)
    if case.get("path_bounces") is not None:   # the bounce definition (ours; see scripts/make_golden.py)
        for _ in range(case["frames"]):
            ref = oracle.path_frame(scene, cam, seed, case["path_bounces"], accum=accum)
        assert np.array_equal(seed, gold["seed_after"])
        for k, v in (("dirT", ref["dirT"]), ("rgba8", ref["rgba8"]), ("accum", accum)):
            assert np.array_equal(np.asarray(v).view(np.uint8), gold[k].view(np.uint8)), f"{name}: oracle {k} drifted from the fixture"
        return
    oracle.set_mode(1 if case.get("release") else 0)
    try:
        pre = None
        for _ in range(case["frames"]):
            ref = oracle.frame(scene, cam, seed, case["samples"], accum=accum, prefill=pre)
            pre = {k: ref[k] for k in ("dirT", "uvN", "bits", "lighting")}
    finally:
        oracle.set_mode(0)
    ref["accum"] = accum
    assert np.array_equal(seed, gold["seed_after"])
    for k in KEYS:
        assert np.array_equal(np.asarray(ref[k]).view(np.uint8), gold[k].view(np.uint8)), f"{name}: oracle {k} drifted from the fixture"


@pytest.mark.gpu
@pytest.mark.parametrize("accel", [0, 1, 2])
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_equals_golden(rtb, name, accel):
    """No oracle on this path: the fixture bytes are the expectation.  Budget: the 1-ulp binary64 transcendental
    differences between CUDA and glibc — at these sizes that is 0 texels almost always; allow 2."""
    case, gold = CASES[name], load(name)
    sky = synthetic_sky() if case["sky"] else None
    w, h = case["w"], case["h"]
    ctx = rtb.Context()
    ctx.resize(w, h, case["samples"])
    scene = rtb.niels_scene(case["time"])
    if case.get("degenerate"):
        scene["triangles"] = degenerate_triangles(scene["triangles"])
    ctx.upload_scene(scene, sky)
    ctx.set_option(rtb.OPT_SHADER_BUILD, rtb.SHADER_RELEASE if case.get("release") else rtb.SHADER_DEBUG)
    ctx.build_accel(accel)
    cam = rtb.pack_camera(w, h, **case["cam"])
    assert np.array_equal(cam, gold["camera"]), "host camera packing differs from the fixture"
    ctx.upload(rtb.BUF_CAMERA, cam)
    ctx.upload(rtb.BUF_SEED, rtb.make_seed(tuple(case["off"])))
    if case.get("path_bounces") is not None:
        for _ in range(case["frames"]):
            ctx.path_frame(case["path_bounces"])
        got = dict(dirT=ctx.readback(rtb.TGT_DIR_T), rgba8=ctx.readback(rtb.TGT_RGBA8), accum=ctx.readback(rtb.TGT_ACCUM))
        seed = ctx.readback(rtb.TGT_SEED)
        ctx.close()
        assert np.array_equal(seed, gold["seed_after"])
        for k in ("dirT", "rgba8", "accum"):
            g, r = np.asarray(got[k]).view(np.uint8).reshape(-1, 4), gold[k].view(np.uint8).reshape(-1, 4)
            bad = int((g != r).any(axis=-1).sum())
            assert bad <= 4, f"{name}: {bad} 32-bit words of {k} differ from the fixture"
        return
    for _ in range(case["frames"]):
        ctx.dispatch(rtb.PASS_FRAME)
    got = dict(dirT=ctx.readback(rtb.TGT_DIR_T), uvN=ctx.readback(rtb.TGT_UV_NORMAL), bits=ctx.readback(rtb.TGT_SHADOW_BITS),
               lighting=ctx.readback(rtb.TGT_LIGHTING), rgba8=ctx.readback(rtb.TGT_RGBA8), accum=ctx.readback(rtb.TGT_ACCUM))
    seed = ctx.readback(rtb.TGT_SEED)
    ctx.close()
    assert np.array_equal(seed, gold["seed_after"])
    ids_g, ids_r = got["dirT"][..., 3].view(np.uint32), gold["dirT"][..., 3].view(np.uint32)
    assert np.array_equal(ids_g, ids_r), f"{int((ids_g != ids_r).sum())} hit ids differ from the fixture"
    for k in KEYS:
        if k == "accum" and case["frames"] == 1 and not (case["cam"].get("flags", 0) & 2):
            continue   # accumulation target is only defined under USE_SUPERSAMPLING
        g, r = np.asarray(got[k]).view(np.uint8).reshape(-1, 4), gold[k].view(np.uint8).reshape(-1, 4)
        bad = int((g != r).any(axis=-1).sum())
        assert bad <= 4, f"{name}: {bad} 32-bit words of {k} differ from the fixture"
