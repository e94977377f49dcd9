"""Writes tests/golden/*.npz: frames rendered by THE REFERENCE'S OWN SHADERS on fixed inputs.

The generator is oracle/_ref (oracle/ref_shim/: the reference's res/shaders sources, read where they lie under
/root/reference, adapted for syntax and compiled for the host; -DDEBUG build = what the shipped .spv are, -DRELEASE for the
cases marked "release").  The hand-written oracle must reproduce every fixture bit for bit, or this script refuses to write.
The fixtures are committed so that

  * the oracle cannot drift from the reference unnoticed, also where /root/reference does not exist (tests/test_golden.py, CPU), and
  * the CUDA path is compared against bytes that came out of the reference's code, not out of our restatement
    (tests/test_golden.py::test_cuda_equals_golden, GPU).

    python scripts/make_golden.py            # rewrites every fixture (needs /root/reference)
    python scripts/make_golden.py --check    # regenerates in memory and compares with the committed files

The only reference-owned known answers for HOST packing are the 22 f32 -> f16 conversions of core2/test/test.cpp:8-29; they
are stored in f16_kat.json next to the frames.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# name -> scene time, size, shadow samples, camera kwargs, cpuOffset, frames, skybox?, [degenerate: triangle 1 gets p1 = p0],
#         [release: the -DRELEASE build of the shaders, targets zero before the frame]
CASES = {
    "niels_default_96x54": dict(time=0.0, w=96, h=54, samples=1, cam=dict(eye=(4, 2, -2)), off=(0.0, 0.0), frames=1, sky=True),
    "niels_all_objects_128x72": dict(time=0.0, w=128, h=72, samples=2, cam=dict(eye=(6, 5, 12)), off=(0.0, 0.0), frames=1, sky=True),
    "niels_inside_cube_64x36": dict(time=0.0, w=64, h=36, samples=1, cam=dict(eye=(0.5, 0.5, 0.5)), off=(3.0, 9.0), frames=1, sky=False),
    "niels_rotated_ragged_77x45": dict(time=0.7, w=77, h=45, samples=3, cam=dict(eye=(6, 5, 12), pitch=0.2, yaw=0.4, roll=0.1),
                                       off=(12.5, -431.25), frames=1, sky=True),
    "niels_accumulate4_64x36": dict(time=0.0, w=64, h=36, samples=1, cam=dict(eye=(6, 5, 12), flags=2), off=(3.0, 9.0), frames=4, sky=True),
    "niels_omni_64x32": dict(time=0.0, w=64, h=32, samples=1, cam=dict(eye=(6, 5, 12), projection=1, yaw=0.3), off=(0.0, 0.0), frames=1, sky=True),
    "niels_stereo_tb_64x32": dict(time=0.0, w=64, h=32, samples=1, cam=dict(eye=(6, 5, 12), projection=2, yaw=0.3), off=(0.0, 0.0), frames=1, sky=True),
    "niels_degenerate_tri_96x54": dict(time=0.0, w=96, h=54, samples=1, cam=dict(eye=(2, 5, 6), pitch=-0.25), off=(0.0, 0.0), frames=1, sky=True,
                                       degenerate=True),
    "niels_release_96x54": dict(time=0.0, w=96, h=54, samples=2, cam=dict(eye=(6, 5, 12)), off=(0.0, 0.0), frames=1, sky=True, release=True),
    # diffuse bounces (rtb_path_frame): NO reference semantics beyond depth 0 (the reference traces no secondary rays), so this fixture
    # is the ORACLE's statement of our definition (oracle.h orc_path_frame); its depth 0 is checked against the reference's raygen.comp
    "niels_path4_64x36": dict(time=0.0, w=64, h=36, samples=1, cam=dict(eye=(6, 5, 12), flags=2), off=(3.0, 9.0), frames=2, sky=True, path_bounces=4),
    "niels_release_degenerate_80x45": dict(time=0.0, w=80, h=45, samples=1, cam=dict(eye=(2, 5, 6), pitch=-0.25), off=(0.0, 0.0), frames=1,
                                           sky=True, degenerate=True, release=True),
}

F16_KAT = [   # core2/test/test.cpp:8-29 (f32 bits, expected f16 bits)
    (0x00000000, 0x0000), (0x80000000, 0x8000), (0x3f800000, 0x3c00), (0xbf800000, 0xbc00), (0x3f000000, 0x3800),
    (0x3e800000, 0x3400), (0x3e000000, 0x3000), (0x3eaaaaab, 0x3555), (0x38002000, 0x0001), (0x47000000, 0x7800),
    (0x477fe000, 0x7bff), (0x10001999, 0x0000), (0x00002000, 0x0000), (0x48000000, 0x7c00), (0x477ff000, 0x7c00),
    (0x40490fdb, 0x4248), (0x402d70a4, 0x416b), (0x7f800000, 0x7c00), (0xff800000, 0xfc00), (0xff800001, 0xffff),
    (0x7f800001, 0x7fff), (0x40a9999a, 0x454c),
]


def render(engine, oracle, case):
    """engine: oracle.ref.Ref or oracle.oracle.Oracle (same frame() signature); scenes and cameras come from the oracle's
    host packing (pinned separately by the f16 known answers and the SURVEY Appendix-B anchors)."""
    from conftest import case_scene
    scene = case_scene(oracle, case)
    w, h = case["w"], case["h"]
    cam = oracle.camera(w, h, **case["cam"])
    seed = oracle.seed(tuple(case["off"]))
    accum = np.zeros((h, w, 4), np.float32)
    pre = None
    for _ in range(case["frames"]):
        ref = engine.frame(scene, cam, seed, case["samples"], accum=accum, prefill=pre)
        pre = {k: ref[k] for k in ("dirT", "uvN", "bits", "lighting")}
    return dict(camera=np.frombuffer(bytes(cam), np.uint8).copy(), seed_after=np.frombuffer(bytes(seed), np.uint8).copy(),
                dirT=ref["dirT"], uvN=ref["uvN"], bits=ref["bits"], lighting=ref["lighting"], rgba8=ref["rgba8"], accum=accum)


def render_path(oracle, ref, case):
    """rtb_path_frame's definition as the oracle states it; depth 0 (the G-buffer) must be the reference's raygen.comp output"""
    from conftest import case_scene
    scene = case_scene(oracle, case)
    w, h = case["w"], case["h"]
    cam = oracle.camera(w, h, **case["cam"])
    seed = oracle.seed(tuple(case["off"]))
    accum = np.zeros((h, w, 4), np.float32)
    for _ in range(case["frames"]):
        out = oracle.path_frame(scene, cam, seed, case["path_bounces"], accum=accum)
        ref_dirT, _ = ref.raygen(scene, cam, seed)   # seed is the state after this frame's init pass
        if not np.array_equal(ref_dirT.view(np.uint32), out["dirT"].view(np.uint32)):
            sys.exit("path fixture: depth 0 differs from the reference's raygen.comp")
    return dict(camera=np.frombuffer(bytes(cam), np.uint8).copy(), seed_after=np.frombuffer(bytes(seed), np.uint8).copy(),
                dirT=out["dirT"], rgba8=out["rgba8"], accum=accum)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    from oracle.oracle import Oracle
    from oracle import ref as refmod
    orc = Oracle()
    if not refmod.build():
        sys.exit("scripts/make_golden.py needs the reference tree (oracle/_ref is built from it)")
    refs = {False: refmod.Ref(debug=True), True: refmod.Ref(debug=False)}
    os.makedirs(GOLDEN, exist_ok=True)
    bad = 0
    for name, case in CASES.items():
        release = bool(case.get("release"))
        if case.get("path_bounces") is not None:
            out = render_path(orc, refs[False], case)
        else:
            out = render(refs[release], orc, case)
            orc.set_mode(1 if release else 0)
            mine = render(orc, orc, case)
            orc.set_mode(0)
            for k, v in out.items():
                if not np.array_equal(np.asarray(v).view(np.uint8), np.asarray(mine[k]).view(np.uint8)):
                    sys.exit(f"{name}: the oracle's {k} differs from the reference shaders — fix the oracle before writing fixtures")
        path = os.path.join(GOLDEN, name + ".npz")
        if args.check:
            have = np.load(path)
            for k, v in out.items():
                if not np.array_equal(np.asarray(v).view(np.uint8), have[k].view(np.uint8)):
                    print(f"{name}: {k} differs from the committed fixture")
                    bad += 1
        else:
            np.savez_compressed(path, **out)
            print(f"{path}: {os.path.getsize(path) / 1024:.0f} KiB")
    kat = os.path.join(GOLDEN, "f16_kat.json")
    if not args.check:
        with open(kat, "w") as f:
            json.dump({"source": "igx/igxi-tool/igxi/ignis/core2/test/test.cpp:8-29", "f32_bits_to_f16_bits": F16_KAT}, f, indent=1)
        with open(os.path.join(GOLDEN, "cases.json"), "w") as f:
            json.dump(dict(generator="oracle/_ref (the reference's shaders compiled for the host; scripts/make_golden.py)", cases=CASES), f, indent=1)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
