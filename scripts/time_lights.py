"""RTB_OPT_LIGHTS: ms per frame of NielsScene at 1920x1080 with N small point lights, every light evaluated (mode 1) against the
same through per-tile light lists (mode 2).  Prints one JSON line.  (No reference counterpart: the reference samples light 0 only.)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from igx_raytracing_b200 import rtb
    n_lights = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    rng = np.random.default_rng(9)
    lights = [rtb.pack_light_directional((-0.5, -2.0, -1.0), (0.3, 0.3, 0.3))]
    for _ in range(n_lights):
        lights.append(rtb.pack_light_point(rng.uniform([-40, 0.2, -40], [40, 3, 40]), rng.uniform(0.2, 1.0, 3), float(rng.uniform(1.0, 2.5)), 0.05, 1.0))
    scene = rtb.niels_scene(0.0)
    scene["lights"] = np.concatenate([np.asarray(l).view(np.uint8).reshape(-1) for l in lights])
    info = np.asarray(scene["info"]).copy()
    info[0], info[6], info[8] = n_lights + 1, 1, n_lights
    scene["info"] = info
    w, h = 1920, 1080
    out = {"lights": n_lights + 1, "frame": f"NielsScene {w}x{h}, 1 shadow sample per light"}
    frames = {}
    for mode in (0, 1, 2):
        ctx = rtb.Context(max_lights=n_lights + 1)
        ctx.set_option(rtb.OPT_LIGHTS, mode)
        ctx.resize(w, h, 1)
        ctx.upload_scene(scene, None)
        ctx.build_accel(rtb.ACCEL_BVH)
        ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, eye=(6, 5, 12)))
        ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
        for _ in range(3):
            ctx.dispatch(rtb.PASS_FRAME)
        ctx.sync()
        t0 = time.perf_counter()
        n = 10 if mode != 1 or n_lights <= 2000 else 3
        for _ in range(n):
            ctx.dispatch(rtb.PASS_FRAME)
        ctx.sync()
        out[{0: "reference_light0_ms", 1: "all_lights_ms", 2: "tile_lists_ms"}[mode]] = (time.perf_counter() - t0) / n * 1e3
        ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
        ctx.dispatch(rtb.PASS_FRAME)
        frames[mode] = ctx.readback(rtb.TGT_RGBA8)
        ctx.close()
    out["tile_lists_equal_all_lights"] = bool(np.array_equal(frames[1], frames[2]))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
