"""What one rank of an N-GPU run does, measured on ONE GPU: the 4K soup frame with RTB_OPT_TILE_COUNT = N, RTB_OPT_TILE_RANK = 0
(1/N of the 32x32-pixel blocks), per launch and per frame, against the ideal 1/N of the single-GPU time.  One JSON line per N.
Extra arguments KEY=VALUE are passed to rtb_set_option by number (e.g. 8=2 for two lanes)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from igx_raytracing_b200 import rtb
    opts = [tuple(int(v) for v in a.split("=")) for a in sys.argv[1:] if "=" in a]
    n = 1_000_000
    tris = rtb.gen_soup(n, 0xB200)
    sun = rtb.niels_scene()["lights"][:32]
    mat = rtb.pack_material((0.8, 0.8, 0.8), (0.05, 0.05, 0.05), (0, 0, 0), 0.0, 1.0, 1.0)
    scene = dict(triangles=tris, lights=sun, materials=mat, material_indices=np.zeros(n, np.uint32), info=np.array([1, 1, n, 0, 0, 0, 1, 0, 0], np.uint32))
    w, h = 3840, 2160
    base = None
    for tiles in (1, 2, 4, 8):
        ctx = rtb.Context(max_triangles=n)
        ctx.set_option(rtb.OPT_TILE_COUNT, tiles)
        ctx.set_option(rtb.OPT_TILE_RANK, 0)
        for k, v in opts:
            ctx.set_option(k, v)
        ctx.resize(w, h, 1)
        ctx.upload_scene(scene, None)
        ctx.build_accel(rtb.ACCEL_BVH)
        ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, eye=(0.0, 0.0, 13.9)))
        ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
        for _ in range(5):
            ctx.dispatch(rtb.PASS_FRAME)
        ctx.sync()
        frames = 200
        t0 = time.perf_counter()
        for _ in range(frames):
            ctx.dispatch(rtb.PASS_FRAME)
        ctx.sync()
        graph_ms = (time.perf_counter() - t0) / frames * 1e3
        ctx.set_option(rtb.OPT_FRAME_LANES, 1)
        ctx.set_option(rtb.OPT_FRAME_GRAPH, 0)
        ph = np.zeros(8)
        for _ in range(10):
            ctx.dispatch(rtb.PASS_FRAME)
            ph += np.array(ctx.last_frame_ms())
        ph /= 10
        ctx.close()
        if tiles == 1:
            base = (graph_ms, ph.copy())
        names = ["init", "raygen", "trace_primary", "finish", "shadowgen", "trace_shadow", "shade", "total"]
        print(json.dumps({"tiles": tiles, "frame_ms": round(graph_ms, 4), "ideal_ms": round(base[0] / tiles, 4), "efficiency": round(base[0] / tiles / graph_ms, 3),
                          "phase_ms": {k: round(float(v), 4) for k, v in zip(names, ph)},
                          "phase_efficiency": {k: round(float(b / tiles / v), 3) for k, v, b in zip(names, ph, base[1]) if v > 1e-4}}), flush=True)


if __name__ == "__main__":
    main()
