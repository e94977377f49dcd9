"""How much would overlapping consecutive frames buy?  Two independent contexts render the same rank-local frame on their own
streams, dispatches interleaved from one host thread: frames per second of the pair against one context alone."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from igx_raytracing_b200 import rtb
    n = 1_000_000
    tris = rtb.gen_soup(n, 0xB200)
    sun = rtb.niels_scene()["lights"][:32]
    mat = rtb.pack_material((0.8, 0.8, 0.8), (0.05, 0.05, 0.05), (0, 0, 0), 0.0, 1.0, 1.0)
    scene = dict(triangles=tris, lights=sun, materials=mat, material_indices=np.zeros(n, np.uint32), info=np.array([1, 1, n, 0, 0, 0, 1, 0, 0], np.uint32))
    w, h = 3840, 2160
    for tiles in (1, 8):
        ctxs = []
        for k in range(2):
            ctx = rtb.Context(max_triangles=n)
            ctx.set_option(rtb.OPT_TILE_COUNT, tiles)
            ctx.set_option(rtb.OPT_TILE_RANK, 0)
            ctx.set_option(rtb.OPT_ACCEL_BUILDER, 1)
            ctx.resize(w, h, 1)
            ctx.upload_scene(scene, None)
            ctx.build_accel(rtb.ACCEL_BVH)
            ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, eye=(0.0, 0.0, 13.9)))
            ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
            for _ in range(5):
                ctx.dispatch(rtb.PASS_FRAME)
            ctx.sync()
            ctxs.append(ctx)
        frames = 200
        t0 = time.perf_counter()
        for _ in range(frames):
            ctxs[0].dispatch(rtb.PASS_FRAME)
        ctxs[0].sync()
        one = (time.perf_counter() - t0) / frames * 1e3
        t0 = time.perf_counter()
        for _ in range(frames):
            ctxs[0].dispatch(rtb.PASS_FRAME)
            ctxs[1].dispatch(rtb.PASS_FRAME)
        ctxs[0].sync(); ctxs[1].sync()
        two = (time.perf_counter() - t0) / (2 * frames) * 1e3
        print(json.dumps({"tiles": tiles, "ms_per_frame_one_context": round(one, 4), "ms_per_frame_two_contexts_interleaved": round(two, 4), "gain": round(one / two, 3)}), flush=True)
        for c in ctxs:
            c.close()


if __name__ == "__main__":
    main()
