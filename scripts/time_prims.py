"""RTB_OPT_PRIMITIVE_TREES: ms per 1920x1080 frame (1 shadow sample) over the reference's maximum primitive counts — 32768 spheres and
16384 cubes (ref: igx/include/helpers/scene_graph.hpp:137-142) plus NielsScene's triangles and plane — searched through the sphere /
cube trees against the reference's linear loops on the same GPU.  Prints one JSON line.  The frames must be identical."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from igx_raytracing_b200 import rtb
    n_sph = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    n_cub = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    rng = np.random.default_rng(5)
    base = rtb.niels_scene(0.0)
    sph = np.concatenate([rng.uniform([-40, 0.2, -40], [40, 12, 40], (n_sph, 3)), rng.uniform(0.1, 0.5, (n_sph, 1))], axis=1).astype(np.float32)
    lo = rng.uniform([-40, 0, -40], [40, 12, 40], (n_cub, 3)).astype(np.float32)
    cub = np.concatenate([lo, lo + rng.uniform(0.1, 0.9, (n_cub, 3)).astype(np.float32)], axis=1).astype(np.float32)
    n_obj = 3 + n_sph + n_cub + 1
    scene = dict(base, spheres=sph.view(np.uint8).reshape(-1), cubes=cub.view(np.uint8).reshape(-1),
                 material_indices=(np.arange(n_obj) % 8).astype(np.uint32), info=np.array([3, 8, 3, n_sph, n_cub, 1, 1, 0, 2], np.uint32))
    w, h = 1920, 1080
    out = {"spheres": n_sph, "cubes": n_cub, "frame": f"{w}x{h}, 1 shadow sample"}
    frames = {}
    for trees in (64, 0):
        ctx = rtb.Context(max_spheres=max(n_sph, 64), max_cubes=max(n_cub, 64))
        ctx.set_option(rtb.OPT_PRIMITIVE_TREES, trees)
        ctx.resize(w, h, 1)
        ctx.upload_scene(scene, None)
        ctx.build_accel(rtb.ACCEL_BVH)
        ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, eye=(6, 5, 30)))
        ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
        t0 = time.perf_counter()
        ctx.dispatch(rtb.PASS_FRAME)
        ctx.sync()
        first = (time.perf_counter() - t0) * 1e3
        for _ in range(2):
            ctx.dispatch(rtb.PASS_FRAME)
        ctx.sync()
        n = 20 if trees else 3
        t0 = time.perf_counter()
        for _ in range(n):
            ctx.dispatch(rtb.PASS_FRAME)
        ctx.sync()
        key = "trees" if trees else "loops"
        out[key + "_ms"] = (time.perf_counter() - t0) / n * 1e3
        out[key + "_first_frame_ms"] = first
        info = ctx.accel_info()
        if trees:
            out["sphere_tree_nodes"], out["cube_tree_nodes"] = info.sphere_tree_nodes, info.cube_tree_nodes
            moved = sph.copy()
            moved[:, 1] += 0.1
            t0 = time.perf_counter()
            ctx.upload(rtb.BUF_SPHERES, moved.view(np.uint8).reshape(-1))
            ctx.dispatch(rtb.PASS_FRAME)
            ctx.sync()
            out["frame_after_sphere_upload_ms"] = (time.perf_counter() - t0) * 1e3
            ctx.upload(rtb.BUF_SPHERES, sph.view(np.uint8).reshape(-1))
        ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
        ctx.dispatch(rtb.PASS_FRAME)
        frames[trees] = (ctx.readback(rtb.TGT_DIR_T).copy(), ctx.readback(rtb.TGT_RGBA8).copy())
        ctx.close()
    out["hit_fraction"] = float((frames[64][0][..., 3].view(np.uint32) != 0xFFFFFFFF).mean())
    out["frames_identical"] = bool(np.array_equal(frames[64][0].view(np.uint32), frames[0][0].view(np.uint32)) and np.array_equal(frames[64][1], frames[0][1]))
    out["speedup"] = out["loops_ms"] / out["trees_ms"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
