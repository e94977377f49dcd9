"""Renders a few frames of a bench.py workload and nothing else: the process ncu wraps (see B200_PROFILING.md).

    ncu --set full --clock-control none --import-source on -k regex:k_trace --launch-skip S -c C -o gpurun_out/x python scripts/profile_frame.py --workload soup1m --frames 2
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from igx_raytracing_b200 import rtb
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="soup1m", choices=list(bench.WORKLOADS))
    ap.add_argument("--frames", type=int, default=2)
    ap.add_argument("--shadow-order", type=int, default=1)
    args = ap.parse_args()
    wl = bench.WORKLOADS[args.workload]
    scene, limits = bench.build_scene(rtb, wl)
    w, h = wl["width"], wl["height"]
    ctx = rtb.Context(**limits)
    ctx.set_option(rtb.OPT_SHADOW_ORDER, args.shadow_order)
    ctx.resize(w, h, wl["samples"])
    ctx.upload_scene(scene, None)
    ctx.build_accel(rtb.ACCEL_BVH)
    ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, **bench.camera_kwargs(wl)))
    ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
    for _ in range(args.frames):
        if wl.get("bounces"):
            ctx.path_frame(wl["bounces"])
        else:
            ctx.dispatch(rtb.PASS_FRAME)
        ctx.sync()
    ctx.close()


if __name__ == "__main__":
    main()
