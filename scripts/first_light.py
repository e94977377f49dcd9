"""Quick GPU timing of the soup scene (not the bench): python scripts/first_light.py [ntri] [w] [h] [frames]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from igx_raytracing_b200 import rtb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
w = int(sys.argv[2]) if len(sys.argv) > 2 else 3840
h = int(sys.argv[3]) if len(sys.argv) > 3 else 2160
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 10
ez = float(sys.argv[5]) if len(sys.argv) > 5 else 13.9
mode = int(sys.argv[6]) if len(sys.argv) > 6 else 1   # 1 = 8-wide compressed BVH, 2 = binary BVH
t0 = time.time()
tris = rtb.gen_soup(n)
mat = rtb.pack_material((0.8, 0.8, 0.8), (0.05, 0.05, 0.05), (0, 0, 0), 0.0, 1.0, 1.0)
scene = dict(triangles=tris, lights=rtb.niels_scene()["lights"][:32], materials=mat, material_indices=np.zeros(n, np.uint32),
             info=np.array([1, 1, n, 0, 0, 0, 1, 0, 0], np.uint32))
ctx = rtb.Context(max_triangles=n)
ctx.resize(w, h, 1)
ctx.upload_scene(scene, None)
t1 = time.time()
ctx.build_accel(mode)
info = ctx.accel_info()
print(f"gen+upload {t1 - t0:.2f}s  build {info.build_ms:.0f} ms nodes {info.node_count} leaves {info.leaf_count} depth {info.max_depth} sah {info.sah_cost:.1f}", flush=True)
ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, eye=(0, 0, ez)))
ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
ctx.set_option(rtb.OPT_COUNTERS, 1)
ctx.dispatch(rtb.PASS_FRAME); ctx.sync()
c = ctx.counters()
print(f"primary rays {c.primary_rays} nodes/ray {c.primary_nodes / max(c.primary_rays,1):.1f} tris/ray {c.primary_tris / max(c.primary_rays,1):.2f} hits {c.primary_hits}")
print(f"shadow  rays {c.shadow_rays} nodes/ray {c.shadow_nodes / max(c.shadow_rays,1):.1f} tris/ray {c.shadow_tris / max(c.shadow_rays,1):.2f} occluded {c.shadow_occluded}")
print("instrumented ms", [round(x, 3) for x in ctx.last_frame_ms()])
ctx.set_option(rtb.OPT_COUNTERS, 0)
rays = c.primary_rays + c.shadow_rays
for i in range(frames):
    ctx.dispatch(rtb.PASS_FRAME)
    ms = ctx.last_frame_ms()
    print(f"frame {i}: init {ms[0]:.3f} raygen {ms[1]:.3f} trace {ms[2]:.3f} finish {ms[3]:.3f} shgen {ms[4]:.3f} shtrace {ms[5]:.3f} shade {ms[6]:.3f} total {ms[7]:.3f} ms -> {rays / ms[7] / 1e3:.1f} Mrays/s", flush=True)
