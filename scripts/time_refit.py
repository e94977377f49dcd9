"""Times rtb_refit_accel against rtb_build_accel on the 1M-triangle soup (B200): python scripts/time_refit.py [n_triangles]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from igx_raytracing_b200 import rtb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
tris = rtb.gen_soup(n, 0xB200)
mat = rtb.pack_material((0.8, 0.8, 0.8), (0.05, 0.05, 0.05), (0, 0, 0), 0.0, 1.0, 1.0)
scene = dict(triangles=tris, lights=rtb.niels_scene()["lights"][:32], materials=mat, material_indices=np.zeros(n, np.uint32),
             info=np.array([1, 1, n, 0, 0, 0, 1, 0, 0], np.uint32))
ctx = rtb.Context(max_triangles=n)
ctx.resize(64, 64, 1)
ctx.upload_scene(scene, None)
t0 = time.perf_counter(); ctx.build_accel(rtb.ACCEL_BVH); ctx.sync(); t_build = time.perf_counter() - t0
f = tris.copy().view(np.float32).reshape(n, 12)
f.reshape(n, 3, 4)[:, :, :3] += np.random.default_rng(0).uniform(-0.05, 0.05, (n, 1, 3)).astype(np.float32)
moved = f.reshape(-1).view(np.uint8)
times = []
for i in range(6):
    t0 = time.perf_counter(); ctx.upload(rtb.BUF_TRIANGLES, moved); ctx.sync(); t_up = time.perf_counter() - t0
    t0 = time.perf_counter(); ctx.refit_accel(); ctx.sync(); times.append(time.perf_counter() - t0)
info = ctx.accel_info()
print(f"triangles {n}  nodes {info.node_count}  host build {t_build * 1e3:.1f} ms  upload of all triangles {t_up * 1e3:.2f} ms  "
      f"device refit (wall, incl. launches and a 24-byte read-back) min {min(times) * 1e3:.3f} ms  median {sorted(times)[len(times) // 2] * 1e3:.3f} ms  refits {info.refits}")
ctx.close()
