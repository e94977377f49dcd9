"""Diffuse bounces (BASELINE.json configs[3], rtb_path_frame).  The reference traces no secondary rays, so there is nothing of the
reference to compare deeper vertices with ("parity unpinned" beyond depth 0): depth 0 is held to the reference path
(the oracle, itself pinned by oracle/_ref), and the whole path to (a) the oracle's CPU restatement of the same definition on small
scenes with every primitive type, and (b) the reference's linear triangle loop (RTB_ACCEL_BRUTE) on a 1/64 tile subsample of the
10M-triangle frame at every depth."""
import numpy as np
import pytest

from conftest import synthetic_sky


def gpu_path(rtb, scene, sky, w, h, cam_kw, bounces, accel, seed_off=(0.0, 0.0), frames=1, limits=None, tile=None):
    ctx = rtb.Context(**(limits or {}))
    if tile:
        ctx.set_option(rtb.OPT_TILE_COUNT, tile[1])
        ctx.set_option(rtb.OPT_TILE_RANK, tile[0])
    ctx.resize(w, h, 1)
    ctx.upload_scene(scene, sky)
    ctx.build_accel(accel)
    ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, **cam_kw))
    ctx.upload(rtb.BUF_SEED, rtb.make_seed(seed_off))
    for _ in range(frames):
        ctx.path_frame(bounces)
    out = dict(rgba8=ctx.readback(rtb.TGT_RGBA8), accum=ctx.readback(rtb.TGT_ACCUM), dirT=ctx.readback(rtb.TGT_DIR_T), stats=ctx.path_stats().as_dict())
    if tile:
        out["tiled"] = ctx.readback(rtb.TGT_RGBA8_TILED)
    ctx.close()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("accel", [0, 1, 2])
@pytest.mark.parametrize("bounces", [0, 1, 4])
def test_path_frame_equals_the_oracle_statement_on_nielsscene(rtb, oracle, accel, bounces):
    """every primitive type in the bounce loop (spheres, cubes, a plane, triangles), sun first; radiance compared as floats through
    the accumulation target (USE_SUPERSAMPLING), the frame as rgba8"""
    w, h = 160, 90
    cam_kw = dict(eye=(6, 5, 12), flags=2)
    sky = synthetic_sky()
    want = oracle.path_frame(oracle.niels_scene(0.0, sky), oracle.camera(w, h, **cam_kw), oracle.seed((3.0, 9.0)), bounces, accum=(acc := np.zeros((h, w, 4), np.float32)))
    got = gpu_path(rtb, rtb.niels_scene(0.0), sky, w, h, cam_kw, bounces, accel, seed_off=(3.0, 9.0))
    assert got["stats"]["closest_rays"] + got["stats"]["shadow_rays"] == want["rays"]
    assert np.array_equal(got["dirT"].view(np.uint32), want["dirT"].view(np.uint32))
    # budget: 1-ulp differences of the binary64 transcendentals between CUDA and glibc (DESIGN.md numerics) can move a bounce direction
    bad = int((got["accum"][..., :3].view(np.uint32) != acc[..., :3].view(np.uint32)).any(-1).sum())
    assert bad <= 3, f"{bad} pixels whose path radiance differs from the oracle"
    d = np.abs(got["rgba8"].view(np.uint8).astype(int) - want["rgba8"].view(np.uint8).astype(int))
    assert int((d > 0).any(-1).sum() if d.ndim > 1 else (d > 0).sum()) <= 3 * 4
    if bounces == 4:
        assert got["stats"]["closest_rays_at_depth"][4] > 0


@pytest.mark.gpu
def test_path_frame_point_light_first_and_progressive(rtb, oracle):
    """lights[0] a point light (range rule of shadow.comp on every vertex) and three accumulated frames"""
    w, h, bounces = 128, 72, 3
    cam_kw = dict(eye=(6, 5, 12), flags=2)
    sky = synthetic_sky()
    osc = oracle.niels_scene(0.0, sky)
    osc.lights = np.ascontiguousarray(osc.lights.reshape(3, 32)[[2, 1, 0]]).reshape(-1)
    scene = rtb.niels_scene(0.0)
    scene["lights"] = np.ascontiguousarray(np.asarray(scene["lights"]).reshape(3, 32)[[2, 1, 0]]).reshape(-1)
    seed = oracle.seed((1.0, 2.0))
    acc = np.zeros((h, w, 4), np.float32)
    for _ in range(3):
        want = oracle.path_frame(osc, oracle.camera(w, h, **cam_kw), seed, bounces, accum=acc)
    got = gpu_path(rtb, scene, sky, w, h, cam_kw, bounces, rtb.ACCEL_BVH, seed_off=(1.0, 2.0), frames=3)
    bad = int((got["accum"][..., :3].view(np.uint32) != acc[..., :3].view(np.uint32)).any(-1).sum())
    assert bad <= 6, f"{bad} pixels whose accumulated radiance differs from the oracle"
    assert (np.abs(got["rgba8"].view(np.uint8).astype(int) - want["rgba8"].view(np.uint8).astype(int)) > 1).sum() == 0


@pytest.mark.gpu
def test_path_frame_soup_equals_the_oracle_statement(rtb, oracle):
    """incoherent bounces through a triangle soup: BVH traversal of queued rays against the oracle's linear loops"""
    from oracle.oracle import Scene
    n, w, h, bounces = 20000, 96, 54, 4
    tris = rtb.gen_soup(n, 0xB200)
    sun = rtb.niels_scene()["lights"][:32]
    mat = rtb.pack_material((0.8, 0.7, 0.6), (0.05, 0.05, 0.05), (0.0, 0.0, 0.02), 0.0, 1.0, 1.0)
    info = np.array([1, 1, n, 0, 0, 0, 1, 0, 0], np.uint32)
    scene = dict(triangles=tris, lights=sun, materials=mat, material_indices=np.zeros(n, np.uint32), info=info)
    cam_kw = dict(eye=(0.0, 0.0, 13.9), flags=2)
    acc = np.zeros((h, w, 4), np.float32)
    want = oracle.path_frame(Scene(tris, None, None, None, sun, mat, np.zeros(n, np.uint32), info, synthetic_sky()), oracle.camera(w, h, **cam_kw),
                             oracle.seed((0.0, 0.0)), bounces, accum=acc)
    got = gpu_path(rtb, scene, synthetic_sky(), w, h, cam_kw, bounces, rtb.ACCEL_BVH, limits=dict(max_triangles=n))
    bad = int((got["accum"][..., :3].view(np.uint32) != acc[..., :3].view(np.uint32)).any(-1).sum())
    assert bad <= 5, f"{bad} of {w * h} pixels differ (edge / tie rays and 1-ulp transcendental differences are the only allowed causes)"
    assert got["stats"]["closest_rays"] + got["stats"]["shadow_rays"] == pytest.approx(want["rays"], abs=10)


@pytest.mark.gpu
def test_path_frame_10m_mesh_bvh_equals_brute_on_a_tile_subsample(rtb):
    """configs[3] itself: the 10M-triangle height field at 1080p, 4 bounces.  Tile 0 of 64 (1/64 of the 32x32 blocks of the SAME
    frame: global pixel coordinates feed the RNG) rendered with the 8-wide BVH and with the reference's linear loop over all
    triangles; every depth's ray count and every pixel must agree."""
    grid, w, h, bounces = 2236, 1920, 1080, 4
    n = 2 * grid * grid
    scene = dict(triangles=rtb.gen_heightfield(grid, 0xB200), lights=rtb.niels_scene()["lights"][:32],
                 materials=rtb.pack_material((0.8, 0.8, 0.8), (0.05, 0.05, 0.05), (0, 0, 0), 0.0, 1.0, 1.0),
                 material_indices=np.zeros(n, np.uint32), info=np.array([1, 1, n, 0, 0, 0, 1, 0, 0], np.uint32))
    cam_kw = dict(eye=(0.0, 7.5, 0.0), pitch=1.5707964, left_fov=120.0, right_fov=120.0, flags=2)   # bench.py's heightfield10m_b4 camera
    a = gpu_path(rtb, scene, None, w, h, cam_kw, bounces, rtb.ACCEL_BVH, limits=dict(max_triangles=n), tile=(0, 64))
    b = gpu_path(rtb, scene, None, w, h, cam_kw, bounces, rtb.ACCEL_BRUTE, limits=dict(max_triangles=n), tile=(0, 64))
    assert a["stats"]["closest_rays_at_depth"][0] > 30000
    hit0 = (a["dirT"][..., 3].view(np.uint32) != 0xFFFFFFFF).sum() / max(a["stats"]["closest_rays_at_depth"][0], 1)
    assert hit0 > 0.8, f"the mesh should fill the frame, {hit0:.2f} of the tile's pixels hit"
    for d in range(bounces + 1):
        assert abs(a["stats"]["closest_rays_at_depth"][d] - b["stats"]["closest_rays_at_depth"][d]) <= 2, d
        assert abs(a["stats"]["shadow_rays_at_depth"][d] - b["stats"]["shadow_rays_at_depth"][d]) <= 2, d
    bad = int((a["accum"][..., :3].view(np.uint32) != b["accum"][..., :3].view(np.uint32)).any(-1).sum())
    assert bad <= 4, f"{bad} pixels differ between the BVH and the linear loop"
    assert np.array_equal(a["tiled"], b["tiled"]) or bad > 0


@pytest.mark.gpu
@pytest.mark.parametrize("tile", [None, (2, 3)])
def test_overlapped_path_frames_identical(rtb, tile):
    """RTB_OPT_FRAME_OVERLAP for path frames: init + depth 0 of frame k+1 on the front stream under the deeper launches of frame k, and
    the occlusion launch of depth d beside the nearest-hit launch of depth d+1.  Accumulating path frames mixed with RTB_PASS_FRAME,
    seed uploads and read-backs: the same bytes as with one launch after the other."""
    import torch
    w, h = 480, 270
    scene = dict(triangles=rtb.gen_heightfield(245, 0xB200), lights=rtb.niels_scene()["lights"][:32],
                 materials=rtb.pack_material((0.8, 0.8, 0.8), (0.05, 0.05, 0.05), (0, 0, 0), 0.0, 1.0, 1.0),
                 material_indices=np.zeros(2 * 245 * 245, np.uint32), info=np.array([1, 1, 2 * 245 * 245, 0, 0, 0, 1, 0, 0], np.uint32))

    def run(overlap):
        ctx = rtb.Context(max_triangles=2 * 245 * 245)
        ctx.set_option(rtb.OPT_FRAME_OVERLAP, overlap)
        if tile:
            ctx.set_option(rtb.OPT_TILE_COUNT, tile[1])
            ctx.set_option(rtb.OPT_TILE_RANK, tile[0])
        ctx.resize(w, h, 1)
        ctx.upload_scene(scene, synthetic_sky())
        ctx.build_accel(rtb.ACCEL_BVH)
        ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, eye=(0.0, 7.5, 0.0), pitch=1.5707964, left_fov=120.0, right_fov=120.0, flags=2))
        ctx.upload(rtb.BUF_SEED, rtb.make_seed((4.0, 1.0)))
        pins = [torch.empty(w * h, dtype=torch.int32).pin_memory() for _ in range(2)]
        out = []
        for k in range(10):
            if k == 4:
                ctx.upload(rtb.BUF_SEED, rtb.make_seed((9.0, 2.0))[:20])
            if k == 6:
                ctx.dispatch(rtb.PASS_FRAME)
                out.append(ctx.readback(rtb.TGT_SHADOW_BITS).copy())
            ctx.path_frame(4 if k != 7 else 1)
            ctx.readback_async_into(rtb.TGT_RGBA8, pins[k & 1].data_ptr(), w * h * 4)
            if k in (2, 8):
                out.append(ctx.readback(rtb.TGT_DIR_T).copy())
                out.append(np.array(ctx.path_stats().as_dict()["closest_rays_at_depth"]))
            if k % 3 == 2:
                ctx.readback_wait()
                out.append(pins[k & 1].numpy().copy())
        ctx.readback_wait()
        out += [ctx.readback(t).copy() for t in (rtb.TGT_DIR_T, rtb.TGT_UV_NORMAL, rtb.TGT_ACCUM, rtb.TGT_RGBA8, rtb.TGT_SEED)]
        ctx.close()
        return out

    a, b = run(0), run(1)
    assert len(a) == len(b)
    for i, (x, y) in enumerate(zip(a, b)):
        assert np.array_equal(np.asarray(x).view(np.uint8), np.asarray(y).view(np.uint8)), f"output {i} differs between overlapped and sequential path frames"
    assert a[1][1] > 1000, "the bounce rays must hit the field again"
