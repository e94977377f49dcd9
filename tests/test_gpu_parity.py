"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): nearest-hit ids exact except rays the oracle flags (edge / tie / parallel /
NaN); t within 1e-5 relative; final RGB within 1/255.  The kernels evaluate the shader arithmetic in the
oracle's operation order with -fmad=false, so in practice everything below is compared bit-for-bit and
the tolerances are only a budget for the rare 1-ulp difference between CUDA's and glibc's binary64
transcendentals (each test states its budget).
"""
import numpy as np
import pytest

from conftest import synthetic_sky

pytestmark = pytest.mark.gpu

NO_RAY_HIT = 0xFFFFFFFF


def to_oracle_scene(scene, sky=None):
    from oracle.oracle import Scene
    return Scene(scene.get("triangles"), scene.get("spheres"), scene.get("cubes"), scene.get("planes"), scene.get("lights"),
                 scene.get("materials"), scene.get("material_indices"), scene.get("info"), sky)


def make_ctx(rtb, scene, sky, w, h, samples, accel, **limits):
    ctx = rtb.Context(**limits)
    ctx.resize(w, h, samples)
    ctx.upload_scene(scene, sky)
    ctx.build_accel(accel)
    return ctx


def frame_both(rtb, oracle, scene, sky, cam_kwargs, w, h, samples, accel, cpu_offset=(0.0, 0.0), frames=1, limits=None, packets=None):
    ctx = make_ctx(rtb, scene, sky, w, h, samples, accel, **(limits or {}))
    if packets is not None:
        ctx.set_option(rtb.OPT_PRIMARY_PACKETS, packets)
    cam = rtb.pack_camera(w, h, **cam_kwargs)
    ocam = oracle.camera(w, h, **cam_kwargs)
    assert np.array_equal(cam, ocam), "camera packing differs from the oracle"
    ctx.upload(rtb.BUF_CAMERA, cam)
    ctx.upload(rtb.BUF_SEED, rtb.make_seed(cpu_offset))
    oseed = oracle.seed(cpu_offset)
    osc = to_oracle_scene(scene, sky)
    accum = np.zeros((h, w, 4), np.float32)
    for _ in range(frames):
        ctx.dispatch(rtb.PASS_FRAME)
        ref = oracle.frame(osc, ocam, oseed, samples, accum=accum)
    got = dict(dirT=ctx.readback(rtb.TGT_DIR_T), uvN=ctx.readback(rtb.TGT_UV_NORMAL), bits=ctx.readback(rtb.TGT_SHADOW_BITS),
               lighting=ctx.readback(rtb.TGT_LIGHTING), rgba8=ctx.readback(rtb.TGT_RGBA8), accum=ctx.readback(rtb.TGT_ACCUM),
               seed=ctx.readback(rtb.TGT_SEED))
    ref["accum"] = accum
    ref["seed"] = oseed
    ctx.close()
    return got, ref


def check_frame(got, ref, w, h, budget=2e-5, rgb_budget=2e-5):
    n = w * h
    assert np.array_equal(got["seed"], ref["seed"]), "Seed after init.comp differs"
    gid, rid = got["dirT"][..., 3].view(np.uint32), ref["dirT"][..., 3].view(np.uint32)
    bad_id = int((gid != rid).sum())
    assert bad_id <= budget * n, f"{bad_id} of {n} hit ids differ"
    same = gid == rid
    gd, rd = got["dirT"][..., :3].view(np.uint32), ref["dirT"][..., :3].view(np.uint32)
    bad_bits = int((gd[same] != rd[same]).any(axis=-1).sum())
    assert bad_bits <= budget * n, f"{bad_bits} of {n} dirT vectors differ bitwise"
    # stated bar: |t| within 1e-5 relative
    gt, rt_ = np.linalg.norm(got["dirT"][..., :3][same].astype(np.float64), axis=-1), np.linalg.norm(ref["dirT"][..., :3][same].astype(np.float64), axis=-1)
    rel = np.abs(gt - rt_) / np.maximum(np.abs(rt_), 1e-30)
    assert int((rel > 1e-5).sum()) <= budget * n
    bad_uvn = int((got["uvN"].view(np.uint32)[same] != ref["uvN"].view(np.uint32)[same]).any(axis=-1).sum())
    assert bad_uvn <= budget * n, f"{bad_uvn} uvObjectNormal texels differ"
    bad_words = int((got["bits"] != ref["bits"]).sum())
    assert bad_words <= max(2, budget * n), f"{bad_words} shadow-mask words differ"
    bad_l = int((got["lighting"] != ref["lighting"]).any(axis=-1).sum())
    assert bad_l <= max(4, 4 * budget * n), f"{bad_l} lighting texels differ"
    g8, r8 = got["rgba8"].view(np.uint8).reshape(h, w, 4).astype(np.int32), ref["rgba8"].view(np.uint8).reshape(h, w, 4).astype(np.int32)
    diff = np.abs(g8 - r8).max(axis=-1)
    assert int((diff > 1).sum()) <= max(4, 4 * rgb_budget * n), f"{int((diff > 1).sum())} pixels differ by more than 1/255"
    assert int((diff > 0).sum()) <= max(8, 50 * rgb_budget * n), f"{int((diff > 0).sum())} pixels differ at all"


POSES = {
    "default": dict(eye=(4, 2, -2)),
    "all_objects": dict(eye=(6, 5, 12)),
    "inside_cube": dict(eye=(0.5, 0.5, 0.5)),
    "grazing_plane": dict(eye=(4, 1e-3, -2)),
    "rotated": dict(eye=(6, 5, 12), pitch=0.2, yaw=0.4, roll=0.1),
}


ACCEL = {"brute": 0, "bvh": 1, "bvh2": 2}


@pytest.mark.parametrize("accel", ["brute", "bvh", "bvh2"])
@pytest.mark.parametrize("pose", list(POSES))
def test_niels_frame(rtb, oracle, sky, pose, accel):
    """BASELINE config 1: NielsScene t=0, 640x360, 1 primary + 1 shadow sample; full-frame compare of every target."""
    w, h = 640, 360
    scene = rtb.niels_scene(0.0)
    osc = oracle.niels_scene(0.0)
    for k in ("triangles", "spheres", "cubes", "planes", "lights", "materials"):
        assert np.array_equal(scene[k], getattr(osc, k)), f"NielsScene {k} differ from the oracle's"
    got, ref = frame_both(rtb, oracle, scene, sky, POSES[pose], w, h, 1, ACCEL[accel])
    check_frame(got, ref, w, h)
    if pose == "all_objects":
        ids = set(np.unique(got["dirT"][..., 3].view(np.uint32)).tolist()) - {NO_RAY_HIT}
        assert ids == set(range(13)), "pose must exercise all 13 objects"


def test_niels_no_skybox_two_samples(rtb, oracle):
    w, h = 320, 180
    got, ref = frame_both(rtb, oracle, rtb.niels_scene(0.7), None, POSES["all_objects"], w, h, 2, rtb.ACCEL_BVH, cpu_offset=(12.5, -431.25))
    check_frame(got, ref, w, h)


def test_ragged_size(rtb, oracle, sky):
    """Sizes that are not multiples of the 32x32 block, the 16x2 shadow strip or the 8x4 warp patch."""
    w, h = 333, 127
    got, ref = frame_both(rtb, oracle, rtb.niels_scene(0.0), sky, POSES["all_objects"], w, h, 3, rtb.ACCEL_BVH)
    check_frame(got, ref, w, h)


def test_progressive_accumulation(rtb, oracle, sky):
    """USE_SUPERSAMPLING: replaying the frame accumulates in fp32 (composite.comp:249-257)."""
    w, h = 320, 180
    cam = dict(POSES["all_objects"], flags=2)
    got, ref = frame_both(rtb, oracle, rtb.niels_scene(0.0), sky, cam, w, h, 1, rtb.ACCEL_BVH, cpu_offset=(3.0, 9.0), frames=4)
    assert got["seed"][16:].view(np.uint32)[0] == 4
    check_frame(got, ref, w, h, budget=1e-4, rgb_budget=1e-4)
    bad = int((got["accum"].view(np.uint32) != ref["accum"].view(np.uint32)).any(axis=-1).sum())
    assert bad <= 1e-3 * w * h, f"{bad} accumulation texels differ"


@pytest.mark.parametrize("packets", [2, 3])   # 2: the auto rule keeps these projections per ray; 3: frustum packets forced (eyes differ per half /
@pytest.mark.parametrize("projection", [1, 2, 3, 4, 5])   # per pixel in the stereo modes: packets with several origins take the generic walk)
def test_projection_modes(rtb, oracle, sky, projection, packets):
    w, h = 256, 128
    cam = dict(eye=(6, 5, 12), projection=projection, yaw=0.3)
    got, ref = frame_both(rtb, oracle, rtb.niels_scene(0.0), sky, cam, w, h, 1, rtb.ACCEL_BVH, packets=packets)
    check_frame(got, ref, w, h, budget=1e-4, rgb_budget=1e-4)


def test_point_light_first(rtb, oracle, sky):
    """lights[0] a point light: exercises the point branch of getDirToLight and the ranged occlusion test."""
    w, h = 320, 180
    scene = rtb.niels_scene(0.0)
    lights = scene["lights"].reshape(3, 32)
    scene["lights"] = np.concatenate([lights[2], lights[1], lights[0]])   # the host would never order them so; the shaders do not care
    got, ref = frame_both(rtb, oracle, scene, sky, POSES["all_objects"], w, h, 2, rtb.ACCEL_BVH)
    check_frame(got, ref, w, h, budget=1e-4, rgb_budget=1e-4)
    assert got["bits"].any(), "some pixel must be shadowed from the point light"


def soup_scene(rtb, n, seed=0xB200):
    tris = rtb.gen_soup(n, seed)
    mat = rtb.pack_material((0.8, 0.8, 0.8), (0.05, 0.05, 0.05), (0, 0, 0), 0.0, 1.0, 1.0)
    sun = rtb.niels_scene()["lights"][:32]
    return dict(triangles=tris, lights=sun, materials=mat, material_indices=np.zeros(n, np.uint32),
                info=np.array([1, 1, n, 0, 0, 0, 1, 0, 0], np.uint32))


def camera_rays(oracle, w, h, n, eye, rng):
    """n primary rays of a w x h view (oracle camera maths), returned as (n, 6)."""
    cam = oracle.camera(w, h, eye=eye)
    from oracle.oracle import Scene
    empty = Scene(info=[0] * 9)
    seed = oracle.init_pass(oracle.seed((1.0, 2.0)))
    _, _, rays, _ = oracle.raygen(empty, cam, seed, want_rays=True)
    rays = rays.reshape(-1, 6)
    return rays[rng.choice(rays.shape[0], n, replace=False)]


@pytest.mark.parametrize("accel", ["brute", "bvh", "bvh2"])
def test_soup_rays_in(rtb, oracle, accel):
    """Random-soup triangles, explicit rays: ids exact on unflagged rays, t bit-exact where ids agree."""
    n_tri, n_rays = 300_000, 8192
    scene = soup_scene(rtb, n_tri)
    rng = np.random.default_rng(1)
    rays = camera_rays(oracle, 512, 288, n_rays, (0, 0, 13.9), rng)   # every ray enters the soup volume
    osc = to_oracle_scene(scene)
    oid, ot, ouv, _, flags = oracle.trace_rays(osc, rays, want_flags=True)
    ctx = rtb.Context(max_triangles=n_tri)
    ctx.upload_scene(scene)
    ctx.build_accel(ACCEL[accel])
    gid, gt, guv = ctx.trace_rays(rays)
    ctx.close()
    assert (oid != NO_RAY_HIT).mean() > 0.05, "the sample must hit something"
    unflagged = flags == 0
    assert unflagged.mean() > 0.97, f"flagged fraction {1 - unflagged.mean():.4f} is implausibly high"
    assert np.array_equal(gid[unflagged], oid[unflagged]), f"{int((gid[unflagged] != oid[unflagged]).sum())} unflagged ids differ"
    same = gid == oid
    assert same.mean() > 0.999
    assert np.array_equal(gt[same].view(np.uint32), ot[same].view(np.uint32)), "t must be bit-exact where ids agree"
    assert np.array_equal(guv[same].view(np.uint32), ouv[same].view(np.uint32))


def test_soup_occlusion_rays(rtb, oracle):
    n_tri, n_rays = 50_000, 8192
    scene = soup_scene(rtb, n_tri)
    rng = np.random.default_rng(2)
    o = rng.uniform(-9, 9, (n_rays, 3)).astype(np.float32)
    d = rng.normal(size=(n_rays, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    rays = np.concatenate([o, d], axis=1).astype(np.float32)
    max_dist = np.where(rng.random(n_rays) < 0.5, np.float32(3.4028235e38), rng.uniform(0.1, 6.0, n_rays)).astype(np.float32)
    prev = rng.integers(0, n_tri, n_rays).astype(np.uint32)
    want = oracle.occlusion_rays(to_oracle_scene(scene), rays, max_dist, prev)
    ctx = rtb.Context(max_triangles=n_tri)
    ctx.upload_scene(scene)
    for mode in (rtb.ACCEL_BRUTE, rtb.ACCEL_BVH, rtb.ACCEL_BVH2):
        ctx.build_accel(mode)
        got = ctx.occlusion_rays(rays, max_dist, prev)
        assert int((got != want).sum()) <= 2, f"mode {mode}: {int((got != want).sum())} occlusion results differ"
    ctx.close()
    assert 0.01 < want.mean() < 0.95


def test_niels_occlusion_rays_all_primitives(rtb, oracle):
    """Occlusion through spheres, cubes (incl. origin inside: negative tmin) and the plane."""
    scene = rtb.niels_scene(0.0)
    rng = np.random.default_rng(3)
    n = 20000
    o = rng.uniform(-6, 8, (n, 3)).astype(np.float32)
    o[: n // 10] = rng.uniform(0.05, 0.95, (n // 10, 3)).astype(np.float32)   # inside cube [0,1]^3
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    rays = np.concatenate([o, d], axis=1).astype(np.float32)
    max_dist = np.where(rng.random(n) < 0.5, np.float32(3.4028235e38), rng.uniform(0.0, 10.0, n)).astype(np.float32)
    prev = rng.integers(0, 14, n).astype(np.uint32)
    want = oracle.occlusion_rays(to_oracle_scene(scene), rays, max_dist, prev)
    ctx = rtb.Context()
    ctx.upload_scene(scene)
    for mode in (rtb.ACCEL_BRUTE, rtb.ACCEL_BVH, rtb.ACCEL_BVH2):
        ctx.build_accel(mode)
        got = ctx.occlusion_rays(rays, max_dist, prev)
        assert np.array_equal(got, want), f"mode {mode}: {int((got != want).sum())} differ"
        gid, gt, guv = ctx.trace_rays(rays, prev)
        oid, ot, ouv, _, fl = oracle.trace_rays(to_oracle_scene(scene), rays, prev, want_flags=True)
        ok = fl == 0
        assert np.array_equal(gid[ok], oid[ok])
        same = gid == oid
        assert np.array_equal(gt[same].view(np.uint32), ot[same].view(np.uint32))
    ctx.close()


def test_soup_full_size_bvh_equals_brute(rtb):
    """BASELINE config 3 geometry (1M-triangle soup): the BVH search returns what the reference's linear loop
    returns, on a reduced frame (the brute-force kernel is the at-scale stand-in for the oracle)."""
    n_tri, w, h = 1_000_000, 480, 270
    scene = soup_scene(rtb, n_tri)
    out = {}
    for mode in (rtb.ACCEL_BVH, rtb.ACCEL_BVH2, rtb.ACCEL_BRUTE):
        ctx = make_ctx(rtb, scene, None, w, h, 1, mode, max_triangles=n_tri)
        ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, eye=(0, 0, 13.9)))
        for packets in ((0, 1, 3) if mode == rtb.ACCEL_BVH else (2,)):   # camera rays one by one / union packets / frustum packets
            ctx.set_option(rtb.OPT_PRIMARY_PACKETS, packets)
            ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
            ctx.dispatch(rtb.PASS_FRAME)
            out[mode, packets] = (ctx.readback(rtb.TGT_DIR_T), ctx.readback(rtb.TGT_SHADOW_BITS), ctx.readback(rtb.TGT_RGBA8))
            if mode != rtb.ACCEL_BRUTE:
                info = ctx.accel_info()
                assert info.node_count > n_tri // 16 and info.max_depth <= 60
                assert mode != rtb.ACCEL_BVH or info.primary_packets == packets
        ctx.close()
    n = w * h
    for key in ((rtb.ACCEL_BVH, 0), (rtb.ACCEL_BVH, 1), (rtb.ACCEL_BVH, 3), (rtb.ACCEL_BVH2, 2)):
        compare_frames(out[key], out[rtb.ACCEL_BRUTE, 2], n)
    # the three 8-wide kernels run the same triangle arithmetic with the same tie rule: identical, not just close
    for pk in (1, 3):
        for a, b in zip(out[rtb.ACCEL_BVH, 0], out[rtb.ACCEL_BVH, pk]):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("packets", [1, 3])
@pytest.mark.parametrize("pose", ["all_objects", "inside_cube"])
def test_niels_frame_packets(rtb, oracle, sky, pose, packets):
    """NielsScene through the packet kernel (forced on: the auto rule would pick it anyway for three large triangles)."""
    w, h = 333, 127
    scene = rtb.niels_scene(0.0)
    ctx = make_ctx(rtb, scene, sky, w, h, 1, rtb.ACCEL_BVH)
    ctx.set_option(rtb.OPT_PRIMARY_PACKETS, packets)
    cam = rtb.pack_camera(w, h, **POSES[pose])
    ctx.upload(rtb.BUF_CAMERA, cam)
    ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
    ctx.dispatch(rtb.PASS_FRAME)
    assert ctx.accel_info().primary_packets == packets
    got = dict(dirT=ctx.readback(rtb.TGT_DIR_T), uvN=ctx.readback(rtb.TGT_UV_NORMAL), bits=ctx.readback(rtb.TGT_SHADOW_BITS),
               lighting=ctx.readback(rtb.TGT_LIGHTING), rgba8=ctx.readback(rtb.TGT_RGBA8), accum=ctx.readback(rtb.TGT_ACCUM),
               seed=ctx.readback(rtb.TGT_SEED))
    ctx.close()
    oseed = oracle.seed((0.0, 0.0))
    ref = oracle.frame(to_oracle_scene(scene, sky), oracle.camera(w, h, **POSES[pose]), oseed, 1)
    ref["seed"] = oseed
    check_frame(got, ref, w, h)


def test_soup_rays_in_packets(rtb, oracle):
    """Rays-in mode with packets forced on, on INCOHERENT rays (32 unrelated rays per packet): the union walk must
    still return exactly the per-ray result — coherence is a performance assumption, never a correctness one."""
    n_tri, n_rays = 20000, 4096
    scene = soup_scene(rtb, n_tri)
    rng = np.random.default_rng(5)
    o = rng.uniform(-12, 12, (n_rays, 3)).astype(np.float32)
    d = rng.normal(size=(n_rays, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d], axis=1).astype(np.float32)
    ctx = make_ctx(rtb, scene, None, 64, 64, 1, rtb.ACCEL_BVH, max_triangles=n_tri)
    res = {}
    for packets in (0, 1, 3):   # 3: the packets share no origin, every axis is left unconstrained — slow, still exact
        ctx.set_option(rtb.OPT_PRIMARY_PACKETS, packets)
        res[packets] = ctx.trace_rays(rays)
    ctx.close()
    for pk in (1, 3):
        for a, b in zip(res[0], res[pk]):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    obj, t, uv, nrm, flags = oracle.trace_rays(to_oracle_scene(scene), rays, want_flags=True)
    ok = flags == 0
    assert np.array_equal(res[1][0][ok], obj[ok])


def compare_frames(a, b, n):
    ids_a, ids_b = a[0][..., 3].view(np.uint32), b[0][..., 3].view(np.uint32)
    assert (ids_a != NO_RAY_HIT).mean() > 0.3
    assert int((ids_a != ids_b).sum()) <= 1e-4 * n, f"{int((ids_a != ids_b).sum())} ids differ between BVH and brute force"
    same = ids_a == ids_b
    assert np.array_equal(a[0].view(np.uint32)[same], b[0].view(np.uint32)[same])
    assert int((a[1] != b[1]).sum()) <= 1e-4 * n
    assert int((a[2] != b[2]).sum()) <= 2e-4 * n


def test_tile_partition_matches_single(rtb, sky):
    """Two contexts rendering interleaved screen blocks + rtb_untile == one context rendering everything."""
    import ctypes as C
    w, h = 333, 200
    scene = rtb.niels_scene(0.0)
    cam = rtb.pack_camera(w, h, eye=(6, 5, 12))

    def render(rank, count):
        ctx = rtb.Context()
        ctx.set_option(rtb.OPT_TILE_COUNT, count)
        ctx.set_option(rtb.OPT_TILE_RANK, rank)
        ctx.resize(w, h, 2)
        ctx.upload_scene(scene, sky)
        ctx.build_accel(rtb.ACCEL_BVH)
        ctx.upload(rtb.BUF_CAMERA, cam)
        ctx.upload(rtb.BUF_SEED, rtb.make_seed((5.0, 6.0)))
        ctx.dispatch(rtb.PASS_FRAME)
        return ctx

    single = render(0, 1)
    want = single.readback(rtb.TGT_RGBA8)
    single.close()
    import torch
    parts = [render(r, 3) for r in range(3)]
    slots = max(p.device_ptr(rtb.TGT_RGBA8_TILED)[1] // 4 for p in parts)
    gathered = torch.zeros(3 * slots, dtype=torch.int32, device="cuda")
    for r, p in enumerate(parts):
        ptr, nbytes = p.device_ptr(rtb.TGT_RGBA8_TILED)
        p.sync()
        tmp = np.zeros(nbytes // 4, np.uint32)
        p.readback_into(rtb.TGT_RGBA8_TILED, tmp.ctypes.data, nbytes)
        gathered[r * slots: r * slots + tmp.size] = torch.from_numpy(tmp.view(np.int32)).cuda()
    out = torch.zeros(h * w, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    parts[0].untile(gathered.data_ptr(), 3, slots, out.data_ptr())
    parts[0].sync()
    got = out.cpu().numpy().view(np.uint32).reshape(h, w)
    for p in parts:
        p.close()
    assert np.array_equal(got, want)
    # the host mirror of the partition (used by the gloo tests and bench.py) agrees with the CUDA side
    from igx_raytracing_b200 import tiles
    assert slots == tiles.slots_per_rank(w, h, 3)
    assert np.array_equal(tiles.untile(gathered.cpu().numpy().view(np.uint32).reshape(3, slots), w, h, 3), want)


def test_cpp_facade_renders_like_the_oracle(rtb, oracle, tmp_path):
    """The C++ host side (include/igx_rt.hpp: SceneGraph + RaytracingInterface + tasks) drives the same frame through the
    C ABI: NielsScene built with add(), one sphere moved with update<T>() (dirty-range upload, accel rebuild), readPixels."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "scene_graph_check")
    libdir = os.path.dirname(rtb.LIB_PATH)
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-ffp-contract=off", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tests", "cpp", "scene_graph_check.cpp"), "-L", libdir, "-lrtb200", f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    out, out3, exp = str(tmp_path / "frame.bin"), str(tmp_path / "frame3.bin"), str(tmp_path / "export0")
    r = subprocess.run([exe, "render", out, out3, exp, str(tmp_path / "tiny.png")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "OK render" in r.stdout and "OK refit frame" in r.stdout and "OK export" in r.stdout, r.stdout
    got = np.fromfile(out, np.uint32).reshape(360, 640)
    scene = rtb.niels_scene(0.0)
    sph = scene["spheres"].view(np.float32).reshape(7, 4).copy()
    sph[5] = [-4, 2.5, 0, 1]
    scene["spheres"] = sph.view(np.uint8).reshape(-1)
    ref = oracle.frame(to_oracle_scene(scene), oracle.camera(640, 360, eye=(6, 5, 12)), oracle.seed((0.0, 0.0)), 1)
    d = np.abs(got.view(np.uint8).astype(np.int32) - ref["rgba8"].view(np.uint8).reshape(-1).astype(np.int32).reshape(got.view(np.uint8).shape))
    assert int((d > 1).sum()) <= 4 and int((d > 0).sum()) <= 64, f"{int((d > 0).sum())} channel values differ"
    # third frame: triangle 1 moved through update<Triangle>() -> device refit; the oracle sees the moved triangle
    got3 = np.fromfile(out3, np.uint32).reshape(360, 640)
    tri = scene["triangles"].copy()
    tri[48:96] = oracle.triangle_flat([[-3, 6, 1], [2, 5, 0.5], [1, 2, 3]])
    scene["triangles"] = tri
    ref3 = oracle.frame(to_oracle_scene(scene), oracle.camera(640, 360, eye=(6, 5, 12)), oracle.seed((0.0, 0.0)), 1)
    d3 = np.abs(got3.view(np.uint8).astype(np.int32) - ref3["rgba8"].view(np.uint8).reshape(-1).astype(np.int32).reshape(got3.view(np.uint8).shape))
    assert int((d3 > 1).sum()) <= 4 and int((d3 > 0).sum()) <= 64, f"{int((d3 > 0).sum())} channel values differ after the refit"
    assert int((got3 != got).sum()) > 500, "the moved triangle must change the picture"
    # export: SD preset (720x480), 4 accumulated samples, written as <targetOutput>.png with the rows flipped
    from PIL import Image
    png = np.array(Image.open(exp + ".png"))
    assert png.shape == (480, 720, 4)
    ocam = oracle.camera(720, 480, eye=(6, 5, 12), flags=2)
    oseed = oracle.seed((0.0, 0.0))
    accum = np.zeros((480, 720, 4), np.float32)
    osc = to_oracle_scene(scene)
    for _ in range(4):
        ref4 = oracle.frame(osc, ocam, oseed, 1, accum=accum)
    want = ref4["rgba8"].view(np.uint8).reshape(480, 720, 4)[::-1]
    dp = np.abs(png.astype(np.int32) - want.astype(np.int32))
    assert int((dp > 1).sum()) <= 16 and int((dp > 0).sum()) <= 400, f"{int((dp > 0).sum())} channel values of the exported PNG differ"


def test_refit_unmoved_reproduces_the_build(rtb):
    """rtb_refit_accel on unmoved triangles rewrites every node and traversal triangle bit for bit: the device refit
    and the host builder share one encoder (csrc/rtb_node8_encode.h) and one padding rule."""
    n_tri = 50000
    scene = soup_scene(rtb, n_tri)
    ctx = make_ctx(rtb, scene, None, 64, 64, 1, rtb.ACCEL_BVH, max_triangles=n_tri)
    nodes0, tris0 = ctx.accel_bytes(rtb.TGT_ACCEL_NODES), ctx.accel_bytes(rtb.TGT_ACCEL_TRIANGLES)
    assert nodes0.size == ctx.accel_info().node_count * 128 and tris0.size == n_tri * 48
    cost0 = ctx.accel_info().sah_cost
    ctx.refit_accel()
    assert ctx.accel_info().refits == 1
    assert abs(ctx.accel_info().sah_cost - cost0) <= 1e-3 * cost0, "the refit recomputes the builder's SAH cost"
    assert np.array_equal(ctx.accel_bytes(rtb.TGT_ACCEL_NODES), nodes0)
    assert np.array_equal(ctx.accel_bytes(rtb.TGT_ACCEL_TRIANGLES), tris0)
    ctx.close()


@pytest.mark.parametrize("packets", [0, 1, 3])
def test_refit_after_motion_equals_brute(rtb, packets):
    """Triangles move (a dirty-range upload, as SceneGraph::update does), the tree is refitted on the device, and the
    nearest hits / shadow bits / pixels equal the reference's linear loop over the moved triangles."""
    n_tri, w, h = 100000, 320, 180
    scene = soup_scene(rtb, n_tri)
    ctx = make_ctx(rtb, scene, None, w, h, 1, rtb.ACCEL_BVH, max_triangles=n_tri)
    ctx.set_option(rtb.OPT_PRIMARY_PACKETS, packets)
    cost_built = ctx.accel_info().sah_cost
    tris = scene["triangles"].copy().view(np.float32).reshape(n_tri, 12)
    rng = np.random.default_rng(3)
    lo, hi = 20000, 60000                                    # the dirty range
    shift = rng.uniform(-0.3, 0.3, (hi - lo, 1, 3)).astype(np.float32)
    shift[:100] *= 40.0                                      # a few fly far outside the old bounds
    pts = tris[lo:hi].reshape(-1, 3, 4)
    pts[:, :, :3] += shift
    moved = tris.reshape(-1).view(np.uint8)
    ctx.upload(rtb.BUF_TRIANGLES, moved[lo * 48:hi * 48], offset=lo * 48)
    with pytest.raises(rtb.RtbError):                        # stale tree: the dispatch refuses
        ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, eye=(0, 0, 13.9)))
        ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
        ctx.dispatch(rtb.PASS_FRAME)
    ctx.refit_accel()
    ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
    ctx.dispatch(rtb.PASS_FRAME)
    got = (ctx.readback(rtb.TGT_DIR_T), ctx.readback(rtb.TGT_SHADOW_BITS), ctx.readback(rtb.TGT_RGBA8))
    assert ctx.accel_info().refits == 1
    cost_refit = ctx.accel_info().sah_cost   # relative to the area of the scene bounds, which the outliers enlarged
    assert np.isfinite(cost_refit) and cost_refit > 0 and abs(cost_refit - cost_built) > 0.05 * cost_built, "sah_cost follows the deformation"
    ctx.close()
    scene2 = dict(scene, triangles=moved.copy())
    ref_ctx = make_ctx(rtb, scene2, None, w, h, 1, rtb.ACCEL_BRUTE, max_triangles=n_tri)
    ref_ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, eye=(0, 0, 13.9)))
    ref_ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
    ref_ctx.dispatch(rtb.PASS_FRAME)
    ref = (ref_ctx.readback(rtb.TGT_DIR_T), ref_ctx.readback(rtb.TGT_SHADOW_BITS), ref_ctx.readback(rtb.TGT_RGBA8))
    ref_ctx.close()
    ids_a, ids_b = got[0][..., 3].view(np.uint32), ref[0][..., 3].view(np.uint32)
    assert (ids_a != NO_RAY_HIT).mean() > 0.03
    assert int((ids_a != ids_b).sum()) <= 1e-4 * w * h
    same = ids_a == ids_b
    assert np.array_equal(got[0].view(np.uint32)[same], ref[0].view(np.uint32)[same])
    assert int((got[1] != ref[1]).sum()) <= 1e-4 * w * h
    assert int((got[2] != ref[2]).sum()) <= 2e-4 * w * h


def mesh_scene(rtb, grid, seed=0xB200):
    n = 2 * grid * grid
    tris = rtb.gen_heightfield(grid, seed)
    mat = rtb.pack_material((0.7, 0.6, 0.5), (0.05, 0.05, 0.05), (0, 0, 0), 0.1, 0.6, 1.0)
    sun = rtb.niels_scene()["lights"][:32]
    return dict(triangles=tris, lights=sun, materials=mat, material_indices=np.zeros(n, np.uint32),
                info=np.array([1, 1, n, 0, 0, 0, 1, 0, 0], np.uint32)), n


@pytest.mark.parametrize("packets", [0, 1, 3])
def test_heightfield_frame_vs_oracle(rtb, oracle, sky, packets):
    """BASELINE config 4 geometry at oracle size: a connected mesh with smooth vertex normals (the 6-argument Triangle
    constructor, interpolated un-normalised as SH/trace.glsl:52-60 does).  Shared edges are where tie-breaking shows: a ray
    through an edge meets two triangles at the same t and the reference keeps the lower index; rays the oracle flags
    (edge / tie / parallel) are the only ones allowed to differ."""
    scene, n_tri = mesh_scene(rtb, 48)
    w, h = 320, 180
    cam_kwargs = dict(eye=(0.0, 4.0, 9.0), pitch=0.3)
    ctx = make_ctx(rtb, scene, sky, w, h, 1, rtb.ACCEL_BVH, max_triangles=n_tri)
    ctx.set_option(rtb.OPT_PRIMARY_PACKETS, packets)
    ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, **cam_kwargs))
    ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
    ctx.dispatch(rtb.PASS_FRAME)
    got = dict(dirT=ctx.readback(rtb.TGT_DIR_T), uvN=ctx.readback(rtb.TGT_UV_NORMAL), bits=ctx.readback(rtb.TGT_SHADOW_BITS),
               lighting=ctx.readback(rtb.TGT_LIGHTING), rgba8=ctx.readback(rtb.TGT_RGBA8), seed=ctx.readback(rtb.TGT_SEED))
    ctx.close()
    osc = to_oracle_scene(scene, sky)
    ocam, oseed = oracle.camera(w, h, **cam_kwargs), oracle.seed((0.0, 0.0))
    ref = oracle.frame(osc, ocam, oseed, 1)
    ref["seed"] = oseed
    ids_g, ids_r = got["dirT"][..., 3].view(np.uint32), ref["dirT"][..., 3].view(np.uint32)
    assert (ids_r != NO_RAY_HIT).mean() > 0.3, "the pose must look at the mesh"
    # flagged rays: recompute the primary rays' flags with the oracle's own ray generator
    _, _, rays, flags = oracle.raygen(osc, ocam, oracle.init_pass(oracle.seed((0.0, 0.0))), want_rays=True, want_flags=True)
    clean = flags == 0
    assert clean.mean() > 0.98
    assert np.array_equal(ids_g[clean], ids_r[clean]), f"{int((ids_g[clean] != ids_r[clean]).sum())} unflagged hit ids differ"
    same = ids_g == ids_r
    assert int((~same).sum()) <= 1e-3 * w * h
    assert np.array_equal(got["dirT"].view(np.uint32)[same], ref["dirT"].view(np.uint32)[same])
    bad_uvn = int((got["uvN"].view(np.uint32)[same] != ref["uvN"].view(np.uint32)[same]).any(axis=-1).sum())
    assert bad_uvn <= 2e-5 * w * h + 2, f"{bad_uvn} interpolated normals / uv differ"
    g8, r8 = got["rgba8"].view(np.uint8).reshape(h, w, 4).astype(np.int32), ref["rgba8"].view(np.uint8).reshape(h, w, 4).astype(np.int32)
    diff = np.abs(g8 - r8).max(axis=-1)
    assert int((diff[same] > 1).sum()) <= 8, f"{int((diff[same] > 1).sum())} pixels differ by more than 1/255"


def test_heightfield_full_size_bvh_equals_brute(rtb):
    """BASELINE config 4 geometry at full size (10M triangles, HBM-resident tree): the BVH search returns what the
    reference's linear loop returns on a reduced frame; the per-ray and the packet kernel agree exactly."""
    scene, n_tri = mesh_scene(rtb, 2236)
    w, h = 192, 108
    out = {}
    for mode in (rtb.ACCEL_BVH, rtb.ACCEL_BRUTE):
        ctx = make_ctx(rtb, scene, None, w, h, 1, mode, max_triangles=n_tri)
        ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, eye=(0.0, 4.0, 9.0), pitch=0.3))
        for packets in ((0, 1, 3) if mode == rtb.ACCEL_BVH else (2,)):
            ctx.set_option(rtb.OPT_PRIMARY_PACKETS, packets)
            ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
            ctx.dispatch(rtb.PASS_FRAME)
            out[mode, packets] = (ctx.readback(rtb.TGT_DIR_T), ctx.readback(rtb.TGT_SHADOW_BITS), ctx.readback(rtb.TGT_RGBA8))
        ctx.close()
    for pk in (1, 3):
        for a, b in zip(out[rtb.ACCEL_BVH, 0], out[rtb.ACCEL_BVH, pk]):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # a connected mesh: rays through shared edges / vertices are exact ties, resolved by index in both searches
    compare_frames(out[rtb.ACCEL_BVH, 1], out[rtb.ACCEL_BRUTE, 2], w * h)


def test_packets_identical_at_full_4k(rtb):
    """The headline frame itself (1M-triangle soup, 3840x2160): per-ray, union-packet and frustum-packet searches return
    the same 8.3 M hit records and the same pixels, bit for bit."""
    n_tri, w, h = 1_000_000, 3840, 2160
    scene = soup_scene(rtb, n_tri)
    ctx = make_ctx(rtb, scene, None, w, h, 1, rtb.ACCEL_BVH, max_triangles=n_tri)
    ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, eye=(0, 0, 13.9)))
    out = {}
    for packets in (0, 1, 3, 2):
        ctx.set_option(rtb.OPT_PRIMARY_PACKETS, packets)
        ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
        ctx.dispatch(rtb.PASS_FRAME)
        out[packets] = (ctx.readback(rtb.TGT_DIR_T), ctx.readback(rtb.TGT_UV_NORMAL), ctx.readback(rtb.TGT_RGBA8))
        if packets == 2:
            assert ctx.accel_info().primary_packets == 3, "the auto rule takes frustum packets on this frame"
    ctx.close()
    assert (out[0][0][..., 3].view(np.uint32) != NO_RAY_HIT).mean() > 0.4
    for pk in (1, 3, 2):
        for a, b in zip(out[0], out[pk]):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_async_readback_pipelines_frames(rtb, sky):
    """rtb_readback_async (the reference's presentToCpu + fence): frame N is copied out while frame N+1 renders; the pass
    that overwrites the target waits for the copy, so each buffer holds exactly its own frame."""
    w, h = 640, 360
    scene = rtb.niels_scene(0.0)
    ctx = make_ctx(rtb, scene, sky, w, h, 1, rtb.ACCEL_BVH)
    poses = [dict(eye=(6, 5, 12)), dict(eye=(4, 2, -2)), dict(eye=(6, 5, 12), yaw=0.4)]
    want = []
    for pose in poses:   # reference frames, synchronous read-back
        ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, **pose))
        ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
        ctx.dispatch(rtb.PASS_FRAME)
        want.append(ctx.readback(rtb.TGT_RGBA8).copy())
    assert not np.array_equal(want[0], want[1])
    bufs = [np.zeros((h, w), np.uint32) for _ in poses]
    for pose, buf in zip(poses, bufs):   # pipelined: no host wait between frames
        ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, **pose))
        ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
        ctx.dispatch(rtb.PASS_FRAME)
        ctx.readback_async_into(rtb.TGT_RGBA8, buf.ctypes.data, buf.nbytes)
    ctx.readback_wait()
    for got, ref in zip(bufs, want):
        assert np.array_equal(got, ref)
    # a synchronous read-back and a resize with a copy in flight
    ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
    ctx.dispatch(rtb.PASS_FRAME)
    ctx.readback_async_into(rtb.TGT_RGBA8, bufs[0].ctypes.data, bufs[0].nbytes)
    assert np.array_equal(ctx.readback(rtb.TGT_RGBA8), want[2])
    ctx.resize(320, 180, 1)
    assert np.array_equal(bufs[0], want[2])
    ctx.close()


def test_api_state_and_argument_errors(rtb):
    """Error behaviour of the entry points added for refit / packets / read-back: codes, not aborts."""
    scene = soup_scene(rtb, 2000)
    ctx = rtb.Context(max_triangles=2000)
    ctx.resize(64, 64, 1)
    ctx.upload_scene(scene, None)
    ctx.refit_accel()                                   # no tree yet: falls back to a build of the current mode (brute = nothing to do)
    assert ctx.accel_info().refits == 0
    ctx.build_accel(rtb.ACCEL_BVH2)
    ctx.refit_accel()                                   # binary tree: rebuilt, not refitted
    assert ctx.accel_info().refits == 0 and ctx.accel_info().mode == rtb.ACCEL_BVH2
    ctx.build_accel(rtb.ACCEL_BVH)
    ctx.refit_accel()
    assert ctx.accel_info().refits == 1
    with pytest.raises(rtb.RtbError):
        ctx.set_option(rtb.OPT_PRIMARY_PACKETS, 4)
    buf = np.zeros(64 * 64 + 1, np.uint32)
    with pytest.raises(rtb.RtbError):
        ctx.readback_async_into(rtb.TGT_RGBA8, buf.ctypes.data, buf.nbytes)   # more bytes than the target holds
    with pytest.raises(rtb.RtbError):
        ctx.probe_l2_read_gbs(1024)
    assert ctx.probe_l2_read_gbs(8 << 20) > 1000.0
    ctx.close()


def test_example_app_exports_a_png(rtb, tmp_path):
    """examples/niels_export.cpp end to end: animated spheres for a few frames, then the export path at 480x270, 4 samples."""
    import os
    import subprocess
    from PIL import Image
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "niels_export")
    libdir = os.path.dirname(rtb.LIB_PATH)
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "niels_export.cpp"),
                    "-L", libdir, "-lrtb200", f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    out = str(tmp_path / "frame")
    r = subprocess.run([exe, out, "480", "270", "4"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "wrote" in r.stdout, r.stdout
    im = np.array(Image.open(out + ".png"))
    assert im.shape == (270, 480, 4)
    assert len(np.unique(im.reshape(-1, 4), axis=0)) > 500, "a real picture: ground plane, spheres, cubes, sky"


@pytest.mark.gpu
@pytest.mark.parametrize("scene_kind", ["niels_sun", "niels_point_first", "soup"])
def test_shadow_order_modes_identical(rtb, scene_kind):
    """RTB_OPT_SHADOW_ORDER: slot order (one record per pixel and sample), queue of live rays, queue sorted in light space —
    the shadow words and the frame must be the same bits whatever order the occlusion rays are traced in."""
    if scene_kind == "soup":
        n = 200_000
        scene = dict(triangles=rtb.gen_soup(n, 0xB200), lights=rtb.niels_scene()["lights"][:32],
                     materials=rtb.pack_material((0.8, 0.8, 0.8), (0.05, 0.05, 0.05), (0, 0, 0), 0.0, 1.0, 1.0),
                     material_indices=np.zeros(n, np.uint32), info=np.array([1, 1, n, 0, 0, 0, 1, 0, 0], np.uint32))
        w, h, samples, eye, limits = 640, 360, 2, (0.0, 0.0, 13.9), dict(max_triangles=n)
    else:
        scene = rtb.niels_scene(0.0)
        if scene_kind == "niels_point_first":
            scene["lights"] = np.ascontiguousarray(np.asarray(scene["lights"]).reshape(3, 32)[[2, 1, 0]]).reshape(-1)
        w, h, samples, eye, limits = 333, 187, 3, (6, 5, 12), dict()
    outs = []
    for order in (0, 1, 2, 3):
        ctx = rtb.Context(**limits)
        ctx.resize(w, h, samples)
        ctx.upload_scene(scene, synthetic_sky())
        ctx.build_accel(rtb.ACCEL_BVH)
        ctx.set_option(rtb.OPT_SHADOW_ORDER, order)
        ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, eye=eye))
        ctx.upload(rtb.BUF_SEED, rtb.make_seed((2.0, 5.0)))
        for _ in range(2):
            ctx.dispatch(rtb.PASS_FRAME)
        outs.append((ctx.readback(rtb.TGT_SHADOW_BITS), ctx.readback(rtb.TGT_RGBA8), ctx.readback(rtb.TGT_LIGHTING)))
        ctx.close()
    assert outs[0][0].any(), "no occluded ray in the test frame"
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["soup", "heightfield", "niels", "two_triangles", "clustered"])
def test_device_builder_returns_the_brute_force_hits(rtb, kind):
    """RTB_OPT_ACCEL_BUILDER = 1: the 8-wide tree built on the device (Morton sort, radix tree, greedy collapse, boxes by the refit
    kernels) must return exactly the hits of the reference's linear loop, like the host-built tree does."""
    rng = np.random.default_rng(5)
    sun = rtb.niels_scene()["lights"][:32]
    mat = rtb.pack_material((0.8, 0.8, 0.8), (0.05, 0.05, 0.05), (0, 0, 0), 0.0, 1.0, 1.0)
    if kind == "niels":
        scene, limits, w, h, eye, pitch = rtb.niels_scene(0.4), dict(), 320, 180, (6, 5, 12), 0.0
    else:
        if kind == "soup":
            tris, eye, pitch = rtb.gen_soup(300_000, 0xB200), (0.0, 0.0, 13.9), 0.0
        elif kind == "heightfield":
            tris, eye, pitch = rtb.gen_heightfield(400, 0xB200), (0.0, 6.0, 13.0), 0.45
        elif kind == "two_triangles":
            tris, eye, pitch = rtb.gen_soup(2, 7), (0.0, 0.0, 13.9), 0.0
        else:   # many triangles on the same few Morton cells: a deep radix tree
            base = np.asarray(rtb.gen_soup(64, 3)).view(np.float32).reshape(-1, 12).copy()
            reps = np.repeat(base, 600, axis=0)
            reps[:, [0, 1, 2, 4, 5, 6, 8, 9, 10]] += rng.normal(0, 1e-4, (reps.shape[0], 1)).astype(np.float32)
            tris, eye, pitch = np.ascontiguousarray(reps).view(np.uint8).reshape(-1), (0.0, 0.0, 13.9), 0.0
        n = np.asarray(tris).view(np.uint8).size // 48
        scene = dict(triangles=tris, lights=sun, materials=mat, material_indices=np.zeros(n, np.uint32), info=np.array([1, 1, n, 0, 0, 0, 1, 0, 0], np.uint32))
        limits, w, h = dict(max_triangles=n), 480, 270
    outs, infos = [], []
    for mode, builder in ((rtb.ACCEL_BRUTE, 0), (rtb.ACCEL_BVH, 0), (rtb.ACCEL_BVH, 1)):
        ctx = rtb.Context(**limits)
        ctx.set_option(rtb.OPT_ACCEL_BUILDER, builder)
        ctx.resize(w, h, 1)
        ctx.upload_scene(scene, None)
        ctx.build_accel(mode)
        ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, eye=eye, pitch=pitch))
        ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
        ctx.dispatch(rtb.PASS_FRAME)
        outs.append((ctx.readback(rtb.TGT_DIR_T), ctx.readback(rtb.TGT_SHADOW_BITS), ctx.readback(rtb.TGT_RGBA8)))
        infos.append(ctx.accel_info())
        if builder == 1:   # a refit of the device-built tree reproduces it (same encoder, same topology)
            before = ctx.accel_bytes()
            ctx.refit_accel()
            assert np.array_equal(before, ctx.accel_bytes())
            ctx.build_accel(mode)   # a second build: the first one in a process also pays for loading the sort's kernels
            infos[-1] = ctx.accel_info()
        ctx.close()
    brute, host, dev = outs
    for a, b in zip(brute, dev):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "device-built tree differs from the linear loop"
    for a, b in zip(host, dev):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    if kind in ("soup", "heightfield"):
        assert infos[2].builder == 1
        assert infos[2].sah_cost < 1.6 * infos[1].sah_cost, (infos[1].sah_cost, infos[2].sah_cost)
        # host wall clock of a few milliseconds of launches: one run in ~25 sees a scheduling hiccup larger than the host build (tens of ms)
        assert infos[2].build_ms < max(infos[1].build_ms, 150.0), (infos[1].build_ms, infos[2].build_ms)


@pytest.mark.gpu
@pytest.mark.parametrize("release", [0, 1])
@pytest.mark.parametrize("scene_kind", ["niels", "soup"])
def test_frame_lanes_identical(rtb, scene_kind, release):
    """RTB_OPT_FRAME_LANES: a frame as one lane or as two half-frame lanes on two streams must be the same bits in every target
    (also tiled, also with the RELEASE shader build, whose shadow-word clear needs both lanes' G-buffer)."""
    if scene_kind == "soup":
        n = 150_000
        scene = dict(triangles=rtb.gen_soup(n, 0xB200), lights=rtb.niels_scene()["lights"][:32],
                     materials=rtb.pack_material((0.8, 0.8, 0.8), (0.05, 0.05, 0.05), (0, 0, 0), 0.0, 1.0, 1.0),
                     material_indices=np.zeros(n, np.uint32), info=np.array([1, 1, n, 0, 0, 0, 1, 0, 0], np.uint32))
        w, h, samples, eye, limits = 1000, 600, 2, (0.0, 0.0, 13.9), dict(max_triangles=n)
    else:
        scene, w, h, samples, eye, limits = rtb.niels_scene(0.2), 1111, 555, 3, (6, 5, 12), dict()
    outs = []
    for lanes, tile in ((1, None), (2, None), (2, (1, 3))):
        ctx = rtb.Context(**limits)
        if tile:
            ctx.set_option(rtb.OPT_TILE_COUNT, tile[1])
            ctx.set_option(rtb.OPT_TILE_RANK, tile[0])
        ctx.set_option(rtb.OPT_FRAME_LANES, lanes)
        ctx.set_option(rtb.OPT_SHADER_BUILD, release)
        ctx.resize(w, h, samples)
        ctx.upload_scene(scene, synthetic_sky())
        ctx.build_accel(rtb.ACCEL_BVH)
        ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, eye=eye, flags=2))
        ctx.upload(rtb.BUF_SEED, rtb.make_seed((2.0, 5.0)))
        for _ in range(3):
            ctx.dispatch(rtb.PASS_FRAME)
        o = {k: ctx.readback(t) for k, t in (("dirT", rtb.TGT_DIR_T), ("uvN", rtb.TGT_UV_NORMAL), ("bits", rtb.TGT_SHADOW_BITS), ("lighting", rtb.TGT_LIGHTING),
                                                 ("accum", rtb.TGT_ACCUM), ("rgba8", rtb.TGT_RGBA8))}
        if tile:
            o["tiled"] = ctx.readback(rtb.TGT_RGBA8_TILED)
        outs.append(o)
        ctx.close()
    one, two, tiled = outs
    for k in ("dirT", "uvN", "bits", "lighting", "accum", "rgba8"):
        assert np.array_equal(one[k].view(np.uint8), two[k].view(np.uint8)), k
    # the tiled rank's own pixels equal the full frame's, and its tiled buffer holds them in the rank's slot order
    bx, by = (w + 31) // 32, (h + 31) // 32
    own = (np.arange(bx * by) % 3 == 1).reshape(by, bx)
    px = np.kron(own, np.ones((32, 32), bool))[:h, :w]
    assert np.array_equal(one["rgba8"][px], tiled["rgba8"][px])
    from igx_raytracing_b200 import tiles
    x, y, valid = tiles.slot_pixels(w, h, 1, 3)
    assert np.array_equal(tiled["tiled"][: valid.size][valid], one["rgba8"][y[valid], x[valid]])


@pytest.mark.gpu
def test_frame_graph_replay_identical_and_rerecorded_on_change(rtb, oracle):
    """RTB_OPT_FRAME_GRAPH: from the third unchanged frame on RTB_PASS_FRAME replays two CUDA graphs.  Six accumulated frames
    equal the oracle's (and the directly launched ones); a camera change, a moved triangle (refit) and a new shadow-sample count
    re-record instead of replaying stale launches."""
    w, h = 320, 180
    sky = synthetic_sky()
    cam_kw = dict(eye=(6, 5, 12), flags=2)

    def run(graph):
        ctx = rtb.Context()
        ctx.set_option(rtb.OPT_FRAME_GRAPH, graph)
        ctx.resize(w, h, 2)
        scene = rtb.niels_scene(0.0)
        ctx.upload_scene(scene, sky)
        ctx.build_accel(rtb.ACCEL_BVH)
        ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, **cam_kw))
        ctx.upload(rtb.BUF_SEED, rtb.make_seed((3.0, 9.0)))
        outs = []
        for _ in range(6):
            ctx.dispatch(rtb.PASS_FRAME)
        outs.append((ctx.readback(rtb.TGT_ACCUM), ctx.readback(rtb.TGT_RGBA8), ctx.readback(rtb.TGT_SEED)))
        # a new camera: the recorded launches hold the old one by value
        ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, eye=(5, 4, 11), flags=0))
        for _ in range(4):
            ctx.dispatch(rtb.PASS_FRAME)
        outs.append((ctx.readback(rtb.TGT_DIR_T), ctx.readback(rtb.TGT_RGBA8)))
        # a moved triangle + refit, then a different shadow-sample count
        tris = np.asarray(scene["triangles"]).view(np.float32).reshape(-1, 12).copy()
        tris[0, [1, 5, 9]] += 0.75
        ctx.upload(rtb.BUF_TRIANGLES, tris.view(np.uint8).reshape(-1))
        ctx.refit_accel()
        for _ in range(3):
            ctx.dispatch(rtb.PASS_FRAME)
        outs.append((ctx.readback(rtb.TGT_DIR_T), ctx.readback(rtb.TGT_RGBA8)))
        ctx.resize(w, h, 3)
        for _ in range(3):
            ctx.dispatch(rtb.PASS_FRAME)
        outs.append((ctx.readback(rtb.TGT_SHADOW_BITS), ctx.readback(rtb.TGT_RGBA8)))
        ctx.close()
        return outs

    direct, replayed = run(0), run(1)
    for a, b in zip(direct, replayed):
        for x, y in zip(a, b):
            assert np.array_equal(np.asarray(x).view(np.uint8), np.asarray(y).view(np.uint8))
    # and the six accumulated frames are the oracle's
    seed = oracle.seed((3.0, 9.0))
    acc = np.zeros((h, w, 4), np.float32)
    osc = oracle.niels_scene(0.0, sky)
    for _ in range(6):
        want = oracle.frame(osc, oracle.camera(w, h, **cam_kw), seed, 2, accum=acc)
    assert np.array_equal(replayed[0][2], seed)
    bad = int((replayed[0][0].view(np.uint32) != acc.view(np.uint32)).any(-1).sum())
    assert bad <= 4
    assert int((replayed[0][1] != want["rgba8"]).sum()) <= 4


@pytest.mark.gpu
@pytest.mark.parametrize("tiles", [1, 4])
def test_overlapped_frames_identical(rtb, tiles):
    """RTB_OPT_FRAME_OVERLAP: the camera rays of frame k+1 run under the shadow + shade launches of frame k, on two sets of
    G-buffer / wavefront / Seed buffers.  A sequence that exercises every ordering rule — per-frame seed uploads, an asynchronous
    read-back of every frame, G-buffer read-backs in between, a sphere upload (no re-recording), per-pass dispatches and rays-in
    calls between overlapped frames, progressive accumulation — gives the same bytes as one frame after the other."""
    import torch
    w, h = 640, 360
    scene = soup_scene(rtb, 150_000)
    base = rtb.niels_scene(0.0)
    scene.update(spheres=base["spheres"], cubes=base["cubes"], planes=base["planes"],
                 material_indices=np.zeros(150_000 + 10, np.uint32), info=np.array([1, 1, 150_000, 7, 2, 1, 1, 0, 0], np.uint32))
    rays = np.random.default_rng(4).normal(size=(4096, 6)).astype(np.float32)

    def run(overlap):
        ctx = rtb.Context(max_triangles=150_000)
        ctx.set_option(rtb.OPT_FRAME_OVERLAP, overlap)
        ctx.set_option(rtb.OPT_TILE_COUNT, tiles)
        ctx.set_option(rtb.OPT_TILE_RANK, tiles - 1)
        ctx.resize(w, h, 2)
        ctx.upload_scene(scene, None)
        ctx.build_accel(rtb.ACCEL_BVH)
        ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, eye=(0, 0, 13.9), flags=2))
        pins = [torch.empty(w * h, dtype=torch.int32).pin_memory() for _ in range(2)]
        out = []
        seed = rtb.make_seed((1.0, 5.0))
        ctx.upload(rtb.BUF_SEED, seed)
        for k in range(14):
            if k in (5, 6, 7):   # the host rewrites the first 20 bytes of Seed, as CompositeTask::update does every frame
                now = ctx.readback(rtb.TGT_SEED).copy() if k == 5 else None
                s2 = rtb.make_seed((1.0 + k, 5.0))
                ctx.upload(rtb.BUF_SEED, s2[:20])
            if k == 9:       # moved spheres: an upload the recorded launches need not know about
                sph = base["spheres"].view(np.float32).reshape(-1, 4).copy()
                sph[:, 1] += 0.5
                ctx.upload(rtb.BUF_SPHERES, sph.view(np.uint8).reshape(-1))
            if k == 11:      # per-pass dispatches and a rays-in call between recorded frames
                ctx.dispatch(rtb.PASS_INIT)
                ctx.dispatch(rtb.PASS_RAYGEN)
                out.append(ctx.trace_rays(rays)[0].copy())
            ctx.dispatch(rtb.PASS_FRAME)
            ctx.readback_async_into(rtb.TGT_RGBA8, pins[k & 1].data_ptr(), w * h * 4)
            if k in (3, 8, 12):
                out.append(ctx.readback(rtb.TGT_DIR_T).copy())
                out.append(ctx.readback(rtb.TGT_SEED).copy())
            if k >= 1 and k % 2 == 0:
                ctx.readback_wait()
                out.append(pins[k & 1].numpy().copy())
        ctx.readback_wait()
        out += [ctx.readback(t).copy() for t in (rtb.TGT_DIR_T, rtb.TGT_UV_NORMAL, rtb.TGT_SHADOW_BITS, rtb.TGT_LIGHTING, rtb.TGT_ACCUM, rtb.TGT_RGBA8, rtb.TGT_SEED)]
        ctx.close()
        return out

    a, b = run(0), run(1)
    assert len(a) == len(b)
    for i, (x, y) in enumerate(zip(a, b)):
        assert np.array_equal(np.asarray(x).view(np.uint8), np.asarray(y).view(np.uint8)), f"output {i} differs between overlapped and sequential frames"


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["sun_first", "point_first", "all_lights", "tile_rank"])
def test_light_cache_does_not_change_a_bit(rtb, sky, case):
    """RTB_OPT_LIGHT_CACHE: lighting.comp's per-pixel random pair (and the direction to a directional light 0) read from a cache
    filled once per frame size instead of being evaluated every frame.  Cache on == cache off in every target, over several frames,
    after light 0 was rewritten (a new sun direction: the cached directions are stale) and after a resize."""
    scene = rtb.niels_scene(0.3)
    if case == "point_first":
        lights = np.asarray(scene["lights"]).reshape(3, 32)
        scene["lights"] = np.concatenate([lights[1], lights[2], lights[0]])   # the host would never order them so; the shaders do not care
    outs = {}
    for cache in (1, 0):
        ctx = rtb.Context()
        ctx.set_option(rtb.OPT_LIGHT_CACHE, cache)
        if case == "all_lights":
            ctx.set_option(rtb.OPT_LIGHTS, 1)
        if case == "tile_rank":
            ctx.set_option(rtb.OPT_TILE_COUNT, 3)
            ctx.set_option(rtb.OPT_TILE_RANK, 1)
        w, h = 333, 190
        ctx.resize(w, h, 3)
        ctx.upload_scene(scene, sky)
        ctx.build_accel(rtb.ACCEL_BVH)
        ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, eye=(6, 5, 12), flags=2))
        ctx.upload(rtb.BUF_SEED, rtb.make_seed((2.0, 7.0)))
        got = []
        for step in range(3):
            if step == 1:   # a different first light: same type, other direction / position
                l0 = np.asarray(scene["lights"]).reshape(-1, 32)[0].copy()
                if case == "point_first":
                    l0[:12].view(np.float32)[:] = [1.0, 2.5, 0.5]
                else:
                    l0 = rtb.pack_light_directional((0.3, -1.0, 0.4), (0.9, 0.8, 0.7))
                ctx.upload(rtb.BUF_LIGHTS, np.asarray(l0).view(np.uint8).reshape(-1)[:32])
            if step == 2:
                w, h = 200, 120
                ctx.resize(w, h, 2)
                ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, eye=(6, 5, 12), flags=2))
            for _ in range(4):
                ctx.dispatch(rtb.PASS_FRAME)
            got.append([ctx.readback(t).copy() for t in (rtb.TGT_LIGHTING, rtb.TGT_ACCUM, rtb.TGT_RGBA8)])
            ctx.dispatch(rtb.PASS_INIT); ctx.dispatch(rtb.PASS_RAYGEN); ctx.dispatch(rtb.PASS_SHADOW); ctx.dispatch(rtb.PASS_LIGHTING)
            got[-1].append(ctx.readback(rtb.TGT_LIGHTING).copy())
        ctx.close()
        outs[cache] = got
    for step in range(3):
        for a, b in zip(outs[1][step], outs[0][step]):
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f"{case}, step {step}: the cache changed a target"
    assert not np.array_equal(outs[1][0][0].view(np.uint8), outs[1][1][0].view(np.uint8)), "the new first light must change the lighting"


@pytest.mark.gpu
@pytest.mark.parametrize("release", [0, 1])
def test_per_pass_dispatches_equal_the_frame_pass(rtb, sky, release):
    """The reference records INIT, RAYGEN, SHADOW, LIGHTING, COMPOSITE as five dispatches; RTB_PASS_FRAME is the same five with
    lighting + composite fused, recorded and overlapped.  Five accumulating frames either way (the C++ facade replays its command
    list through RTB_PASS_FRAME when the five stand together): every target identical."""
    w, h = 300, 170
    scene = rtb.niels_scene(0.6)
    outs = []
    for per_pass in (True, False):
        ctx = rtb.Context()
        ctx.set_option(rtb.OPT_SHADER_BUILD, release)
        ctx.resize(w, h, 2)
        ctx.upload_scene(scene, sky)
        ctx.build_accel(rtb.ACCEL_BVH)
        ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, eye=(6, 5, 12), flags=2))
        ctx.upload(rtb.BUF_SEED, rtb.make_seed((8.0, 3.0)))
        for _ in range(5):
            if per_pass:
                for p in (rtb.PASS_INIT, rtb.PASS_RAYGEN, rtb.PASS_SHADOW, rtb.PASS_LIGHTING, rtb.PASS_COMPOSITE):
                    ctx.dispatch(p)
            else:
                ctx.upload(rtb.BUF_SHADOW_PROPS, np.array([2], np.uint32))   # what the facade flushes before every frame: no re-recording
                ctx.dispatch(rtb.PASS_FRAME)
        outs.append([ctx.readback(t).copy() for t in (rtb.TGT_DIR_T, rtb.TGT_UV_NORMAL, rtb.TGT_SHADOW_BITS, rtb.TGT_LIGHTING, rtb.TGT_ACCUM, rtb.TGT_RGBA8, rtb.TGT_SEED)])
        ctx.close()
    for i, (a, b) in enumerate(zip(*outs)):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f"target {i} differs between five dispatches and RTB_PASS_FRAME"
