"""GPU parity of the sphere / cube trees (RTB_OPT_PRIMITIVE_TREES): thousands of analytic primitives, searched through their own
8-wide trees, must return what the reference's linear loops return (ref: res/shaders/trace.glsl:31-40, :83-90 through the oracle) —
including the loops' tie rules: the FIRST sphere wins on equal distances, the LAST cube wins on equal tmin, a cube entered from
inside has a negative distance.  Ids and bits are compared exactly; t / uv bitwise where the ids agree."""
import numpy as np
import pytest

from test_gpu_parity import check_frame, frame_both, to_oracle_scene

pytestmark = pytest.mark.gpu

NO_RAY_HIT = 0xFFFFFFFF
FLT_MAX = np.float32(3.4028235e38)


def prim_scene(rtb, n_sph, n_cub, n_tri=0, seed=5, planes=True, spread=10.0):
    rng = np.random.default_rng(seed)
    base = rtb.niels_scene()
    sph = np.concatenate([rng.uniform(-spread, spread, (n_sph, 3)), rng.uniform(0.05, 0.45, (n_sph, 1))], axis=1).astype(np.float32)
    lo = rng.uniform(-spread, spread, (n_cub, 3)).astype(np.float32)
    cub = np.concatenate([lo, lo + rng.uniform(0.05, 0.8, (n_cub, 3)).astype(np.float32)], axis=1).astype(np.float32)
    tris = rtb.gen_soup(n_tri, 77) if n_tri else np.zeros(0, np.uint8)
    n_pl = 1 if planes else 0
    n_obj = n_tri + n_sph + n_cub + n_pl
    return dict(triangles=tris, spheres=sph.view(np.uint8).reshape(-1), cubes=cub.view(np.uint8).reshape(-1),
                planes=base["planes"] if planes else np.zeros(0, np.uint8), lights=base["lights"], materials=base["materials"],
                material_indices=(np.arange(n_obj) % 8).astype(np.uint32),
                info=np.array([3, 8, n_tri, n_sph, n_cub, n_pl, 1, 0, 2], np.uint32))


def random_rays(n, seed, spread=10.0, inside=None):
    rng = np.random.default_rng(seed)
    o = rng.uniform(-spread, spread, (n, 3)).astype(np.float32)
    if inside is not None:   # a tenth of the rays start inside some cube
        k = n // 10
        pick = inside[rng.integers(0, inside.shape[0], k)]
        o[:k] = (pick[:, :3] + (pick[:, 3:] - pick[:, :3]) * rng.uniform(0.1, 0.9, (k, 3))).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    return np.concatenate([o, d], axis=1).astype(np.float32)


def ctx_for(rtb, scene, trees, accel=None):
    info = scene["info"]
    ctx = rtb.Context(max_triangles=max(int(info[2]), 1024), max_spheres=32768, max_cubes=32768)
    ctx.set_option(rtb.OPT_PRIMITIVE_TREES, trees)
    ctx.upload_scene(scene)
    ctx.build_accel(rtb.ACCEL_BVH if accel is None else accel)
    return ctx


@pytest.mark.parametrize("n_tri", [0, 3000])
def test_many_spheres_and_cubes_rays_in(rtb, oracle, n_tri):
    scene = prim_scene(rtb, 6000, 5000, n_tri)
    cubes = scene["cubes"].view(np.float32).reshape(-1, 6)
    rays = random_rays(16384, 11, inside=cubes)
    rng = np.random.default_rng(12)
    n_obj = n_tri + 6000 + 5000 + 1
    prev = np.where(rng.random(rays.shape[0]) < 0.5, NO_RAY_HIT, rng.integers(0, n_obj, rays.shape[0])).astype(np.uint32)
    osc = to_oracle_scene(scene)
    oid, ot, ouv, _, flags = oracle.trace_rays(osc, rays, prev, want_flags=True)
    max_dist = np.where(rng.random(rays.shape[0]) < 0.5, FLT_MAX, rng.uniform(0.0, 8.0, rays.shape[0])).astype(np.float32)
    want_occ = oracle.occlusion_rays(osc, rays, max_dist, prev)
    kinds = np.digitize(oid, [n_tri, n_tri + 6000, n_tri + 11000, n_tri + 11001])
    assert (kinds == 1).sum() > 500 and (kinds == 2).sum() > 500, "the sample must hit spheres and cubes"
    res = {}
    for trees in (64, 0):   # trees / the loops
        ctx = ctx_for(rtb, scene, trees)
        gid, gt, guv = ctx.trace_rays(rays, prev)
        occ = ctx.occlusion_rays(rays, max_dist, prev)
        info = ctx.accel_info()
        ctx.close()
        assert (info.sphere_tree_nodes > 1) == (trees != 0) and (info.cube_tree_nodes > 1) == (trees != 0)
        ok = flags == 0
        assert np.array_equal(gid[ok], oid[ok]), f"trees={trees}: {int((gid[ok] != oid[ok]).sum())} unflagged ids differ"
        same = gid == oid
        assert same.mean() > 0.999
        assert np.array_equal(gt[same].view(np.uint32), ot[same].view(np.uint32)), f"trees={trees}: t differs"
        assert np.array_equal(guv[same].view(np.uint32), ouv[same].view(np.uint32)), f"trees={trees}: uv differs"
        assert int((occ != want_occ).sum()) <= (2 if n_tri else 0), f"trees={trees}: {int((occ != want_occ).sum())} occlusion results differ"
        res[trees] = (gid, gt, guv, occ)
    for a, b in zip(res[64], res[0]):
        assert np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a, b.view(np.uint32) if b.dtype == np.float32 else b), \
            "trees and loops must agree on every ray, flagged ones included"


def test_tie_rules_duplicates_and_nested(rtb, oracle):
    """Exact ties: every sphere and cube exists three times (the loop's first sphere / last cube wins), cubes nested in cubes
    with the ray origin inside all of them (most negative tmin... the loop keeps the SMALLEST tmin, last on ties), cubes sharing
    faces on an integer lattice."""
    rng = np.random.default_rng(21)
    base_s = np.concatenate([rng.integers(-8, 9, (400, 3)), rng.integers(1, 3, (400, 1)) * 0.25], axis=1).astype(np.float32)
    sph = np.concatenate([base_s, base_s, base_s])[rng.permutation(1200)]
    lo = rng.integers(-8, 8, (300, 3)).astype(np.float32)
    base_c = np.concatenate([lo, lo + 1], axis=1)
    nested = np.array([[-k, -k, -k, k, k, k] for k in (0.5, 1, 2, 3, 20)], np.float32)
    cub = np.concatenate([base_c, base_c, nested, base_c, nested])[rng.permutation(910)]
    scene = prim_scene(rtb, 0, 0, 0, planes=False)
    scene["spheres"] = np.ascontiguousarray(sph).view(np.uint8).reshape(-1)
    scene["cubes"] = np.ascontiguousarray(cub).view(np.uint8).reshape(-1)
    scene["info"] = np.array([3, 8, 0, 1200, 910, 0, 1, 0, 2], np.uint32)
    scene["material_indices"] = (np.arange(2110) % 8).astype(np.uint32)
    n = 12000
    rays = random_rays(n, 22, spread=9.0)
    rays[:2000, :3] = rng.uniform(-0.4, 0.4, (2000, 3)).astype(np.float32)            # inside every nested cube
    rays[2000:4000, :3] = rng.integers(-8, 9, (2000, 3)).astype(np.float32) + 0.5     # lattice cell centres
    axis = rng.integers(0, 3, 2000)
    rays[4000:6000, 3:] = 0.0                                                          # axis-parallel rays: infinite reciprocals
    rays[np.arange(4000, 6000), 3 + axis] = rng.choice([-1.0, 1.0], 2000).astype(np.float32)
    prev = np.full(n, NO_RAY_HIT, np.uint32)
    osc = to_oracle_scene(scene)
    oid, ot, ouv, _, flags = oracle.trace_rays(osc, rays, prev, want_flags=True)
    max_dist = np.where(rng.random(n) < 0.5, FLT_MAX, rng.uniform(0.0, 6.0, n)).astype(np.float32)
    want_occ = oracle.occlusion_rays(osc, rays, max_dist, prev)
    ctx = ctx_for(rtb, scene, 64)
    gid, gt, guv = ctx.trace_rays(rays, prev)
    occ = ctx.occlusion_rays(rays, max_dist, prev)
    info = ctx.accel_info()
    ctx.close()
    assert (oid != NO_RAY_HIT).mean() > 0.5
    assert np.array_equal(gid, oid), f"{int((gid != oid).sum())} ids differ (ties must resolve as the loops resolve them)"
    assert np.array_equal(gt.view(np.uint32), ot.view(np.uint32))
    assert np.array_equal(guv.view(np.uint32), ouv.view(np.uint32))
    assert np.array_equal(occ, want_occ)
    assert info.sphere_tree_nodes > 1 and info.cube_tree_nodes > 1, "the duplicates must not push the builder past its depth limit"


def test_frame_with_many_primitives(rtb, oracle, sky):
    """A whole frame (raygen, shadow, lighting, composite) over 1500 spheres + 1500 cubes + triangles + the plane."""
    scene = prim_scene(rtb, 1500, 1500, 400, spread=6.0)
    w, h = 96, 54
    got, ref = frame_both(rtb, oracle, scene, sky, dict(eye=(6, 5, 12)), w, h, 2, rtb.ACCEL_BVH,
                          limits=dict(max_spheres=4096, max_cubes=4096))
    check_frame(got, ref, w, h, budget=2e-4, rgb_budget=2e-4)
    ids = got["dirT"][..., 3].view(np.uint32)
    assert ((ids >= 400) & (ids < 1900)).sum() > 200 and ((ids >= 1900) & (ids < 3400)).sum() > 200


@pytest.mark.parametrize("order", [0, 1, 2])
@pytest.mark.parametrize("lanes", [1, 2])
def test_frame_trees_equal_loops(rtb, sky, order, lanes):
    """Larger frame, every shadow-ray order and both lane counts: trees on == trees off, bit for bit, and again after the
    spheres moved (the tree is rebuilt from the uploaded buffer)."""
    scene = prim_scene(rtb, 3000, 2500, 2000, spread=7.0)
    w, h = 320, 180
    outs = {}
    for trees in (64, 0):
        ctx = rtb.Context(max_spheres=4096, max_cubes=4096)
        ctx.set_option(rtb.OPT_PRIMITIVE_TREES, trees)
        ctx.set_option(rtb.OPT_SHADOW_ORDER, order)
        ctx.set_option(rtb.OPT_FRAME_LANES, lanes)
        ctx.resize(w, h, 2)
        ctx.upload_scene(scene, sky)
        ctx.build_accel(rtb.ACCEL_BVH)
        ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, eye=(6, 5, 12)))
        ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
        frames = []
        for step in range(3):
            if step == 2:
                moved = scene["spheres"].view(np.float32).reshape(-1, 4).copy()
                moved[:, 1] += 0.37
                ctx.upload(rtb.BUF_SPHERES, moved.view(np.uint8).reshape(-1))
            ctx.dispatch(rtb.PASS_FRAME)
            frames.append((ctx.readback(rtb.TGT_DIR_T).copy(), ctx.readback(rtb.TGT_SHADOW_BITS).copy(), ctx.readback(rtb.TGT_RGBA8).copy()))
        ctx.close()
        outs[trees] = frames
    for step in range(3):
        for a, b, name in zip(outs[64][step], outs[0][step], ("dirT", "shadow bits", "rgba8")):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"step {step}: {name} differs between trees and loops"
    assert not np.array_equal(outs[64][1][0].view(np.uint32), outs[64][2][0].view(np.uint32)), "the moved spheres must change the frame"
