"""The headline configuration itself (BASELINE.json configs[2]: 1M-triangle soup, 3840x2160, eye z = 13.9) against the checkers,
at full size — not a smaller stand-in:

  * 4096 random pixels of the exact frame through the CPU oracle (itself bit-equal to the reference's shaders, tests/
    test_oracle_vs_ref.py): hit ids exact off the flagged rays, t bit-equal, G-buffer texel, shadow bit, rgba8 within 1/255;
  * the shadow words (and everything else) of 1/64 of the SAME frame's 32x32 blocks against the reference's linear loop
    (RTB_ACCEL_BRUTE) — the occlusion launch, the dominant kernel, at its real size;
  * a 4096x2048 skybox (the size of the reference's qwantani_4k.hdr) seen through the omnidirectional projection: the +-pi seam,
    both poles and the border texels of sampleEquirect / clamp-to-border.
"""
import numpy as np
import pytest

W, H, N_TRI, EYE = 3840, 2160, 1_000_000, (0.0, 0.0, 13.9)


@pytest.fixture(scope="module")
def soup(rtb):
    return dict(triangles=rtb.gen_soup(N_TRI, 0xB200), lights=rtb.niels_scene()["lights"][:32],
                materials=rtb.pack_material((0.8, 0.8, 0.8), (0.05, 0.05, 0.05), (0, 0, 0), 0.0, 1.0, 1.0),
                material_indices=np.zeros(N_TRI, np.uint32), info=np.array([1, 1, N_TRI, 0, 0, 0, 1, 0, 0], np.uint32))


@pytest.fixture(scope="module")
def headline_frame(rtb, soup):
    ctx = rtb.Context(max_triangles=N_TRI)
    ctx.resize(W, H, 1)
    ctx.upload_scene(soup, None)
    ctx.build_accel(rtb.ACCEL_BVH)
    ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(W, H, eye=EYE))
    ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
    ctx.dispatch(rtb.PASS_FRAME)
    out = dict(dirT=ctx.readback(rtb.TGT_DIR_T), bits=ctx.readback(rtb.TGT_SHADOW_BITS), rgba8=ctx.readback(rtb.TGT_RGBA8),
               lighting=ctx.readback(rtb.TGT_LIGHTING), uvN=ctx.readback(rtb.TGT_UV_NORMAL), packets=ctx.accel_info().primary_packets)
    ctx.close()
    return out


def shadow_bit(bits, x, y, w, h):
    tiles_x = (w + 15) // 16
    word = bits[(x >> 4) + (y >> 1) * tiles_x]
    return (word >> ((x & 15) | ((y & 1) << 4))) & 1


@pytest.mark.gpu
def test_headline_frame_against_the_oracle_on_4096_pixels(rtb, oracle, soup, headline_frame):
    from oracle.oracle import Scene, FLAG_EDGE, FLAG_TIE, FLAG_PARALLEL, FLAG_NAN
    rng = np.random.default_rng(2024)
    idx = rng.choice(W * H, size=4096, replace=False)
    xy = np.stack([idx % W, idx // W], axis=1).astype(np.uint32)
    osc = Scene(soup["triangles"], None, None, None, soup["lights"], soup["materials"], soup["material_indices"], soup["info"], None)
    cam = oracle.camera(W, H, eye=EYE)
    seed = oracle.init_pass(oracle.seed((0.0, 0.0)))
    want = oracle.frame_pixels_ex(osc, cam, seed, 1, xy)
    x, y = xy[:, 0].astype(np.int64), xy[:, 1].astype(np.int64)
    got_dirT = headline_frame["dirT"][y, x]
    got_ids = got_dirT[:, 3].view(np.uint32)
    flagged = (want["flags"] & (FLAG_EDGE | FLAG_TIE | FLAG_PARALLEL | FLAG_NAN)) != 0
    assert headline_frame["packets"] == 3, "the headline frame runs the frustum-packet kernel"
    assert (want["object"] != 0xFFFFFFFF).mean() > 0.3
    bad = (got_ids != want["object"]) & ~flagged
    assert not bad.any(), f"{int(bad.sum())} unflagged pixels with another hit id ({int(flagged.sum())} flagged)"
    same = got_ids == want["object"]
    # the G-buffer texel (ray direction x t, id) is the oracle's bit for bit
    assert np.array_equal(got_dirT[same].view(np.uint32), want["dirT"][same].view(np.uint32))
    hit = same & (want["object"] != 0xFFFFFFFF)
    t_got = np.sqrt((got_dirT[hit, :3].astype(np.float64) ** 2).sum(-1))
    assert np.all(np.abs(t_got - want["t"][hit]) <= 1e-5 * want["t"][hit])
    # the occlusion launch: sample 0's shadow bit of every one of those pixels
    got_shadow = shadow_bit(headline_frame["bits"], x, y, W, H)
    sb = (got_shadow != want["shadowed"]) & same
    assert int(sb.sum()) <= 1, f"{int(sb.sum())} shadow bits differ from the oracle"   # 1-ulp binary64 sin/cos budget (DESIGN.md numerics)
    assert 0.1 < want["shadowed"][hit].mean() < 0.9   # both outcomes are well represented (29 % of the shadow rays are occluded)
    d = np.abs(headline_frame["rgba8"][y, x].view(np.uint8).reshape(-1, 4).astype(int) - want["rgba8"].view(np.uint8).reshape(-1, 4).astype(int)).max(-1)
    assert int((d[same & ~sb] > 1).sum()) == 0


@pytest.mark.gpu
def test_headline_shadow_words_against_the_linear_loop_on_a_64th_of_the_frame(rtb, soup, headline_frame):
    """tile 0 of 64 of the SAME frame rendered with RTB_ACCEL_BRUTE (the reference's loop over all 1M triangles, verbatim)"""
    ctx = rtb.Context(max_triangles=N_TRI)
    ctx.set_option(rtb.OPT_TILE_COUNT, 64)
    ctx.set_option(rtb.OPT_TILE_RANK, 0)
    ctx.resize(W, H, 1)
    ctx.upload_scene(soup, None)
    ctx.build_accel(rtb.ACCEL_BRUTE)
    ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(W, H, eye=EYE))
    ctx.upload(rtb.BUF_SEED, rtb.make_seed((0.0, 0.0)))
    ctx.dispatch(rtb.PASS_FRAME)
    brute = dict(dirT=ctx.readback(rtb.TGT_DIR_T), bits=ctx.readback(rtb.TGT_SHADOW_BITS), rgba8=ctx.readback(rtb.TGT_RGBA8), lighting=ctx.readback(rtb.TGT_LIGHTING))
    ctx.close()
    bx, by = (W + 31) // 32, (H + 31) // 32
    g = np.arange(bx * by)
    own = (g % 64 == 0).reshape(by, bx)
    px_mask = np.kron(own, np.ones((32, 32), bool))[:H, :W]
    assert px_mask.sum() > 120_000
    assert np.array_equal(headline_frame["dirT"][px_mask].view(np.uint32), brute["dirT"][px_mask].view(np.uint32))
    # shadow words: 16x2-pixel strips, each inside one block
    tiles_x, tiles_y = (W + 15) // 16, (H + 1) // 2
    word_mask = px_mask[::2, ::16][:tiles_y, :tiles_x].reshape(-1)
    a, b = headline_frame["bits"][: tiles_x * tiles_y][word_mask], brute["bits"][: tiles_x * tiles_y][word_mask]
    assert b.any()
    assert np.array_equal(a, b), f"{int((a != b).sum())} of {a.size} shadow words differ between the BVH occlusion launch and the linear loop"
    assert np.array_equal(headline_frame["lighting"][px_mask], brute["lighting"][px_mask])
    assert np.array_equal(headline_frame["rgba8"][px_mask], brute["rgba8"][px_mask])


def big_sky(w=4096, h=2048, seed=3):
    """deterministic 4096x2048 rgba16f equirect: smooth gradients plus per-texel noise, bright rows at both poles and bright
    columns either side of the +-pi seam, so that a wrong border / seam / pole rule shows"""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    sky = np.zeros((h, w, 4), np.float16)
    base = 0.2 + 0.6 * (yy / (h - 1)) + 0.15 * np.sin(xx / w * 12.566)
    for c, k in enumerate((1.0, 0.8, 0.6)):
        sky[..., c] = (base * k + rng.random((h, w), dtype=np.float32) * 0.2).astype(np.float16)
    sky[:2], sky[-2:] = np.float16(3.0), np.float16(5.0)
    sky[:, :2, 0], sky[:, -2:, 2] = np.float16(7.0), np.float16(9.0)
    return sky.view(np.uint16)


@pytest.mark.gpu
@pytest.mark.parametrize("pose", ["omni", "screen_zenith_a", "screen_zenith_b"])
def test_reference_sized_skybox_seam_poles_and_border(rtb, oracle, pose):
    sky = big_sky()
    w, h = 512, 256
    cam_kw = {"omni": dict(eye=(6, 5, 12), projection=1, yaw=0.3),                      # every direction: seam, both poles
              "screen_zenith_a": dict(eye=(0.3, 2.5, 0.2), pitch=1.4, yaw=3.0),          # Default projection looking up past the pole
              "screen_zenith_b": dict(eye=(0.3, 2.5, 0.2), pitch=-1.4, yaw=0.0)}[pose]
    want = oracle.frame(oracle.niels_scene(0.0, sky), oracle.camera(w, h, **cam_kw), oracle.seed((5.0, 1.0)), 1)
    ctx = rtb.Context()
    ctx.resize(w, h, 1)
    ctx.upload_scene(rtb.niels_scene(0.0), sky)
    ctx.build_accel(rtb.ACCEL_BVH)
    ctx.upload(rtb.BUF_CAMERA, rtb.pack_camera(w, h, **cam_kw))
    ctx.upload(rtb.BUF_SEED, rtb.make_seed((5.0, 1.0)))
    ctx.dispatch(rtb.PASS_FRAME)
    got = ctx.readback(rtb.TGT_RGBA8)
    ids = ctx.readback(rtb.TGT_DIR_T)[..., 3].view(np.uint32)
    ctx.close()
    assert np.array_equal(ids, want["dirT"][..., 3].view(np.uint32))
    miss = ids == 0xFFFFFFFF
    assert miss.mean() > 0.2, "the pose must look at the sky"
    bad = int((got != want["rgba8"]).sum())
    assert bad <= 3, f"{bad} pixels differ from the oracle with the 4096x2048 sky"
    if pose == "omni":   # the frame really contains the seam columns and the pole rows of the texture
        rgb = want["rgba8"].view(np.uint8).reshape(h, w, 4)
        assert (rgb[miss][:, 0] == 255).any() or (rgb[miss][:, 2] == 255).any()
