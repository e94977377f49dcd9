"""CPU-side tests (no GPU): the oracle against the reference's only known-answer vectors, the product's host packing
against the oracle, the .hdr loader, the C-ABI surface, the BVH builder, the C++ facade's SceneGraph logic."""
import ctypes as C
import os
import re
import struct
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# igx/igxi-tool/igxi/ignis/core2/test/test.cpp:8-29 — the only known-answer vectors in the reference tree: f32 bits -> f16 bits
F16_KAT = [
    (0x00000000, 0x0000), (0x80000000, 0x8000), (0x3f800000, 0x3c00), (0xbf800000, 0xbc00), (0x3f000000, 0x3800),
    (0x3e800000, 0x3400), (0x3e000000, 0x3000), (0x3eaaaaab, 0x3555), (0x38002000, 0x0001), (0x47000000, 0x7800),
    (0x477fe000, 0x7bff), (0x10001999, 0x0000), (0x00002000, 0x0000), (0x48000000, 0x7c00), (0x477ff000, 0x7c00),
    (0x40490fdb, 0x4248), (0x402d70a4, 0x416b), (0x7f800000, 0x7c00), (0xff800000, 0xfc00), (0xff800001, 0xffff),
    (0x7f800001, 0x7fff), (0x40a9999a, 0x454c),
]


def f32(bits):
    return struct.unpack("<f", struct.pack("<I", bits))[0]


def test_oracle_f16_known_answers(oracle):
    """Pins the oracle's packing row on the reference's own vectors."""
    for bits, want in F16_KAT:
        v = np.array([bits], np.uint32).view(np.float32)[0]
        got = int(oracle.lib.orc_f16_trunc(C.c_float(v)))
        assert got == want, f"f32 {bits:#010x}: oracle {got:#06x}, reference {want:#06x}"


def test_product_f16_known_answers(rtb):
    """The facade's igx::f16 (through Material's albedo channel) on the same vectors."""
    for bits, want in F16_KAT:
        v = np.array([bits], np.uint32).view(np.float32)
        out = np.zeros(32, np.uint8)
        z = np.zeros(3, np.float32)
        a = np.array([v[0], 0, 0], np.float32)
        rtb.lib().rtb_pack_material(a.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p),
                                    C.c_float(0), C.c_float(0), C.c_float(0), out.ctypes.data_as(C.c_void_p))
        got = int(out.view(np.uint16)[0])
        assert got == want, f"f32 {bits:#010x}: facade {got:#06x}, reference {want:#06x}"


def test_f16_roundtrip_and_rtne(oracle):
    rng = np.random.default_rng(0)
    for h in rng.integers(0, 0x7C00, 2000):   # every finite non-negative half survives half -> float -> half
        f = oracle.f16_to_f32(int(h))
        assert oracle.f32_to_f16_rtne(f) == int(h)
        if h >= 0x0400:   # normal halves; below 2^-14 core2's conversion is not a proper subnormal encoding (see the 0x38002000 vector)
            assert oracle.f16_trunc(f) == int(h)
    x = rng.uniform(-70000, 70000, 5000).astype(np.float32)
    want = x.astype(np.float16).view(np.uint16)
    got = np.array([oracle.f32_to_f16_rtne(v) for v in x], np.uint16)
    assert np.array_equal(got, want)


def test_packing_matches_oracle(rtb, oracle):
    rng = np.random.default_rng(1)
    for _ in range(200):
        p = rng.uniform(-20, 20, 9).astype(np.float32)
        assert np.array_equal(rtb.pack_triangle(p), oracle.triangle_flat(p))
        n = rng.normal(size=9).astype(np.float32)
        assert np.array_equal(rtb.pack_triangle(p, n), oracle.triangle_normals(p, n))
        d, c = rng.normal(size=3).astype(np.float32), rng.uniform(0, 3, 3).astype(np.float32)
        ae = np.float32(rng.uniform(0.001, 0.03))
        assert np.array_equal(rtb.pack_light_directional(d, c, ae), oracle.light_directional(d, c, ae))
        r, o, s = (np.float32(v) for v in rng.uniform(0.05, 50, 3))
        assert np.array_equal(rtb.pack_light_point(p[:3], c, r, o, s), oracle.light_point(p[:3], c, r, o, s))
        a, b, e = (rng.uniform(0, 2, 3).astype(np.float32) for _ in range(3))
        m, ro, t = (np.float32(v) for v in rng.uniform(0, 1, 3))
        assert np.array_equal(rtb.pack_material(a, b, e, m, ro, t), oracle.material(a, b, e, m, ro, t))
    # axis-aligned normals hit normalize(0) in spheremapTransform (NaN halves): must agree too
    for nrm in ([0, 0, 1], [0, 0, -1], [1, 0, 0]):
        n = np.array(nrm * 3, np.float32)
        p = np.arange(9, dtype=np.float32)
        assert np.array_equal(rtb.pack_triangle(p, n), oracle.triangle_normals(p, n))


def test_camera_matches_oracle_and_survey_anchor(rtb, oracle):
    rng = np.random.default_rng(2)
    for proj in range(6):
        for _ in range(20):
            kw = dict(eye=tuple(rng.uniform(-10, 10, 3)), pitch=float(rng.uniform(0, 6.2)), yaw=float(rng.uniform(0, 6.2)),
                      roll=float(rng.uniform(0, 6.2)), left_fov=float(rng.uniform(1, 179)), right_fov=float(rng.uniform(1, 179)),
                      ipd=float(rng.uniform(50, 80)), projection=proj, flags=int(rng.integers(0, 4)), exposure=float(rng.uniform(0.1, 4)))
            w, h = int(rng.integers(1, 4000)), int(rng.integers(1, 4000))
            assert np.array_equal(rtb.pack_camera(w, h, **kw), oracle.camera(w, h, **kw))
    cam = rtb.pack_camera(640, 360).view(np.float32)
    assert np.allclose(cam[4:7], [2.2222223, 3, -2.7002075], rtol=0, atol=1e-6)
    assert np.allclose(cam[8:11], [5.7777777, 3, -2.7002075], rtol=0, atol=1e-6)
    assert np.allclose(cam[12:15], [2.2222223, 1, -2.7002075], rtol=0, atol=1e-6)


def test_niels_scene_matches_oracle(rtb, oracle):
    for t in (0.0, 0.7, 3.0):
        a, b = rtb.niels_scene(t), oracle.niels_scene(t)
        for k in ("triangles", "spheres", "cubes", "planes", "lights", "materials", "material_indices", "info"):
            assert np.array_equal(np.asarray(a[k]).view(np.uint8).reshape(-1), np.asarray(getattr(b, k)).view(np.uint8).reshape(-1)), k
    sun = rtb.niels_scene()["lights"][:32]   # packed sun
    assert sun.view(np.uint32)[4] == 0x64111045 and sun.view(np.uint32)[5] == 0x4822 and sun.view(np.uint16)[6] == 0x20c3
    assert list(sun.view(np.uint16)[12:16]) == [0x3b33, 0x3b33, 0x3b33, 0]


def write_hdr(path, w, h, rng, rle):
    px = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    px[0, 0] = [0, 0, 0, 0]
    px[0, 1] = [255, 255, 255, 200]   # overflows half: replaced by 65504
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n" + f"-Y {h} +X {w}\n".encode())
        for y in range(h):
            if not rle:
                f.write(px[y].tobytes())
                continue
            f.write(bytes([2, 2, w >> 8, w & 255]))
            for c in range(4):
                row, x = px[y, :, c], 0
                while x < w:
                    run = 1
                    while x + run < w and run < 127 and row[x + run] == row[x]:
                        run += 1
                    if run >= 3:
                        f.write(bytes([128 + run, row[x]]))
                        x += run
                    else:
                        n = min(w - x, 5)
                        f.write(bytes([n]) + row[x:x + n].tobytes())
                        x += n
    return px


@pytest.mark.parametrize("rle", [False, True])
def test_hdr_loader(rtb, oracle, tmp_path, rle):
    rng = np.random.default_rng(3)
    path = str(tmp_path / "sky.hdr")
    px = write_hdr(path, 40, 9, rng, rle)
    got, ref = rtb.load_hdr(path), oracle.load_hdr(path)
    assert got.shape == (9, 40, 4) and np.array_equal(got, ref)
    assert list(got[0, 0]) == [0, 0, 0, 0] and list(got[0, 1][:3]) == [0x7BFF] * 3
    r, g, b, e = (int(v) for v in px[3, 7])
    want = np.float32(r) * np.float32(2.0 ** (e - 136))
    assert got[3, 7, 0] == oracle.f16_trunc(want)
    with pytest.raises(IOError):
        rtb.load_hdr(str(tmp_path / "missing.hdr"))


@pytest.mark.skipif(not os.path.exists("/root/reference/res/textures/qwantani_4k.hdr"), reason="reference data not on this machine")
def test_hdr_loader_reference_skybox(rtb, oracle):
    path = "/root/reference/res/textures/qwantani_4k.hdr"
    got = rtb.load_hdr(path)
    assert got.shape == (2048, 4096, 4)
    assert np.array_equal(got, oracle.load_hdr(path))


def test_abi_exports_every_declared_symbol(rtb):
    """Every function include/rtb200.h declares is exported by the shared library (and bound by rtb.py)."""
    hdr = open(os.path.join(ROOT, "include", "rtb200.h")).read()
    declared = sorted(set(re.findall(r"\b(rtb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 25
    out = subprocess.run(["nm", "-D", "--defined-only", rtb.LIB_PATH], stdout=subprocess.PIPE, text=True, check=True).stdout
    exported = set(re.findall(r" T (rtb_[a-z0-9_]+)", out))
    missing = [d for d in declared if d not in exported]
    assert not missing, f"declared but not exported: {missing}"
    assert set(rtb.EXPORTS) == set(declared), f"rtb.py binding list out of date: {set(rtb.EXPORTS) ^ set(declared)}"
    for name in declared:
        getattr(rtb.lib(), name)


def test_sass_is_sm100a(rtb):
    out = subprocess.run(["cuobjdump", "-lelf", rtb.LIB_PATH], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    assert "sm_100a" in out, out


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_no_cpu_fallback(rtb):
    """Without a CUDA device the product refuses to run: there is no fallback path."""
    with pytest.raises(rtb.RtbError) as e:
        rtb.Context()
    assert "CUDA" in str(e.value) or "device" in str(e.value)


def test_null_context_is_an_argument_error(rtb):
    """Every context-taking entry point rejects a NULL context with RTB_ERR_ARG instead of touching it (no GPU needed)."""
    L = rtb.lib()
    null = C.c_void_p(None)
    buf = (C.c_uint8 * 16)()
    d = C.c_double()
    assert L.rtb_refit_accel(null) == 1
    assert L.rtb_build_accel(null, 1) == 1
    assert L.rtb_dispatch(null, 5) == 1
    assert L.rtb_readback(null, 5, buf, 16) == 1
    L.rtb_readback_async.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    L.rtb_readback_wait.argtypes = [C.c_void_p]
    L.rtb_probe_l2_read_gbs.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_double)]
    assert L.rtb_readback_async(null, 5, buf, 16) == 1
    assert L.rtb_readback_wait(null) == 1
    assert L.rtb_probe_l2_read_gbs(null, 1 << 20, C.byref(d)) == 1
    assert L.rtb_set_option(null, 4, 3) == 1
    assert L.rtb_sync(null) == 1
    L.rtb_destroy(null)   # a no-op


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "igx_raytracing_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oracle" not in src.lower().replace("test infrastructure", ""), f"{fn} mentions the oracle"
    for fn in ("rtb200.h", "igx_rt.hpp"):
        assert "oracle" not in open(os.path.join(ROOT, "include", fn)).read().lower()


def build_cpp(tmp_path, name, sources, extra=()):
    exe = str(tmp_path / name)
    cmd = ["g++", "-O2", "-std=c++17", "-pthread", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "igx_raytracing_b200", "csrc")] + [os.path.join(ROOT, s) for s in sources] + list(extra) + ["-o", exe]
    subprocess.run(cmd, check=True)
    return exe


@pytest.mark.parametrize("n,threads", [(1, 1), (2, 1), (5, 2), (4000, 1), (60000, 4)])
def test_bvh_builder_host(tmp_path, n, threads):
    """Builder + a CPU walk that mirrors the kernel's slab test, links and tie rule == linear search (incl. duplicates, axis-parallel rays)."""
    exe = build_cpp(tmp_path, "bvh_check", ["tests/cpp/bvh_check.cpp", "igx_raytracing_b200/csrc/rtb_bvh.cpp"])
    r = subprocess.run([exe, str(n), str(threads), "600"], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout


def test_scene_graph_facade_host_logic(rtb, tmp_path):
    """igx::SceneGraph of include/igx_rt.hpp: handles, capacity, dirty upload ranges, hole compaction, light ordering,
    material-index table — checked against the behaviour of igx/src/helpers/scene_graph.cpp (no GPU: uploads fail, logic runs)."""
    exe = build_cpp(tmp_path, "scene_graph_check", ["tests/cpp/scene_graph_check.cpp"],
                    ["-L", os.path.dirname(rtb.LIB_PATH), "-lrtb200", f"-Wl,-rpath,{os.path.dirname(rtb.LIB_PATH)}"])
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    assert "OK" in r.stdout, r.stdout


def test_crmath_matches_binary64_libm(tmp_path):
    """rtb_crmath.h (sin / cos / pow5 of binary32 arguments without the slow general path) returns the binary32 value the
    oracle's `(float)std::sin((double)x)` returns: ~1.6e8 arguments incl. the RNG's own, at most 1e-7 may differ."""
    exe = str(tmp_path / "crmath_check")
    subprocess.check_call(["g++", "-O2", "-march=native", "-fopenmp", "-ffp-contract=off", "-I", os.path.join(ROOT, "igx_raytracing_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cpp", "crmath_check.cpp"), "-o", exe])
    r = subprocess.run([exe, "37"], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout


@pytest.mark.parametrize("stored", [False, True])
def test_png_export_roundtrip(rtb, tmp_path, monkeypatch, stored):
    """rtb_write_png (the export step after the hot path, ref: igxi convert.cpp:747-781): 8-bit RGBA, rows flipped on write.
    Both encodings (libz looked up at run time / stored deflate blocks) decode to the same pixels."""
    from PIL import Image
    if stored:
        monkeypatch.setenv("RTB_PNG_STORED", "1")
    rng = np.random.default_rng(1)
    frame = rng.integers(0, 2**32, (37, 53), dtype=np.uint32)
    frame[:8] = 0xFF336699          # compressible rows too
    path = tmp_path / "f.png"
    rtb.write_png(path, frame)
    im = np.array(Image.open(path))
    assert im.shape == (37, 53, 4)
    assert np.array_equal(im, frame.view(np.uint8).reshape(37, 53, 4)[::-1])
    rtb.write_png(path, frame, flip_vertically=False)
    assert np.array_equal(np.array(Image.open(path)), frame.view(np.uint8).reshape(37, 53, 4))
    big = np.full((300, 400), 0xFF102030, np.uint32)   # > 65535 bytes: several stored blocks
    rtb.write_png(path, big)
    assert np.array_equal(np.array(Image.open(path)).view(np.uint32)[..., 0], big)
    size = os.path.getsize(path)
    assert (size > big.nbytes) == stored


def test_example_app_builds_and_fails_loudly_without_a_gpu(rtb, tmp_path):
    """examples/niels_export.cpp (the reference's demo, headless) compiles against the facade; without a device it says so."""
    exe = str(tmp_path / "niels_export")
    libdir = os.path.dirname(rtb.LIB_PATH)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "niels_export.cpp"),
                           "-L", libdir, "-lrtb200", f"-Wl,-rpath,{libdir}", "-o", exe])
    r = subprocess.run([exe, str(tmp_path / "frame")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    import torch
    if not torch.cuda.is_available():
        assert r.returncode == 2 and "no usable CUDA device" in r.stdout


def test_oracle_generators_equal_the_product_generators(oracle, rtb):
    """bench.py's CPU legs build their scenes from the oracle's own generators (nothing of the product is loaded there):
    the two statements of the synthetic scenes must be the same bytes."""
    assert np.array_equal(oracle.gen_soup(30000), np.asarray(rtb.gen_soup(30000, 0xB200)).view(np.uint8).reshape(-1))
    assert np.array_equal(oracle.gen_heightfield(113), np.asarray(rtb.gen_heightfield(113, 0xB200)).view(np.uint8).reshape(-1))
    a, b = oracle.niels_scene(0.7, None), rtb.niels_scene(0.7)
    for k in ("triangles", "spheres", "cubes", "planes", "lights", "materials"):
        assert np.array_equal(getattr(a, k), np.asarray(b[k]).view(np.uint8).reshape(-1)), k
    assert np.array_equal(oracle.material((0.8, 0.8, 0.8), (0.05, 0.05, 0.05), (0, 0, 0), 0.0, 1.0, 1.0),
                          np.asarray(rtb.pack_material((0.8, 0.8, 0.8), (0.05, 0.05, 0.05), (0, 0, 0), 0.0, 1.0, 1.0)).view(np.uint8).reshape(-1))


def test_reference_arm_runs_without_the_product_package():
    """`bench.py --impl reference` must not import igx_raytracing_b200 (nor map librtb200.so): run it with the package blocked."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, runpy; sys.modules['igx_raytracing_b200'] = None; "
            "sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'niels360', '--steps', '2', '--warmup', '1']; "
            "runpy.run_path(%r, run_name='__main__'); "
            "import os; maps = open('/proc/self/maps').read(); assert 'librtb200' not in maps, 'product library mapped'" % os.path.join(root, "bench.py"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0


def test_python_binding_constants_follow_the_header(rtb):
    """Every enumerator of include/rtb200.h that the ctypes binding names (options, passes, targets, buffers, accel modes, status
    codes) carries the header's value: the binding is written by hand."""
    import os
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "rtb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    enums = {m.group(1): int(m.group(2)) for m in re.finditer(r"\b(RTB_[A-Z0-9_]+)\s*=\s*(-?\d+)", hdr)}
    assert len(enums) > 40
    checked = 0
    for name, value in enums.items():
        short = name[4:]   # RTB_OPT_LIGHTS -> OPT_LIGHTS
        if hasattr(rtb, short):
            assert getattr(rtb, short) == value, f"{name}: header {value}, binding {getattr(rtb, short)}"
            checked += 1
    assert checked >= 30, checked
    for opt in ("OPT_FRAME_OVERLAP", "OPT_LIGHT_CACHE", "OPT_PRIMITIVE_TREES", "OPT_FRAME_GRAPH", "OPT_FRAME_LANES", "OPT_LIGHTS", "OPT_HISTORY_ALPHA"):
        assert hasattr(rtb, opt), opt
