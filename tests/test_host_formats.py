"""CPU-only tests of the host side: C-ABI surface, model constructors, parsers, text and image
formats.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re
import struct

import numpy as np
import pytest

import ppmpa_b200 as P
from ppmpa_b200 import _capi as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "examples")
REF_EX = "/root/reference/example"


def test_cabi_exports_every_declared_symbol():
    """Every function include/ppm.h declares is exported by the .so and bound in _capi.SIGNATURES."""
    hdr = open(os.path.join(ROOT, "include", "ppm.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ppm_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) > 45
    for name in declared:
        assert hasattr(K.lib, name), f"{name} is declared in ppm.h but not exported"
    assert declared == set(K.SIGNATURES), declared ^ set(K.SIGNATURES)
    assert K.lib.ppm_abi_version() == 2


def test_no_device_is_a_loud_error():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert K.lib.ppm_create(0, C.byref(h)) == -7          # PPM_ERR_NODEVICE: no CPU fallback
    with pytest.raises(P.PPMError):
        P.Engine(0)


def test_struct_layout_matches_header():
    assert C.sizeof(K.Prim) == 112 and C.sizeof(K.Photon) == 56
    assert C.sizeof(K.Material) == 9 * 8 + 8 + 6 * 8 + 5 * 8
    assert C.sizeof(K.Light) == 8 + 3 * 8 + 8 + 15 * 8
    assert K.PHOTON_DTYPE.itemsize == 56


def test_format_f64_is_rust_display():
    f = P.format_f64
    # Rust `{}`: shortest round-trip digits, never an exponent, no trailing ".0"
    assert f(1.0) == "1" and f(-0.0) == "-0" and f(0.1) == "0.1" and f(250.0) == "250"
    assert f(3.3000000000000003) == "3.3000000000000003"          # algebra.rs:275
    assert f(0.5773502691896258) == "0.5773502691896258"          # geometry.rs:218
    assert f(1e-7) == "0.0000001" and f(1e21) == "1000000000000000000000"
    assert f(5e-05) == "0.00005" and f(123456.789) == "123456.789"
    assert f(float("inf")) == "inf" and f(float("nan")) == "NaN"
    # Rust `{:e}`
    assert f(0.0, True) == "0e0" and f(1.0, True) == "1e0" and f(0.0015, True) == "1.5e-3"
    assert f(123.456, True) == "1.23456e2" and f(-2.5e-10, True) == "-2.5e-10"
    rng = np.random.default_rng(0)
    for v in np.concatenate([rng.normal(size=200), rng.normal(size=200) * 1e-12, rng.normal(size=200) * 1e15]):
        assert float(f(v)) == v and float(f(v, True)) == v        # round trip


def test_builtin_scene_is_the_reference_literal():
    sc = P.read_scene()
    assert (sc.nprims, sc.nmats, sc.nlights) == (17, 14, 1)
    types = [p.type for p in sc.prims]
    assert types == [K.SHAPE_PLAIN] * 6 + [K.SHAPE_SPHERE] * 10 + [K.SHAPE_PARALLELOGRAM]
    assert [sc.prims[i].scalar for i in range(6)] == [0.0, 4.0, 2.0, 2.0, 6.0, 5.0]        # scene.rs:360-383
    assert list(sc.prims[6].position) == [-1.6, 1.5, 3.0] and sc.prims[6].scalar == 0.4     # ball_1
    surf = [sc.mats[sc.prims[i].material].surface for i in range(6, 16)]
    assert surf == [K.SURF_TS, K.SURF_SIMPLE, K.SURF_SIMPLE, K.SURF_SIMPLE, K.SURF_TS, K.SURF_TS,
                    K.SURF_SIMPLE, K.SURF_SIMPLE, K.SURF_SIMPLE, K.SURF_TS]                 # ball1..ball10
    l = sc.lights[0]
    assert l.type == K.LIGHT_PARALLELOGRAM and l.flux == 5.0 and list(l.color) == [1 / 3, 1 / 3, 1 / 3]
    quad = sc.prims[16]
    assert list(quad.dir1) == [0.67 - (-0.67), 0.0, 0.0] and list(quad.dir2) == [0.0, 0.0, 3.67 - 2.33]
    assert list(sc.mats[quad.material].emittance) == [0.15, 0.15, 0.15]
    power, ns = sc.photon_budget(100000)
    assert power == 5.0 / 100000 and ns == [100000]


def test_constructors_match_oracle(oracle):
    L = oracle.L
    for rough in (0.0, 0.2, 0.5, 0.9, 1.0):
        m = K.Material()
        z = K.D3(0, 0, 0)
        K.lib.ppm_material_simple(C.byref(m), z, z, z, z, z, 0.5, 0.0, rough)
        assert m.density_pow == L.orc_density_pow(rough)          # surface.rs:52
        K.lib.ppm_material_ts(C.byref(m), z, z, z, z, z, 0.5, 0.0, rough)
        assert m.density_pow == L.orc_density_pow(rough) and m.alpha == rough * rough * rough * rough
    a, b = K.Prim(), K.Prim()
    p0, p1, p2 = K.D3(-1.2, 1.8, 2.6), K.D3(-0.4, 1.8, 4.2), K.D3(0.4, 1.8, 1.8)
    assert K.lib.ppm_prim_polygon(C.byref(a), p0, p1, p2, 1, 0) == 0 and L.orc_new_polygon(p0, p1, p2, 1, C.byref(b)) == 1
    assert bytes(a) == bytes(b)
    assert K.lib.ppm_prim_polygon(C.byref(a), p0, p0, p2, 1, 0) == -1                       # degenerate: error, no panic
    out, ref = K.D3(), K.D3()
    K.lib.ppm_color_normalize(K.D3(0.71, 0.49, 0.36), out); L.orc_color_normalize(K.D3(0.71, 0.49, 0.36), ref)
    assert list(out) == list(ref)


@pytest.mark.parametrize("path,kw", [(None, {}), ("camera0.scr", {}), ("screen1.scr", dict(xreso=1920, yreso=1080)),
                                     ("camera0.scr", dict(xreso=1024, yreso=1024, blur=0))])
def test_camera_finalize_matches_oracle(oracle, path, kw):
    cam = P.read_camera(None if path is None else os.path.join(EX, path), **kw)
    ref = K.Camera.from_buffer_copy(cam)
    for f in ("origin", "esx", "esy", "eex", "eey", "eye_dir"):
        setattr(ref, f, K.D3(9, 9, 9))
    assert oracle.L.orc_camera_finalize(C.byref(ref)) == 1
    assert bytes(cam) == bytes(ref)


def test_camera_defaults_and_dialects(tmp_path):
    d = P.read_camera()
    assert (d.xreso, d.yreso, d.progressive, d.antialias, d.blur, d.pfilter) == (256, 256, 1, 1, 1, K.FILTER_NONE)
    assert d.radius == 0.2 * 0.2 and d.focal_len == 50.0 / 1000.0 and list(d.eye_pos) == [1.0, 2.0, -4.5]
    c0 = P.read_camera(os.path.join(EX, "camera0.scr"))
    assert (c0.xreso, c0.progressive, c0.use_classic, c0.pfilter) == (512, 0, 0, K.FILTER_GAUSS) and c0.radius == 0.1 * 0.1
    # EBNF dialect (doc/ebnf-camera.txt) == legacy dialect
    p = tmp_path / "new.scr"
    p.write_text("x_resolution : 512\ny_resolution: 512 # c\nprogressive: no\nantialias: yes\nuse_classic: no\n"
                 "estimate_radius: 0.1\nambient: [ 0.001, 0.001, 0.001 ]\nmax_radiance: 0.01\n"
                 "eye_position: [ 0.0, 2.0, -4.5 ]\ntarget_position: [ 0.0, 2.0, 0.0 ]\nupper_direction: [ 0.0, 1.0, 0.0 ]\n"
                 "focus: 7.0\nphoton_filter: gauss\nsamplephoton: 500\n")
    assert bytes(P.read_camera(str(p))) == bytes(c0)
    bad = tmp_path / "bad.scr"
    bad.write_text("xresolution: many\n")
    with pytest.raises(P.PPMError):
        P.read_camera(str(bad))
    with pytest.raises(P.PPMError):
        P.read_camera(str(tmp_path / "missing.scr"))


def test_scene_parser_reproduces_hardcoded_room():
    """example/ex-11.9.scene describes the room of the hard-coded scene (SURVEY 8c fixtures row)."""
    a, b = P.read_scene(os.path.join(EX, "ex-11.9.scene")), P.read_scene()
    assert bytes(a.lights)[:C.sizeof(K.Light)].replace(b"\x80", b"\x00") == bytes(b.lights)[:C.sizeof(K.Light)].replace(b"\x80", b"\x00")
    for i in range(6):                                            # the six walls: same normals (up to -0.0) and dist
        assert [abs(x) for x in a.prims[i].nvec] == [abs(x) for x in b.prims[i].nvec]
        assert a.prims[i].scalar == b.prims[i].scalar and a.prims[i].type == K.SHAPE_PLAIN
    ma, mb = a.mats[a.prims[3].material], b.mats[b.prims[3].material]          # mwallr
    assert list(ma.color_a) == list(mb.color_a) == [0.4, 0.1, 0.1]
    ma, mb = a.mats[a.prims[2].material], b.mats[b.prims[2].material]          # mwallb
    assert list(ma.color_a) == list(mb.color_a) == [0.1, 0.1, 0.4]


def test_scene_parser_details_and_errors(tmp_path):
    g = P.read_scene(os.path.join(EX, "ex-glassbox.scene"))
    assert [p.type for p in g.prims] == [1] * 6 + [4] * 7 and g.nlights == 1
    glass = g.mats[g.prims[7].material]
    assert list(glass.ior) == [1.5, 1.5, 1.5] and glass.p0 == 0.0 and list(glass.color_b) == [0.08, 0.08, 0.08]
    s = P.read_scene(os.path.join(EX, "ex-sunwindow.scene"))
    sun = s.lights[0]
    assert sun.type == K.LIGHT_SUN and sun.flux == 20.0
    assert abs(sum(sun.color) - 1.0) < 1e-15 and abs(np.linalg.norm(list(sun.dir)) - 1.0) < 1e-15
    assert list(sun.nvec) == [0.0, -1.0, 0.0]                     # normalize(dir1 x dir2) = -EY like scene.rs:27
    t = P.read_scene(os.path.join(EX, "sample1.scene"))
    assert [p.type for p in t.prims].count(K.SHAPE_POLYGON) == 4
    assert list(t.mats[t.prims[6].material].ior) == [1.3, 1.5, 1.7] or any(list(m.ior) == [1.3, 1.5, 1.7] for m in t.mats)
    for text, frag in [("light:\n  - type: point\n    color: [1,1,1]\n    flux: 1\n    position: [0,1,0]\n", "no object"),
                       ("object:\n  - type: sphere\n    material: nope\n    center: [0,0,0]\n    radius: 1\n", "unknown material"),
                       ("stuff\n", "section")]:
        p = tmp_path / "bad.scene"
        p.write_text(text)
        with pytest.raises(P.PPMError) as e:
            P.read_scene(str(p))
        assert frag in str(e.value)


@pytest.mark.skipif(not os.path.isdir(REF_EX), reason="reference examples not mounted")
def test_fixtures_parse_identically_to_reference_files():
    for n in ["sample1", "ex-glassbox", "ex-sunwindow", "mirror-ball", "coral-ball", "ex-11.9"]:
        a, b = P.read_scene(os.path.join(REF_EX, n + ".scene")), P.read_scene(os.path.join(EX, n + ".scene"))
        assert a.nprims == b.nprims and a.nlights == b.nlights
        for pa, pb in zip(a.prims, b.prims):
            assert bytes(a.mats[pa.material]) == bytes(b.mats[pb.material])
            qa, qb = K.Prim.from_buffer_copy(pa), K.Prim.from_buffer_copy(pb)
            qa.material = qb.material = 0
            assert bytes(qa) == bytes(qb)
        assert bytes(a.lights)[:C.sizeof(K.Light) * a.nlights] == bytes(b.lights)[:C.sizeof(K.Light) * a.nlights]
    for n in ["camera0", "screen1"]:
        assert bytes(P.read_camera(os.path.join(REF_EX, n + ".scr"))) == bytes(P.read_camera(os.path.join(EX, n + ".scr")))
    # every parseable reference scene loads (ex-scene1 is an unfinished draft, materials.scene has no objects)
    import glob
    bad = []
    for f in sorted(glob.glob(os.path.join(REF_EX, "*.scene"))):
        try:
            P.read_scene(f)
        except P.PPMError:
            bad.append(os.path.basename(f))
    assert bad == ["ex-scene1.scene", "materials.scene"]


def test_radius_schedule_and_budget(oracle):
    r = P.radius_schedule(0.1, 6)
    ref = np.zeros(6)
    oracle.L.orc_radius_schedule(0.1, 6, ref.ctypes.data)
    assert np.array_equal(r, ref)
    assert all(K.lib.ppm_radius_at(0.1, i) == r[i] for i in range(6))
    sc = P.read_scene(os.path.join(EX, "ex-sunwindow.scene"))
    power, ns = sc.photon_budget(1000)
    assert power == 20.0 / 1000 and ns == [1000]


def test_photon_dump_roundtrip(tmp_path):
    rng = np.random.default_rng(1)
    ph = np.zeros(50, K.PHOTON_DTYPE)
    ph["pos"] = rng.normal(size=(50, 3)) * 3
    d = rng.normal(size=(50, 3))
    ph["dir"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    ph["wl"] = rng.integers(0, 3, 50)
    p = str(tmp_path / "map.txt")
    assert K.lib.ppm_write_photon_dump(p.encode(), 100000, 5e-05, ph.ctypes.data, 50) == 0
    lines = open(p).read().splitlines()
    assert lines[0] == "100000" and lines[1] == "0.00005" and len(lines) == 52          # pm.rs:44-45
    w, *nums = lines[2].split(" ")
    assert w in ("Red", "Green", "Blue") and [float(x) for x in nums[:3]] == ph["pos"][0].tolist()
    buf, n, pw = C.c_void_p(), C.c_uint64(), C.c_double()
    assert K.lib.ppm_read_photon_dump(p.encode(), C.byref(buf), C.byref(n), C.byref(pw)) == 0
    back = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_uint8)), (n.value * 56,)).view(K.PHOTON_DTYPE).copy()
    K.lib.ppm_free(buf)
    assert n.value == 50 and pw.value == 5e-05
    assert np.array_equal(back["pos"], ph["pos"]) and np.array_equal(back["wl"], ph["wl"])
    assert np.allclose(back["dir"], ph["dir"], rtol=0, atol=5e-16)        # re-normalised on read, geometry.rs:46-54


def test_image_writers(tmp_path, oracle):
    cam = P.read_camera(None, xreso=4, yreso=3)
    img = np.arange(36, dtype=np.float64).reshape(12, 3) * 1e-3
    img[5] = [0.0, 1.5e-3, 2.0]
    img[11] = 0.0                                                 # a pixel no pass ever lit
    p = str(tmp_path / "a.ppmf")
    assert K.lib.ppm_write_image(p.encode(), C.byref(cam), img.ctypes.data, 1) == 0
    lines = open(p).read().splitlines()
    assert lines[:5] == ["P3", "## max radiance = 0.01", "## image parameters = 1/250, F4, ISO100", "4 3", "255"]   # camera.rs:77-90
    assert lines[5] == "0e0 1e-3 2e-3" and lines[10] == "0e0 1.5e-3 2e0" and len(lines) == 17
    assert K.lib.ppm_write_image(p.encode(), C.byref(cam), img.ctypes.data, 0) == 0
    rgb = (C.c_int32 * 3)()
    oracle.L.orc_radiance_to_rgb(0.01, K.D3(*img[2]), rgb)
    assert open(p).read().splitlines()[7] == " ".join(str(x) for x in rgb)
    # averager2.rb:84-110 mean PPM over 4 passes
    assert K.lib.ppm_write_mean_ppm(p.encode(), C.byref(cam), (img * 4).copy().ctypes.data, 4) == 0
    lines = open(p).read().splitlines()
    assert lines[:4] == ["P3", "## max radiance = 0.01", "4 3", "255"] and lines[15] == "0 0 0"
    assert lines[6] == " ".join(str(oracle.L.orc_averager_clip(v * 4, 4, 0.01)) for v in img[2])
    # averager2.rb:154-218 float32 OpenEXR
    e = str(tmp_path / "a.exr")
    assert K.lib.ppm_write_mean_exr(e.encode(), C.byref(cam), (img * 4).copy().ctypes.data, 4) == 0
    raw = open(e, "rb").read()
    assert struct.unpack("<ii", raw[:8]) == (20000630, 2) and raw[8:17] == b"channels\0"
    hdr_end = raw.index(b"screenWindowWidth\0float\0") + len(b"screenWindowWidth\0float\0") + 4 + 4 + 1
    offs = struct.unpack("<3Q", raw[hdr_end:hdr_end + 24])
    assert offs[0] == hdr_end + 24 and offs[1] - offs[0] == 8 + 4 * 4 * 3
    y, nbytes = struct.unpack("<ii", raw[offs[1]:offs[1] + 8])
    assert (y, nbytes) == (1, 48)
    bplane = np.frombuffer(raw[offs[1] + 8:offs[1] + 8 + 16], "<f4")
    assert np.array_equal(bplane, (img[4:8, 2] * 4 / (4 * 0.01)).astype(np.float32))
