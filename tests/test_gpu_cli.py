"""GPU tests of the drop-in CLIs (same argv / stdin / stdout protocols as src/bin/{ppmpa,pm,rt}.rs)."""
import os
import subprocess

import numpy as np
import pytest

import ppmpa_b200 as P
from ppmpa_b200 import _capi as K

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "ppmpa_b200", "bin")
EX = os.path.join(ROOT, "examples")
ENV = dict(os.environ, PPM_SEED="12345", PPM_PASS="3")


def parse_ppmf(text):
    lines = text.splitlines()
    w, h = [int(x) for x in lines[3].split()]
    img = np.array([[float(x) for x in l.split()] for l in lines[5:5 + w * h]])
    return lines[:5], img


def test_pm_rt_pipeline_matches_library(engine, tmp_path):
    scene, cam_file = os.path.join(EX, "ex-glassbox.scene"), os.path.join(EX, "screen1.scr")
    pm = subprocess.run([os.path.join(BIN, "pm"), scene, "30000"], capture_output=True, env=ENV, check=True)
    dump = pm.stdout.decode().splitlines()
    assert dump[0] == "30000" and float(dump[1]) == 5.0 / 30000          # pm.rs:44-45
    assert dump[2].split(" ")[0] in ("Red", "Green", "Blue") and len(dump[2].split(" ")) == 7
    # the same photons through the library
    sc = P.read_scene(scene)
    engine.set_scene(sc)
    power, ns = sc.photon_budget(30000)
    engine.trace_photons(12345, 3, True, ns, power)
    ph, _ = engine.export_photons()
    assert len(dump) - 2 == len(ph)
    key = lambda a: np.sort(np.round(a, 12))
    cli_x = np.array([float(l.split(" ")[1]) for l in dump[2:]])
    assert np.array_equal(key(cli_x), key(ph["pos"][:, 0]))
    # rt: photon dump on stdin -> image on stdout (screen1.scr is not progressive -> RGB ints)
    rt = subprocess.run([os.path.join(BIN, "rt"), scene, cam_file, "0.2"], input=pm.stdout, capture_output=True, env=ENV, check=True)
    out = rt.stdout.decode().splitlines()
    assert out[:5] == ["P3", "## max radiance = 0.01", "## image parameters = 1/250, F4, ISO100", "256 256", "255"]
    assert len(out) == 5 + 256 * 256 and b"finished reading map" in rt.stderr
    rgb = np.array([[int(x) for x in l.split()] for l in out[5:]])
    assert rgb.min() >= 0 and rgb.max() <= 255 and rgb.mean() > 5
    # library path on the dump read back (directions re-normalised exactly as read_map does)
    p = tmp_path / "map.txt"
    p.write_bytes(pm.stdout)
    cam = P.read_camera(cam_file)
    engine.set_camera(cam)
    engine.read_map(str(p), 0.2 * 0.2)
    img = engine.trace_rays(engine.generate_rays(12345, 3), 12345, 3, True)
    want = np.floor(np.minimum(img / cam.max_radiance, 1.0) ** (1 / 2.2) * 255).astype(int)
    assert np.mean(np.any(np.abs(want - rgb) > 1, axis=1)) < 1e-3


def test_ppmpa_matches_render_pass(engine):
    scene, cam_file = os.path.join(EX, "mirror-ball.scene"), os.path.join(EX, "screen1.scr")
    for flags, uc in ([], True), (["-nc"], False):
        r = subprocess.run([os.path.join(BIN, "ppmpa")] + flags + ["20000", "0.15", cam_file, scene], capture_output=True, env=ENV, check=True)
        hdr, img = parse_ppmf(r.stdout.decode())
        assert hdr[3] == "256 256"
        engine.set_scene(P.read_scene(scene)); engine.set_camera(P.read_camera(cam_file))
        engine.iteration(12345, 3, 20000, 0.15 ** 2, uc=uc)
        assert np.array_equal(img, engine.pass_image())                  # `{:e}` text round-trips every f64
    # builtin = what the reference hard-codes; bad numbers fall back to the defaults (ppmpa.rs:56-63)
    r = subprocess.run([os.path.join(BIN, "ppmpa"), "x", "y", "builtin", "builtin"], capture_output=True, env=ENV, check=True)
    hdr, img = parse_ppmf(r.stdout.decode())
    engine.set_scene(P.read_scene()); engine.set_camera(P.read_camera())
    engine.iteration(12345, 3, 100000, 0.1 * 0.1, uc=True)
    assert hdr[3] == "256 256" and np.array_equal(img, engine.pass_image())
    # usage on too few arguments, exit status 0 like the reference
    r = subprocess.run([os.path.join(BIN, "ppmpa"), "-h"], capture_output=True)
    assert r.returncode == 0 and b"Usage: ppmpa" in r.stderr


def test_rtc_matches_library(engine):
    """rtc <screen file> <scene file> (src/bin/rtc.rs): classic tracer, header first, RGB ints when not progressive."""
    scene, cam_file = os.path.join(EX, "coral-ball.scene"), os.path.join(EX, "screen1.scr")
    r = subprocess.run([os.path.join(BIN, "rtc"), cam_file, scene], capture_output=True, env=ENV, check=True)
    out = r.stdout.decode().splitlines()
    assert out[:5] == ["P3", "## max radiance = 0.01", "## image parameters = 1/250, F4, ISO100", "256 256", "255"]
    rgb = np.array([[int(x) for x in l.split()] for l in out[5:]])
    cam = P.read_camera(cam_file)
    engine.set_scene(P.read_scene(scene)); engine.set_camera(cam)
    img = engine.trace_rays_classic(engine.generate_rays(12345, 3), 12345, 3)
    want = np.floor(np.minimum(img / cam.max_radiance, 1.0) ** (1 / 2.2) * 255).astype(int)
    assert rgb.shape == (256 * 256, 3) and np.array_equal(rgb, want)
    r = subprocess.run([os.path.join(BIN, "rtc")], capture_output=True)
    assert r.returncode == 0 and b"Usage: rtc" in r.stdout


def _read_p3(path):
    tok = open(path).read().split("\n")
    body = [l for l in tok if l and not l.startswith("#")]
    assert body[0] == "P3"
    w, h = [int(x) for x in body[1].split()]
    vals = np.array(" ".join(body[3:]).split(), dtype=np.int64)
    return vals.reshape(h * w, 3)


def test_ppmpa_frame_matches_library(engine, tmp_path):
    """ppmpa_frame = util/iterator.rb:90-117 + util/averager2.rb:49-110 in one process over the C ABI: the frame it
    writes is the mean of the same pass ids rendered through the library.  With 2 GPUs present the same job sharded
    over both (NCCL reduce inside libppm_b200.so, no Python) must give the same picture."""
    import ctypes as C
    import torch
    scene, cam_file = os.path.join(EX, "mirror-ball.scene"), os.path.join(EX, "screen1.scr")
    out1 = tmp_path / "frame1.ppm"
    r = subprocess.run([os.path.join(BIN, "ppmpa_frame"), "5", "20000", "0.2", cam_file, scene, str(out1)], capture_output=True, env=ENV)
    assert r.returncode == 0, r.stderr.decode()
    assert b"5 passes x 20000 photons at 256x256 on 1 GPU(s)" in r.stderr
    cam = P.read_camera(cam_file)
    engine.set_scene(P.read_scene(scene)); engine.set_camera(cam)
    engine.accum_reset()
    radii = P.radius_schedule(0.2, 5)
    engine.iterate(12345, 0, 5, 20000, radii ** 2, uc=True)
    acc, n = engine.accum_read()
    want = tmp_path / "want.ppm"
    assert K.lib.ppm_write_mean_ppm(str(want).encode(), C.byref(cam), acc.ctypes.data, n) == 0
    a, b = _read_p3(out1), _read_p3(want)
    assert a.shape == (256 * 256, 3) and np.array_equal(a, b)
    assert a.max() > 50
    if torch.cuda.device_count() >= 2:
        out2 = tmp_path / "frame2.ppm"
        r = subprocess.run([os.path.join(BIN, "ppmpa_frame"), "-g", "2", "5", "20000", "0.2", cam_file, scene, str(out2)], capture_output=True, env=ENV)
        assert r.returncode == 0, r.stderr.decode()
        c = _read_p3(out2)
        # same pass images, another summation order: 8-bit values may differ by one level on a rounding boundary
        assert np.abs(c - a).max() <= 1 and np.mean(c != a) < 1e-3
