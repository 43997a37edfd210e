"""Pins the oracle against every still-valid known answer the reference's own
unit tests hold for this path.  CPU only."""
import ctypes as C

import numpy as np

import oracle_lib
from oracle_lib import d3
from ppmpa_b200 import _capi as K


def v(x):
    return [x[0], x[1], x[2]]


def test_vector3_goldens(oracle):
    L = oracle.L
    out = K.D3()
    # algebra.rs:268  normalize(1,-2,3)
    assert L.orc_normalize(d3([1.0, -2.0, 3.0]), out) == 1
    assert v(out) == [0.2672612419124244, -0.5345224838248488, 0.8017837257372732]
    # algebra.rs:275  v1 * 1.1 -> 3.3000000000000003
    L.orc_scale(d3([1.0, 2.0, 3.0]), 1.1, out)
    assert v(out) == [1.1, 2.2, 3.3000000000000003]
    # zero vector -> None
    assert L.orc_normalize(d3([0.0, 0.0, 0.0]), out) == 0


def test_ray_goldens(oracle):
    L = oracle.L
    out = K.D3()
    # geometry.rs:217-218  new_dir(1,1,1)
    L.orc_normalize(d3([1.0, 1.0, 1.0]), out)
    assert v(out) == [0.5773502691896258] * 3
    # geometry.rs:224-228  Ray(1,1,1 ; -1,-1,-1).target(2.0)
    nd = K.D3()
    L.orc_normalize(d3([-1.0, -1.0, -1.0]), nd)
    L.orc_ray_target(d3([1.0, 1.0, 1.0]), nd, 2.0, out)
    assert v(out) == [-0.15470053837925168] * 3


def test_getnormal_goldens(oracle):
    L = oracle.L
    out = K.D3()
    # geometry.rs:231-243
    pl = K.Prim(); pl.type = K.SHAPE_PLAIN; pl.nvec = d3([0.0, 1.0, 0.0]); pl.scalar = 1.0
    assert L.orc_shape_normal(C.byref(pl), d3([0.0, 1.0, 0.0]), out) == 1 and v(out) == [0.0, 1.0, 0.0]
    sp = K.Prim(); sp.type = K.SHAPE_SPHERE; sp.position = d3([0.0, 0.0, 0.0]); sp.scalar = 2.0
    assert L.orc_shape_normal(C.byref(sp), d3([2.0, 0.0, 0.0]), out) == 1 and v(out) == [1.0, 0.0, 0.0]
    for para in (0, 1):
        po = K.Prim()
        assert L.orc_new_polygon(d3([0.0, 0.0, 0.0]), d3([2.0, 1.0, 0.0]), d3([0.0, 1.0, 2.0]), para, C.byref(po)) == 1
        assert L.orc_shape_normal(C.byref(po), d3([0.0, 1.0, 0.0]), out) == 1
        assert v(out) == [0.4082482904638631, -0.8164965809277261, 0.4082482904638631]
    pt = K.Prim(); pt.type = K.SHAPE_POINT
    assert L.orc_shape_normal(C.byref(pt), d3([0.0, 1.0, 0.0]), out) == 0


def test_filter_goldens(oracle):
    L = oracle.L
    r = 0.1 * 0.1
    # tracer.rs:369-372 (valid)
    assert L.orc_filter_cone(0.0, r) == 2.538461538461538
    assert L.orc_filter_cone(r, r) == 0.23076923076923078
    # tracer.rs:373-376 are STALE (they pin CORR=0.355); the code at :211 has CORR=0.5.
    # The oracle follows the code; the stale values differ by exactly the CORR delta.
    g0, g1 = L.orc_filter_gauss(0.0, r), L.orc_filter_gauss(r, r)
    assert g0 == 1.4180000000000001
    assert abs(g1 - 0.7511526928041553) < 1e-15
    assert abs((g0 - 1.2730000000000001) - (0.5 - 0.355)) < 1e-15
    assert abs((g1 - 0.6061526928041553) - (0.5 - 0.355)) < 1e-15


def test_color_goldens(oracle):
    L = oracle.L
    out = K.D3()
    # physics.rs:372-377
    L.orc_color_normalize(d3([0.4, 0.78, 1.0]), out)
    assert v(out) == [0.1834862385321101, 0.35779816513761464, 0.4587155963302752]
    assert out[0] + out[1] + out[2] == 1.0
    assert L.orc_decide_wavelength(out, 0.1) == K.WL_RED
    assert L.orc_decide_wavelength(out, 0.3) == K.WL_GREEN
    assert L.orc_decide_wavelength(out, 0.7) == K.WL_BLUE
    L.orc_color_normalize(d3([0.0, -1.0, 0.0]), out)
    assert v(out) == [1.0 / 3.0] * 3


def test_check_under_goldens(oracle):
    # physics.rs:409-415
    ps = (C.c_double * 5)(0.1, 0.2, 0.3, 0.5, 0.8)
    for p, want in [(0.03, 0), (0.12, 1), (0.28, 2), (0.4, 3), (0.64, 4), (0.99, 5)]:
        assert oracle.L.orc_check_under(ps, 5, p) == want


def test_relative_ior(oracle):
    # physics.rs:200-205: eta = n2/n1, 1.0 when n1 == 0 (the reference's test_ior asserts a stale 0.0)
    L = oracle.L
    assert L.orc_relative_ior_average(d3([1.0, 1.0, 1.0]), d3([1.5, 1.5, 1.5])) == 1.5
    assert L.orc_relative_ior_average(d3([0.0, 0.0, 0.0]), d3([1.5, 1.5, 1.5])) == 1.0


def test_radius_schedule(oracle):
    # util/iterator.rb:34-38, alpha = 0.5
    out = np.zeros(4)
    oracle.L.orc_radius_schedule(0.1, 4, out.ctypes.data)
    r = 0.1
    want = []
    for i in range(4):
        want.append(r)
        r = np.sqrt(((i + 1) + 0.5) / ((i + 1) + 1.0)) * r
    assert out.tolist() == want
    assert out[1] == np.sqrt(1.5 / 2.0) * 0.1


def test_tonemap(oracle):
    rgb = (C.c_int32 * 3)()
    oracle.L.orc_radiance_to_rgb(0.01, d3([0.0, 0.005, 1.0]), rgb)
    assert list(rgb) == [0, int(np.floor((0.5 ** (1 / 2.2)) * 255)), 255]
    assert oracle.L.orc_averager_clip(0.5, 100, 0.01) == int(((0.5 / 100 / 0.01) ** (1 / 2.2)) * 255)


def test_within_grid_equals_bruteforce(oracle):
    """The oracle's grid accelerator must return exactly the brute-force neighbour set."""
    from ppmpa_b200.synth import wall_photons
    ph, power = wall_photons(20000, seed=7)
    for r in (0.05, 0.1, 0.3):
        m = oracle.map_build(ph, power, r * r)
        rng = np.random.default_rng(1)
        for q in ph["pos"][rng.integers(0, len(ph), 40)] + rng.normal(scale=0.02, size=(40, 3)):
            a, da, ka = m.within(q, brute=False)
            b, db, kb = m.within(q, brute=True)
            assert ka == kb and np.array_equal(a, b) and np.array_equal(da, db)
            assert np.all(np.diff(da) >= 0)
