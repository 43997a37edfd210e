"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Tiers (BASELINE.json north_star):
  bit-exact   ray hit indices / t / positions / normals, neighbour-photon sets,
              camera rays, emitted photon origins
  <= 1e-5 rel per-pixel radiance estimates from an identical photon map
              (asserted much tighter here: 1e-9; the contract is RTOL_CONTRACT)
  structural  photon paths with identical Philox draws: same records, positions
              to 1e-9 (libdevice vs glibc sin/cos/pow differ in the last bits)
"""
import os

import numpy as np
import pytest

import ppmpa_b200 as P
from ppmpa_b200 import _capi as K
from ppmpa_b200.synth import wall_photons

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "examples")
RTOL_CONTRACT = 1e-5     # north_star: radiance estimates agree within 1e-5 relative
RTOL = 1e-9              # what we actually hold
SEED = 0x5EED0001

SCENES = [None, "ex-glassbox", "sample1", "ex-sunwindow", "mirror-ball", "coral-ball"]


def load_scene(name):
    return P.read_scene(None if name is None else os.path.join(EX, name + ".scene"))


def random_rays(n, seed, scene=None):
    rng = np.random.default_rng(seed)
    pos = rng.uniform([-1.9, 0.1, -5.9], [1.9, 3.9, 4.9], size=(n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([pos, d], axis=1)


WT_MAX = {K.FILTER_NONE: 1.0, K.FILTER_CONE: 1.0 / (1.0 - 2.0 / (3.0 * 1.1)), K.FILTER_GAUSS: 0.918 + 0.5}   # tracer.rs:198-216


def photon_quantum(power, r2, pfilter):
    """Largest contribution of ONE photon to a radiance estimate: wt_max * power / (pi r^2) (tracer.rs:179-195)."""
    return WT_MAX[pfilter] * power / (np.pi * r2)


def assert_outliers_bounded(g, o, quantum, flips=4):
    """A pixel may deviate because one photon falls on the other side of a d2 <= r2 test after a last-bit libm
    difference: every deviation must stay within `flips` photon contributions (throughput <= 1)."""
    worst = np.abs(g - o).max() / quantum
    assert worst <= flips, f"a pixel is off by {worst:.2f} photon contributions (quantum {quantum:.3e})"


def assert_rel(a, b, rtol, atol=0.0):
    a = np.asarray(a); b = np.asarray(b)
    err = np.abs(a - b)
    tol = atol + rtol * np.maximum(np.abs(a), np.abs(b))
    bad = err > tol
    assert not bad.any(), f"{bad.sum()} of {bad.size} differ; max abs err {err.max():.3e}"


# ---------------------------------------------------------------------------
# calc_intersection  (bit-exact tier)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", SCENES)
def test_intersect_bit_exact(engine, oracle, name):
    sc = load_scene(name)
    engine.set_scene(sc)
    rays = random_rays(20000, 11)
    # edge cases: axis-parallel rays (cos0 == 0 for some planes), rays that start
    # on a wall (t ~ 0 must be skipped by NEARLY0), rays from inside the spheres,
    # tangent-ish rays, rays leaving through nothing
    extra = []
    for ax in range(3):
        for s in (-1.0, 1.0):
            d = np.zeros(3); d[ax] = s
            extra.append(np.concatenate([[0.3, 1.7, 2.9], d]))
            extra.append(np.concatenate([[0.0, 0.0, 0.0], d]))          # starts on the floor plane
            extra.append(np.concatenate([[-1.6, 1.5, 3.0], d]))         # centre of a builtin sphere
            extra.append(np.concatenate([[0.0, 1.9, 3.0], d]))          # tangent to the r=0.4 sphere at y=1.5
    rays = np.concatenate([rays, np.array(extra)])
    g = engine.calc_intersection(rays)
    o = oracle.intersect(sc, rays)
    for a, b, what in zip(g, o, ["hit", "t", "pos", "nvec", "io"]):
        assert np.array_equal(a, b), f"{what}: {np.sum(a != b)} mismatches"
    assert (g[0] >= 0).sum() > 0.9 * len(rays)


def test_intersect_tie_break_lowest_index(engine, oracle):
    """Two coincident planes: the stable sort keeps the earlier object (tracer.rs:335-336)."""
    sc = load_scene(None)
    import ctypes as C
    prims = (K.Prim * 3)()
    for i in range(3):
        K.lib.ppm_prim_plain(C.byref(prims[i]), K.D3(0.0, 1.0, 0.0), 0.0, i)
    sc.prims, sc.nprims = prims, 3
    engine.set_scene(sc)
    rays = random_rays(500, 5)
    g = engine.calc_intersection(rays)
    o = oracle.intersect(sc, rays)
    assert np.array_equal(g[0], o[0]) and set(np.unique(g[0])) <= {-1, 0}


def test_intersect_empty_and_errors(engine):
    sc = load_scene(None)
    engine.set_scene(sc)
    hit, t, pos, nrm, io = engine.calc_intersection(np.zeros((0, 6)))
    assert len(hit) == 0
    # zero direction: every plane has cos0 == 0 -> no NaN poisoning, defined output
    r = np.array([[0.0, 1.0, 0.0, 0.0, 0.0, 0.0]])
    hit, *_ = engine.calc_intersection(r)
    assert hit[0] == -1 or hit[0] >= 0


# ---------------------------------------------------------------------------
# emission + photon tracing
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", [None, "ex-sunwindow", "ex-glassbox"])
def test_emit_parity(engine, oracle, name):
    sc = load_scene(name)
    engine.set_scene(sc)
    _, ns = sc.photon_budget(5000)
    g = engine.generate_photons(SEED, 3, ns)
    o = oracle.emit_photons(sc, SEED, 3, ns)
    assert np.array_equal(g["wl"], o["wl"])
    assert np.array_equal(g["pos"], o["pos"])                 # pure arithmetic on shared draws: bit-exact
    assert_rel(g["dir"], o["dir"], 0.0, atol=1e-14)           # sin/cos: last-bit differences allowed


def sort_by_tag(ph, tags):
    order = np.argsort(tags, kind="stable")
    return ph[order], tags[order]


@pytest.mark.parametrize("name", SCENES)
@pytest.mark.parametrize("uc", [True, False])
def test_trace_photons_parity(engine, oracle, name, uc):
    sc = load_scene(name)
    engine.set_scene(sc)
    power, ns = sc.photon_budget(20000)
    n = engine.trace_photons(SEED, 1, uc, ns, power)
    g, gp, gt = engine.export_photons(with_tags=True)
    o, ot = oracle.trace_photons(sc, SEED, 1, uc, ns)
    assert gp == power
    g, gt = sort_by_tag(g, gt)
    o, ot = sort_by_tag(o, ot)
    # Same Philox draws on both sides: paths agree unless a last-bit sin/cos/pow
    # difference flips a roulette / hit decision (never observed; allow 1e-4).
    common, gi, oi = np.intersect1d(gt, ot, return_indices=True)
    assert len(common) >= (1 - 1e-4) * max(len(gt), len(ot)), (n, len(ot), len(common))
    assert np.array_equal(g["wl"][gi], o["wl"][oi])
    assert_rel(g["pos"][gi], o["pos"][oi], 0.0, atol=1e-9)
    assert_rel(g["dir"][gi], o["dir"][oi], 0.0, atol=1e-9)
    if uc:
        assert np.all((gt & 15) > 0)           # depth-0 hits are not stored with use_classic (tracer.rs:77)
    assert n > 0


def test_trace_photons_pass_streams_disjoint(engine):
    sc = load_scene(None)
    engine.set_scene(sc)
    power, ns = sc.photon_budget(4000)
    engine.trace_photons(SEED, 0, False, ns, power); a, _ = engine.export_photons()
    engine.trace_photons(SEED, 1, False, ns, power); b, _ = engine.export_photons()
    engine.trace_photons(SEED, 0, False, ns, power); c, _ = engine.export_photons()
    key = lambda x: np.sort(x["pos"][:, 0])
    assert len(a) == len(c) and np.array_equal(key(a), key(c))      # same (seed, pass) -> same photons
    assert len(a) != len(b) or not np.array_equal(key(a), key(b))   # another pass -> another stream


# ---------------------------------------------------------------------------
# photon map: neighbour sets (bit-exact tier) and radiance estimate
# ---------------------------------------------------------------------------
def query_points(ph, n, seed, jitter):
    rng = np.random.default_rng(seed)
    q = ph["pos"][rng.integers(0, len(ph), n)] + rng.normal(scale=jitter, size=(n, 3))
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    return q, nrm


@pytest.mark.parametrize("r", [0.025, 0.1, 0.3])
def test_within_sets_identical(engine, oracle, r):
    ph, power = wall_photons(50000, seed=3)
    engine.import_photons(ph, power)
    engine.build_photonmap(r * r)
    m = oracle.map_build(ph, power, r * r)
    q, _ = query_points(ph, 300, 4, r / 3)
    # also far-away queries (outside the grid) and exact photon positions (d2 == 0)
    q = np.concatenate([q, [[100.0, 100.0, 100.0], [-50.0, 0.0, 0.0]], ph["pos"][:20]])
    cap = 4096
    idx, cnt = engine.within(q, cap)
    for i in range(len(q)):
        oi, _, k = m.within(q[i])
        assert cnt[i] == k
        assert np.array_equal(idx[i, :k], np.sort(oi)), f"query {i}: neighbour set differs"
    assert cnt[-1] >= 1 and cnt[len(q) - 22] == 0


def test_within_outliers_are_clamped_not_lost(engine, oracle):
    """The reference leaks ~1e-4 of its photons through wall corners, far outside the room.  The
    grid covers a trimmed region and clamps outliers into boundary cells: neighbour sets of
    queries inside, on the boundary and far outside must still be exact."""
    ph, power = wall_photons(60000, seed=13)
    rng = np.random.default_rng(2)
    far = rng.integers(0, len(ph), 40)                      # < n/1024 outliers: they WILL be trimmed
    ph["pos"][far] += rng.normal(scale=30.0, size=(40, 3))
    ph["pos"][far[:3]] = [[1e6, -1e6, 3e5], [2.5, 4.5, 5.5], [-2.05, -0.05, -6.05]]
    r = 0.1
    engine.import_photons(ph, power)
    engine.build_photonmap(r * r)
    m = oracle.map_build(ph, power, r * r)
    q = np.concatenate([ph["pos"][far] + rng.normal(scale=r / 10, size=(40, 3)),       # around the outliers
                        ph["pos"][:200] + rng.normal(scale=r / 3, size=(200, 3)),      # inside the room
                        [[-2.0, 0.0, -6.0], [2.0, 4.0, 5.0], [2.04, 4.04, 5.04]]])      # room corners
    idx, cnt = engine.within(q, 2048)
    g, gc = engine.estimate_radiance(q, np.tile([0.0, 1.0, 0.0], (len(q), 1)))
    o, oc = m.gather(q, np.tile([0.0, 1.0, 0.0], (len(q), 1)), K.FILTER_NONE)
    for i in range(len(q)):
        oi, _, k = m.within(q[i])
        assert cnt[i] == k == gc[i] == oc[i], i
        assert np.array_equal(idx[i, :k], np.sort(oi)), i
    assert cnt[:40].min() >= 1
    assert_rel(g, o, RTOL)


@pytest.mark.parametrize("pfilter", [K.FILTER_NONE, K.FILTER_CONE, K.FILTER_GAUSS])
@pytest.mark.parametrize("r", [0.05, 0.1])
def test_gather_parity(engine, oracle, pfilter, r):
    ph, power = wall_photons(200000, seed=5)
    engine.import_photons(ph, power)
    engine.build_photonmap(r * r)
    m = oracle.map_build(ph, power, r * r)
    q, nrm = query_points(ph, 5000, 6, r / 4)
    g, gc = engine.estimate_radiance(q, nrm, pfilter)
    o, oc = m.gather(q, nrm, pfilter, nthreads=4)
    assert np.array_equal(gc, oc)                      # |neighbour set| bit-exact
    assert gc.max() > 10
    assert_rel(g, o, RTOL)
    assert RTOL <= RTOL_CONTRACT


@pytest.mark.parametrize("pfilter", [K.FILTER_NONE, K.FILTER_GAUSS])
def test_gather_heavy_groups(engine, oracle, pfilter, monkeypatch):
    """Work concentrated in few queries (BASELINE config 3: a sunlit patch holding most photons): groups whose
    candidate stream exceeds GATHER_HEAVY_MIN are split over 8 warps by k_gather_heavy.  Same neighbour counts
    as the single-warp path and the oracle, radiance equal up to the summation order."""
    rng = np.random.default_rng(11)
    n = 400000
    ph = np.zeros(n, K.PHOTON_DTYPE)
    dense = rng.random(n) < 0.9                                        # 90 % of the photons on a 0.5 x 0.5 patch of the floor
    ph["pos"][:, 0] = np.where(dense, rng.uniform(-0.25, 0.25, n), rng.uniform(-2, 2, n))
    ph["pos"][:, 2] = np.where(dense, rng.uniform(2.0, 2.5, n), rng.uniform(-6, 5, n))
    d = rng.normal(size=(n, 3)); d[:, 1] = -np.abs(d[:, 1]) - 0.1; d /= np.linalg.norm(d, axis=1, keepdims=True)
    ph["dir"] = d
    ph["wl"] = rng.integers(0, 3, n)
    power = 5.0 / n
    r = 0.1
    q = np.zeros((6000, 3))
    q[:, 0] = rng.uniform(-1.0, 1.0, 6000); q[:, 2] = rng.uniform(1.0, 3.5, 6000)
    nrm = np.tile([0.0, 1.0, 0.0], (len(q), 1))
    engine.import_photons(ph, power); engine.build_photonmap(r * r)
    g, gc = engine.estimate_radiance(q, nrm, pfilter)
    engine.set_option("gather_heavy", 0)
    try:
        g1, gc1 = engine.estimate_radiance(q, nrm, pfilter)
    finally:
        engine.set_option("gather_heavy", 1)
    assert gc.max() > 20000                                            # really heavy: > 2e4 neighbours per query in the patch
    assert np.array_equal(gc, gc1)
    assert_rel(g, g1, 1e-12)
    assert not np.array_equal(g, g1)                                   # the heavy path did run (different summation order)
    sub = np.concatenate([np.argsort(gc)[-150:], np.arange(150)])
    m = oracle.map_build(ph, power, r * r)
    o, oc = m.gather(q[sub], nrm[sub], pfilter, nthreads=4)
    assert np.array_equal(gc[sub], oc)
    assert_rel(g[sub], o, RTOL)
    # k-NN with k larger than any neighbourhood takes the same (heavy) path and reproduces the fixed radius exactly
    same, r2k, c2 = engine.estimate_radiance_knn(q, nrm, 10 ** 7, pfilter)
    assert np.array_equal(same, g) and np.array_equal(c2, gc)


def test_gather_empty_map_and_state_errors(engine):
    eng2 = P.Engine(0)
    try:
        with pytest.raises(P.PPMError) as e:
            eng2.estimate_radiance(np.zeros((1, 3)), np.zeros((1, 3)))
        assert e.value.code == -2                      # PPM_ERR_STATE: map not built
        eng2.import_photons(np.zeros(0, K.PHOTON_DTYPE), 1.0)
        eng2.build_photonmap(0.01)
        g, c = eng2.estimate_radiance(np.array([[0.0, 1.0, 2.0]]), np.array([[0.0, 1.0, 0.0]]))
        assert np.all(g == 0) and c[0] == 0
        g, c = eng2.estimate_radiance(np.zeros((0, 3)), np.zeros((0, 3)))
        assert g.shape == (0, 3)
        with pytest.raises(P.PPMError):
            eng2.build_photonmap(0.0)
    finally:
        eng2.close()


def test_gather_single_photon_known_answer(engine):
    """One photon straight down onto an up-facing point: L = power * 1 / (pi r^2)."""
    ph = np.zeros(1, K.PHOTON_DTYPE)
    ph["pos"][0] = [0.0, 0.0, 0.0]; ph["dir"][0] = [0.0, -1.0, 0.0]; ph["wl"][0] = K.WL_GREEN
    engine.import_photons(ph, 0.5)
    r2 = 0.01
    engine.build_photonmap(r2)
    g, c = engine.estimate_radiance(np.array([[0.05, 0.0, 0.0], [0.2, 0.0, 0.0]]), np.array([[0.0, 1.0, 0.0]] * 2))
    assert c.tolist() == [1, 0]
    assert g[0].tolist() == [0.0, (0.5 * 1.0) * ((1.0 / np.pi) / r2), 0.0]
    assert g[1].tolist() == [0.0, 0.0, 0.0]
    # a photon arriving from below (n.dir > 0) contributes nothing but is still counted (optics.rs:226)
    g, c = engine.estimate_radiance(np.array([[0.05, 0.0, 0.0]]), np.array([[0.0, -1.0, 0.0]]))
    assert c[0] == 1 and np.all(g == 0)


def test_gather_properties_large(engine):
    """Size-independent properties at a BASELINE-sized map (1 M photons):
    sum_q K_q is symmetric under swapping the roles of photons and queries, radiance is
    linear in the photon power, and K is monotone in r."""
    ph, power = wall_photons(1_000_000, seed=SEED)
    qs, _ = wall_photons(200_000, seed=9)
    nrm = -qs["dir"]
    r = 0.05
    engine.import_photons(ph, power); engine.build_photonmap(r * r)
    g1, c1 = engine.estimate_radiance(qs["pos"], nrm)
    engine.import_photons(ph, 2.0 * power); engine.build_photonmap(r * r)
    g2, c2 = engine.estimate_radiance(qs["pos"], nrm)
    assert np.array_equal(c1, c2) and np.array_equal(g2, 2.0 * g1)       # exact: power enters as one factor of 2
    engine.import_photons(qs, power); engine.build_photonmap(r * r)
    _, crev = engine.estimate_radiance(ph["pos"], -ph["dir"])
    assert int(c1.sum()) == int(crev.sum())
    engine.import_photons(ph, power); engine.build_photonmap((2 * r) ** 2)
    _, c4 = engine.estimate_radiance(qs["pos"], nrm)
    assert np.all(c4 >= c1) and c4.sum() > 3 * c1.sum()


# ---------------------------------------------------------------------------
# camera rays, eye paths, direct light, whole pass
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("kw", [dict(), dict(blur=0), dict(antialias=0), dict(blur=0, progressive=0)])
def test_generate_rays_bit_exact(engine, oracle, kw):
    cam = P.read_camera(os.path.join(EX, "camera0.scr"), xreso=96, yreso=64, **kw)
    engine.set_camera(cam)
    g = engine.generate_rays(SEED, 7)
    o = oracle.generate_rays(cam, SEED, 7)
    assert np.array_equal(g, o)


@pytest.mark.parametrize("name,uc,pfilter", [
    (None, True, K.FILTER_NONE), (None, False, K.FILTER_CONE), ("ex-glassbox", True, K.FILTER_GAUSS),
    ("ex-glassbox", False, K.FILTER_NONE), ("sample1", True, K.FILTER_NONE), ("mirror-ball", True, K.FILTER_NONE),
    ("coral-ball", False, K.FILTER_NONE), ("ex-sunwindow", False, K.FILTER_NONE), ("ex-sunwindow", True, K.FILTER_NONE)])
def test_trace_rays_parity(engine, oracle, name, uc, pfilter):
    """trace_ray on an IDENTICAL photon map (exported from the engine, fed to the oracle)."""
    sc = load_scene(name)
    cam = P.read_camera(os.path.join(EX, "camera0.scr"), xreso=48, yreso=48, pfilter=pfilter, progressive=1)
    engine.set_scene(sc); engine.set_camera(cam)
    power, ns = sc.photon_budget(60000)
    engine.trace_photons(SEED, 2, uc, ns, power)
    ph, pw = engine.export_photons()
    r2 = 0.15 ** 2
    engine.build_photonmap(r2)
    rays = oracle.generate_rays(cam, SEED, 2)
    g = engine.trace_rays(rays, SEED, 2, uc)
    m = oracle.map_build(ph, pw, r2)
    o, stats = oracle.trace_rays(sc, m, pfilter, rays, SEED, 2, uc, nthreads=4)
    assert o.max() > 0
    # glossy directions differ in the last bits (pow/sin/cos), which can move a secondary
    # hit point by ~1e-15 and flip one photon in or out of a neighbour set on rare pixels
    err = np.abs(g - o) / np.maximum(np.maximum(np.abs(g), np.abs(o)), 1e-300)
    frac_bad = np.mean(np.any(err > RTOL, axis=1))
    assert frac_bad <= 2e-3, f"{frac_bad:.4%} pixels differ by more than {RTOL}"
    # ... and each of them by no more than a few photon contributions, or -- where a last-bit difference sends a glossy
    # secondary ray to another surface -- by no more than the brightest thing an eye path can see
    bound = max(4 * photon_quantum(pw, r2, pfilter), 0.0)
    worst = np.abs(g - o).max()
    assert worst <= max(bound, o.max()), (worst, bound)
    if frac_bad == 0.0:
        assert worst <= 1e-9 * o.max()
    assert np.median(err) < 1e-13


def test_direct_light_off_by_one_quirk(engine, oracle):
    """Point/sun lights contribute zero classic direct light (SURVEY B-6); area light matches the oracle."""
    sc = load_scene("ex-sunwindow")
    cam = P.read_camera(None, xreso=32, yreso=32, blur=0, antialias=0)
    engine.set_scene(sc); engine.set_camera(cam)
    engine.import_photons(np.zeros(0, K.PHOTON_DTYPE), 1.0); engine.build_photonmap(0.01)
    rays = oracle.generate_rays(cam, 1, 0)
    g = engine.trace_rays(rays, 1, 0, True)
    m = oracle.map_build(np.zeros(0, K.PHOTON_DTYPE), 1.0, 0.01)
    o, _ = oracle.trace_rays(sc, m, K.FILTER_NONE, rays, 1, 0, True)
    assert_rel(g, o, RTOL)
    # only the emitter quad's own emittance/(2 pi) term survives
    assert np.count_nonzero(g) == np.count_nonzero(o)


@pytest.mark.parametrize("name", [None, "ex-glassbox", "sample1", "mirror-ball"])
def test_direct_light_matches_oracle(engine, oracle, name):
    """Classic direct light alone (empty photon map): 25 shadow rays per node incl. the off-by-one pairing."""
    sc = load_scene(name)
    cam = P.read_camera(os.path.join(EX, "camera0.scr"), xreso=160, yreso=160, progressive=1, pfilter=K.FILTER_NONE)
    engine.set_scene(sc); engine.set_camera(cam)
    engine.import_photons(np.zeros(0, K.PHOTON_DTYPE), 1.0); engine.build_photonmap(0.01)
    rays = oracle.generate_rays(cam, 7, 0)
    first = len(rays) // 3
    sub = slice(first, first + 4000)                         # contiguous: pixel ids (RNG streams) stay aligned
    g = engine.trace_rays(rays[sub], 7, 0, True, first_pixel=first)
    m = oracle.map_build(np.zeros(0, K.PHOTON_DTYPE), 1.0, 0.01)
    o, _ = oracle.trace_rays(sc, m, K.FILTER_NONE, rays[sub], 7, 0, True, first_pixel=first, nthreads=4)
    assert o.max() > 0
    err = np.abs(g - o) / np.maximum(np.maximum(np.abs(o), np.abs(g)), 1e-300)
    assert np.mean(np.any(err > RTOL, axis=1)) <= 2e-3
    assert np.isfinite(g).all() and np.abs(g - o).max() <= o.max()      # a flipped shadow-ray decision moves a pixel by at most one node's direct light


ADVERSARIAL_SCENE = """
# open scene (no enclosing room), tilted planes with un-normalised normals, a triangle, spheres,
# two area lights (one tilted, one tiny): stresses every branch of the shadow-ray culling
light:
  - type: parallelogram
    color: [ 1.0, 0.8, 0.6 ]
    flux: 5.0
    position: [ -0.5, 3.0, 2.5 ]
    dir1: [ 1.0, 0.2, 0.0 ]
    dir2: [ 0.0, 0.3, 1.0 ]
  - type: parallelogram
    color: [ 0.2, 0.3, 1.0 ]
    flux: 2.0
    position: [ 1.5, 1.0, 0.0 ]
    dir1: [ 0.0, 0.05, 0.0 ]
    dir2: [ 0.0, 0.0, 0.05 ]
material:
  - type: solid
    name: m
    emittance: [ 0.0, 0.0, 0.0 ]
    reflectance: [ 0.5, 0.5, 0.5 ]
    transmittance: [ 0.0, 0.0, 0.0 ]
    specularrefl: [ 0.0, 0.0, 0.0 ]
    ior: [ 0.0, 0.0, 0.0 ]
    diffuseness: 1.0
    metalness: 0.0
    smoothness: 0.0
object:
  - type: plain
    name: floor
    normal: [ 0.0, 1.0, 0.0 ]
    position: [ 0.0, 0.0, 0.0 ]
    material: m
  - type: plain
    name: tilted
    normal: [ 3.0, 1.0, 0.5 ]
    position: [ -2.0, 0.0, 0.0 ]
    material: m
  - type: plain
    name: above_the_light
    normal: [ 0.0, -2.0, 0.1 ]
    position: [ 0.0, 3.6, 0.0 ]
    material: m
  - type: sphere
    name: ball
    center: [ 0.2, 1.0, 2.8 ]
    radius: 0.7
    material: m
  - type: sphere
    name: small_ball
    center: [ 1.0, 1.0, 0.3 ]
    radius: 0.2
    material: m
  - type: polygon
    name: tri
    pos1: [ -1.0, 1.5, 2.0 ]
    pos2: [ 1.0, 2.0, 2.2 ]
    pos3: [ -0.5, 1.8, 4.0 ]
    material: m
  - type: parallelogram
    name: card
    pos1: [ -1.5, 0.0, 1.0 ]
    pos2: [ -1.5, 1.0, 1.0 ]
    pos3: [ -1.2, 0.0, 2.0 ]
    material: m
"""


def _cull_probe_nodes(engine, seed):
    """Surface points as the eye paths produce them (hits of random rays) plus free points with
    arbitrary normals, points a hair above/below every kind of surface and points around the lights."""
    rays = random_rays(30000, seed)
    hit, t, pos, nrm, io = engine.calc_intersection(rays)
    ok = hit >= 0
    pos, nrm = pos[ok], nrm[ok]
    rng = np.random.default_rng(seed + 1)
    free = rng.uniform([-2.5, -0.5, -6.5], [2.5, 4.5, 5.5], size=(15000, 3))
    fn = rng.normal(size=(15000, 3)); fn /= np.linalg.norm(fn, axis=1, keepdims=True)
    # offsets straddling the culling margins (1e-6 sign margin, 1e-2 certificate gap, NEARLY0)
    k = min(len(pos), 12000)
    off = rng.choice([0.0, 1e-12, -1e-12, 1e-7, -1e-7, 9e-7, 1.1e-6, -1.1e-6, 1e-4, -1e-4, 5e-3, 1.1e-2, -1.1e-2], size=(k, 1))
    near = pos[:k] + off * nrm[:k]
    lights = rng.uniform([-1.2, 2.8, 2.3], [1.2, 4.2, 3.7], size=(6000, 3))
    ln = rng.normal(size=(6000, 3)); ln /= np.linalg.norm(ln, axis=1, keepdims=True)
    P3 = np.concatenate([pos, free, near, lights])
    N3 = np.concatenate([nrm, fn, nrm[:k], ln])
    return np.ascontiguousarray(P3), np.ascontiguousarray(N3)


@pytest.mark.parametrize("coherent", [False, True])
@pytest.mark.parametrize("name", SCENES + ["adversarial"])
def test_direct_light_cull_is_exact(engine, oracle, name, coherent, tmp_path, monkeypatch):
    """k_direct_light's conservative per-node culling must not change a decision: compare with the
    culling switched off (option dl_cull = 0) and with the oracle's get_radiance_from_light.
    coherent=True feeds the probe nodes sorted by a 5 cm grid, as ppm_render_pass does (cell-sorted order): the 32
    nodes of a warp are then neighbours, which is the case the warp-wide mask OR (and any per-warp classification)
    is built for; the unsorted order makes every warp a random mix."""
    if name == "adversarial":
        f = tmp_path / "adversarial.scene"
        f.write_text(ADVERSARIAL_SCENE)
        sc = P.read_scene(str(f))
    else:
        sc = load_scene(name)
    engine.set_scene(sc)
    pos, nrm = _cull_probe_nodes(engine, 77)
    if coherent:
        cell = np.floor(pos / 0.05).astype(np.int64)
        order = np.lexsort((cell[:, 0], cell[:, 1], cell[:, 2]))
        pos, nrm = np.ascontiguousarray(pos[order]), np.ascontiguousarray(nrm[order])
    engine.set_option("dl_cull", 1)
    a = engine.direct_light(pos, nrm)
    engine.set_option("dl_cull", 0)
    try:
        b = engine.direct_light(pos, nrm)
    finally:
        engine.set_option("dl_cull", 1)
    # Culling may only remove work, never change a decision: the same samples survive with and without it.  Nodes that
    # have nothing to test take the division-lean arithmetic of k_direct_light (cos0^2 = (n.d)^2 / |d|^2 instead of
    # normalising d), so the values agree to a few ulp rather than bit for bit.
    assert np.array_equal(a > 0, b > 0), f"{np.sum(np.any((a > 0) != (b > 0), axis=1))} of {len(a)} nodes lit differently with culling"
    assert_rel(a, b, 1e-12, atol=1e-15)               # grazing samples: the reference's own cos0 = n . (d/|d|) cancels to ~1e-16 absolute
    sub = np.random.default_rng(5).choice(len(pos), 12000, replace=False)
    o = oracle.direct_light(sc, pos[sub], nrm[sub])
    if name not in ("ex-sunwindow",):
        assert o.max() > 0 and np.count_nonzero(np.any(o > 0, axis=1)) > 500
    assert np.array_equal(a[sub] > 0, o > 0)           # same lit/occluded decisions
    assert_rel(a[sub], o, 1e-12)


@pytest.mark.parametrize("name", [None, "ex-glassbox", "sample1", "mirror-ball", "ex-sunwindow"])
def test_trace_rays_classic_parity(engine, oracle, name):
    """trace_ray_classic (tracer.rs:221-259, the `rtc` renderer): deterministic given the rays."""
    sc = load_scene(name)
    cam = P.read_camera(os.path.join(EX, "screen1.scr"), xreso=64, yreso=64)       # ambient 0.001
    engine.set_scene(sc); engine.set_camera(cam)
    rays = oracle.generate_rays(cam, 3, 0)
    g = engine.trace_rays_classic(rays, 3, 0)
    o = oracle.trace_rays_classic(sc, list(cam.ambient), rays)
    assert o.max() > 0 and list(cam.ambient) == [0.001, 0.001, 0.001]
    assert_rel(g, o, 1e-12)


def test_render_pass_matches_oracle_and_accumulates(engine, oracle):
    sc = load_scene("ex-glassbox")
    cam = P.read_camera(os.path.join(EX, "camera0.scr"), xreso=40, yreso=40, pfilter=K.FILTER_NONE, progressive=1)
    engine.set_scene(sc); engine.set_camera(cam)
    engine.accum_reset()
    radii = P.radius_schedule(0.2, 2)
    imgs = []
    for p in range(2):
        engine.iteration(SEED, p, 30000, radii[p] ** 2, uc=True)
        imgs.append(engine.pass_image())
        o, _, ostats = oracle.render_pass(sc, cam, SEED, p, 30000, radii[p] ** 2, True)
        err = np.abs(imgs[-1] - o) / np.maximum(np.maximum(np.abs(o), np.abs(imgs[-1])), 1e-300)
        assert np.mean(np.any(err > 1e-6, axis=1)) <= 5e-3
        assert_outliers_bounded(imgs[-1], o, photon_quantum(sc.photon_budget(30000)[0], radii[p] ** 2, K.FILTER_NONE))
        ms, ct = engine.last_pass_stats()
        assert ct["emitted"] == 30000 and ct["stored"] == int(ostats[0]) and ct["launches"] > 5
        assert abs(ct["sum_k"] - int(ostats[3])) <= 1e-3 * int(ostats[3])
    acc, n = engine.accum_read()
    assert n == 2
    assert np.array_equal(acc, imgs[0] + imgs[1])
    assert np.array_equal(engine.image_mean(), acc / 2.0)


@pytest.mark.parametrize("xres,yres,rows", [(1024, 1024, (640, 644)), (1920, 1080, (700, 703))])
def test_full_size_pass_rows_match_oracle(engine, oracle, xres, yres, rows):
    """BASELINE configs 2 and 5 at FULL size (1 M photons, 1024^2 / 1920x1080): the oracle traces the same
    1 M photon paths and the image rows that cross the glass box; the rest of the image is covered by
    size-independent properties (finite, non-negative, accumulator == pass image, record count)."""
    sc = load_scene("ex-glassbox")
    cam = P.read_camera(os.path.join(EX, "camera0.scr"), xreso=xres, yreso=yres, pfilter=K.FILTER_NONE, progressive=1)
    engine.set_scene(sc); engine.set_camera(cam)
    engine.accum_reset()
    r2 = 0.06 ** 2
    engine.iteration(SEED, 5, 1_000_000, r2, uc=True)
    img = engine.pass_image()
    ms, ct = engine.last_pass_stats()
    o, _, ostats = oracle.render_pass(sc, cam, SEED, 5, 1_000_000, r2, True, rows[0], rows[1])
    assert ct["emitted"] == 1_000_000 and ct["stored"] == int(ostats[0])
    assert int(ostats[2]) > 1.5 * (rows[1] - rows[0]) * xres            # the rows do cross the glass box (eye-path trees)
    g = img.reshape(yres, xres, 3)[rows[0]:rows[1]].reshape(-1, 3)
    err = np.abs(g - o) / np.maximum(np.maximum(np.abs(o), np.abs(g)), 1e-300)
    assert np.mean(np.any(err > 1e-6, axis=1)) <= 5e-3
    assert_outliers_bounded(g, o, photon_quantum(sc.photon_budget(1_000_000)[0], r2, K.FILTER_NONE))
    assert np.median(err) < 1e-12
    assert np.isfinite(img).all() and img.min() >= 0.0 and img.max() > 0.0
    acc, n = engine.accum_read()
    assert n == 1 and np.array_equal(acc, img)


def test_two_contexts_are_independent():
    a, b = P.Engine(0), P.Engine(0)
    try:
        a.set_scene(load_scene(None)); b.set_scene(load_scene("ex-glassbox"))
        rays = random_rays(1000, 2)
        ha = a.calc_intersection(rays)[0]; hb = b.calc_intersection(rays)[0]
        assert ha.max() > 12 and hb.max() <= 12
    finally:
        a.close(); b.close()


def test_render_passes_two_lanes_equal_sequential(engine):
    """ppm_render_passes (two lanes on one GPU) == the same passes rendered one by one."""
    sc = load_scene("ex-glassbox")
    cam = P.read_camera(os.path.join(EX, "camera0.scr"), xreso=64, yreso=64, pfilter=K.FILTER_NONE, progressive=1)
    engine.set_scene(sc); engine.set_camera(cam)
    radii = P.radius_schedule(0.2, 5)
    engine.accum_reset()
    imgs = []
    for p in range(5):
        engine.iteration(SEED, 10 + 2 * p, 20000, radii[p] ** 2, uc=True)
        imgs.append(engine.pass_image())
    engine.accum_reset()
    engine.iterate(SEED, 10, 5, 20000, radii ** 2, uc=True, pass_stride=2)
    acc, n = engine.accum_read()
    assert n == 5
    assert np.array_equal(engine.pass_image(), imgs[4])                    # last pass of the batch
    lane0 = (imgs[0] + imgs[2]) + imgs[4]
    lane1 = imgs[1] + imgs[3]
    assert np.array_equal(acc, lane0 + lane1)                              # per-lane sums, then merged
    ms, ct = engine.last_pass_stats()
    assert ct["emitted"] == 5 * 20000 and ct["launches"] >= 5 * 10
    # odd batch sizes / single pass / empty batch
    engine.accum_reset()
    engine.iterate(SEED, 10, 1, 20000, radii[:1] ** 2)
    assert np.array_equal(engine.pass_image(), imgs[0])
    engine.iterate(SEED, 10, 0, 20000, [])
    assert engine.accum_read()[1] == 1


@pytest.mark.parametrize("pfilter", [K.FILTER_NONE, K.FILTER_CONE, K.FILTER_GAUSS])
@pytest.mark.parametrize("k", [1, 10, 100, 500])
def test_gather_knn_matches_bruteforce(engine, oracle, pfilter, k):
    """k-NN estimate against the brute-force CPU statement (no reference implementation exists):
    the k-th squared distance is found exactly, counts and radiance agree."""
    ph, power = wall_photons(100000, seed=21)
    r = 0.15
    engine.import_photons(ph, power)
    engine.build_photonmap(r * r)
    m = oracle.map_build(ph, power, r * r)
    q, nrm = query_points(ph, 1500, 22, r / 4)
    q = np.concatenate([q, [[50.0, 50.0, 50.0]], ph["pos"][:5]])           # no neighbours / d2 == 0 neighbours
    nrm = np.concatenate([nrm, [[0.0, 1.0, 0.0]] * 6])
    g, gr, gc = engine.estimate_radiance_knn(q, nrm, k, pfilter)
    o, orr, oc = m.gather_knn(q, nrm, k, pfilter)
    assert np.array_equal(gr, orr)                                          # r_k^2 bit-exact
    assert np.array_equal(gc, oc)
    assert_rel(g, o, RTOL)
    full = gc[:1500]
    assert np.all(full[orr[:1500] < r * r] == k) and gc[1500] == 0 and gr[1500] == r * r


def test_gather_knn_config4_sweep(engine):
    """configs[3]: mirror-ball / coral-ball, fixed radius vs k-NN: with k photons found everywhere the
    k-NN estimate has the same mean as the fixed-radius estimate to within Monte-Carlo noise, and a
    huge k reproduces the fixed-radius estimate exactly."""
    for name in ("mirror-ball", "coral-ball"):
        sc = load_scene(name)
        cam = P.read_camera(os.path.join(EX, "camera0.scr"), xreso=64, yreso=64, blur=0, antialias=0, progressive=1)
        engine.set_scene(sc); engine.set_camera(cam)
        power, ns = sc.photon_budget(200000)
        engine.trace_photons(SEED, 0, False, ns, power)
        rays = engine.generate_rays(SEED, 0)
        hit, t, pos, nrm, io = engine.calc_intersection(rays)
        pos, nrm = pos[hit >= 0], nrm[hit >= 0]
        for r in (0.1, 0.2, 0.3):
            engine.build_photonmap(r * r)
            fixed, cnt = engine.estimate_radiance(pos, nrm, K.FILTER_NONE)
            same, r2k, c2 = engine.estimate_radiance_knn(pos, nrm, 10 ** 7, K.FILTER_NONE)
            assert np.array_equal(same, fixed) and np.array_equal(c2, cnt) and np.all(r2k == r * r)
            for k in (100, 500):
                knn, r2k, ck = engine.estimate_radiance_knn(pos, nrm, k, K.FILTER_NONE)
                assert np.all(ck <= np.maximum(cnt, k)) and np.all(r2k <= r * r)
                lit = cnt > 4 * k
                if lit.sum() > 100:
                    assert abs(knn[lit].mean() / fixed[lit].mean() - 1.0) < 0.15


def test_converged_image_statistical_parity(engine, oracle):
    """Third tier of the parity contract: with DIFFERENT random streams the averaged multi-pass image must
    agree with the reference estimator statistically.  GPU passes use seed A, oracle passes seed B; the
    per-pixel standard error comes from the oracle's own pass-to-pass variance (a bare RMSE is dominated by
    a few firefly pixels -- light edges, caustic spikes -- that only one side happens to sample).
    Stated bounds at 48x48, 24 passes x 40 k photons:
      * per-pixel z-scores: mean |z| < 1.25 and fewer than 1 % beyond 5 sigma;
      * no bias: total energy within 2 %, median per-pixel ratio within 0.5 %;
      * trimmed relative RMSE (pixels with standard error < 5 % of their mean, largest 1 % of deviations
        discarded) <= 10 %."""
    sc = load_scene("ex-glassbox")
    cam = P.read_camera(os.path.join(EX, "camera0.scr"), xreso=48, yreso=48, pfilter=K.FILTER_NONE, progressive=1)
    engine.set_scene(sc); engine.set_camera(cam)
    NP, NPH = 24, 40000
    radii = P.radius_schedule(0.25, NP)
    engine.accum_reset()
    engine.iterate(0xA11CE, 0, NP, NPH, radii ** 2, uc=True)
    g = engine.image_mean()
    passes = np.stack([oracle.render_pass(sc, cam, 0xB0B, p, NPH, radii[p] ** 2, True)[0] for p in range(NP)])
    o = passes.mean(0)
    se = passes.std(0, ddof=1) / np.sqrt(NP)
    lum_g, lum_o, lum_se = g.sum(1), o.sum(1), np.sqrt((se ** 2).sum(1))
    ok = lum_se > 0
    z = (lum_g[ok] - lum_o[ok]) / (lum_se[ok] * np.sqrt(2.0))      # both means carry the same variance
    assert np.mean(np.abs(z)) < 1.25 and np.mean(np.abs(z) > 5) < 0.01, (np.mean(np.abs(z)), np.mean(np.abs(z) > 5))
    assert abs(lum_g.mean() / lum_o.mean() - 1.0) < 0.02
    smooth = ok & (lum_se < 0.05 * np.maximum(lum_o, 1e-300))
    assert smooth.mean() > 0.5
    assert abs(np.median(lum_g[smooth] / lum_o[smooth]) - 1.0) < 0.005
    dev = np.sort(np.abs(lum_g[smooth] - lum_o[smooth]))
    dev = dev[: int(len(dev) * 0.99)]
    trimmed = np.sqrt(np.mean(dev ** 2)) / lum_o[smooth].mean()
    assert trimmed <= 0.10, trimmed
