"""GPU parity of the BVH path (scenes beyond 64 primitives, or the "bvh" option): the traversal must return exactly
what calc_intersection's scan over every object returns (tracer.rs:306-350), so every downstream stage agrees with
the oracle as it does on the small scenes.

  bit-exact   hit index / t / position / normal / io against the oracle's brute-force scan, including rays aimed at
              shared edges and vertices of a mesh (ties go to the lower object index, tracer.rs:335-336), rays parallel
              to the axes and rays starting on a surface
  bit-exact   BVH mode == brute-force mode on the six example scenes (same engine, option switched)
  structural  photon records, direct light and a whole pass on a 973-primitive glass-mesh scene against the oracle
"""
import contextlib
import os

import numpy as np
import pytest

import ppmpa_b200 as P
from ppmpa_b200 import _capi as K
from ppmpa_b200 import synth

from test_gpu_parity import (EX, SCENES, SEED, assert_outliers_bounded, assert_rel, load_scene, photon_quantum,
                             random_rays, sort_by_tag)

pytestmark = pytest.mark.gpu

GLASS = 4          # material index of `glass` in ex-glassbox.scene
WALL = 0


@contextlib.contextmanager
def bvh_forced(engine):
    engine.set_option("bvh", 1)
    try:
        yield
    finally:
        engine.set_option("bvh", 0)


def mesh_scene(nlat=16, nlon=32, material=GLASS, spheres=()):
    base = load_scene("ex-glassbox")
    tris = synth.uv_sphere_triangles((0.3, 2.6, 1.0), 0.7, nlat, nlon)
    return synth.mesh_scene(base, tris, material, spheres), tris


def mesh_rays(tris, seed):
    """Rays from two viewpoints through the vertices, edge midpoints and centroids of the mesh (vertices and edges are
    shared by several triangles: equal distances), plus rays parallel to the axes through vertices."""
    rng = np.random.default_rng(seed)
    v = tris.reshape(-1, 3)
    targets = np.concatenate([v, 0.5 * (tris[:, 0] + tris[:, 1]), 0.5 * (tris[:, 1] + tris[:, 2]), tris.mean(axis=1)])
    rays = []
    for eye in ([1.0, 2.0, -4.5], [-1.5, 0.5, 4.0], [0.3, 2.6, 1.0]):         # the last one is the mesh's centre
        eye = np.asarray(eye)
        d = targets - eye
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        rays.append(np.concatenate([np.broadcast_to(eye, d.shape), d], axis=1))
    for ax in range(3):
        d = np.zeros(3); d[ax] = 1.0
        o = v[rng.integers(0, len(v), 300)].copy()
        o[:, ax] = -1.5
        rays.append(np.concatenate([o, np.broadcast_to(d, o.shape)], axis=1))
        rays.append(np.concatenate([o - 3.0 * d, np.broadcast_to(-d, o.shape)], axis=1))
    # rays that start ON the mesh (t ~ 0 must be skipped by NEARLY0) in random directions
    o = tris.mean(axis=1)
    d = rng.normal(size=o.shape)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays.append(np.concatenate([o, d], axis=1))
    return np.concatenate(rays)


def assert_hits_equal(g, o):
    for a, b, what in zip(g, o, ["hit", "t", "pos", "nvec", "io"]):
        assert np.array_equal(a, b), f"{what}: {np.sum(a != b)} mismatches"


@pytest.mark.parametrize("name", SCENES)
def test_bvh_mode_equals_bruteforce_on_example_scenes(engine, oracle, name):
    sc = load_scene(name)
    rays = random_rays(20000, 21)
    par = random_rays(3000, 22)                          # rays parallel (or within 1e-300 of parallel) to the axes
    par[:1000, 3:] = [0.0, 1.0, 0.0]
    par[1000:2000, 3:] = [-1.0, 0.0, 1e-300]
    par[2000:, 3:] = [0.0, -0.0, -1.0]
    rays = np.concatenate([rays, par])
    engine.set_scene(sc)
    brute = engine.calc_intersection(rays)
    with bvh_forced(engine):
        engine.set_scene(sc)
        assert engine.get_option("bvh") == 1
        g = engine.calc_intersection(rays)
    engine.set_scene(sc)
    assert_hits_equal(g, brute)
    assert_hits_equal(g, oracle.intersect(sc, rays))


def test_bvh_mesh_hits_bit_exact(engine, oracle):
    sc, tris = mesh_scene(spheres=[((-1.2, 0.5, 2.0), 0.5), ((1.2, 0.4, 3.0), 0.4)])
    assert sc.nprims > 64
    engine.set_scene(sc)
    rays = np.concatenate([random_rays(30000, 31), mesh_rays(tris, 32)])
    g = engine.calc_intersection(rays)
    o = oracle.intersect(sc, rays)
    assert_hits_equal(g, o)
    assert (g[0] >= 13).sum() > 3000                      # the mesh and the spheres are hit, not only the room
    # ties: some rays through shared edges reach two triangles at the same distance
    assert len(np.unique(g[0])) > 500


def test_bvh_degenerate_meshes(engine, oracle):
    """Coincident triangles (all centroids equal: the builder's median fallback), a single bounded primitive (root with
    one child) and a scene without planes."""
    base = load_scene("ex-glassbox")
    t = np.array([[[0.0, 1.0, 1.0], [1.0, 1.0, 1.0], [0.0, 2.0, 1.5]]])
    for tris in (np.repeat(t, 100, axis=0), t):
        sc = synth.mesh_scene(base, tris, WALL)
        with bvh_forced(engine):
            engine.set_scene(sc)
            rays = random_rays(5000, 41)
            assert_hits_equal(engine.calc_intersection(rays), oracle.intersect(sc, rays))
    # only bounded primitives
    tris = synth.uv_sphere_triangles((0.0, 1.0, 0.0), 1.0, 8, 16)
    sc = synth.mesh_scene(base, tris, WALL)
    sc.prims = (K.Prim * len(tris))(*[sc.prims[base.nprims + i] for i in range(len(tris))])
    sc.nprims = len(tris)
    engine.set_scene(sc)
    rays = random_rays(5000, 42)
    g = engine.calc_intersection(rays)
    assert_hits_equal(g, oracle.intersect(sc, rays))
    assert (g[0] < 0).any() and (g[0] >= 0).any()
    engine.set_scene(base)


@pytest.mark.parametrize("uc", [True, False])
def test_bvh_trace_photons_parity(engine, oracle, uc):
    sc, _ = mesh_scene()
    engine.set_scene(sc)
    power, ns = sc.photon_budget(20000)
    n = engine.trace_photons(SEED, 1, uc, ns, power)
    g, gp, gt = engine.export_photons(with_tags=True)
    o, ot = oracle.trace_photons(sc, SEED, 1, uc, ns)
    g, gt = sort_by_tag(g, gt)
    o, ot = sort_by_tag(o, ot)
    common, gi, oi = np.intersect1d(gt, ot, return_indices=True)
    assert len(common) >= (1 - 1e-4) * max(len(gt), len(ot)), (n, len(ot), len(common))
    assert np.array_equal(g["wl"][gi], o["wl"][oi])
    assert_rel(g["pos"][gi], o["pos"][oi], 0.0, atol=1e-9)
    assert_rel(g["dir"][gi], o["dir"][oi], 0.0, atol=1e-9)
    assert n > 0


def test_bvh_direct_light_matches_oracle(engine, oracle):
    sc, _ = mesh_scene(material=WALL)                   # an opaque mesh: it casts a shadow
    cam = P.read_camera(os.path.join(EX, "camera0.scr"), xreso=96, yreso=96, progressive=1, pfilter=K.FILTER_NONE)
    engine.set_scene(sc); engine.set_camera(cam)
    engine.import_photons(np.zeros(0, K.PHOTON_DTYPE), 1.0); engine.build_photonmap(0.01)
    rays = oracle.generate_rays(cam, 7, 0)
    g = engine.trace_rays(rays, 7, 0, True)
    m = oracle.map_build(np.zeros(0, K.PHOTON_DTYPE), 1.0, 0.01)
    o, _ = oracle.trace_rays(sc, m, K.FILTER_NONE, rays, 7, 0, True, nthreads=8)
    assert o.max() > 0 and (o.sum(axis=1) == 0).any()      # lit and shadowed pixels
    assert_rel(g, o, 1e-12)


def test_bvh_render_pass_matches_oracle(engine, oracle):
    """A whole pass (photon tracing through the glass mesh, map, eye paths, direct light, gather) on 973 primitives."""
    sc, _ = mesh_scene()
    cam = P.read_camera(os.path.join(EX, "camera0.scr"), xreso=48, yreso=48, pfilter=K.FILTER_NONE, progressive=1)
    engine.set_scene(sc); engine.set_camera(cam)
    engine.accum_reset()
    r2 = 0.15 ** 2
    engine.iteration(SEED, 3, 30000, r2, uc=True)
    img = engine.pass_image()
    o, _, ostats = oracle.render_pass(sc, cam, SEED, 3, 30000, r2, True)
    ms, ct = engine.last_pass_stats()
    assert ct["emitted"] == 30000 and ct["stored"] == int(ostats[0])
    err = np.abs(img - o) / np.maximum(np.maximum(np.abs(o), np.abs(img)), 1e-300)
    assert np.mean(np.any(err > 1e-6, axis=1)) <= 5e-3
    assert_outliers_bounded(img, o, photon_quantum(sc.photon_budget(30000)[0], r2, K.FILTER_NONE))
    assert np.median(err) < 1e-12
    engine.set_scene(load_scene("ex-glassbox"))


def test_bvh_large_mesh_properties(engine):
    """65 k triangles (beyond what the oracle scans in seconds): size-independent properties.  Every primary ray from
    outside that passes clearly inside the silhouette hits the mesh, on the tessellated sphere, at the analytic
    sphere's distance within the tessellation error; rays clearly outside never do."""
    base = load_scene("ex-glassbox")
    c, r = np.array([0.3, 2.4, 0.8]), 0.9
    tris = synth.uv_sphere_triangles(c, r, 128, 256)
    sc = synth.mesh_scene(base, tris, GLASS)
    engine.set_scene(sc)
    rng = np.random.default_rng(5)
    eye = np.array([1.0, 2.0, -4.5])
    tgt = c + rng.normal(size=(200000, 3)) * 0.5
    d = tgt - eye
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([np.broadcast_to(eye, d.shape), d], axis=1)
    hit, t, pos, nrm, io = engine.calc_intersection(rays)
    oc = c - eye
    b = d @ oc
    disc = r * r - (oc @ oc - b * b)
    mesh = hit >= base.nprims
    inner = disc > (0.05 * r) ** 2                      # clearly inside the silhouette
    assert mesh[inner].all()
    t_an = b[mesh] - np.sqrt(np.maximum(disc[mesh], 0.0))
    assert np.all(np.abs(np.linalg.norm(pos[mesh] - c, axis=1) - r) < 1e-3 * r)      # on the tessellated sphere
    assert np.median(np.abs(t[mesh] - t_an)) < 1e-3
    assert (np.einsum("ij,ij->i", nrm[mesh], d[mesh]) <= 0.0).all()                  # the normal faces the ray
    assert (~mesh[disc < -(0.05 * r) ** 2]).all()          # clearly outside the silhouette: never the mesh
    engine.set_scene(base)


def test_ppmpa_cli_on_a_mesh_scene_file(engine, tmp_path):
    """The drop-in CLI (`ppmpa <#photon> <radius> <camera> <scene>`, ppmpa.rs:15) on a 973-object scene FILE: the
    parser, the BVH build and the pass give the image the library gives for the same scene built through the ABI."""
    import subprocess
    from test_gpu_cli import BIN, ENV, parse_ppmf
    path = os.path.join(EX, "ex-glassbox.scene")
    tris = synth.uv_sphere_triangles((0.3, 2.6, 1.0), 0.7, 16, 32)
    f = tmp_path / "mesh.scene"
    f.write_text(synth.mesh_scene_text(open(path).read(), tris, "glass"))
    cam_file = os.path.join(EX, "screen1.scr")
    r = subprocess.run([os.path.join(BIN, "ppmpa"), "20000", "0.15", cam_file, str(f)], capture_output=True, env=ENV, check=True)
    hdr, img = parse_ppmf(r.stdout.decode())
    sc, _ = mesh_scene()
    engine.set_scene(sc); engine.set_camera(P.read_camera(cam_file))
    engine.iteration(12345, 3, 20000, 0.15 ** 2, uc=True)
    assert hdr[3] == "256 256" and np.array_equal(img, engine.pass_image())
    engine.set_scene(load_scene("ex-glassbox"))


@pytest.mark.parametrize("coherent", [False, True])
@pytest.mark.parametrize("name", ["mesh-opaque", "mesh-glass", None, "ex-glassbox", "sample1", "adversarial"])
def test_bvh_direct_light_cull_is_exact(engine, oracle, name, coherent, tmp_path):
    """BVH mode's shadow-ray culling (planes classified per node as in brute-force mode, the hierarchy against the
    node -> light-quad shaft, kernels_eye.cuh: bvh_shaft_classify) must not change a decision: culled == unculled,
    and both == the oracle's get_radiance_from_light.  Mesh scenes, and example scenes with the hierarchy forced
    (ex-glassbox and sample1 have an emitter polygon lying in the light's plane: class (b))."""
    from test_gpu_parity import ADVERSARIAL_SCENE, _cull_probe_nodes
    forced = contextlib.nullcontext()
    if name == "mesh-opaque":
        sc, _ = mesh_scene(material=WALL, spheres=[((-1.0, 0.6, 0.5), 0.6)])
    elif name == "mesh-glass":
        sc, _ = mesh_scene(material=GLASS)
    else:
        if name == "adversarial":
            f = tmp_path / "adversarial.scene"
            f.write_text(ADVERSARIAL_SCENE)
            sc = P.read_scene(str(f))
        else:
            sc = load_scene(name)
        forced = bvh_forced(engine)
    with forced:
        engine.set_scene(sc)
        pos, nrm = _cull_probe_nodes(engine, 78)
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
    assert np.array_equal(a > 0, b > 0), f"{np.sum(np.any((a > 0) != (b > 0), axis=1))} of {len(a)} nodes lit differently with culling"
    assert_rel(a, b, 1e-12, atol=1e-15)
    sub = np.random.default_rng(5).choice(len(pos), 6000, replace=False)
    o = oracle.direct_light(sc, pos[sub], nrm[sub])
    assert o.max() > 0 and np.count_nonzero(np.any(o > 0, axis=1)) > 300
    assert np.array_equal(a[sub] > 0, o > 0)
    assert_rel(a[sub], o, 1e-12)
    engine.set_scene(load_scene("ex-glassbox"))


def test_bvh_two_lanes_equal_sequential_and_scene_switch(engine):
    """ppm_render_passes on two lanes (the second lane reads the first one's hierarchy) == the same passes one by one,
    on a mesh scene; then switching back to a small scene (brute-force mode) and to the mesh again re-renders the
    same images (graphs and lanes follow the scene version)."""
    sc, _ = mesh_scene()
    small = load_scene("ex-glassbox")
    cam = P.read_camera(os.path.join(EX, "camera0.scr"), xreso=64, yreso=64, pfilter=K.FILTER_NONE, progressive=1)
    engine.set_scene(sc); engine.set_camera(cam)
    radii = P.radius_schedule(0.2, 4)
    engine.accum_reset()
    imgs = []
    for p in range(4):
        engine.iteration(SEED, 20 + p, 20000, radii[p] ** 2, uc=True)
        imgs.append(engine.pass_image())
    engine.accum_reset()
    engine.iterate(SEED, 20, 4, 20000, radii ** 2, uc=True)
    acc, n = engine.accum_read()
    assert n == 4 and np.array_equal(engine.pass_image(), imgs[3])
    assert np.array_equal(acc, (imgs[0] + imgs[2]) + (imgs[1] + imgs[3]))
    engine.set_scene(small)
    engine.accum_reset()
    engine.iterate(SEED, 20, 2, 20000, radii[:2] ** 2, uc=True)
    small_img = engine.pass_image()
    assert not np.array_equal(small_img, imgs[1])
    engine.set_scene(sc)
    engine.accum_reset()
    engine.iterate(SEED, 20, 2, 20000, radii[:2] ** 2, uc=True)
    assert np.array_equal(engine.pass_image(), imgs[1])
    engine.set_scene(small)


@pytest.mark.parametrize("material", [GLASS, WALL])
def test_bvh_trace_rays_classic_parity(engine, oracle, material):
    """trace_ray_classic (tracer.rs:221-259, the `rtc` renderer) through the hierarchy: deterministic given the rays."""
    sc, _ = mesh_scene(material=material)
    cam = P.read_camera(os.path.join(EX, "screen1.scr"), xreso=64, yreso=64)
    engine.set_scene(sc); engine.set_camera(cam)
    rays = oracle.generate_rays(cam, 3, 0)
    g = engine.trace_rays_classic(rays, 3, 0)
    o = oracle.trace_rays_classic(sc, list(cam.ambient), rays)
    assert o.max() > 0
    assert_rel(g, o, 1e-12)
    engine.set_scene(load_scene("ex-glassbox"))


@pytest.mark.parametrize("seed,scale,offset", [(1, 1.0, 0.0), (2, 1.0, 0.0), (3, 1e-3, 0.0), (4, 1e3, 0.0), (5, 1.0, 0.0),
                                               (6, 37.5, 0.0), (7, 1.0, 1e5), (8, 0.01, -3e3)])
def test_bvh_random_scenes_match_bruteforce_and_oracle(engine, oracle, seed, scale, offset):
    """Random soups of overlapping triangles, parallelograms (some needle-thin), spheres (some nested) and tilted
    planes at scene scales from 1e-3 to 1e3, also far from the origin (coordinates 1e5 x the scene's size: the hit
    arithmetic loses 5 digits there, the padded boxes must still hold every accepted hit): the hierarchy must return
    the scan's hit for every ray."""
    import ctypes as C
    rng = np.random.default_rng(seed)
    base = load_scene("ex-glassbox")
    prims = []

    def poly(p0, p1, p2, para):
        q = K.Prim()
        if K.lib.ppm_prim_polygon(C.byref(q), K.D3(*p0), K.D3(*p1), K.D3(*p2), int(para), int(rng.integers(0, base.nmats))) == 0:
            prims.append(q)

    for _ in range(34):
        c = rng.uniform(-2, 2, 3) * scale + offset
        e = rng.normal(size=(2, 3)) * scale * rng.choice([1.0, 0.3, 1e-3])          # 1e-3: needles
        poly(c, c + e[0], c + e[1], rng.random() < 0.3)
    for _ in range(8):
        q = K.Prim()
        K.lib.ppm_prim_sphere(C.byref(q), K.D3(*(rng.uniform(-1, 1, 3) * scale + offset)), float(rng.uniform(0.05, 1.5) * scale), 0)
        prims.append(q)
    for _ in range(4):
        q = K.Prim()
        n = rng.normal(size=3)
        K.lib.ppm_prim_plain(C.byref(q), K.D3(*n), float(rng.uniform(1, 4) * scale - n.sum() * offset), 0)
        prims.append(q)
    order = rng.permutation(len(prims))                      # planes in between the bounded primitives
    arr = (K.Prim * len(prims))(*[prims[i] for i in order])
    sc = synth.ArrayScene(arr, base.mats, base.lights)
    sc.nmats, sc.nlights = base.nmats, base.nlights
    assert 40 <= sc.nprims <= 64
    o = rng.uniform(-3, 3, size=(30000, 3)) * scale + offset
    d = rng.normal(size=(30000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    aimed = (rng.uniform(-2, 2, size=(15000, 3)) * scale + offset) - o[:15000]       # towards the cluster
    d[:15000] = aimed / np.linalg.norm(aimed, axis=1, keepdims=True)
    rays = np.concatenate([o, d], axis=1)
    engine.set_scene(sc)
    brute = engine.calc_intersection(rays)
    with bvh_forced(engine):
        engine.set_scene(sc)
        walked = engine.calc_intersection(rays)
    assert_hits_equal(walked, brute)
    assert_hits_equal(walked, oracle.intersect(sc, rays))
    assert (walked[0] >= 0).mean() > 0.5
    engine.set_scene(base)


def test_bvh_edge_scenes(engine, oracle):
    """Planes only with the hierarchy forced (nothing to index), a Point primitive among the triangles (never hit,
    calc_distance has no root for it), and more than 64 infinite planes (refused with PPM_ERR_CAPACITY, not indexed)."""
    import ctypes as C
    base = load_scene("ex-glassbox")
    # planes only, forced
    planes = synth.ArrayScene((K.Prim * 6)(*[base.prims[i] for i in range(6)]), base.mats, base.lights)
    planes.nmats, planes.nlights = base.nmats, base.nlights
    rays = random_rays(4000, 51)
    with bvh_forced(engine):
        engine.set_scene(planes)
        assert_hits_equal(engine.calc_intersection(rays), oracle.intersect(planes, rays))
        pos = rays[:500, :3].copy()
        nrm = np.tile([0.0, 1.0, 0.0], (500, 1))
        assert_rel(engine.direct_light(pos, nrm), oracle.direct_light(planes, pos, nrm), 1e-12)
    # a Point primitive in a mesh scene
    tris = synth.uv_sphere_triangles((0.3, 2.6, 1.0), 0.7, 8, 16)
    sc = synth.mesh_scene(base, tris, WALL)
    sc.prims[base.nprims + 5].type = K.SHAPE_POINT
    engine.set_scene(sc)
    g = engine.calc_intersection(rays)
    assert_hits_equal(g, oracle.intersect(sc, rays))
    assert not (g[0] == base.nprims + 5).any()
    # 70 planes + 1 triangle: too many unbounded primitives for the constant list
    many = (K.Prim * 71)()
    for i in range(70):
        K.lib.ppm_prim_plain(C.byref(many[i]), K.D3(0.0, 1.0, 0.0), float(i), 0)
    many[70] = sc.prims[base.nprims]
    bad = synth.ArrayScene(many, base.mats, base.lights)
    bad.nmats, bad.nlights = base.nmats, base.nlights
    with pytest.raises(P.PPMError) as e:
        engine.set_scene(bad)
    assert e.value.code == -4
    engine.set_scene(base)
