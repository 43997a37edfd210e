"""BVH path: build figures, traversal throughput (device-resident rays, CUDA events on the engine's stream) and
whole-pass times for a glass sphere mesh of growing size inside the ex-glassbox room; diagnostic.

usage: python tools/bvh_bench.py [--sizes 16x32,64x128,256x512] [--pass-res 1920x1080] [--no-pass]
The first row is the 13-primitive ex-glassbox scene itself, brute force against BVH-forced (same hits)."""
import argparse, ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ppmpa_b200 as P
from ppmpa_b200 import _capi as K, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "examples")
ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="16x32,64x128,256x512")
ap.add_argument("--pass-res", default="1920x1080")
ap.add_argument("--no-pass", action="store_true")
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()

eng = P.Engine(0)
stream = torch.cuda.ExternalStream(eng.stream)
base = P.read_scene(os.path.join(EX, "ex-glassbox.scene"))
xres, yres = [int(x) for x in a.pass_res.split("x")]
cam = P.read_camera(os.path.join(EX, "camera0.scr"), xreso=xres, yreso=yres, progressive=1, pfilter=0)
eng.set_camera(cam)
prim_rays = torch.from_numpy(eng.generate_rays(7, 0)).cuda()
rng = np.random.default_rng(3)
n_rand = 2_000_000
pos = rng.uniform([-1.9, 0.1, -5.9], [1.9, 3.9, 4.9], size=(n_rand, 3))
d = rng.normal(size=(n_rand, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
rand_rays = torch.from_numpy(np.concatenate([pos, d], axis=1)).cuda()


def rays_per_s(rays):
    n = rays.shape[0]
    hit = torch.empty(n, dtype=torch.int32, device="cuda"); t = torch.empty(n, dtype=torch.float64, device="cuda")
    best = 1e30
    for _ in range(a.reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record()
            rc = K.lib.ppm_intersect(eng._h, rays.data_ptr(), n, hit.data_ptr(), t.data_ptr(), None, None, None)
            e1.record()
        assert rc == 0
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return n / best / 1e3, int(hit.to(torch.int64).sum().item())      # M rays/s, checksum


def one_pass(label):
    if a.no_pass:
        return ""
    for i in range(3):
        eng.iteration(0x5EED0001, 600 + i, 1_000_000, 0.0215 ** 2, True)
    ms, ct = eng.last_pass_stats()
    return (f"pass {ms['total']:.2f} ms (trace {ms['photon_trace']:.2f} expand {ms['eye_expand']:.2f} direct {ms['direct_light']:.2f} "
            f"gather {ms['gather']:.2f}) nodes {ct['gather_nodes']}")


def row(label, sc, nprims, build_s, info):
    p, cs1 = rays_per_s(prim_rays)
    r, cs2 = rays_per_s(rand_rays)
    print(f"{label:>22}  prims {nprims:>8}  build {build_s:6.2f} s  {info:<34} primary {p:8.1f} M rays/s  random {r:8.1f} M rays/s  "
          f"{one_pass(label)}  [{cs1} {cs2}]", flush=True)


eng.set_option("lanes", 1)
eng.set_scene(base)
row("ex-glassbox brute", base, base.nprims, 0.0, "")
eng.set_option("bvh", 1)
eng.set_scene(base)
row("ex-glassbox BVH", base, base.nprims, 0.0, "")
eng.set_option("bvh", 0)
for s in a.sizes.split(","):
    nlat, nlon = [int(x) for x in s.split("x")]
    tris = synth.uv_sphere_triangles((0.3, 2.6, 1.0), 0.7, nlat, nlon)
    sc = synth.mesh_scene(base, tris, 4)
    nn, nl, dep, cost = C.c_int64(), C.c_int64(), C.c_int32(), C.c_double()
    K.lib.ppm_bvh_inspect(sc.prims, sc.nprims, C.byref(nn), C.byref(nl), C.byref(dep), C.byref(cost))
    t0 = time.time()
    eng.set_scene(sc)
    build_s = time.time() - t0
    row(f"glass mesh {nlat}x{nlon}", sc, sc.nprims, build_s, f"nodes {nn.value} depth {dep.value} sah {cost.value:.1f}")
