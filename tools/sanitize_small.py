"""Small end-to-end run for compute-sanitizer: every kernel once (pass, batch with two lanes, probes, k-NN, classic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ppmpa_b200 as P
from ppmpa_b200 import _capi as K
from ppmpa_b200.synth import wall_photons
EX = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples")
eng = P.Engine(0)
for scene in (None, os.path.join(EX, "ex-glassbox.scene")):
    sc = P.read_scene(scene)
    cam = P.read_camera(None, xreso=24, yreso=24)
    eng.set_scene(sc); eng.set_camera(cam)
    eng.accum_reset()
    eng.iteration(1, 0, 3000, 0.3 ** 2, True)
    eng.iterate(1, 1, 3, 3000, [0.09, 0.08, 0.07], uc=False)
    img = eng.image_mean()
    rays = eng.generate_rays(1, 0)
    eng.calc_intersection(rays)
    eng.trace_rays_classic(rays, 1, 0)
ph, power = wall_photons(3000, seed=1)
eng.import_photons(ph, power); eng.build_photonmap(0.04)
q = ph["pos"][:200] + 0.01
nrm = np.tile([0.0, 1.0, 0.0], (200, 1))
for f in (0, 1, 2):
    eng.estimate_radiance(q, nrm, f)
eng.estimate_radiance_knn(q, nrm, 20, 1)
eng.within(q, 64)
eng.set_scene(P.read_scene(os.path.join(EX, "sample1.scene")))
eng.direct_light(q, nrm)                              # culled shadow rays (spheres, planes, emitter quad)
eng.set_option("dl_cull", 0)
eng.direct_light(q, nrm)
eng.set_option("dl_cull", 1)
# heavy gather groups: 6000 photons in one cell neighbourhood -> parts + k_gather_heavy, all modes
rng = np.random.default_rng(3)
hp = np.zeros(6000, K.PHOTON_DTYPE)
hp["pos"][:, 0] = rng.uniform(-0.05, 0.05, 6000); hp["pos"][:, 2] = rng.uniform(-0.05, 0.05, 6000)
hp["dir"][:, 1] = -1.0
eng.import_photons(hp, 1e-3); eng.build_photonmap(0.04)
hq = np.zeros((100, 3)); hq[:, 0] = rng.uniform(-0.1, 0.1, 100)
g, cnt = eng.estimate_radiance(hq, nrm[:100], 2)
assert cnt.max() > 4000
eng.estimate_radiance_knn(hq, nrm[:100], 50, 0)
# BVH mode: a 221-triangle mesh + spheres (hierarchy, shaft classification, any-hit shadow rays), then the forced mode
# on a small scene, then back; checkpoint add
from ppmpa_b200 import synth
base = P.read_scene(os.path.join(EX, "ex-glassbox.scene"))
mesh = synth.mesh_scene(base, synth.uv_sphere_triangles((0.3, 2.6, 1.0), 0.7, 8, 16), 4, spheres=[((-1.0, 0.5, 1.0), 0.4)])
eng.set_scene(mesh)
eng.accum_reset()
eng.iteration(1, 0, 3000, 0.3 ** 2, True)
eng.iterate(1, 1, 2, 3000, [0.09, 0.08], uc=False)
eng.calc_intersection(rays)
eng.trace_rays_classic(rays, 1, 0)
eng.direct_light(q, nrm)
eng.set_option("bvh", 1)
eng.set_scene(base)
eng.iteration(1, 5, 3000, 0.2 ** 2, True)
eng.set_option("bvh", 0)
eng.set_scene(base)
acc, n = eng.accum_read()
eng.accum_add(acc, n)
print("sanitize run ok", float(img.sum()))
eng.close()
