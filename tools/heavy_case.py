import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import ppmpa_b200 as P
from ppmpa_b200.synth import wall_photons
ROOT='/root/repo'
eng = P.Engine(0)
eng.set_scene(P.read_scene(os.path.join(ROOT, "examples", "ex-glassbox.scene")))
cam = P.read_camera(os.path.join(ROOT, "examples", "camera0.scr"), xreso=1024, yreso=1024, blur=0, antialias=0, progressive=1)
eng.set_camera(cam)
rays = eng.generate_rays(1, 0)
hit, t, pos, nrm, io = eng.calc_intersection(rays)
q = torch.from_numpy(pos[hit >= 0]).cuda(); qn = torch.from_numpy(nrm[hit >= 0]).cuda()
out = torch.empty_like(q); cnt = torch.empty(len(q), dtype=torch.int32, device="cuda")
nph=int(sys.argv[1]); r=float(sys.argv[2])
ph, power = wall_photons(nph)
eng.import_photons(ph, power)
eng.build_photonmap(r*r)
for _ in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter()
    eng.estimate_radiance(q, qn, 0, out=out, counts=cnt, n=len(q))
    torch.cuda.synchronize(); print("ms", (time.perf_counter()-t0)*1e3)
