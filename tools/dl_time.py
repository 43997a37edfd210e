"""Times k_dl_classify + k_direct_light through the ppm_direct_light probe on the nodes of a config-2 frame; diagnostic."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, ppmpa_b200 as P
eng = P.Engine(0)
sc, cam = bench.load_workload()
eng.set_scene(sc); eng.set_camera(cam)
rays = eng.generate_rays(1, 0)
hit, t, pos, nrm, io = eng.calc_intersection(rays)
ok = hit >= 0
q = torch.from_numpy(np.ascontiguousarray(pos[ok])).cuda(); qn = torch.from_numpy(np.ascontiguousarray(nrm[ok])).cuda()
out = torch.empty_like(q)
stream = torch.cuda.ExternalStream(eng.stream)
from ppmpa_b200._capi import lib
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream)
    for _ in range(10):
        lib.ppm_direct_light(eng._h, q.data_ptr(), qn.data_ptr(), len(q), out.data_ptr())
    e1.record(stream); torch.cuda.synchronize()
print("primary-hit nodes", len(q), "classify + direct light ms", e0.elapsed_time(e1) / 10)
