"""Kernel-level gather measurements on the seed-fixed synthetic sets of SURVEY.md 8d:
N photons on the room's walls, queries = primary hits of camera0.scr (blur/AA off)."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ppmpa_b200 as P
from ppmpa_b200.synth import wall_photons
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
eng = P.Engine(0)
eng.set_scene(P.read_scene(os.path.join(ROOT, "examples", "ex-glassbox.scene")))
stream = torch.cuda.ExternalStream(eng.stream)
rows = []
for res in [(1024, 1024), (1920, 1080)]:
    cam = P.read_camera(os.path.join(ROOT, "examples", "camera0.scr"), xreso=res[0], yreso=res[1], blur=0, antialias=0, progressive=1)
    eng.set_camera(cam)
    rays = eng.generate_rays(1, 0)
    hit, t, pos, nrm, io = eng.calc_intersection(rays)
    q = torch.from_numpy(pos[hit >= 0]).cuda(); qn = torch.from_numpy(nrm[hit >= 0]).cuda()
    out = torch.empty_like(q); cnt = torch.empty(len(q), dtype=torch.int32, device="cuda")
    for nph in [1_000_000, 4_000_000, 16_000_000]:
        ph, power = wall_photons(nph)
        eng.import_photons(ph, power)
        for r in [0.1, 0.05, 0.025]:
            eng.build_photonmap(r * r)
            for _ in range(2):
                eng.estimate_radiance(q, qn, 0, out=out, counts=cnt, n=len(q))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(5):
                eng.estimate_radiance(q, qn, 0, out=out, counts=cnt, n=len(q))
            e1.record(stream); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            sumk = int(cnt.to(torch.int64).sum().item())
            b = 49.0 * sumk + 72.0 * len(q)
            rows.append(dict(res=f"{res[0]}x{res[1]}", photons=nph, r=r, queries=len(q), sum_k=sumk, kbar=sumk / len(q), ms_sort_plus_kernel=ms,
                             logical_gbs=b / ms / 1e6, frac_of_hbm_peak=b / ms / 1e6 / peak))
            print(json.dumps(rows[-1]), flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "gather_roofline.json"), "w"), indent=1)
