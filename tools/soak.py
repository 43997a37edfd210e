"""Long-run check of the BASELINE job sizes on one GPU: configs[4] (1920x1080, 1000 passes x 1M) and configs[2]
(ex-sunwindow, 500 passes x 1M): wall time, pass counter, finiteness, GPU memory before/after (no growth)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ppmpa_b200 as P
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
eng = P.Engine(0)
for scene, res, npass, uc in [("ex-glassbox", (1920, 1080), 1000, True), ("ex-sunwindow", (1024, 1024), 500, False)]:
    eng.set_scene(P.read_scene(os.path.join(ROOT, "examples", scene + ".scene")))
    eng.set_camera(P.read_camera(os.path.join(ROOT, "examples", "camera0.scr"), xreso=res[0], yreso=res[1], progressive=1, pfilter=0))
    eng.accum_reset()
    radii = P.radius_schedule(0.1, npass)
    eng.iterate(0x5EED0001, 0, 8, 1_000_000, [float(r) ** 2 for r in radii[:8]], uc)      # warm-up / allocations
    eng.accum_reset()
    torch.cuda.synchronize()
    free0, total = torch.cuda.mem_get_info()
    t0 = time.perf_counter()
    eng.iterate(0x5EED0001, 0, npass, 1_000_000, [float(r) ** 2 for r in radii], uc)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    free1, _ = torch.cuda.mem_get_info()
    acc, n = eng.accum_read()
    mean = acc / n
    print(f"{scene} {res[0]}x{res[1]} uc={uc}: {npass} passes x 1M photons in {dt:.3f} s ({dt / npass * 1e3:.3f} ms/pass), pass counter {n}, "
          f"finite {bool(np.isfinite(acc).all())}, mean radiance {mean.mean():.6e}, max {mean.max():.4e}, "
          f"GPU memory in use {(total - free1) / 2**30:.2f} GiB (change during the run {(free0 - free1) / 2**20:+.1f} MiB)")
eng.close()
