"""Config 5 smoke: ex-glassbox at 1920x1080, 1M photons, a few passes; prints phases and memory."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ppmpa_b200 as P
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
eng = P.Engine(0)
eng.set_scene(P.read_scene(os.path.join(ROOT, "examples", "ex-glassbox.scene")))
eng.set_camera(P.read_camera(os.path.join(ROOT, "examples", "camera0.scr"), xreso=1920, yreso=1080, progressive=1, pfilter=0))
radii = P.radius_schedule(0.1, 8)
for i in range(8):
    eng.iteration(0x5EED0001, i, 1_000_000, float(radii[i]) ** 2, True)
    ms, ct = eng.last_pass_stats()
    print(i, " ".join(f"{k}={v:.2f}" for k, v in ms.items()), ct["gather_nodes"], ct["sum_k"])
acc, n = eng.accum_read()
free, total = torch.cuda.mem_get_info()
print("passes", n, "mean radiance", acc.mean() / n, "finite", bool(np.isfinite(acc).all()), "GPU memory used GB", (total - free) / 2**30)
for scene, uc in [("ex-sunwindow", False), ("mirror-ball", True), ("coral-ball", True), ("sample1", True)]:
    eng.set_scene(P.read_scene(os.path.join(ROOT, "examples", scene + ".scene")))
    eng.accum_reset()
    for i in range(3):
        eng.iteration(0x5EED0001, i, 1_000_000, float(radii[i]) ** 2, uc)
    ms, ct = eng.last_pass_stats()
    img = eng.pass_image()
    print(scene, "uc" if uc else "nc", f"total={ms['total']:.2f} ms", "nodes", ct["gather_nodes"], "stored", ct["stored"], "mean", img.mean(), "finite", bool(np.isfinite(img).all()))
