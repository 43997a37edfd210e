"""Timeline of one pass (phase boundaries in ms since the pass began) at a few points of the schedule; diagnostic.
usage: [PPM_LANES=1] python tools/timeline.py [xres yres]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppmpa_b200 as P
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
xres = int(sys.argv[1]) if len(sys.argv) > 1 else 1920
yres = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
eng = P.Engine(0)
eng.set_scene(P.read_scene(os.path.join(ROOT, "examples", "ex-glassbox.scene")))
eng.set_camera(P.read_camera(os.path.join(ROOT, "examples", "camera0.scr"), xreso=xres, yreso=yres, progressive=1, pfilter=0))
radii = P.radius_schedule(0.1, 1000)
for g in (0, 0, 300, 999):
    eng.iteration(0x5EED0001, g, 1_000_000, float(radii[g]) ** 2, True)
    t = eng.last_pass_timeline()
    ms, ct = eng.last_pass_stats()
    print(f"pass {g:4d} r={radii[g]:.4f} |", " ".join(f"{k}={v:.3f}" for k, v in t.items()), "| occ", ct["gather_nodes"], ct["stored"], ct["launches"])
eng.close()
