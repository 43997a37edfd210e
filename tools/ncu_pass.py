"""A few whole passes in stream mode (one lane) for ncu: warm-up, then passes at a late point of the schedule."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppmpa_b200 as P
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = int(sys.argv[1]) if len(sys.argv) > 1 else 600
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = P.Engine(0)
eng.set_option("graph", 0); eng.set_option("lanes", 1)
eng.set_scene(P.read_scene(os.path.join(ROOT, "examples", "ex-glassbox.scene")))
eng.set_camera(P.read_camera(os.path.join(ROOT, "examples", "camera0.scr"), xreso=1920, yreso=1080, progressive=1, pfilter=0))
radii = P.radius_schedule(0.1, 1000)
for i in range(n):
    eng.iteration(0x5EED0001, g + i, 1_000_000, float(radii[g + i]) ** 2, True)
eng.close()
