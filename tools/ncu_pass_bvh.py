"""Whole passes of a mesh scene (BVH mode) in stream mode, one lane, for ncu: a 261 k-triangle glass sphere in the
ex-glassbox room at 1920x1080, pass 600 of the schedule.  usage: python tools/ncu_pass_bvh.py [nlat nlon npasses]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppmpa_b200 as P
from ppmpa_b200 import synth
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
nlat = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nlon = int(sys.argv[2]) if len(sys.argv) > 2 else 512
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
eng = P.Engine(0)
eng.set_option("graph", 0); eng.set_option("lanes", 1)
base = P.read_scene(os.path.join(ROOT, "examples", "ex-glassbox.scene"))
eng.set_scene(synth.mesh_scene(base, synth.uv_sphere_triangles((0.3, 2.6, 1.0), 0.7, nlat, nlon), 4))
eng.set_camera(P.read_camera(os.path.join(ROOT, "examples", "camera0.scr"), xreso=1920, yreso=1080, progressive=1, pfilter=0))
radii = P.radius_schedule(0.1, 1000)
for i in range(n):
    eng.iteration(0x5EED0001, 600 + i, 1_000_000, float(radii[600 + i]) ** 2, True)
eng.close()
