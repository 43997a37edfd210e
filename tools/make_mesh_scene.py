"""Writes a .scene file (SURVEY.md Appendix A.1 grammar) with a tessellated sphere of polygons inside the ex-glassbox room:
a scene beyond the 64-primitive limit of the brute-force hit test, i.e. one that runs through the BVH path.
usage: python tools/make_mesh_scene.py out.scene [nlat nlon [material]]      (default 64 x 128 = 16 128 triangles of glass)
then:  ppmpa_b200/bin/ppmpa 1000000 0.1 examples/camera0.scr out.scene > pass.ppmf"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ppmpa_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = sys.argv[1]
nlat = int(sys.argv[2]) if len(sys.argv) > 2 else 64
nlon = int(sys.argv[3]) if len(sys.argv) > 3 else 128
mat = sys.argv[4] if len(sys.argv) > 4 else "glass"
tris = synth.uv_sphere_triangles((0.3, 2.6, 1.0), 0.7, nlat, nlon)
with open(out, "w") as f:
    f.write(synth.mesh_scene_text(open(os.path.join(ROOT, "examples", "ex-glassbox.scene")).read(), tris, mat))
print(f"{out}: {len(tris)} triangles + the 13 primitives of ex-glassbox")
