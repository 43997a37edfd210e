"""Per-pass phase times and gather statistics for any example scene at a given resolution; diagnostic.
usage: python tools/scene_phases.py <scene> <uc 0|1> [xres yres npasses]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppmpa_b200 as P
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
scene, uc = sys.argv[1], bool(int(sys.argv[2]))
xres = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
yres = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
n = int(sys.argv[5]) if len(sys.argv) > 5 else 6
eng = P.Engine(0)
eng.set_scene(P.read_scene(os.path.join(ROOT, "examples", scene + ".scene")))
eng.set_camera(P.read_camera(os.path.join(ROOT, "examples", "camera0.scr"), xreso=xres, yreso=yres, progressive=1, pfilter=0))
radii = P.radius_schedule(0.1, n)
for i in range(n):
    eng.iteration(0x5EED0001, i, 1_000_000, float(radii[i]) ** 2, uc)
    ms, ct = eng.last_pass_stats()
    gb = (49.0 * ct["sum_k"] + 72.0 * ct["gather_nodes"]) / 1e9
    print(i, " ".join(f"{k}={v:.2f}" for k, v in ms.items()), "| stored", ct["stored"], "nodes", ct["gather_nodes"], "sum_k", ct["sum_k"],
          f"kbar {ct['sum_k'] / max(ct['gather_nodes'], 1):.1f} gather logical {gb / (ms['gather_kernel'] / 1e3):.0f} GB/s")
