"""Whole-pass time and gather roofline along the radius schedule of BASELINE configs[4] (1920x1080, 1000 passes):
batches of passes around pass 0, 100, 300, 600 and 990.  The late passes have small radii (r -> 0.019, K ~ 6 photons
per query), where the fixed per-query / per-group cost of k_gather dominates its logical bytes.
usage (on a B200): [PPM_B200_LIB=variant.so] python tools/schedule_regimes.py [xres yres batch]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ppmpa_b200 as P
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
xres = int(sys.argv[1]) if len(sys.argv) > 1 else 1920
yres = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 10
peak = 6650.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except (OSError, ValueError, KeyError):
    pass
eng = P.Engine(0)
eng.set_scene(P.read_scene(os.path.join(ROOT, "examples", "ex-glassbox.scene")))
eng.set_camera(P.read_camera(os.path.join(ROOT, "examples", "camera0.scr"), xreso=xres, yreso=yres, progressive=1, pfilter=0))
radii = P.radius_schedule(0.1, 1000)
eng.iterate(0x5EED0001, 0, 4, 1_000_000, [float(r) ** 2 for r in radii[:4]], True)          # allocations, lanes
for first in (0, 100, 300, 600, 1000 - batch):
    r2 = [float(r) ** 2 for r in radii[first:first + batch]]
    eng.iterate(0x5EED0001, first, batch, 1_000_000, r2, True)                            # warm the cell tables for this radius
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.iterate(0x5EED0001, first, batch, 1_000_000, r2, True)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / batch * 1e3
    ms, ct = eng.last_pass_stats()                                                         # batch totals
    nodes, sumk = ct["gather_nodes"] / batch, ct["sum_k"] / batch
    gk = ms["gather_kernel"] / batch
    logical = (49.0 * sumk + 72.0 * nodes) / (gk * 1e-3) / 1e9
    print(f"passes {first:4d}..{first + batch - 1:4d}  r = {radii[first]:.4f}  {wall:6.3f} ms/pass | k_gather {gk:.3f} ms, K = {sumk / max(nodes, 1):5.1f} "
          f"per query, logical {logical:6.0f} GB/s = {logical / peak:.2f} of the HBM peak | direct light {ms['direct_light'] / batch:.3f} "
          f"expand {ms['eye_expand'] / batch:.3f} trace {ms['photon_trace'] / batch:.3f} build {ms['map_build'] / batch:.3f} (ms, overlapped lanes)")
eng.close()
