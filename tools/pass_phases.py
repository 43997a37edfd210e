"""Prints per-pass phase times (CUDA events inside the engine) for config 2; diagnostic."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, ppmpa_b200 as P
eng = P.Engine(0)
sc, cam = bench.load_workload()
eng.set_scene(sc); eng.set_camera(cam)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 14
radii = P.radius_schedule(0.1, n)
for rep in range(2):
    for s in range(n):
        t = time.perf_counter()
        eng.iteration(bench.SEED, s, bench.NPHOTON, float(radii[s]) ** 2, True)
        wall = (time.perf_counter() - t) * 1e3
        ms, ct = eng.last_pass_stats()
        print(rep, s, f"wall {wall:7.2f} |", " ".join(f"{k}={v:6.2f}" for k, v in ms.items()), "| nodes", ct["gather_nodes"], "sumk", ct["sum_k"])
