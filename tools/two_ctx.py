"""Upper-bound experiment for cross-pass pipelining: K passes on one context vs the same K passes
split over T contexts driven by T host threads on the SAME GPU."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, ppmpa_b200 as P
K = 40
radii = P.radius_schedule(0.1, K + 4)
def make():
    e = P.Engine(0); sc, cam = bench.load_workload(); e.set_scene(sc); e.set_camera(cam); e._keep = (sc, cam); return e
def run(e, ids):
    for i in ids:
        e.iteration(bench.SEED, i, bench.NPHOTON, float(radii[i]) ** 2, True)
for T in (1, 2, 3):
    engs = [make() for _ in range(T)]
    for e in engs: run(e, range(3))                      # warm-up
    t0 = time.perf_counter()
    th = [threading.Thread(target=run, args=(engs[t], range(4 + t, 4 + K, T))) for t in range(T)]
    for x in th: x.start()
    for x in th: x.join()
    dt = time.perf_counter() - t0
    print(f"contexts={T}: {K} passes in {dt*1e3:.1f} ms -> {dt*1e3/K:.3f} ms/pass", flush=True)
    for e in engs: e.close()
