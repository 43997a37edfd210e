"""How much of k_gather is per-query overhead?  The north-star pass at a late radius with fewer and fewer photons: the
queries (5.4 M gather nodes) stay the same, the candidates vanish.  usage: python tools/gather_floor.py [radius]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppmpa_b200 as P
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
r = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0189
eng = P.Engine(0)
eng.set_option("lanes", 1)
eng.set_scene(P.read_scene(os.path.join(ROOT, "examples", "ex-glassbox.scene")))
eng.set_camera(P.read_camera(os.path.join(ROOT, "examples", "camera0.scr"), xreso=1920, yreso=1080, progressive=1, pfilter=0))
for nph in (1_000_000, 300_000, 100_000, 10_000, 1_000):
    for i in range(3):
        eng.iteration(0x5EED0001, 900 + i, nph, r * r, True)
    ms, ct = eng.last_pass_stats()
    print(f"photons {nph:>8}  stored {ct['stored']:>7}  queries {ct['gather_nodes']}  K = {ct['sum_k'] / ct['gather_nodes']:.3f}  "
          f"candidates/query {ct['candidates'] / ct['gather_nodes']:.2f}  k_gather {ms['gather_kernel']:.3f} ms  (pass {ms['total']:.2f} ms)", flush=True)
