"""One frame over 2 GPUs through the C ABI alone (no torch.distributed): rank r renders passes r, r+2, ... of a 6-pass
job on GPU r, ppm_accum_reduce (NCCL inside libppm_b200.so) sums the accumulators onto rank 0, and rank 0 also renders
all 6 passes on its own for comparison.  usage: python tools/two_rank_frame.py out.npz   (needs 2 GPUs)"""
import multiprocessing as mp
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NP, NPH, SEED, WORLD = 6, 50000, 0x5EED0001, 2


def rank_main(rank, uid_q, out_path):
    import numpy as np
    import ppmpa_b200 as P
    from ppmpa_b200.parallel import passes_for_rank
    sc = P.read_scene(os.path.join(ROOT, "examples", "ex-glassbox.scene"))
    cam = P.read_camera(os.path.join(ROOT, "examples", "camera0.scr"), xreso=96, yreso=64, pfilter=P.FILTER_NONE, progressive=1)
    radii = P.radius_schedule(0.15, NP)
    eng = P.Engine(rank)
    eng.set_scene(sc); eng.set_camera(cam)
    if rank == 0:
        uid = P.Engine.comm_unique_id()
        for _ in range(WORLD - 1):
            uid_q.put(uid)
    else:
        uid = uid_q.get(timeout=120)
    eng.comm_init(WORLD, rank, uid)
    mine = passes_for_rank(NP, WORLD, rank)
    eng.accum_reset()
    eng.iterate(SEED, mine[0], len(mine), NPH, [radii[p] ** 2 for p in mine], uc=True, pass_stride=WORLD)
    eng.accum_reduce(root=0)
    if rank == 0:
        reduced, n_red = eng.accum_read()
        single = P.Engine(0)
        single.set_scene(sc); single.set_camera(cam)
        single.iterate(SEED, 0, NP, NPH, radii ** 2, uc=True)
        one, n_one = single.accum_read()
        single.close()
        np.savez(out_path, reduced=reduced, single=one, n_reduced=n_red, n_single=n_one)
    eng.comm_destroy()
    eng.close()


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else "two_rank_frame.npz"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=rank_main, args=(r, q, out)) for r in range(WORLD)]
    for p in ps:
        p.start()
    rc = 0
    for p in ps:
        p.join(240)
        if p.is_alive():
            p.terminate(); rc = 1
        rc = rc or (p.exitcode or 0)
    sys.exit(rc)
