"""CPU check of the MATH behind the experimental per-warp shadow-ray classification (-DPPM_DL_REGION=1,
cull_region_bounded in ppmpa_b200/csrc/kernels_eye.cuh).  Test infrastructure: uses the oracle.

A numpy restatement of the region criteria (same formulas, same margins, tables as engine.cu:build_cull) classifies the
bounded primitives for clusters of 32 nearby nodes; every claim is then checked by brute force with the oracle's own
calc_intersection restricted to ONE primitive, for all 25 shadow rays of every node of the cluster:
  class (a)  "no candidate":            the primitive yields no root t >= NEARLY0 on any ray;
  class (b)  "only beyond the light":   every root has t >= ldist (1 - 1e-9)  (never an occluder).
It does not exercise the CUDA code (the GPU test test_direct_light_cull_is_exact[*-True] does); it guards the geometry:
translation argument for the cone test, side-plane distance margin for the pyramid test, coplanar emitter geometry.

  python tools/check_region_cull.py [clusters_per_scene]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib                      # noqa: E402
import ppmpa_b200 as P                 # noqa: E402
from ppmpa_b200 import _capi as K      # noqa: E402

EX = os.path.join(ROOT, "examples")
TS5 = np.array([0.1, 0.3, 0.5, 0.7, 0.9])


def v3(a):
    return np.array([a[0], a[1], a[2]], dtype=np.float64)


def quad_sphere(p0, d1, d2):
    s, d = d1 + d2, d1 - d2
    c = p0 + 0.5 * s
    rr = 0.5 * max(np.linalg.norm(s), np.linalg.norm(d))
    return c, rr * (1.0 + 1e-6) + 1e-6 * (1.0 + np.linalg.norm(c))


def build_tables(sc):
    """engine.cu:build_cull, bounded primitives and parallelogram lights only."""
    prims = []
    for o in range(sc.nprims):
        s = sc.prims[o]
        e = dict(kind=0)
        if s.type == K.SHAPE_SPHERE:
            c = v3(s.position)
            e = dict(kind=2, c=c, r=abs(s.scalar) * (1.0 + 1e-6) + 1e-6 * (1.0 + np.linalg.norm(c)), vtx=None)
        elif s.type in (K.SHAPE_POLYGON, K.SHAPE_PARALLELOGRAM):
            p0, d1, d2 = v3(s.position), v3(s.dir1), v3(s.dir2)
            c, r = quad_sphere(p0, d1, d2)
            e = dict(kind=2, c=c, r=r, vtx=np.array([p0, p0 + d1, p0 + d1 + d2, p0 + d2]))
        elif s.type == K.SHAPE_PLAIN:
            e = dict(kind=1)
        prims.append(e)
    lights = []
    for li in range(sc.nlights):
        l = sc.lights[li]
        if l.type != K.LIGHT_PARALLELOGRAM:
            lights.append(None)
            continue
        p0, d1, d2 = v3(l.pos), v3(l.dir1), v3(l.dir2)
        c, r = quad_sphere(p0, d1, d2)
        cx = np.cross(d1, d2)
        nl = cx / np.linalg.norm(cx)
        coplanar = 0
        for o in range(sc.nprims):
            s = sc.prims[o]
            if s.type not in (K.SHAPE_POLYGON, K.SHAPE_PARALLELOGRAM):
                continue
            q0, e1, e2 = v3(s.position), v3(s.dir1), v3(s.dir2)
            scale = 1.0 + np.linalg.norm(p0) + np.linalg.norm(q0) + np.linalg.norm(e1) + np.linalg.norm(e2)
            if all(abs(np.dot(nl, (q0 + (j & 1) * e1 + ((j >> 1) & 1) * e2) - p0)) <= 1e-12 * scale for j in range(4)):
                coplanar |= 1 << o
        lights.append(dict(c=c, r=r, nl=nl, corner=np.array([p0, p0 + d1, p0 + d1 + d2, p0 + d2]), coplanar=coplanar,
                           p0=p0, d1=d1, d2=d2))
    return prims, lights


def region_classify(prims, cl, c, rho):
    """cull_region_bounded: returns (keep, harmless) bit masks over the bounded primitives."""
    u = cl["c"] - c
    uu = float(np.dot(u, u))
    sane = uu < 1e6 and rho < 1e3
    rl = cl["r"] + rho
    ucone = uu - rl * rl
    cone = sane and ucone > 0.0
    L = np.sqrt(uu) + rl
    off_light_plane = sane and abs(float(np.dot(cl["nl"], u))) > rho + 1e-6 * (1.0 + L)
    a = cl["corner"] - c
    pn, pnn = [], []
    pyr_ok = off_light_plane
    for j in range(4):
        n = np.cross(a[j], a[(j + 1) & 3])
        nn = float(np.dot(n, n))
        s = float(np.dot(n, a[(j + 2) & 3]))
        pyr_ok = pyr_ok and nn > 1e-12 * (np.dot(a[j], a[j]) * np.dot(a[(j + 1) & 3], a[(j + 1) & 3])) and \
            s * s > 1e-12 * (nn * np.dot(a[(j + 2) & 3], a[(j + 2) & 3]))
        pn.append(-n if s < 0.0 else n)
        pnn.append(nn)
    rm = rho * (1.0 + 1e-6) + 1e-7
    rm2 = rm * rm
    keep_mask = harm_mask = 0
    for o, cp in enumerate(prims):
        keep = harmless = False
        if cp["kind"] == 2 and not sane:
            keep = True
        elif cp["kind"] == 2:
            if off_light_plane and (cl["coplanar"] >> o) & 1:
                harmless = True
            else:
                keep = True
                if cone:
                    v = cp["c"] - c
                    vv = float(np.dot(v, v))
                    R = cp["r"] + rho
                    vcone = vv - R * R
                    if vcone > 0.0 and vv < 1e12:
                        rhs = (np.sqrt(ucone * vcone) - rl * R) - 1e-7 * (uu + vv)
                        keep = not (float(np.dot(u, v)) < rhs)
                if keep and cp["vtx"] is not None and pyr_ok:
                    w = cp["vtx"] - c
                    ww = np.einsum("ij,ij->i", w, w)
                    out_any = False
                    if np.all(ww < 1e12):
                        for j in range(4):
                            d = w @ pn[j]
                            if np.all((d < 0.0) & (d * d > rm2 * pnn[j]) & (d * d > 1e-12 * (pnn[j] * ww))):
                                out_any = True
                    if out_any:
                        keep, harmless = False, True
        keep_mask |= int(keep) << o
        harm_mask |= int(harmless) << o
    return keep_mask, harm_mask


def single_prim_roots(orc_lib, sc, o, rays):
    """calc_intersection with primitive o alone: (has a root t >= NEARLY0, t)."""
    n = len(rays)
    hit = np.empty(n, np.int32); t = np.empty(n); pos = np.empty((n, 3)); nrm = np.empty((n, 3)); io = np.empty(n, np.int32)
    prim_ptr = C.cast(C.byref(sc.prims, o * C.sizeof(K.Prim)), C.POINTER(K.Prim))
    orc_lib.orc_intersect(prim_ptr, 1, sc.mats, sc.nmats, rays.ctypes.data, n, hit.ctypes.data, t.ctypes.data, pos.ctypes.data,
                          nrm.ctypes.data, io.ctypes.data)
    return hit >= 0, t


def check_scene(name, sc, orc, nclusters, rng):
    L = oracle_lib.load()
    prims, lights = build_tables(sc)
    rays = np.concatenate([rng.uniform([-1.9, 0.1, -5.9], [1.9, 3.9, 4.9], size=(nclusters, 3)), rng.normal(size=(nclusters, 3))], axis=1)
    rays[:, 3:] /= np.linalg.norm(rays[:, 3:], axis=1, keepdims=True)
    hit, t, pos, nrm, io = orc.intersect(sc, rays)
    centres = np.where((hit >= 0)[:, None] & (rng.random(nclusters) < 0.7)[:, None], pos, rays[:, :3])
    stats = dict(clusters=0, culled_a=0, culled_b=0, kept=0, rays=0)
    for ci in range(nclusters):
        spread = rng.choice([0.005, 0.02, 0.06, 0.15, 0.5])
        off = rng.normal(size=(32, 3))
        off *= (spread * rng.random(32) ** (1.0 / 3.0) / np.linalg.norm(off, axis=1))[:, None]
        if hit[ci] >= 0 and rng.random() < 0.6:                      # keep the cluster on the surface (as eye-path nodes are)
            off -= np.outer(off @ nrm[ci], nrm[ci])
        nodes = centres[ci] + off
        lo, hi = nodes.min(axis=0), nodes.max(axis=0)
        c = 0.5 * (lo + hi)
        h, h2 = hi - c, c - lo
        rho = np.sqrt(max(np.dot(h, h), np.dot(h2, h2))) * (1.0 + 1e-6) + 1e-9
        for cl in lights:
            if cl is None:
                continue
            keep, harm = region_classify(prims, cl, c, rho)
            gp = np.array([cl["p0"] + TS5[s // 5] * cl["d1"] + TS5[s % 5] * cl["d2"] for s in range(25)])
            d = gp[None, :, :] - nodes[:, None, :]
            ldist = np.linalg.norm(d, axis=2)
            ok = ldist > 0
            dirs = d / np.where(ok, ldist, 1.0)[:, :, None]
            r6 = np.ascontiguousarray(np.concatenate([np.repeat(nodes[:, None, :], 25, axis=1), dirs], axis=2).reshape(-1, 6))
            ld = ldist.reshape(-1)
            stats["clusters"] += 1
            stats["rays"] += len(r6)
            for o, cp in enumerate(prims):
                if cp["kind"] != 2:
                    continue
                if (keep >> o) & 1:
                    stats["kept"] += 1
                    continue
                has, tt = single_prim_roots(L, sc, o, r6)
                if (harm >> o) & 1:
                    stats["culled_b"] += 1
                    bad = has & ~(tt >= ld * (1.0 - 1e-9))
                    assert not bad.any(), (f"{name}: primitive {o} classified 'only beyond the light' but a ray hits it at t = "
                                           f"{tt[bad][0]} < ldist = {ld[bad][0]} (cluster {ci}, rho = {rho})")
                else:
                    stats["culled_a"] += 1
                    assert not has.any(), (f"{name}: primitive {o} classified 'no candidate' but {int(has.sum())} rays hit it "
                                           f"(cluster {ci}, rho = {rho})")
    return stats


def main():
    ncl = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    orc = oracle_lib.Oracle()
    rng = np.random.default_rng(20261017)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    scenes = [("builtin", P.read_scene())] + [(n, P.read_scene(os.path.join(EX, n + ".scene")))
                                              for n in ("ex-glassbox", "sample1", "mirror-ball", "coral-ball", "ex-sunwindow")]
    try:
        import tempfile
        from test_gpu_parity import ADVERSARIAL_SCENE
        with tempfile.NamedTemporaryFile("w", suffix=".scene", delete=False) as f:
            f.write(ADVERSARIAL_SCENE)
        scenes.append(("adversarial", P.read_scene(f.name)))
        os.unlink(f.name)
    except ImportError:
        pass
    for name, sc in scenes:
        st = check_scene(name, sc, orc, ncl, rng)
        print(f"{name:14s} {st['clusters']:5d} cluster x light regions, {st['rays']:8d} shadow rays: bounded primitives kept {st['kept']}, "
              f"culled 'no candidate' {st['culled_a']}, culled 'only beyond the light' {st['culled_b']} -- all claims hold")


if __name__ == "__main__":
    main()
