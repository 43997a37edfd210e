#!/usr/bin/env python
"""Generates tests/golden/*.npz / *.json: seeded input/output vectors of the hot path.

The reference is Rust and cannot be run here (no cargo/rustc), so the vectors come from the oracle
(oracle/ppm_oracle.cpp, the function-by-function restatement that is pinned by the reference's own
known answers, see reference_known_answers.json).  They freeze today's results: the CPU suite checks that
the oracle still reproduces them (compiler / refactoring drift), the GPU suite checks the CUDA path against
them (bit-exact tiers exactly, tolerance tiers to 1e-9).

usage: python tools/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402
import ppmpa_b200 as P  # noqa: E402
from ppmpa_b200 import _capi as K  # noqa: E402
from ppmpa_b200.synth import wall_photons  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
EX = os.path.join(ROOT, "examples")
SEED = 0x5EED0001


def main():
    os.makedirs(OUT, exist_ok=True)
    orc = oracle_lib.Oracle()
    vec = {}
    rng = np.random.default_rng(2024)
    for tag, path in (("builtin", None), ("glassbox", os.path.join(EX, "ex-glassbox.scene")), ("sample1", os.path.join(EX, "sample1.scene"))):
        sc = P.read_scene(path)
        pos = rng.uniform([-1.9, 0.1, -5.9], [1.9, 3.9, 4.9], size=(192, 3))
        d = rng.normal(size=(192, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        rays = np.concatenate([pos, d], axis=1)
        hit, t, hp, hn, io = orc.intersect(sc, rays)
        vec.update({f"{tag}_rays": rays, f"{tag}_hit": hit, f"{tag}_t": t, f"{tag}_pos": hp, f"{tag}_nvec": hn, f"{tag}_io": io})
        power, ns = sc.photon_budget(300)
        for uc in (0, 1):
            ph, tags = orc.trace_photons(sc, SEED, 5, bool(uc), ns)
            order = np.argsort(tags, kind="stable")
            vec[f"{tag}_photons_uc{uc}"] = ph[order]; vec[f"{tag}_tags_uc{uc}"] = tags[order]
        vec[f"{tag}_emit"] = orc.emit_photons(sc, SEED, 5, ns)
    # gather on a synthetic map, all filters
    ph, power = wall_photons(2500, seed=77)
    q = ph["pos"][rng.integers(0, len(ph), 96)] + rng.normal(scale=0.03, size=(96, 3))
    nrm = rng.normal(size=(96, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    vec["map_photons"] = ph; vec["map_power"] = np.array([power]); vec["map_r2"] = np.array([0.2 * 0.2])
    vec["map_q"] = q; vec["map_nrm"] = nrm
    m = orc.map_build(ph, power, 0.2 * 0.2)
    for f in (0, 1, 2):
        rad, cnt = m.gather(q, nrm, f)
        vec[f"map_rad_f{f}"] = rad; vec["map_cnt"] = cnt
    vec["map_within0"] = m.within(q[0])[0]
    # camera rays + eye paths on a tiny screen
    cam = P.read_camera(os.path.join(EX, "camera0.scr"), xreso=12, yreso=10, progressive=1, pfilter=K.FILTER_NONE)
    vec["cam_rays"] = orc.generate_rays(cam, SEED, 9)
    sc = P.read_scene(os.path.join(EX, "ex-glassbox.scene"))
    power, ns = sc.photon_budget(5000)
    eph, _ = orc.trace_photons(sc, SEED, 9, True, ns)
    em = orc.map_build(eph, power, 0.3 * 0.3)
    vec["eye_photons"] = eph; vec["eye_power"] = np.array([power])
    vec["eye_rad"], _ = orc.trace_rays(sc, em, K.FILTER_NONE, vec["cam_rays"], SEED, 9, True)
    vec["eye_rad_classic"] = orc.trace_rays_classic(sc, [0.001, 0.001, 0.001], vec["cam_rays"])
    np.savez_compressed(os.path.join(OUT, "oracle_vectors.npz"), **vec)
    # the reference's own still-valid known answers, one place, with citations
    kat = {
        "normalize(1,-2,3)": {"src": "src/ray/algebra.rs:268", "value": [0.2672612419124244, -0.5345224838248488, 0.8017837257372732]},
        "(1,2,3)*1.1": {"src": "src/ray/algebra.rs:275", "value": [1.1, 2.2, 3.3000000000000003]},
        "new_dir(1,1,1)": {"src": "src/ray/geometry.rs:217-218", "value": [0.5773502691896258] * 3},
        "Ray(1,1,1;-1,-1,-1).target(2)": {"src": "src/ray/geometry.rs:227", "value": [-0.15470053837925168] * 3},
        "polygon_normal((0,0,0),(2,1,0),(0,1,2))": {"src": "src/ray/geometry.rs:240-242", "value": [0.4082482904638631, -0.8164965809277261, 0.4082482904638631]},
        "filter_cone(0,0.01)": {"src": "src/tracer.rs:370", "value": 2.538461538461538},
        "filter_cone(0.01,0.01)": {"src": "src/tracer.rs:372", "value": 0.23076923076923078},
        "filter_gauss(0,0.01)": {"src": "src/tracer.rs:211 (code, CORR=0.5; the test at :374 pins the stale 0.355)", "value": 1.4180000000000001},
        "Color(0.4,0.78,1.0).normalize()": {"src": "src/ray/physics.rs:373", "value": [0.1834862385321101, 0.35779816513761464, 0.4587155963302752]},
        "check_under([0.1,0.2,0.3,0.5,0.8])": {"src": "src/ray/physics.rs:409-415", "value": {"0.03": 0, "0.12": 1, "0.28": 2, "0.4": 3, "0.64": 4, "0.99": 5}},
    }
    json.dump(kat, open(os.path.join(OUT, "reference_known_answers.json"), "w"), indent=1)
    print("wrote", os.path.join(OUT, "oracle_vectors.npz"), os.path.getsize(os.path.join(OUT, "oracle_vectors.npz")), "bytes")


if __name__ == "__main__":
    main()
