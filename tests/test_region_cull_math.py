"""CPU check of the geometry behind the experimental per-warp shadow-ray classification (-DPPM_DL_REGION=1, DESIGN.md
section 11): tools/check_region_cull.py restates the region criteria in numpy and verifies every 'no candidate' /
'only beyond the light' claim by brute force with the oracle's calc_intersection, one primitive at a time."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import ppmpa_b200 as P  # noqa: E402


@pytest.mark.parametrize("name", [None, "ex-glassbox", "sample1"])
def test_region_claims_hold(oracle, name):
    import check_region_cull as R
    sc = P.read_scene(None if name is None else os.path.join(R.EX, name + ".scene"))
    st = R.check_scene(str(name), sc, oracle, 120, np.random.default_rng(7))
    assert st["culled_a"] > 0 and st["culled_b"] > 0 and st["kept"] > 0


def test_check_is_sensitive(oracle, monkeypatch):
    """Ignoring the spread of the nodes (rho = 0) must be caught."""
    import check_region_cull as R
    orig = R.region_classify
    monkeypatch.setattr(R, "region_classify", lambda prims, cl, c, rho: orig(prims, cl, c, 0.0))
    sc = P.read_scene(os.path.join(R.EX, "ex-glassbox.scene"))
    with pytest.raises(AssertionError):
        R.check_scene("ex-glassbox", sc, oracle, 1500, np.random.default_rng(1))
