"""CPU tests of the host-side BVH build (ppm_bvh_inspect: counts + the builder's self-check)."""
import ctypes as C
import os

import numpy as np
import pytest

import ppmpa_b200 as P
from ppmpa_b200 import _capi as K
from ppmpa_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "examples")


def inspect(sc):
    nn, nl, d, cost = C.c_int64(), C.c_int64(), C.c_int32(), C.c_double()
    rc = K.lib.ppm_bvh_inspect(sc.prims, sc.nprims, C.byref(nn), C.byref(nl), C.byref(d), C.byref(cost))
    return rc, nn.value, nl.value, d.value, cost.value


def bounded(sc):
    return sum(1 for i in range(sc.nprims) if sc.prims[i].type in (K.SHAPE_SPHERE, K.SHAPE_POLYGON, K.SHAPE_PARALLELOGRAM))


@pytest.mark.parametrize("name", [None, "ex-glassbox", "sample1", "ex-sunwindow", "mirror-ball", "coral-ball"])
def test_example_scenes_index(name):
    sc = P.read_scene(None if name is None else os.path.join(EX, name + ".scene"))
    rc, nn, nl, d, cost = inspect(sc)
    assert rc == 0 and nl == bounded(sc) and nn >= 1 and d < 64


@pytest.mark.parametrize("nlat,nlon", [(2, 3), (16, 32), (96, 192)])
def test_mesh_index(nlat, nlon):
    base = P.read_scene(os.path.join(EX, "ex-glassbox.scene"))
    tris = synth.uv_sphere_triangles((0.3, 2.6, 1.0), 0.7, nlat, nlon)
    sc = synth.mesh_scene(base, tris, 4, spheres=[((1.0, 0.5, 0.0), 0.3)])
    rc, nn, nl, d, cost = inspect(sc)
    assert rc == 0 and nl == len(tris) + 7 + 1
    assert nn <= nl and d <= 8 + 2 * int(np.ceil(np.log2(nl)))
    assert 0.0 < cost < 40.0                      # SAH cost in units of the root box's area: a usable tree


def test_degenerate_inputs():
    base = P.read_scene(os.path.join(EX, "ex-glassbox.scene"))
    t = np.array([[[0.0, 1.0, 1.0], [1.0, 1.0, 1.0], [0.0, 2.0, 1.5]]])
    sc = synth.mesh_scene(base, np.repeat(t, 1000, axis=0), 0)        # 1000 coincident triangles: median fallback
    rc, nn, nl, d, cost = inspect(sc)
    assert rc == 0 and nl == 1007 and d <= 20
    # planes only: nothing to index
    sc.nprims = 6
    rc, nn, nl, d, cost = inspect(sc)
    assert rc == 0 and nn == 0 and nl == 0
    # a non-finite primitive is refused, not indexed
    sc = synth.mesh_scene(base, t, 0)
    sc.prims[sc.nprims - 1].position[0] = float("inf")
    assert inspect(sc)[0] == -4                   # PPM_ERR_CAPACITY
    assert K.lib.ppm_bvh_inspect(None, 0, None, None, None, None) != 0


def test_mesh_scene_file_parses_to_the_same_primitives(tmp_path):
    """A 973-object scene FILE (the drop-in format) gives the same primitive array as building it through the ABI."""
    path = os.path.join(EX, "ex-glassbox.scene")
    base = P.read_scene(path)
    tris = synth.uv_sphere_triangles((0.3, 2.6, 1.0), 0.7, 16, 32)
    f = tmp_path / "mesh.scene"
    f.write_text(synth.mesh_scene_text(open(path).read(), tris, "glass"))
    parsed = P.read_scene(str(f))
    built = synth.mesh_scene(base, tris, 4)
    assert parsed.nprims == built.nprims == 973
    a = np.frombuffer(bytes(parsed.prims), np.uint8).reshape(parsed.nprims, -1)
    b = np.frombuffer(bytes(built.prims), np.uint8).reshape(built.nprims, -1)
    assert np.array_equal(a, b)
    assert inspect(parsed)[0] == 0
