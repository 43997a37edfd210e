"""Seed-fixed synthetic workloads: photons spread over the
walls of the example room and deterministic primary-hit queries.  numpy only."""
import numpy as np

from ._capi import PHOTON_DTYPE

ROOM_LO = np.array([-2.0, 0.0, -6.0])
ROOM_HI = np.array([2.0, 4.0, 5.0])
SEED = 0x5EED0001


def wall_photons(n, seed=SEED, flux=5.0):
    """n photons uniform by area on the six walls of the room x[-2,2] y[0,4] z[-6,5];
    direction uniform on the hemisphere pointing INTO the wall; wavelength uniform.
    Returns (photons[PHOTON_DTYPE], power)."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    ext = ROOM_HI - ROOM_LO
    # walls: (axis, side) ; area = product of the other two extents
    areas = np.array([ext[1] * ext[2], ext[1] * ext[2], ext[0] * ext[2], ext[0] * ext[2], ext[0] * ext[1], ext[0] * ext[1]])
    wall = rng.choice(6, size=n, p=areas / areas.sum())
    axis = wall // 2
    side = wall % 2
    pos = ROOM_LO + rng.random((n, 3)) * ext
    pos[np.arange(n), axis] = np.where(side == 0, ROOM_LO[axis], ROOM_HI[axis])
    # uniform sphere direction, flipped to point out of the room (into the wall)
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    outward = np.where(side == 0, -1.0, 1.0)           # outward wall direction along `axis`
    comp = v[np.arange(n), axis]
    flip = np.sign(comp) != outward
    v[flip] = -v[flip]
    ph = np.zeros(n, PHOTON_DTYPE)
    ph["pos"] = pos
    ph["dir"] = v
    ph["wl"] = rng.integers(0, 3, size=n)
    return ph, flux / n


class ArrayScene:
    """A scene held as plain ctypes arrays (the argument shape of ppm_scene_set): what `Scene` exposes, without a
    parsed-file handle behind it."""

    def __init__(self, prims, mats, lights):
        self.prims, self.nprims = prims, len(prims)
        self.mats, self.nmats = mats, len(mats)
        self.lights, self.nlights = lights, (len(lights) if lights is not None else 0)

    def photon_budget(self, nphoton):
        import ctypes as C
        from ._capi import lib
        power = C.c_double()
        ns = (C.c_int64 * max(self.nlights, 1))()
        rc = lib.ppm_photon_budget(self.lights, self.nlights, int(nphoton), C.byref(power), ns)
        if rc:
            raise RuntimeError("photon budget")
        return power.value, [int(x) for x in ns[: self.nlights]]


def uv_sphere_triangles(center, radius, nlat, nlon):
    """Triangles (n, 3, 3) of a latitude/longitude sphere: 2 * nlon * (nlat - 1) of them."""
    c = np.asarray(center, np.float64)
    th = np.linspace(0.0, np.pi, nlat + 1)
    ph = np.linspace(0.0, 2.0 * np.pi, nlon + 1)[:-1]
    pts = np.empty((nlat + 1, nlon, 3))
    pts[..., 0] = np.sin(th)[:, None] * np.cos(ph)[None, :]
    pts[..., 1] = np.cos(th)[:, None] * np.ones(nlon)[None, :]
    pts[..., 2] = np.sin(th)[:, None] * np.sin(ph)[None, :]
    pts = c + radius * pts
    tris = []
    for i in range(nlat):
        for j in range(nlon):
            a, b = pts[i, j], pts[i, (j + 1) % nlon]
            d, e = pts[i + 1, j], pts[i + 1, (j + 1) % nlon]
            if i > 0:
                tris.append((a, d, b))
            if i < nlat - 1:
                tris.append((b, d, e))
    return np.array(tris)


def mesh_scene(base, triangles, material, spheres=()):
    """`base` (a Scene) plus one polygon per triangle (n, 3, 3) and one sphere per (center, radius), all with material
    index `material`: scenes beyond the 64-primitive limit of the brute-force hit test (the BVH path)."""
    from ._capi import Prim, Material, Light, D3, lib
    tri = np.ascontiguousarray(triangles, np.float64)
    n = base.nprims + len(tri) + len(spheres)
    prims = (Prim * n)()
    for i in range(base.nprims):
        prims[i] = base.prims[i]
    k = base.nprims
    for t in tri:
        rc = lib.ppm_prim_polygon(prims[k], D3(*t[0]), D3(*t[1]), D3(*t[2]), 0, int(material))
        if rc:
            raise ValueError("degenerate triangle")
        k += 1
    for cen, rad in spheres:
        lib.ppm_prim_sphere(prims[k], D3(*[float(x) for x in cen]), float(rad), int(material))
        k += 1
    mats = (Material * base.nmats)()
    for i in range(base.nmats):
        mats[i] = base.mats[i]
    lights = (Light * max(base.nlights, 1))()
    for i in range(base.nlights):
        lights[i] = base.lights[i]
    s = ArrayScene(prims, mats, lights)
    s.nlights = base.nlights
    return s


def mesh_scene_text(base_text, triangles, material_name):
    """The text of a .scene file: `base_text` (whose LAST section must be `object:`) plus
    one `polygon` object per triangle with inline vertices, written with repr() so that the doubles round-trip."""
    lines = [base_text.rstrip("\n")]
    for k, t in enumerate(np.asarray(triangles, np.float64)):
        lines.append(f"  - type: polygon\n    name: tri{k}\n    material: {material_name}")
        for i in range(3):
            lines.append(f"    pos{i + 1}: [ {float(t[i][0])!r}, {float(t[i][1])!r}, {float(t[i][2])!r} ]")
    return "\n".join(lines) + "\n"
