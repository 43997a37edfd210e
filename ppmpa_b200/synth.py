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
