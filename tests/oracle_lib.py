"""ctypes loader for oracle/libppm_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs may import this.  The oracle is the CPU restatement of the reference
(see the header of oracle/ppm_oracle.cpp).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from ppmpa_b200 import _capi as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libppm_oracle.so")


def build_oracle(force=False):
    src = os.path.join(ORACLE_DIR, "ppm_oracle.cpp")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return ORACLE_SO


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    build_oracle()
    L = C.CDLL(ORACLE_SO)
    vp, i32, i64, u32, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double
    P = C.POINTER
    D3 = K.D3
    sig = {
        "orc_normalize": (C.c_int, [D3, D3]),
        "orc_cross": (None, [D3, D3, D3]),
        "orc_dot": (dbl, [D3, D3]),
        "orc_scale": (None, [D3, dbl, D3]),
        "orc_ray_target": (None, [D3, D3, dbl, D3]),
        "orc_new_polygon": (C.c_int, [D3, D3, D3, C.c_int, P(K.Prim)]),
        "orc_shape_normal": (C.c_int, [P(K.Prim), D3, D3]),
        "orc_filter_cone": (dbl, [dbl, dbl]),
        "orc_filter_gauss": (dbl, [dbl, dbl]),
        "orc_color_normalize": (None, [D3, D3]),
        "orc_decide_wavelength": (C.c_int, [D3, dbl]),
        "orc_check_under": (C.c_int, [P(dbl), C.c_int, dbl]),
        "orc_schlick": (dbl, [dbl, dbl]),
        "orc_density_pow": (dbl, [dbl]),
        "orc_relative_ior_average": (dbl, [D3, D3]),
        "orc_specular_refraction": (C.c_int, [D3, D3, dbl, D3, P(dbl)]),
        "orc_specular_reflection": (None, [D3, D3, D3, P(dbl)]),
        "orc_camera_finalize": (C.c_int, [P(K.Camera)]),
        "orc_philox_draws": (None, [u64, u32, u32, u64, u32, C.c_int, vp]),
        "orc_radius_schedule": (None, [dbl, C.c_int, vp]),
        "orc_radiance_to_rgb": (None, [dbl, D3, P(i32)]),
        "orc_averager_clip": (i32, [dbl, u32, dbl]),
        "orc_intersect": (None, [P(K.Prim), C.c_int, P(K.Material), C.c_int, vp, i64, vp, vp, vp, vp, vp]),
        "orc_emit_photons": (None, [P(K.Light), C.c_int, u64, u32, P(i64), vp]),
        "orc_trace_photons": (u64, [P(K.Prim), C.c_int, P(K.Material), C.c_int, P(K.Light), C.c_int, u64, u32, C.c_int,
                                    P(i64), vp, vp, u64]),
        "orc_trace_one_photon_seq": (u64, [P(K.Prim), C.c_int, P(K.Material), C.c_int, vp, C.c_int, vp, i64, vp, u64]),
        "orc_map_build": (vp, [vp, u64, dbl, dbl]),
        "orc_map_free": (None, [vp]),
        "orc_within": (u32, [vp, D3, C.c_int, vp, vp, u32]),
        "orc_gather": (None, [vp, vp, vp, i64, C.c_int, vp, vp, C.c_int]),
        "orc_gather_knn": (None, [vp, vp, vp, i64, u32, C.c_int, vp, vp, vp]),
        "orc_generate_rays": (None, [P(K.Camera), u64, u32, vp]),
        "orc_trace_rays": (None, [P(K.Prim), C.c_int, P(K.Material), C.c_int, P(K.Light), C.c_int, vp, C.c_int, vp, i64,
                                  i64, u64, u32, C.c_int, vp, C.c_int, vp]),
        "orc_trace_rays_classic": (None, [P(K.Prim), C.c_int, P(K.Material), C.c_int, P(K.Light), C.c_int, D3, vp, i64, vp]),
        "orc_direct_light": (None, [P(K.Prim), C.c_int, P(K.Material), C.c_int, P(K.Light), C.c_int, vp, vp, i64, vp]),
        "orc_render_pass": (C.c_int, [P(K.Prim), C.c_int, P(K.Material), C.c_int, P(K.Light), C.c_int, P(K.Camera), u64,
                                      u32, i64, dbl, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
        "orc_render_passes_parallel": (C.c_int, [P(K.Prim), C.c_int, P(K.Material), C.c_int, P(K.Light), C.c_int,
                                                 P(K.Camera), u64, u32, C.c_int, i64, vp, C.c_int, C.c_int, C.c_int, vp, vp]),
        "orc_render_passes_bands": (C.c_int, [P(K.Prim), C.c_int, P(K.Material), C.c_int, P(K.Light), C.c_int,
                                              P(K.Camera), u64, vp, C.c_int, i64, vp, C.c_int, vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def d3(v):
    return K.D3(*[float(x) for x in v])


def _p(a):
    return None if a is None else a.ctypes.data


class Oracle:
    """numpy-level convenience wrapper around the orc_* functions."""

    def __init__(self):
        self.L = load()

    # -- scene-level ---------------------------------------------------------
    def intersect(self, scene, rays6):
        rays6 = np.ascontiguousarray(rays6, np.float64)
        n = rays6.shape[0]
        hit = np.empty(n, np.int32); t = np.empty(n); pos = np.empty((n, 3)); nrm = np.empty((n, 3)); io = np.empty(n, np.int32)
        self.L.orc_intersect(scene.prims, scene.nprims, scene.mats, scene.nmats, _p(rays6), n, _p(hit), _p(t), _p(pos), _p(nrm), _p(io))
        return hit, t, pos, nrm, io

    def emit_photons(self, scene, seed, npass, n_per_light):
        ns = (C.c_int64 * len(n_per_light))(*n_per_light)
        out = np.zeros(int(sum(n_per_light)), K.PHOTON_DTYPE)
        self.L.orc_emit_photons(scene.lights, scene.nlights, seed, npass, ns, _p(out))
        return out

    def trace_photons(self, scene, seed, npass, uc, n_per_light):
        ns = (C.c_int64 * len(n_per_light))(*n_per_light)
        cap = int(sum(n_per_light)) * 10 + 1
        out = np.zeros(cap, K.PHOTON_DTYPE); tags = np.zeros(cap, np.uint64)
        n = self.L.orc_trace_photons(scene.prims, scene.nprims, scene.mats, scene.nmats, scene.lights, scene.nlights,
                                     seed, npass, 1 if uc else 0, ns, _p(out), _p(tags), cap)
        return out[:n].copy(), tags[:n].copy()

    def trace_one_photon_seq(self, scene, start, uc, draws):
        draws = np.ascontiguousarray(draws, np.float64)
        st = np.zeros(1, K.PHOTON_DTYPE); st[0] = start
        out = np.zeros(16, K.PHOTON_DTYPE)
        n = self.L.orc_trace_one_photon_seq(scene.prims, scene.nprims, scene.mats, scene.nmats, _p(st), 1 if uc else 0,
                                            _p(draws), len(draws), _p(out), 16)
        return out[:n].copy()

    # -- map -------------------------------------------------------------------
    def map_build(self, photons, power, radius2):
        photons = np.ascontiguousarray(photons)
        assert photons.dtype == K.PHOTON_DTYPE
        return OracleMap(self.L, self.L.orc_map_build(_p(photons), len(photons), power, radius2), len(photons))

    def generate_rays(self, cam, seed, npass):
        out = np.empty((cam.xreso * cam.yreso, 6))
        self.L.orc_generate_rays(C.byref(cam), seed, npass, _p(out))
        return out

    def trace_rays(self, scene, omap, pfilter, rays6, seed, npass, uc, first_pixel=0, nthreads=1):
        rays6 = np.ascontiguousarray(rays6, np.float64)
        n = rays6.shape[0]
        out = np.empty((n, 3)); stats = np.zeros(3, np.uint64)
        self.L.orc_trace_rays(scene.prims, scene.nprims, scene.mats, scene.nmats, scene.lights, scene.nlights, omap.h,
                              pfilter, _p(rays6), n, first_pixel, seed, npass, 1 if uc else 0, _p(out), nthreads, _p(stats))
        return out, stats

    def trace_rays_classic(self, scene, ambient, rays6):
        rays6 = np.ascontiguousarray(rays6, np.float64)
        out = np.empty((len(rays6), 3))
        self.L.orc_trace_rays_classic(scene.prims, scene.nprims, scene.mats, scene.nmats, scene.lights, scene.nlights,
                                      d3(ambient), _p(rays6), len(rays6), _p(out))
        return out

    def direct_light(self, scene, pos3, nrm3):
        pos3 = np.ascontiguousarray(pos3, np.float64); nrm3 = np.ascontiguousarray(nrm3, np.float64)
        out = np.empty_like(pos3)
        self.L.orc_direct_light(scene.prims, scene.nprims, scene.mats, scene.nmats, scene.lights, scene.nlights,
                                _p(pos3), _p(nrm3), len(pos3), _p(out))
        return out

    def render_pass(self, scene, cam, seed, npass, nphoton, radius2, uc, row0=0, row1=None):
        row1 = cam.yreso if row1 is None else row1
        out = np.empty(((row1 - row0) * cam.xreso, 3)); times = np.zeros(3); stats = np.zeros(4, np.uint64)
        self.L.orc_render_pass(scene.prims, scene.nprims, scene.mats, scene.nmats, scene.lights, scene.nlights,
                               C.byref(cam), seed, npass, nphoton, radius2, 1 if uc else 0, row0, row1, _p(out), _p(times), _p(stats))
        return out, times, stats

    def render_passes_parallel(self, scene, cam, seed, pass0, nthreads, nphoton, radius2_per_pass, uc, row0=0, row1=None):
        row1 = cam.yreso if row1 is None else row1
        r2 = np.ascontiguousarray(radius2_per_pass, np.float64)
        times = np.zeros((nthreads, 3)); stats = np.zeros((nthreads, 4), np.uint64)
        self.L.orc_render_passes_parallel(scene.prims, scene.nprims, scene.mats, scene.nmats, scene.lights, scene.nlights,
                                          C.byref(cam), seed, pass0, nthreads, nphoton, _p(r2), 1 if uc else 0, row0, row1,
                                          _p(times), _p(stats))
        return times, stats


    def render_passes_bands(self, scene, cam, seed, pass_ids, nphoton, radius2_per_pass, uc, row0s, row1s):
        """One single-threaded pass per entry, all concurrently; pass t traces image rows row0s[t]..row1s[t]."""
        n = len(pass_ids)
        ids = np.ascontiguousarray(pass_ids, np.uint32); r2 = np.ascontiguousarray(radius2_per_pass, np.float64)
        a = np.ascontiguousarray(row0s, np.int32); b = np.ascontiguousarray(row1s, np.int32)
        times = np.zeros((n, 3)); stats = np.zeros((n, 4), np.uint64)
        self.L.orc_render_passes_bands(scene.prims, scene.nprims, scene.mats, scene.nmats, scene.lights, scene.nlights,
                                       C.byref(cam), seed, _p(ids), n, nphoton, _p(r2), 1 if uc else 0, _p(a), _p(b),
                                       _p(times), _p(stats))
        return times, stats


class OracleMap:
    def __init__(self, L, h, n):
        self.L, self.h, self.n = L, h, n

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_map_free(self.h)
            self.h = None

    def within(self, q, brute=False, cap=None):
        cap = self.n if cap is None else cap
        idx = np.zeros(max(cap, 1), np.uint32); d2 = np.zeros(max(cap, 1))
        k = self.L.orc_within(self.h, d3(q), 1 if brute else 0, _p(idx), _p(d2), cap)
        return idx[:min(k, cap)].copy(), d2[:min(k, cap)].copy(), k

    def gather_knn(self, pos3, nrm3, k, pfilter):
        pos3 = np.ascontiguousarray(pos3, np.float64); nrm3 = np.ascontiguousarray(nrm3, np.float64)
        n = len(pos3)
        out = np.empty((n, 3)); r2k = np.empty(n); cnt = np.empty(n, np.uint32)
        self.L.orc_gather_knn(self.h, _p(pos3), _p(nrm3), n, k, pfilter, _p(out), _p(r2k), _p(cnt))
        return out, r2k, cnt

    def gather(self, pos3, nrm3, pfilter, nthreads=1):
        pos3 = np.ascontiguousarray(pos3, np.float64); nrm3 = np.ascontiguousarray(nrm3, np.float64)
        n = len(pos3)
        out = np.empty((n, 3)); cnt = np.empty(n, np.uint32)
        self.L.orc_gather(self.h, _p(pos3), _p(nrm3), n, pfilter, _p(out), _p(cnt), nthreads)
        return out, cnt
