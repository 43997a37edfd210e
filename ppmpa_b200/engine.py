"""Host-side mirror of the reference's render-path interface over the C ABI.

Names follow the reference (file:line under the reference tree):
  read_scene / read_camera           scene.rs:20, camera.rs:108
  Engine.calc_intersection           tracer.rs:306
  Engine.generate_photons            light.rs:67   (batched over a pass)
  Engine.trace_photons               tracer.rs:31  (batched over a pass)
  Engine.build_photonmap / read_map  photonmap.rs:23,31
  Engine.within                      kdtree.within, tracer.rs:180
  Engine.estimate_radiance           tracer.rs:179
  Engine.generate_rays               camera.rs:58
  Engine.trace_rays                  tracer.rs:129
  Engine.trace_rays_classic          tracer.rs:221 (rtc)
  Engine.iteration                   ppmpa.rs:74   (one whole pass)

Arrays are numpy (host) or anything exposing a CUDA device pointer via
`.data_ptr()` (torch CUDA tensors); device buffers are used in place.
Everything computes on the GPU; there is no CPU path.
"""
import ctypes as C

import numpy as np

from . import _capi as K
from ._capi import lib


class PPMError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__(f"{K.ERR_NAMES.get(code, code)}: {msg}")


class Scene:
    """(lights, objects) as read_scene returns them (scene.rs:443-447)."""

    def __init__(self, handle):
        self._h = handle
        self.nprims = lib.ppm_scene_nprims(handle)
        self.nmats = lib.ppm_scene_nmaterials(handle)
        self.nlights = lib.ppm_scene_nlights(handle)
        self.prims = (K.Prim * self.nprims).from_address(C.addressof(lib.ppm_scene_prims(handle).contents))
        self.mats = (K.Material * self.nmats).from_address(C.addressof(lib.ppm_scene_materials(handle).contents))
        self.lights = (K.Light * max(self.nlights, 1)).from_address(
            C.addressof(lib.ppm_scene_lights(handle).contents)) if self.nlights else (K.Light * 1)()

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.ppm_scene_free(h)

    def photon_budget(self, nphoton):
        """power and per-light photon counts, ppmpa.rs:30-31,70-72."""
        power = C.c_double()
        ns = (C.c_int64 * max(self.nlights, 1))()
        rc = lib.ppm_photon_budget(self.lights, self.nlights, int(nphoton), C.byref(power), ns)
        if rc:
            raise PPMError(rc, "photon budget")
        return power.value, [int(x) for x in ns[: self.nlights]]


def read_scene(path=None):
    """Parse an example/<name>.scene file; `None` returns the scene that the
    reference's read_scene hard-codes (scene.rs:20-448)."""
    h = C.c_void_p()
    if path is None:
        rc = lib.ppm_scene_builtin(C.byref(h))
        if rc:
            raise PPMError(rc, "builtin scene")
    else:
        err = C.create_string_buffer(512)
        rc = lib.ppm_scene_load(str(path).encode(), C.byref(h), err, 512)
        if rc:
            raise PPMError(rc, err.value.decode())
    return Scene(h)


def read_camera(path=None, **overrides):
    """Parse an example/<name>.scr file; `None` gives the defaults read_camera
    hard-codes (camera.rs:109-128).  Keyword overrides set configuration fields
    (e.g. xreso=1024, blur=0) and the derived basis is recomputed."""
    cam = K.Camera()
    if path is None:
        lib.ppm_camera_default(C.byref(cam))
    else:
        err = C.create_string_buffer(512)
        rc = lib.ppm_camera_load(str(path).encode(), C.byref(cam), err, 512)
        if rc:
            raise PPMError(rc, err.value.decode())
    if overrides:
        for k, v in overrides.items():
            if not hasattr(cam, k):
                raise AttributeError(k)
            if isinstance(v, (tuple, list, np.ndarray)):
                v = K.D3(*[float(x) for x in v])
            setattr(cam, k, v)
        rc = lib.ppm_camera_finalize(C.byref(cam))
        if rc:
            raise PPMError(rc, "degenerate camera")
    return cam


def radius_schedule(r0, npass):
    """util/iterator.rb:34-38."""
    out = (C.c_double * npass)()
    lib.ppm_radius_schedule(float(r0), int(npass), out)
    return np.array(out[:], dtype=np.float64)


def format_f64(v, exp_form=False):
    buf = C.create_string_buffer(400)
    rc = lib.ppm_format_f64(float(v), 1 if exp_form else 0, buf, 400)
    if rc:
        raise PPMError(rc, "format")
    return buf.value.decode()


def _ptr(a):
    """raw pointer of a numpy array / torch tensor / None"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous()
        return a.data_ptr()
    raise TypeError(type(a))


def _f64(a, shape_tail):
    if isinstance(a, np.ndarray):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape[1:] == shape_tail, (a.shape, shape_tail)
    return a


class Engine:
    """One engine per GPU (ppm_ctx)."""

    def __init__(self, device=0):
        h = C.c_void_p()
        rc = lib.ppm_create(int(device), C.byref(h))
        if rc:
            raise PPMError(rc, "ppm_create failed (ppmpa_b200 needs a CUDA device; there is no CPU fallback)")
        self._h = h
        self.scene = None
        self.camera = None

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.ppm_destroy(h)

    __del__ = close

    def _ck(self, rc):
        if rc:
            raise PPMError(rc, lib.ppm_last_error(self._h).decode())

    @property
    def stream(self):
        return lib.ppm_stream(self._h)

    # ---- switches ------------------------------------------------------------
    def set_option(self, name, value):
        """Engine switches of include/ppm.h: lanes, dl_cull, gather_heavy, dl_stats, graph."""
        self._ck(lib.ppm_option_set(self._h, name.encode(), int(value)))

    def get_option(self, name):
        v = C.c_int64()
        self._ck(lib.ppm_option_get(self._h, name.encode(), C.byref(v)))
        return v.value

    # ---- model -------------------------------------------------------------
    def set_scene(self, scene):
        self._ck(lib.ppm_scene_set(self._h, scene.prims, scene.nprims, scene.mats, scene.nmats, scene.lights, scene.nlights))
        self.scene = scene

    def set_camera(self, cam):
        self._ck(lib.ppm_camera_set(self._h, C.byref(cam)))
        self.camera = cam

    @property
    def npixels(self):
        return self.camera.xreso * self.camera.yreso

    # ---- probes --------------------------------------------------------------
    def calc_intersection(self, rays6):
        rays6 = _f64(rays6, (6,))
        n = rays6.shape[0]
        hit = np.empty(n, np.int32); t = np.empty(n); pos = np.empty((n, 3)); nrm = np.empty((n, 3)); io = np.empty(n, np.int32)
        self._ck(lib.ppm_intersect(self._h, _ptr(rays6), n, _ptr(hit), _ptr(t), _ptr(pos), _ptr(nrm), _ptr(io)))
        return hit, t, pos, nrm, io

    def generate_photons(self, seed, npass, n_per_light):
        ns = (C.c_int64 * len(n_per_light))(*n_per_light)
        out = np.zeros(int(sum(n_per_light)), K.PHOTON_DTYPE)
        self._ck(lib.ppm_emit_photons(self._h, seed, npass, ns, _ptr(out)))
        return out

    def trace_photons(self, seed, npass, uc, n_per_light, power):
        ns = (C.c_int64 * len(n_per_light))(*n_per_light)
        n = C.c_uint64()
        self._ck(lib.ppm_trace_photons(self._h, seed, npass, 1 if uc else 0, ns, float(power), C.byref(n)))
        return n.value

    def photon_count(self):
        n = C.c_uint64(); pw = C.c_double()
        self._ck(lib.ppm_photons_count(self._h, C.byref(n), C.byref(pw)))
        return n.value, pw.value

    def export_photons(self, with_tags=False):
        n, power = self.photon_count()
        out = np.zeros(n, K.PHOTON_DTYPE)
        tags = np.zeros(n, np.uint64) if with_tags else None
        self._ck(lib.ppm_photons_export(self._h, _ptr(out), n, _ptr(tags)))
        return (out, power, tags) if with_tags else (out, power)

    def import_photons(self, photons, power):
        if isinstance(photons, np.ndarray):
            assert photons.dtype == K.PHOTON_DTYPE
            photons = np.ascontiguousarray(photons)
            n = photons.shape[0]
        else:
            n = photons.numel() * photons.element_size() // 56
        self._ck(lib.ppm_photons_import(self._h, _ptr(photons), n, float(power)))

    def build_photonmap(self, radius2):
        self._ck(lib.ppm_map_build(self._h, float(radius2)))

    def read_map(self, path, radius2):
        """rt's read_map (photonmap.rs:31-74): load a `pm` dump, then build."""
        buf = C.c_void_p(); n = C.c_uint64(); pw = C.c_double()
        rc = lib.ppm_read_photon_dump(str(path).encode() if path else None, C.byref(buf), C.byref(n), C.byref(pw))
        if rc:
            raise PPMError(rc, f"reading {path}")
        try:
            self._ck(lib.ppm_photons_import(self._h, buf, n.value, pw.value))
        finally:
            lib.ppm_free(buf)
        self.build_photonmap(radius2)
        return n.value

    def within(self, q3, cap):
        q3 = _f64(q3, (3,))
        n = q3.shape[0]
        idx = np.zeros((n, cap), np.uint32); cnt = np.zeros(n, np.uint32)
        self._ck(lib.ppm_within(self._h, _ptr(q3), n, _ptr(idx), _ptr(cnt), cap))
        return idx, cnt

    def estimate_radiance(self, pos3, nrm3, pfilter=K.FILTER_NONE, out=None, counts=None, n=None):
        pos3 = _f64(pos3, (3,)); nrm3 = _f64(nrm3, (3,))
        if n is None:
            n = pos3.shape[0]
        if out is None:
            out = np.empty((n, 3))
            counts = np.empty(n, np.uint32)
        self._ck(lib.ppm_gather(self._h, _ptr(pos3), _ptr(nrm3), n, pfilter, _ptr(out), _ptr(counts)))
        return out, counts

    def estimate_radiance_knn(self, pos3, nrm3, k, pfilter=K.FILTER_NONE):
        """k-NN estimate (no reference implementation; semantics in include/ppm.h)."""
        pos3 = _f64(pos3, (3,)); nrm3 = _f64(nrm3, (3,))
        n = pos3.shape[0]
        out = np.empty((n, 3)); r2k = np.empty(n); counts = np.empty(n, np.uint32)
        self._ck(lib.ppm_gather_knn(self._h, _ptr(pos3), _ptr(nrm3), n, int(k), pfilter, _ptr(out), _ptr(r2k), _ptr(counts)))
        return out, r2k, counts

    def generate_rays(self, seed, npass):
        out = np.empty((self.npixels, 6))
        self._ck(lib.ppm_generate_rays(self._h, seed, npass, _ptr(out)))
        return out

    def trace_rays(self, rays6, seed, npass, uc, first_pixel=0):
        rays6 = _f64(rays6, (6,))
        n = rays6.shape[0]
        out = np.empty((n, 3))
        self._ck(lib.ppm_trace_rays(self._h, _ptr(rays6), n, first_pixel, seed, npass, 1 if uc else 0, _ptr(out)))
        return out

    def trace_rays_classic(self, rays6, seed, npass, first_pixel=0):
        """trace_ray_classic (tracer.rs:221), the `rtc` renderer: no photon map."""
        rays6 = _f64(rays6, (6,))
        n = rays6.shape[0]
        out = np.empty((n, 3))
        self._ck(lib.ppm_trace_rays_classic(self._h, _ptr(rays6), n, first_pixel, seed, npass, _ptr(out)))
        return out

    def direct_light(self, pos3, nrm3):
        """get_radiance_from_light summed over the lights (tracer.rs:136-141, 263-290) at surface points."""
        pos3 = _f64(pos3, (3,)); nrm3 = _f64(nrm3, (3,))
        n = pos3.shape[0]
        out = np.empty((n, 3))
        self._ck(lib.ppm_direct_light(self._h, _ptr(pos3), _ptr(nrm3), n, _ptr(out)))
        return out

    # ---- whole pass --------------------------------------------------------------
    def iteration(self, seed, npass, nphoton, radius2, uc=True):
        """One PPM-PA pass (ppmpa.rs:74-84), accumulated on the device."""
        self._ck(lib.ppm_render_pass(self._h, seed, npass, int(nphoton), float(radius2), 1 if uc else 0))

    def iterate(self, seed, first_pass, npass, nphoton, radius2_list, uc=True, pass_stride=1):
        """A batch of passes (the loop of util/iterator.rb:96-117) with cross-pass overlap."""
        r2 = (C.c_double * npass)(*[float(x) for x in radius2_list])
        self._ck(lib.ppm_render_passes(self._h, seed, int(first_pass), int(pass_stride), int(npass), int(nphoton), r2, 1 if uc else 0))

    def pass_image(self, out=None):
        if out is None:
            out = np.empty((self.npixels, 3))
        self._ck(lib.ppm_pass_image_read(self._h, _ptr(out)))
        return out

    def accum_reset(self):
        self._ck(lib.ppm_accum_reset(self._h))

    def accum_read(self):
        out = np.empty((self.npixels, 3)); n = C.c_uint32()
        self._ck(lib.ppm_accum_read(self._h, _ptr(out), C.byref(n)))
        return out, n.value

    def accum_add(self, rgb3, n_pass):
        """Adds a sum image (as accum_read returns it) and its pass count: resume / merge (averager2.rb:49-62)."""
        self._ck(lib.ppm_accum_add(self._h, _ptr(_f64(rgb3, (3,))), int(n_pass)))

    def accum_save(self, path):
        self._ck(lib.ppm_accum_save(self._h, str(path).encode()))

    def accum_load(self, path):
        n = C.c_uint32()
        self._ck(lib.ppm_accum_load(self._h, str(path).encode(), C.byref(n)))
        return n.value

    def accum_device(self):
        """(device pointer, number of doubles) of [sum image | pass count]."""
        p = C.c_void_p(); q = C.c_void_p(); n = C.c_uint64()
        self._ck(lib.ppm_accum_device(self._h, C.byref(p), C.byref(q), C.byref(n)))
        return p.value, n.value

    # ---- multi-GPU frame (one NCCL sum-reduce per frame, util/averager2.rb:49-62,86) ----
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL id made by rank 0; hand it to the other ranks by any means."""
        buf = C.create_string_buffer(128)
        rc = lib.ppm_comm_unique_id(buf)
        if rc:
            raise PPMError(rc, "ppm_comm_unique_id (is libnccl.so.2 loadable?)")
        return buf.raw

    def comm_init(self, nranks, rank, uid):
        assert len(uid) == 128
        buf = C.create_string_buffer(bytes(uid), 128)
        self._ck(lib.ppm_comm_init(self._h, int(nranks), int(rank), C.cast(buf, C.c_void_p)))

    def comm_destroy(self):
        self._ck(lib.ppm_comm_destroy(self._h))

    def accum_reduce(self, root=0, comm=None):
        """In-place sum of [sum image | pass count] over the ranks onto `root` (root < 0: all ranks)."""
        self._ck(lib.ppm_accum_reduce(self._h, comm, int(root)))

    def image_mean(self):
        out = np.empty((self.npixels, 3))
        self._ck(lib.ppm_image_mean(self._h, _ptr(out)))
        return out

    def last_pass_timeline(self):
        """Phase boundaries of the last pass, ms since its begin (include/ppm.h)."""
        t = (C.c_double * 16)()
        self._ck(lib.ppm_last_pass_timeline(self._h, t))
        names = ["begin", "trace_end", "build_end", "expand_begin", "expand_end", "classify_end", "qsort_end", "dl_begin", "dl_end",
                 "gather_begin", "gather_end", "combine_begin", "end"]
        return {k: t[i] for i, k in enumerate(names)}

    def last_pass_stats(self):
        ms = (C.c_double * 8)(); ct = (C.c_uint64 * 8)()
        self._ck(lib.ppm_last_pass_stats(self._h, ms, ct))
        names = ["photon_trace", "map_build", "eye_expand", "direct_light", "gather", "combine", "total", "gather_kernel"]
        cn = ["emitted", "stored", "eye_nodes", "gather_nodes", "sum_k", "launches", "candidates", "retried"]
        return {k: ms[i] for i, k in enumerate(names)}, {k: int(ct[i]) for i, k in enumerate(cn)}
