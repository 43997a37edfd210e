"""ctypes binding of include/ppm.h (libppm_b200.so).

There is no CPU fallback: if the shared library is missing this module raises
at import time, and `Engine()` raises if no CUDA device is present.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PPM_B200_LIB selects another build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("PPM_B200_LIB") or os.path.join(_HERE, "libppm_b200.so")

PPM_OK = 0
ERR_NAMES = {0: "PPM_OK", -1: "PPM_ERR_ARG", -2: "PPM_ERR_STATE", -3: "PPM_ERR_CUDA", -4: "PPM_ERR_CAPACITY",
             -5: "PPM_ERR_IO", -6: "PPM_ERR_PARSE", -7: "PPM_ERR_NODEVICE"}

SHAPE_POINT, SHAPE_PLAIN, SHAPE_SPHERE, SHAPE_POLYGON, SHAPE_PARALLELOGRAM = range(5)
SURF_NOTHING, SURF_SIMPLE, SURF_TS, SURF_DISNEY, SURF_BRADY = range(5)
LIGHT_POINT, LIGHT_PARALLELOGRAM, LIGHT_SUN = range(3)
FILTER_NONE, FILTER_CONE, FILTER_GAUSS = range(3)
WL_RED, WL_GREEN, WL_BLUE = range(3)

D3 = C.c_double * 3


class Prim(C.Structure):
    _fields_ = [("type", C.c_int32), ("material", C.c_int32), ("position", D3), ("nvec", D3), ("dir1", D3),
                ("dir2", D3), ("scalar", C.c_double)]


class Material(C.Structure):
    _fields_ = [("emittance", D3), ("transmittance", D3), ("ior", D3), ("surface", C.c_int32), ("_pad", C.c_int32),
                ("color_a", D3), ("color_b", D3), ("p0", C.c_double), ("metalness", C.c_double),
                ("roughness", C.c_double), ("density_pow", C.c_double), ("alpha", C.c_double)]


class Light(C.Structure):
    _fields_ = [("type", C.c_int32), ("_pad", C.c_int32), ("color", D3), ("flux", C.c_double), ("pos", D3),
                ("nvec", D3), ("dir1", D3), ("dir2", D3), ("dir", D3)]


class Camera(C.Structure):
    _fields_ = [("xreso", C.c_int32), ("yreso", C.c_int32), ("progressive", C.c_int32), ("antialias", C.c_int32),
                ("use_classic", C.c_int32), ("blur", C.c_int32), ("pfilter", C.c_int32), ("n_sample_photon", C.c_int32),
                ("radius", C.c_double), ("max_radiance", C.c_double), ("iso_sens", C.c_double),
                ("shut_speed", C.c_double), ("focal_len", C.c_double), ("f_number", C.c_double), ("focus", C.c_double),
                ("ambient", D3), ("eye_pos", D3), ("target_pos", D3), ("upper_dir", D3),
                ("photon_power", C.c_double), ("eye_dir", D3), ("origin", D3), ("esx", D3), ("esy", D3), ("eex", D3),
                ("eey", D3)]


class Photon(C.Structure):
    _fields_ = [("pos", D3), ("dir", D3), ("wl", C.c_int32), ("_pad", C.c_int32)]


assert C.sizeof(Prim) == 112 and C.sizeof(Photon) == 56

# numpy view of ppm_photon[]
import numpy as np  # noqa: E402

PHOTON_DTYPE = np.dtype([("pos", "<f8", 3), ("dir", "<f8", 3), ("wl", "<i4"), ("_pad", "<i4")])
assert PHOTON_DTYPE.itemsize == 56

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C ppmpa_b200/csrc`). ppmpa_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

vp, i32, i64, u32, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double
P = C.POINTER

# every symbol include/ppm.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "ppm_abi_version": (C.c_int, []),
    "ppm_material_simple": (None, [P(Material), D3, D3, D3, D3, D3, dbl, dbl, dbl]),
    "ppm_material_ts": (None, [P(Material), D3, D3, D3, D3, D3, dbl, dbl, dbl]),
    "ppm_prim_plain": (None, [P(Prim), D3, dbl, i32]),
    "ppm_prim_sphere": (None, [P(Prim), D3, dbl, i32]),
    "ppm_prim_polygon": (C.c_int, [P(Prim), D3, D3, D3, C.c_int, i32]),
    "ppm_color_normalize": (None, [D3, D3]),
    "ppm_camera_default": (None, [P(Camera)]),
    "ppm_camera_finalize": (C.c_int, [P(Camera)]),
    "ppm_scene_builtin": (C.c_int, [P(vp)]),
    "ppm_scene_load": (C.c_int, [C.c_char_p, P(vp), C.c_char_p, C.c_size_t]),
    "ppm_scene_free": (None, [vp]),
    "ppm_scene_nprims": (i32, [vp]),
    "ppm_scene_nmaterials": (i32, [vp]),
    "ppm_scene_nlights": (i32, [vp]),
    "ppm_scene_prims": (P(Prim), [vp]),
    "ppm_scene_materials": (P(Material), [vp]),
    "ppm_scene_lights": (P(Light), [vp]),
    "ppm_camera_load": (C.c_int, [C.c_char_p, P(Camera), C.c_char_p, C.c_size_t]),
    "ppm_photon_budget": (C.c_int, [P(Light), i32, i64, P(dbl), P(i64)]),
    "ppm_radius_schedule": (None, [dbl, i32, P(dbl)]),
    "ppm_radius_at": (dbl, [dbl, u32]),
    "ppm_create": (C.c_int, [C.c_int, P(vp)]),
    "ppm_destroy": (None, [vp]),
    "ppm_last_error": (C.c_char_p, [vp]),
    "ppm_stream": (vp, [vp]),
    "ppm_option_set": (C.c_int, [vp, C.c_char_p, i64]),
    "ppm_option_get": (C.c_int, [vp, C.c_char_p, P(i64)]),
    "ppm_scene_set": (C.c_int, [vp, P(Prim), i32, P(Material), i32, P(Light), i32]),
    "ppm_bvh_inspect": (C.c_int, [P(Prim), i32, P(C.c_int64), P(C.c_int64), P(i32), P(dbl)]),
    "ppm_camera_set": (C.c_int, [vp, P(Camera)]),
    "ppm_intersect": (C.c_int, [vp, vp, i64, vp, vp, vp, vp, vp]),
    "ppm_trace_photons": (C.c_int, [vp, u64, u32, C.c_int, P(i64), dbl, P(u64)]),
    "ppm_emit_photons": (C.c_int, [vp, u64, u32, P(i64), vp]),
    "ppm_photons_count": (C.c_int, [vp, P(u64), P(dbl)]),
    "ppm_photons_export": (C.c_int, [vp, vp, u64, vp]),
    "ppm_photons_import": (C.c_int, [vp, vp, u64, dbl]),
    "ppm_map_build": (C.c_int, [vp, dbl]),
    "ppm_within": (C.c_int, [vp, vp, i64, vp, vp, u32]),
    "ppm_gather": (C.c_int, [vp, vp, vp, i64, C.c_int, vp, vp]),
    "ppm_gather_knn": (C.c_int, [vp, vp, vp, i64, u32, C.c_int, vp, vp, vp]),
    "ppm_generate_rays": (C.c_int, [vp, u64, u32, vp]),
    "ppm_trace_rays": (C.c_int, [vp, vp, i64, i64, u64, u32, C.c_int, vp]),
    "ppm_trace_rays_classic": (C.c_int, [vp, vp, i64, i64, u64, u32, vp]),
    "ppm_direct_light": (C.c_int, [vp, vp, vp, i64, vp]),
    "ppm_render_pass": (C.c_int, [vp, u64, u32, i64, dbl, C.c_int]),
    "ppm_render_passes": (C.c_int, [vp, u64, u32, u32, i32, i64, P(dbl), C.c_int]),
    "ppm_pass_image_read": (C.c_int, [vp, vp]),
    "ppm_accum_reset": (C.c_int, [vp]),
    "ppm_accum_read": (C.c_int, [vp, vp, P(u32)]),
    "ppm_accum_add": (C.c_int, [vp, vp, u32]),
    "ppm_accum_save": (C.c_int, [vp, C.c_char_p]),
    "ppm_accum_load": (C.c_int, [vp, C.c_char_p, P(u32)]),
    "ppm_accum_device": (C.c_int, [vp, P(vp), P(vp), P(u64)]),
    "ppm_image_mean": (C.c_int, [vp, vp]),
    "ppm_comm_unique_id": (C.c_int, [vp]),
    "ppm_comm_init": (C.c_int, [vp, i32, i32, vp]),
    "ppm_comm_destroy": (C.c_int, [vp]),
    "ppm_accum_reduce": (C.c_int, [vp, vp, i32]),
    "ppm_last_pass_timeline": (C.c_int, [vp, P(dbl)]),
    "ppm_last_pass_stats": (C.c_int, [vp, P(dbl), P(u64)]),
    "ppm_format_f64": (C.c_int, [dbl, C.c_int, C.c_char_p, C.c_size_t]),
    "ppm_radiance_to_rgb": (None, [dbl, D3, P(i32)]),
    "ppm_write_photon_dump": (C.c_int, [C.c_char_p, i64, dbl, vp, u64]),
    "ppm_read_photon_dump": (C.c_int, [C.c_char_p, P(vp), P(u64), P(dbl)]),
    "ppm_free": (None, [vp]),
    "ppm_write_image": (C.c_int, [C.c_char_p, P(Camera), vp, C.c_int]),
    "ppm_write_mean_ppm": (C.c_int, [C.c_char_p, P(Camera), vp, u32]),
    "ppm_write_mean_exr": (C.c_int, [C.c_char_p, P(Camera), vp, u32]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)     # AttributeError here = the .so does not export a declared symbol
    _f.restype = _res
    _f.argtypes = _args
