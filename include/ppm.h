/*
 * ppm.h -- C ABI of the B200 progressive-photon-mapping hot path.
 *
 * This is the drop-in boundary for the render path of eijian/ppmpa
 * (photon tracing -> photon map -> radiance gather -> pass accumulation).
 * The reference has no FFI layer: the path sits behind Rust free functions
 * and three CLI text protocols.  Every entry point below cites the reference
 * function (file:line under the reference tree) it replaces.  All types are
 * plain C PODs (doubles, int32, pointers, sizes); nothing from C++/torch/Rust
 * crosses this boundary.
 *
 * Conventions
 *   - every function returns int: 0 = PPM_OK, <0 = error (ppm_last_error()).
 *   - never throws / aborts across the ABI.
 *   - one ppm_ctx per GPU, used from one host thread at a time; calls are
 *     synchronous with respect to the caller (the ctx stream is drained
 *     before a call that produces host-visible results returns).
 *   - "h_or_d" pointers may be host pointers OR device pointers on the ctx's
 *     GPU: the engine inspects them with cudaPointerGetAttributes and skips
 *     the staging copy for device memory.  Host buffers are borrowed for the
 *     duration of the call only.
 *   - all arithmetic is IEEE binary64, compiled without FMA contraction, in
 *     the operation order of the reference.
 */
#ifndef PPM_H_
#define PPM_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPM_ABI_VERSION 2

/* ---- error codes ------------------------------------------------------ */
enum {
  PPM_OK            =  0,
  PPM_ERR_ARG       = -1,   /* bad argument (null, negative size, bad enum) */
  PPM_ERR_STATE     = -2,   /* call order (e.g. gather before map build)    */
  PPM_ERR_CUDA      = -3,   /* CUDA runtime failure                         */
  PPM_ERR_CAPACITY  = -4,   /* a fixed capacity was exceeded                */
  PPM_ERR_IO        = -5,   /* file / stream failure                        */
  PPM_ERR_PARSE     = -6,   /* scene / camera / photon text parse failure   */
  PPM_ERR_NODEVICE  = -7    /* no CUDA device: there is NO CPU fallback     */
};

/* ---- model PODs (reference: all #[derive(Clone, Copy)] structs) ------- */

/* Shape, src/ray/geometry.rs:66-90 */
enum { PPM_SHAPE_POINT = 0, PPM_SHAPE_PLAIN = 1, PPM_SHAPE_SPHERE = 2,
       PPM_SHAPE_POLYGON = 3, PPM_SHAPE_PARALLELOGRAM = 4 };

/* Object{shape, material}, src/ray/object.rs:10-13.  112 bytes. */
typedef struct ppm_prim {
  int32_t type;          /* PPM_SHAPE_*                                     */
  int32_t material;      /* index into the material array                   */
  double  position[3];   /* Point.position | Sphere.center | Polygon.position */
  double  nvec[3];       /* Plain / Polygon / Parallelogram normal          */
  double  dir1[3];       /* Polygon / Parallelogram edge 1 (not normalised) */
  double  dir2[3];       /* Polygon / Parallelogram edge 2                  */
  double  scalar;        /* Plain.dist | Sphere.radius                      */
} ppm_prim;

/* Surface, src/ray/surface.rs:16-41 */
enum { PPM_SURF_NOTHING = 0, PPM_SURF_SIMPLE = 1, PPM_SURF_TS = 2,
       PPM_SURF_DISNEY = 3, PPM_SURF_BRADY = 4 };

/* Material{emittance,transmittance,ior,surface}, src/ray/material.rs:11-16 */
typedef struct ppm_material {
  double  emittance[3];
  double  transmittance[3];
  double  ior[3];
  int32_t surface;       /* PPM_SURF_*                                      */
  int32_t _pad;
  double  color_a[3];    /* Simple.reflectance   | TS.albedo_diff           */
  double  color_b[3];    /* Simple.specular_refl | TS.albedo_spec           */
  double  p0;            /* Simple.diffuseness   | TS.scatterness           */
  double  metalness;
  double  roughness;
  double  density_pow;   /* 1/(10^(5(1-sqrt(rough)))+1), surface.rs:52,63   */
  double  alpha;         /* TS: rough^4, surface.rs:64                      */
} ppm_material;

/* Light, src/ray/light.rs:16-39 */
enum { PPM_LIGHT_POINT = 0, PPM_LIGHT_PARALLELOGRAM = 1, PPM_LIGHT_SUN = 2 };

typedef struct ppm_light {
  int32_t type;
  int32_t _pad;
  double  color[3];      /* already normalised to sum 1 (physics.rs:61-71)  */
  double  flux;
  double  pos[3];
  double  nvec[3];
  double  dir1[3];
  double  dir2[3];
  double  dir[3];        /* SunLight only                                   */
} ppm_light;

/* PhotonFilter, src/ray/optics.rs:17-21 */
enum { PPM_FILTER_NONE = 0, PPM_FILTER_CONE = 1, PPM_FILTER_GAUSS = 2 };

/* Wavelength, src/ray/physics.rs:17-21 */
enum { PPM_WL_RED = 0, PPM_WL_GREEN = 1, PPM_WL_BLUE = 2 };

/* Camera, src/camera.rs:23-50.  The first block is configuration (the keys of
 * doc/ebnf-camera.txt / example/<name>.scr); the second block is derived by
 * ppm_camera_finalize() exactly as camera.rs:151-168 does. */
typedef struct ppm_camera {
  int32_t xreso, yreso;
  int32_t progressive, antialias, use_classic, blur;
  int32_t pfilter;       /* PPM_FILTER_*                                    */
  int32_t n_sample_photon; /* constant 500 in the reference (camera.rs:181), dead */
  double  radius;        /* estimate_radius, stored squared (camera.rs:186) */
  double  max_radiance, iso_sens, shut_speed;
  double  focal_len;     /* metres (file value is mm, camera.rs:142)        */
  double  f_number, focus;
  double  ambient[3];
  double  eye_pos[3], target_pos[3], upper_dir[3];
  /* derived */
  double  photon_power;
  double  eye_dir[3], origin[3], esx[3], esy[3], eex[3], eey[3];
} ppm_camera;

/* Photon{wl, ray{pos,dir}}, src/ray/optics.rs:168-171.  56 bytes, the Rust
 * in-memory size; 49 bytes of information. */
typedef struct ppm_photon {
  double  pos[3];
  double  dir[3];
  int32_t wl;            /* PPM_WL_*                                        */
  int32_t _pad;
} ppm_photon;

/* ---- host-side model constructors (pure CPU, no GPU needed) ----------- */

/* Surface::new_simple, surface.rs:45-54 (+ Material literal, scene.rs:33-44) */
void ppm_material_simple(ppm_material* m, const double emittance[3],
                         const double transmittance[3], const double ior[3],
                         const double reflectance[3], const double specular_refl[3],
                         double diffuseness, double metalness, double roughness);
/* Surface::new_ts, surface.rs:56-66 */
void ppm_material_ts(ppm_material* m, const double emittance[3],
                     const double transmittance[3], const double ior[3],
                     const double albedo_diff[3], const double albedo_spec[3],
                     double scatterness, double metalness, double roughness);
/* Shape::Plain literal; dist = -(normal . position) (scene files) */
void ppm_prim_plain(ppm_prim* p, const double normal[3], double dist, int32_t material);
void ppm_prim_sphere(ppm_prim* p, const double center[3], double radius, int32_t material);
/* Shape::new_polygon / new_parallelogram, geometry.rs:93-115.  Returns
 * PPM_ERR_ARG for a degenerate (zero-area) primitive where the reference panics. */
int  ppm_prim_polygon(ppm_prim* p, const double p0[3], const double p1[3],
                      const double p2[3], int parallelogram, int32_t material);
/* Color::normalize, physics.rs:61-71 */
void ppm_color_normalize(const double in[3], double out[3]);
/* camera defaults of read_camera, camera.rs:109-128, then finalize */
void ppm_camera_default(ppm_camera* c);
/* derived fields, camera.rs:151-168.  PPM_ERR_ARG on a degenerate basis. */
int  ppm_camera_finalize(ppm_camera* c);

/* ---- scene / camera files (the reference's read_scene / read_camera
 *      ignore their file argument, scene.rs:20, camera.rs:108; the file
 *      grammar is SURVEY.md Appendix A) --------------------------------- */
typedef struct ppm_scene ppm_scene;   /* host-side owner of the three arrays */

/* the hard-coded scene of read_scene, scene.rs:20-448 (BASELINE config 1) */
int  ppm_scene_builtin(ppm_scene** out);
/* parse example/<name>.scene */
int  ppm_scene_load(const char* path, ppm_scene** out, char* err, size_t errlen);
void ppm_scene_free(ppm_scene* s);
int32_t ppm_scene_nprims(const ppm_scene* s);
int32_t ppm_scene_nmaterials(const ppm_scene* s);
int32_t ppm_scene_nlights(const ppm_scene* s);
const ppm_prim*     ppm_scene_prims(const ppm_scene* s);
const ppm_material* ppm_scene_materials(const ppm_scene* s);
const ppm_light*    ppm_scene_lights(const ppm_scene* s);
/* parse example/<name>.scr (both key dialects); starts from ppm_camera_default */
int  ppm_camera_load(const char* path, ppm_camera* out, char* err, size_t errlen);

/* per-light photon counts: power = sum(flux)/nphoton; n_l = round(flux_l/power)
 * ppmpa.rs:30-31,70-72 / pm.rs:40-42,57-59 */
int  ppm_photon_budget(const ppm_light* lights, int32_t nlights, int64_t nphoton,
                       double* power, int64_t* n_per_light);
/* PPM-PA radius schedule of util/iterator.rb:34-38 (alpha = 0.5):
 * r[0] = r0, r[i+1] = sqrt(((i+1)+alpha)/((i+1)+1)) * r[i] */
void ppm_radius_schedule(double r0, int32_t npass, double* radius_out);
double ppm_radius_at(double r0, uint32_t pass);

/* ---- engine ------------------------------------------------------------ */
typedef struct ppm_ctx ppm_ctx;

int  ppm_abi_version(void);
int  ppm_create(int device, ppm_ctx** out);
void ppm_destroy(ppm_ctx* ctx);
const char* ppm_last_error(const ppm_ctx* ctx);   /* never NULL */
/* the CUDA stream the ctx launches on (a cudaStream_t), for event timing */
void* ppm_stream(ppm_ctx* ctx);
/* Engine switches (diagnostics and tests; none changes a result except where stated).  The environment
 * variables PPM_<NAME> are read ONCE, at ppm_create, as initial values.
 *   "lanes"         passes of one ppm_render_passes batch rendered concurrently on this GPU (default 2)
 *   "dl_cull"       1 = conservative shadow-ray culling in the classic direct light (bit-identical to 0)
 *   "gather_heavy"  1 = split very long candidate streams over many warps (changes only the summation order)
 *   "dl_stats"      1 = print culling statistics to stderr
 *   "graph"         1 = whole passes run as CUDA graphs without host round trips (default), 0 = stream mode
 *   "bvh"           1 = the NEXT ppm_scene_set indexes the bounded primitives with a bounding-volume hierarchy even
 *                   for a scene of <= 64 primitives (bit-identical hits; larger scenes always use it)
 * Unknown names return PPM_ERR_ARG. */
int  ppm_option_set(ppm_ctx* ctx, const char* name, int64_t value);
int  ppm_option_get(ppm_ctx* ctx, const char* name, int64_t* value);

/* replaces the (lgts, objs) pair read_scene returns, scene.rs:443-447.
 * Up to 64 primitives every ray tests every object, as calc_intersection does (tracer.rs:306-350).  Larger scenes
 * (at most 64 infinite planes, 2^26 bounded primitives, 48 materials, 8 lights) are indexed by a bounding-volume
 * hierarchy built here, on the host, once per scene; the traversal culls with padded boxes and tests the remaining
 * candidates with the same arithmetic and tie rule, so hit indices and distances are the brute-force scan's. */
int  ppm_scene_set(ppm_ctx* ctx, const ppm_prim* prims, int32_t nprims,
                   const ppm_material* mats, int32_t nmats,
                   const ppm_light* lights, int32_t nlights);
/* Host-only inspection of the hierarchy ppm_scene_set builds for a scene in BVH mode: node and leaf-primitive counts,
 * depth, SAH cost relative to the root box, and a self-check (PPM_ERR_STATE unless every bounded primitive sits in
 * exactly one leaf, strictly inside every box on its path, and the depth fits the traversal stack). */
int  ppm_bvh_inspect(const ppm_prim* prims, int32_t nprims, int64_t* n_nodes, int64_t* n_leaf_prims,
                     int32_t* depth, double* sah_cost);
/* replaces the Camera read_camera returns, camera.rs:176-204 */
int  ppm_camera_set(ppm_ctx* ctx, const ppm_camera* cam);

/* -- parity probe: calc_intersection, tracer.rs:306-350 -------------------
 * rays6[n][6] = pos, dir.  hit_idx = object index or -1; t = ray parameter;
 * pos3/nrm3 = Intersection.pos / .nvec (normal flipped to face the ray);
 * io = 0 In, 1 Out (flipped).  Any of t/pos3/nrm3/io may be NULL. */
int  ppm_intersect(ppm_ctx* ctx, const double* rays6_h_or_d, int64_t n,
                   int32_t* hit_idx, double* t, double* pos3, double* nrm3,
                   int32_t* io);

/* -- photon tracing: Light::generate_photon light.rs:67-91 +
 *    trace_photon tracer.rs:31-125 for every emitted photon of one pass
 *    (the loops of ppmpa.rs:74-92 / pm.rs:57-75).  Photon path i of pass
 *    `pass` draws from Philox4x32-10 stream (seed, pass, i) -- see DESIGN.md.
 *    Records stay on the device; *n_stored receives their count. */
int  ppm_trace_photons(ppm_ctx* ctx, uint64_t seed, uint32_t pass, int uc,
                       const int64_t* n_per_light, double power,
                       uint64_t* n_stored);
/* emitted photons only (Light::generate_photon), for parity tests */
int  ppm_emit_photons(ppm_ctx* ctx, uint64_t seed, uint32_t pass,
                      const int64_t* n_per_light, ppm_photon* out_h_or_d);

/* photon records <-> host, AoS f64 (the `pm` dump, pm.rs:62-75 <->
 * photonmap.rs:31-74).  tags (may be NULL) = (photon_index << 4) | depth. */
int  ppm_photons_count(ppm_ctx* ctx, uint64_t* n, double* power);
int  ppm_photons_export(ppm_ctx* ctx, ppm_photon* out, uint64_t cap,
                        uint64_t* tags);
int  ppm_photons_import(ppm_ctx* ctx, const ppm_photon* in_h_or_d, uint64_t n,
                        double power);

/* -- photon map: build_photonmap photonmap.rs:23-29 / read_map :31-74.
 *    radius2 is the SQUARED gather radius (ppmpa.rs:60-63). */
int  ppm_map_build(ppm_ctx* ctx, double radius2);

/* -- parity probe: kdtree.within(pos, r2, squared_euclidean), tracer.rs:180.
 *    For query q writes count[q] and up to `cap` photon indices (indices into
 *    the import/export order) at idx[q*cap ...], ascending by index. */
int  ppm_within(ppm_ctx* ctx, const double* q3_h_or_d, int64_t nq,
                uint32_t* idx, uint32_t* count, uint32_t cap);

/* -- estimate_radiance, tracer.rs:179-195 (+ filters :198-216,
 *    photon_to_radiance optics.rs:224-233).  rgb3[n][3]; counts may be NULL. */
int  ppm_gather(ppm_ctx* ctx, const double* pos3_h_or_d, const double* nrm3_h_or_d,
                int64_t n, int filter, double* rgb3_h_or_d, uint32_t* counts_h_or_d);

/* -- k-NN radiance estimate (config 4's sweep).  The reference only plumbs n_sample_photon
 *    (photonmap.rs:18, camera.rs:181) and never reads it, so there is NO reference
 *    implementation; semantics follow SURVEY.md 8c: the k nearest photons within r; if k are
 *    found, r_k^2 (the k-th smallest squared distance, returned in r2k, exact) replaces r^2 in
 *    the membership test (ties included), the filter and the 1/(pi r^2) normaliser; otherwise
 *    the fixed radius is used (r2k = r^2).  counts = photons used. */
int  ppm_gather_knn(ppm_ctx* ctx, const double* pos3_h_or_d, const double* nrm3_h_or_d,
                    int64_t n, uint32_t k, int filter, double* rgb3_h_or_d,
                    double* r2k_h_or_d, uint32_t* counts_h_or_d);

/* -- Camera::generate_ray, camera.rs:58-75 for every (y,x) of screen_map
 *    (row-major, y outer).  rays6[yreso*xreso][6]. */
int  ppm_generate_rays(ppm_ctx* ctx, uint64_t seed, uint32_t pass,
                       double* rays6_h_or_d);

/* -- trace_ray, tracer.rs:129-177, batched over rays; uses the current map.
 *    ray i draws from stream (seed, pass, pixel = first_pixel + i). */
int  ppm_trace_rays(ppm_ctx* ctx, const double* rays6_h_or_d, int64_t n,
                    int64_t first_pixel, uint64_t seed, uint32_t pass, int uc,
                    double* rgb3_h_or_d);

/* -- parity probe: get_radiance_from_light summed over the lights, tracer.rs:136-141,
 *    263-290 (`illuminated`, Light::get_direction light.rs:93-129, get_radiance :131-150):
 *    classic direct light at n surface points (position, facing normal). */
int  ppm_direct_light(ppm_ctx* ctx, const double* pos3_h_or_d, const double* nrm3_h_or_d,
                      int64_t n, double* rgb3_h_or_d);

/* -- trace_ray_classic, tracer.rs:221-259 (the `rtc` binary, rtc.rs:17-38): Whitted-style
 *    tracing without a photon map: di = classic direct light + camera ambient; mirror
 *    direction without the glossy lobe; Fresnel from cos1. */
int  ppm_trace_rays_classic(ppm_ctx* ctx, const double* rays6_h_or_d, int64_t n,
                            int64_t first_pixel, uint64_t seed, uint32_t pass,
                            double* rgb3_h_or_d);

/* -- one whole PPM-PA pass = `ppmpa` main, ppmpa.rs:21-46,74-84:
 *    trace photons, build map, generate + trace every eye ray, and add the
 *    pass image into the on-device accumulator (util/averager2.rb:49-62). */
int  ppm_render_pass(ppm_ctx* ctx, uint64_t seed, uint32_t pass, int64_t nphoton,
                     double radius2, int uc);
/* -- a batch of passes = the loop of util/iterator.rb:96-117 (pass i of the batch uses Philox
 *    pass id first_pass + i*pass_stride and radius2[i]).  Alternate passes run on two lanes of
 *    the same GPU so that phases of different passes overlap; the result (accumulator, last
 *    pass image) is the same as calling ppm_render_pass npass times, except that the
 *    accumulator is summed per lane first.  ppm_last_pass_stats then returns batch totals. */
int  ppm_render_passes(ppm_ctx* ctx, uint64_t seed, uint32_t first_pass, uint32_t pass_stride,
                       int32_t npass, int64_t nphoton, const double* radius2, int uc);
/* last pass image (what ppmpa / rt print), rgb3[yreso*xreso][3] */
int  ppm_pass_image_read(ppm_ctx* ctx, double* rgb3_h_or_d);
/* accumulator: sum over passes + number of passes summed */
int  ppm_accum_reset(ppm_ctx* ctx);
int  ppm_accum_read(ppm_ctx* ctx, double* rgb3_h_or_d, uint32_t* n_pass);
/* Checkpoint / resume of a frame in progress.  The reference's pass images are files, so a killed job keeps what it
 * rendered and util/averager2.rb:49-62 sums whatever is there; here the sums live on the device:
 *   ppm_accum_add   adds a sum image (as ppm_accum_read returns it) and its pass count to the accumulator
 *   ppm_accum_save  writes ["PPMACC1\n", xreso, yreso, n_pass, 0 (u32 each), 3*W*H f64] atomically (temp file + rename)
 *   ppm_accum_load  ADDS such a file to the accumulator (PPM_ERR_ARG if its resolution is not the camera's,
 *                   PPM_ERR_PARSE if it is not a checkpoint); n_pass (may be NULL) = the passes it held */
int  ppm_accum_add(ppm_ctx* ctx, const double* rgb3_h_or_d, uint32_t n_pass);
int  ppm_accum_save(ppm_ctx* ctx, const char* path);
int  ppm_accum_load(ppm_ctx* ctx, const char* path, uint32_t* n_pass);
/* raw device pointers of the accumulator (3*W*H doubles) and pass counter
 * (1 double, so both reduce in one dtype).  The pointers stay valid until the
 * camera resolution changes or the ctx is destroyed. */
int  ppm_accum_device(ppm_ctx* ctx, void** sum_dev, void** npass_dev, uint64_t* n_doubles);
/* mean image: sum / n_pass (averager2.rb:86).  PPM_ERR_STATE if no pass has been accumulated. */
int  ppm_image_mean(ppm_ctx* ctx, double* rgb3_h_or_d);

/* -- multi-GPU frame: the passes of a frame are independent (util/iterator.rb:90-117 runs them as
 *    separate processes) and the frame is their plain sum (util/averager2.rb:49-62,86).  Every rank
 *    renders its own pass ids into its own accumulator; ONE sum-reduce of [3*W*H sums | pass count]
 *    (f64) per frame combines them.  The collective is NCCL, loaded at run time (libnccl.so.2, the
 *    copy already mapped into the process if there is one); no other exchange exists on this path.
 *    ppm_comm_unique_id: rank 0 makes the 128-byte NCCL id and hands it to the other ranks by the
 *    host's own means (file, pipe, MPI, torch.distributed ...).
 *    ppm_comm_init: ncclCommInitRank on the ctx's GPU; the communicator is owned by the ctx.
 *    ppm_accum_reduce: in-place sum of the accumulator onto `root` (root < 0: onto every rank) on
 *    the ctx stream, over `nccl_comm` (an ncclComm_t made by the caller) or, when NULL, over the
 *    ctx's own communicator.  Returns after the reduce has completed. */
int  ppm_comm_unique_id(void* id128);
int  ppm_comm_init(ppm_ctx* ctx, int32_t nranks, int32_t rank, const void* id128);
int  ppm_comm_destroy(ppm_ctx* ctx);
int  ppm_accum_reduce(ppm_ctx* ctx, void* nccl_comm, int32_t root);

/* per-phase device times (ms) of the last ppm_render_pass, or batch totals over the passes of the last
 * ppm_render_passes (phases of different passes overlap: the sum of [6] exceeds the wall time of a batch).  Taken
 * from %globaltimer stamps the pass writes at its phase boundaries on the device, in stream order; with the "graph"
 * switch off, [7] comes from CUDA events recorded around k_gather instead.
 * [0] photon trace, [1] map build, [2] eye expand, [3] direct light, [4] gather (query sort + kernels),
 * [5] combine+accumulate, [6] whole pass, [7] k_gather (+ heavy parts) alone;
 * counters: [0] emitted, [1] stored records, [2] eye nodes, [3] gather nodes,
 * [4] sum of K (photons within r over all gather nodes), [5] kernel launches,
 * [6] candidate distance tests of k_gather, [7] passes rendered again after a buffer overflow */
int  ppm_last_pass_stats(ppm_ctx* ctx, double ms[8], uint64_t counters[8]);
/* phase boundaries of the LAST pass of the last call, ms since the pass began (0 = not reached):
 * [0] begin, [1] photon trace end, [2] map build end, [3] eye expand begin, [4] eye expand end, [5] shadow-ray
 * classification end, [6] query sort end, [7] direct light begin, [8] direct light end, [9] gather begin,
 * [10] gather end, [11] combine begin, [12] end */
int  ppm_last_pass_timeline(ppm_ctx* ctx, double ms_since_begin[16]);

/* ---- output formats (host) -------------------------------------------- */
/* Rust `{}` / `{:e}` f64 formatting (shortest round-trip digits) */
int  ppm_format_f64(double v, int exp_form, char* buf, size_t buflen);
/* Camera::radiance_to_rgb, camera.rs:92-100 */
void ppm_radiance_to_rgb(double max_radiance, const double rad[3], int32_t rgb[3]);
/* photon dump of pm.rs:44-45,65-74 */
int  ppm_write_photon_dump(const char* path /* NULL = stdout */, int64_t nphoton,
                           double power, const ppm_photon* ph, uint64_t n);
/* read_map, photonmap.rs:31-74 (directions are re-normalised, geometry.rs:46-54).
 * Caller frees *out with ppm_free. */
int  ppm_read_photon_dump(const char* path /* NULL = stdin */, ppm_photon** out,
                          uint64_t* n, double* power);
void ppm_free(void* p);
/* image of ppmpa.rs:40-45 / rt.rs:48-59: 5 header lines (camera.rs:77-90) then
 * `{:e} {:e} {:e}` per pixel if progressive else tone-mapped ints */
int  ppm_write_image(const char* path /* NULL = stdout */, const ppm_camera* cam,
                     const double* rgb3, int progressive);
/* averager2.rb:84-110 (P3, gamma 1/2.2) and :154-218 (float32 OpenEXR) applied
 * to a SUM image over n_pass passes */
int  ppm_write_mean_ppm(const char* path, const ppm_camera* cam, const double* sum_rgb3,
                        uint32_t n_pass);
int  ppm_write_mean_exr(const char* path, const ppm_camera* cam, const double* sum_rgb3,
                        uint32_t n_pass);

#ifdef __cplusplus
}
#endif
#endif /* PPM_H_ */
