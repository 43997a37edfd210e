// host_model.cpp -- host-side model constructors of the C ABI (include/ppm.h).
//
// Pure CPU code: builds the POD scene / camera descriptions the engine uploads.
// Mirrors the reference constructors (file:line cited per function) in their
// f64 operation order; compiled with -ffp-contract=off.
#include "host_common.h"

#include <cmath>
#include <cstring>

namespace ppmhost {

static inline void set3(double o[3], double x, double y, double z) { o[0] = x; o[1] = y; o[2] = z; }
static inline void cp3(double o[3], const double i[3]) { o[0] = i[0]; o[1] = i[1]; o[2] = i[2]; }
double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
void cross3(const double a[3], const double b[3], double o[3]) {
  // algebra.rs:174-180 component order
  double x = a[1] * b[2] - b[1] * a[2];
  double y = a[2] * b[0] - b[2] * a[0];
  double z = a[0] * b[1] - b[0] * a[1];
  set3(o, x, y, z);
}
bool normalize3(const double a[3], double o[3]) {
  // algebra.rs:151-158: multiply by the reciprocal norm
  double n = std::sqrt(dot3(a, a));
  if (n == 0.0) return false;
  double inv = 1.0 / n;
  set3(o, a[0] * inv, a[1] * inv, a[2] * inv);
  return true;
}

static double density_pow(double rough) {
  return 1.0 / (std::pow(10.0, 5.0 * (1.0 - std::sqrt(rough))) + 1.0);
}

}  // namespace ppmhost

using namespace ppmhost;

extern "C" {

int ppm_abi_version(void) { return PPM_ABI_VERSION; }

void ppm_material_simple(ppm_material* m, const double emittance[3], const double transmittance[3],
                         const double ior[3], const double reflectance[3], const double specular_refl[3],
                         double diffuseness, double metalness, double roughness) {
  std::memset(m, 0, sizeof *m);
  cp3(m->emittance, emittance); cp3(m->transmittance, transmittance); cp3(m->ior, ior);
  m->surface = PPM_SURF_SIMPLE;
  cp3(m->color_a, reflectance); cp3(m->color_b, specular_refl);
  m->p0 = diffuseness; m->metalness = metalness; m->roughness = roughness;
  m->density_pow = density_pow(roughness);
  m->alpha = 0.0;
}

void ppm_material_ts(ppm_material* m, const double emittance[3], const double transmittance[3],
                     const double ior[3], const double albedo_diff[3], const double albedo_spec[3],
                     double scatterness, double metalness, double roughness) {
  std::memset(m, 0, sizeof *m);
  cp3(m->emittance, emittance); cp3(m->transmittance, transmittance); cp3(m->ior, ior);
  m->surface = PPM_SURF_TS;
  cp3(m->color_a, albedo_diff); cp3(m->color_b, albedo_spec);
  m->p0 = scatterness; m->metalness = metalness; m->roughness = roughness;
  m->density_pow = density_pow(roughness);
  m->alpha = roughness * roughness * roughness * roughness;
}

void ppm_prim_plain(ppm_prim* p, const double normal[3], double dist, int32_t material) {
  std::memset(p, 0, sizeof *p);
  p->type = PPM_SHAPE_PLAIN; p->material = material;
  cp3(p->nvec, normal); p->scalar = dist;
}

void ppm_prim_sphere(ppm_prim* p, const double center[3], double radius, int32_t material) {
  std::memset(p, 0, sizeof *p);
  p->type = PPM_SHAPE_SPHERE; p->material = material;
  cp3(p->position, center); p->scalar = radius;
}

int ppm_prim_polygon(ppm_prim* p, const double p0[3], const double p1[3], const double p2[3],
                     int parallelogram, int32_t material) {
  std::memset(p, 0, sizeof *p);
  p->type = parallelogram ? PPM_SHAPE_PARALLELOGRAM : PPM_SHAPE_POLYGON;
  p->material = material;
  double d1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
  double d2[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
  double c[3], n[3];
  cross3(d1, d2, c);
  if (!normalize3(c, n)) return PPM_ERR_ARG;
  cp3(p->position, p0); cp3(p->nvec, n); cp3(p->dir1, d1); cp3(p->dir2, d2);
  return PPM_OK;
}

void ppm_color_normalize(const double in[3], double out[3]) {
  double r1 = in[0] < 0.0 ? 0.0 : in[0], g1 = in[1] < 0.0 ? 0.0 : in[1], b1 = in[2] < 0.0 ? 0.0 : in[2];
  double mag = r1 + g1 + b1;
  if (mag == 0.0) { out[0] = out[1] = out[2] = 1.0 / 3.0; }
  else { out[0] = r1 / mag; out[1] = g1 / mag; out[2] = b1 / mag; }
}

void ppm_camera_default(ppm_camera* c) {
  // the config literals of read_camera, camera.rs:109-128
  std::memset(c, 0, sizeof *c);
  c->xreso = 256; c->yreso = 256;
  c->progressive = 1; c->antialias = 1; c->use_classic = 1; c->blur = 1;
  c->pfilter = PPM_FILTER_NONE;
  c->n_sample_photon = 500;
  c->radius = 0.2 * 0.2;
  c->max_radiance = 0.01; c->iso_sens = 100.0; c->shut_speed = 0.004;
  c->focal_len = 50.0 / 1000.0; c->f_number = 4.0; c->focus = 7.0;
  set3(c->ambient, 0.0, 0.0, 0.0);
  set3(c->eye_pos, 1.0, 2.0, -4.5);
  set3(c->target_pos, 0.0, 1.0, 0.0);
  set3(c->upper_dir, 0.0, 1.0, 0.0);
  ppm_camera_finalize(c);
}

int ppm_camera_finalize(ppm_camera* c) {
  if (!c || c->xreso <= 0 || c->yreso <= 0 || c->focal_len == 0.0 || c->f_number == 0.0) return PPM_ERR_ARG;
  double d[3] = {c->target_pos[0] - c->eye_pos[0], c->target_pos[1] - c->eye_pos[1], c->target_pos[2] - c->eye_pos[2]};
  double ez[3], ex[3], ey[3], tmp[3];
  if (!normalize3(d, ez)) return PPM_ERR_ARG;
  cross3(c->upper_dir, ez, tmp);
  if (!normalize3(tmp, ex)) return PPM_ERR_ARG;
  cross3(ex, ez, tmp);
  if (!normalize3(tmp, ey)) return PPM_ERR_ARG;
  const double SENSOR_SIZE = 35.0 / 1000.0;
  double step = (c->focus * SENSOR_SIZE / c->focal_len) / (double)c->xreso;
  double ea = c->focal_len / c->f_number;
  double lx = (double)(c->xreso / 2), ly = (double)(c->yreso / 2);
  for (int i = 0; i < 3; ++i) {
    c->esx[i] = step * ex[i]; c->esy[i] = step * ey[i];
    c->eex[i] = ea * ex[i];   c->eey[i] = ea * ey[i];
    c->eye_dir[i] = ez[i];
  }
  for (int i = 0; i < 3; ++i)
    c->origin[i] = (c->focus * ez[i] - (lx - 0.5) * c->esx[i]) - (ly - 0.5) * c->esy[i];
  c->photon_power = c->blur ? c->iso_sens / 100.0 * 4.9 / c->f_number * c->shut_speed / (1.0 / 250.0) : 1.0;
  return PPM_OK;
}

int ppm_photon_budget(const ppm_light* lights, int32_t nlights, int64_t nphoton, double* power, int64_t* n_per_light) {
  if (!lights || nlights <= 0 || nphoton <= 0 || !power || !n_per_light) return PPM_ERR_ARG;
  double flux = 0.0;
  for (int i = 0; i < nlights; ++i) flux = flux + lights[i].flux;
  double pw = flux / (double)nphoton;
  for (int i = 0; i < nlights; ++i) n_per_light[i] = (int64_t)std::round(lights[i].flux / pw);
  *power = pw;
  return PPM_OK;
}

double ppm_radius_at(double r0, uint32_t pass) {
  const double ALPHA = 0.5;
  double r = r0;
  for (uint32_t i = 0; i < pass; ++i) r = std::sqrt(((double)(i + 1) + ALPHA) / ((double)(i + 1) + 1.0)) * r;
  return r;
}

void ppm_radius_schedule(double r0, int32_t npass, double* out) {
  const double ALPHA = 0.5;
  double r = r0;
  for (int32_t i = 0; i < npass; ++i) {
    out[i] = r;
    r = std::sqrt(((double)(i + 1) + ALPHA) / ((double)(i + 1) + 1.0)) * r;
  }
}

// ---- scene container -----------------------------------------------------
void ppm_scene_free(ppm_scene* s) { delete s; }
int32_t ppm_scene_nprims(const ppm_scene* s) { return s ? (int32_t)s->prims.size() : 0; }
int32_t ppm_scene_nmaterials(const ppm_scene* s) { return s ? (int32_t)s->mats.size() : 0; }
int32_t ppm_scene_nlights(const ppm_scene* s) { return s ? (int32_t)s->lights.size() : 0; }
const ppm_prim* ppm_scene_prims(const ppm_scene* s) { return s ? s->prims.data() : nullptr; }
const ppm_material* ppm_scene_materials(const ppm_scene* s) { return s ? s->mats.data() : nullptr; }
const ppm_light* ppm_scene_lights(const ppm_scene* s) { return s ? s->lights.data() : nullptr; }

// The scene every reference binary renders (read_scene ignores its file
// argument): 1 area light, 6 planes, 10 spheres (4 Torrance-Sparrow, 6 glossy
// metal), 1 emitter quad.  Values are the literals of scene.rs:22-441.
int ppm_scene_builtin(ppm_scene** out) {
  if (!out) return PPM_ERR_ARG;
  ppm_scene* s = new ppm_scene();
  const double Z[3] = {0, 0, 0};

  ppm_light l;
  std::memset(&l, 0, sizeof l);
  l.type = PPM_LIGHT_PARALLELOGRAM;
  const double white[3] = {1.0, 1.0, 1.0};
  ppm_color_normalize(white, l.color);
  l.flux = 5.0;
  set3(l.pos, -0.67, 3.99, 2.33); set3(l.nvec, -0.0, -1.0, -0.0);
  set3(l.dir1, 1.33, 0.0, 0.0); set3(l.dir2, 0.0, 0.0, 1.33);
  s->lights.push_back(l);

  auto simple = [&](const double em[3], const double ior[3], double r0, double r1, double r2,
                    double s0, double s1, double s2, double diff, double metal, double rough) {
    ppm_material m;
    const double refl[3] = {r0, r1, r2}, spec[3] = {s0, s1, s2};
    ppm_material_simple(&m, em, Z, ior, refl, spec, diff, metal, rough);
    s->mats.push_back(m);
    return (int32_t)s->mats.size() - 1;
  };
  auto ts = [&](const double ior[3], double a0, double a1, double a2, double s0, double s1, double s2,
                double scat, double metal, double rough) {
    ppm_material m;
    const double ad[3] = {a0, a1, a2}, as[3] = {s0, s1, s2};
    ppm_material_ts(&m, Z, Z, ior, ad, as, scat, metal, rough);
    s->mats.push_back(m);
    return (int32_t)s->mats.size() - 1;
  };
  const double ior15[3] = {1.5, 1.5, 1.5};
  const double em015[3] = {0.15, 0.15, 0.15};
  int32_t mwall  = simple(Z, Z, 0.5, 0.5, 0.5, 0.8, 0.8, 0.8, 1.0, 0.0, 0.0);
  int32_t mwallb = simple(Z, Z, 0.1, 0.1, 0.4, 0.8, 0.0, 0.8, 1.0, 0.0, 0.0);
  int32_t mwallr = simple(Z, Z, 0.4, 0.1, 0.1, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0);
  int32_t ball1  = ts(ior15, 0.6, 0.35, 0.1, 0.05, 0.05, 0.05, 1.0, 0.0, 0.0);
  int32_t ball2  = simple(Z, Z, 0, 0, 0, 0.78, 0.78, 0.78, 0.0, 1.0, 0.2);
  int32_t ball3  = simple(Z, Z, 0, 0, 0, 0.78, 0.78, 0.78, 0.0, 1.0, 0.3);
  int32_t ball4  = simple(Z, Z, 0, 0, 0, 0.78, 0.78, 0.78, 0.0, 1.0, 0.4);
  int32_t ball5  = ts(Z, 0, 0, 0, 0.78, 0.78, 0.78, 0.0, 1.0, 0.5);
  int32_t ball6  = ts(Z, 0, 0, 0, 0.78, 0.78, 0.78, 0.0, 1.0, 0.6);
  int32_t ball7  = simple(Z, Z, 0, 0, 0, 0.78, 0.78, 0.78, 0.0, 1.0, 0.7);
  int32_t ball8  = simple(Z, Z, 0, 0, 0, 0.78, 0.78, 0.78, 0.0, 1.0, 0.8);
  int32_t ball9  = simple(Z, Z, 0, 0, 0, 0.78, 0.78, 0.78, 0.0, 1.0, 0.9);
  int32_t ball10 = ts(ior15, 1.0, 1.0, 1.0, 0.05, 0.05, 0.05, 0.0, 0.0, 0.8);
  int32_t mparal = simple(em015, Z, 0, 0, 0, 0.8, 0.8, 0.8, 0.0, 0.0, 0.0);

  auto plain = [&](double nx, double ny, double nz, double dist, int32_t m) {
    ppm_prim p; const double n[3] = {nx, ny, nz};
    ppm_prim_plain(&p, n, dist, m); s->prims.push_back(p);
  };
  auto sphere = [&](double cx, double cy, double cz, double r, int32_t m) {
    ppm_prim p; const double c[3] = {cx, cy, cz};
    ppm_prim_sphere(&p, c, r, m); s->prims.push_back(p);
  };
  // -Vector3::EX etc. negate every component (algebra.rs:54-61): -0.0 for the zeros
  plain(0.0, 1.0, 0.0, 0.0, mwall);        // flooring
  plain(-0.0, -1.0, -0.0, 4.0, mwall);     // ceiling
  plain(-1.0, -0.0, -0.0, 2.0, mwallb);    // rsidewall
  plain(1.0, 0.0, 0.0, 2.0, mwallr);       // lsidewall
  plain(0.0, 0.0, 1.0, 6.0, mwall);        // backwall
  plain(-0.0, -0.0, -1.0, 5.0, mwall);     // frontwall
  sphere(-1.6, 1.5, 3.0, 0.4, ball1);
  sphere(-0.8, 1.5, 3.0, 0.4, ball2);
  sphere(0.0, 1.5, 3.0, 0.4, ball3);
  sphere(0.8, 1.5, 3.0, 0.4, ball4);
  sphere(1.6, 1.5, 3.0, 0.4, ball5);
  sphere(-1.6, 0.5, 2.5, 0.4, ball6);
  sphere(-0.8, 0.5, 2.5, 0.4, ball7);
  sphere(0.0, 0.5, 2.5, 0.4, ball8);
  sphere(0.8, 0.5, 2.5, 0.4, ball9);
  sphere(1.6, 0.5, 2.5, 0.4, ball10);
  {
    // ceiling_light: Shape::Parallelogram literal with nvec = -EY and edges given
    // as differences of corner literals (scene.rs:432-439)
    ppm_prim p;
    std::memset(&p, 0, sizeof p);
    p.type = PPM_SHAPE_PARALLELOGRAM; p.material = mparal;
    set3(p.position, -0.67, 3.99, 2.33);
    set3(p.nvec, -0.0, -1.0, -0.0);
    set3(p.dir1, 0.67 - (-0.67), 3.99 - 3.99, 2.33 - 2.33);
    set3(p.dir2, -0.67 - (-0.67), 3.99 - 3.99, 3.67 - 2.33);
    s->prims.push_back(p);
  }
  *out = s;
  return PPM_OK;
}

}  // extern "C"
