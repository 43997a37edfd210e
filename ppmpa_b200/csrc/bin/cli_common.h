// cli_common.h -- shared helpers of the ppmpa / pm / rt command-line binaries.
// The binaries keep the reference's argv and stdout protocols (src/bin/*.rs) and
// drive the GPU engine through the C ABI only.
#ifndef PPM_CLI_COMMON_H_
#define PPM_CLI_COMMON_H_

#include "../../../include/ppm.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

// The reference draws from an OS-seeded thread_rng (every process differs).
// PPM_SEED / PPM_PASS pin the Philox stream for reproducible runs.
static inline uint64_t cli_seed() {
  if (const char* s = std::getenv("PPM_SEED")) return std::strtoull(s, nullptr, 0);
  std::random_device rd;
  return ((uint64_t)rd() << 32) ^ (uint64_t)rd();
}
static inline uint32_t cli_pass() {
  if (const char* s = std::getenv("PPM_PASS")) return (uint32_t)std::strtoul(s, nullptr, 0);
  return 0;
}
static inline int cli_device() {
  if (const char* s = std::getenv("PPM_DEVICE")) return std::atoi(s);
  return 0;
}
// "builtin" (or "-") selects what the reference hard-codes; anything else is parsed.
static inline bool cli_load_scene(const char* path, ppm_scene** sc) {
  if (!std::strcmp(path, "builtin") || !std::strcmp(path, "-")) return ppm_scene_builtin(sc) == PPM_OK;
  char err[512] = {0};
  if (ppm_scene_load(path, sc, err, sizeof err) != PPM_OK) { std::fprintf(stderr, "scene: %s\n", err); return false; }
  return true;
}
static inline bool cli_load_camera(const char* path, ppm_camera* cam) {
  if (!std::strcmp(path, "builtin") || !std::strcmp(path, "-")) { ppm_camera_default(cam); return true; }
  char err[512] = {0};
  if (ppm_camera_load(path, cam, err, sizeof err) != PPM_OK) { std::fprintf(stderr, "camera: %s\n", err); return false; }
  return true;
}
static inline bool cli_engine(ppm_ctx** ctx, const ppm_scene* sc) {
  int rc = ppm_create(cli_device(), ctx);
  if (rc != PPM_OK) { std::fprintf(stderr, "no usable CUDA device (error %d); this engine has no CPU path\n", rc); return false; }
  rc = ppm_scene_set(*ctx, ppm_scene_prims(sc), ppm_scene_nprims(sc), ppm_scene_materials(sc), ppm_scene_nmaterials(sc),
                     ppm_scene_lights(sc), ppm_scene_nlights(sc));
  if (rc != PPM_OK) { std::fprintf(stderr, "scene: %s\n", ppm_last_error(*ctx)); return false; }
  return true;
}
#define CLI_CK(ctx, call) do { if ((call) != PPM_OK) { std::fprintf(stderr, "%s: %s\n", #call, ppm_last_error(ctx)); return 1; } } while (0)

#endif
