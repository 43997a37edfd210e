// rt -- gather renderer: photon map on stdin -> image on stdout.
// Same argv, stdin and stdout as the reference's src/bin/rt.rs:16-59:
//   rt <scene file> <camera file> [<radius>]
#include "cli_common.h"

#include <chrono>

static const char* USAGE = "Usage: rtc <scene file> <camera file> [<radius>]";   // sic, rt.rs:16

int main(int argc, char** argv) {
  if (argc < 3) { std::printf("%s\n", USAGE); return 0; }
  const int uc = 1;
  double radius2 = 0.1 * 0.1;
  if (argc == 4) {
    char* end;
    double r = std::strtod(argv[3], &end);
    if (end != argv[3] && !*end) radius2 = r * r;
  }
  ppm_scene* sc = nullptr;
  ppm_camera cam;
  if (!cli_load_scene(argv[1], &sc) || !cli_load_camera(argv[2], &cam)) return 1;
  ppm_ctx* ctx = nullptr;
  if (!cli_engine(&ctx, sc)) return 1;
  CLI_CK(ctx, ppm_camera_set(ctx, &cam));
  auto t0 = std::chrono::steady_clock::now();
  ppm_photon* ph = nullptr;
  uint64_t n = 0;
  double power = 1.0;
  if (ppm_read_photon_dump(nullptr, &ph, &n, &power) != PPM_OK) { std::fprintf(stderr, "Error in reading photon map\n"); return 1; }
  CLI_CK(ctx, ppm_photons_import(ctx, ph, n, power));
  ppm_free(ph);
  CLI_CK(ctx, ppm_map_build(ctx, radius2));
  double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::fprintf(stderr, "finished reading map: %llu photons, %.6fs.\n", (unsigned long long)n, dt);
  const uint64_t seed = cli_seed();
  const uint32_t pass = cli_pass();
  size_t npix = (size_t)cam.xreso * cam.yreso;
  std::vector<double> rays(npix * 6), img(npix * 3);
  CLI_CK(ctx, ppm_generate_rays(ctx, seed, pass, rays.data()));
  CLI_CK(ctx, ppm_trace_rays(ctx, rays.data(), (int64_t)npix, 0, seed, pass, uc, img.data()));
  if (ppm_write_image(nullptr, &cam, img.data(), cam.progressive) != PPM_OK) return 1;
  ppm_destroy(ctx);
  ppm_scene_free(sc);
  return 0;
}
