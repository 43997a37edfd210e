// rtc -- classic (Whitted-style) ray tracer, no photon map.
// Same argv and stdout as the reference's src/bin/rtc.rs:13-38:
//   rtc <screen file> <scene file>
#include "cli_common.h"

static const char* USAGE = "Usage: rtc <screen file> <scene file>";

int main(int argc, char** argv) {
  if (argc != 3) { std::printf("%s\n", USAGE); return 0; }
  ppm_camera cam;
  ppm_scene* sc = nullptr;
  if (!cli_load_camera(argv[1], &cam) || !cli_load_scene(argv[2], &sc)) return 1;
  ppm_ctx* ctx = nullptr;
  if (!cli_engine(&ctx, sc)) return 1;
  CLI_CK(ctx, ppm_camera_set(ctx, &cam));
  const uint64_t seed = cli_seed();
  const uint32_t pass = cli_pass();
  size_t npix = (size_t)cam.xreso * cam.yreso;
  std::vector<double> rays(npix * 6), img(npix * 3);
  CLI_CK(ctx, ppm_generate_rays(ctx, seed, pass, rays.data()));
  CLI_CK(ctx, ppm_trace_rays_classic(ctx, rays.data(), (int64_t)npix, 0, seed, pass, img.data()));
  if (ppm_write_image(nullptr, &cam, img.data(), cam.progressive) != PPM_OK) return 1;
  ppm_destroy(ctx);
  ppm_scene_free(sc);
  return 0;
}
