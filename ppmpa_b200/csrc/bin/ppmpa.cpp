// ppmpa -- one progressive-photon-mapping pass in one process.
// Same argv and stdout as the reference's src/bin/ppmpa.rs:15-46:
//   ppmpa [-nc|-h] <#photon> <radius> <camera file> <scene file>
// stdout: 5 header lines (camera.rs:77-90) + one "{:e} {:e} {:e}" line per pixel.
#include "cli_common.h"

static const char* USAGE = "Usage: ppmpa [-nc|-h] <#photon> <radius> <camera file> <scene file>";

int main(int argc, char** argv) {
  if (argc < 5 || !std::strcmp(argv[1], "-h")) { std::fprintf(stderr, "%s\n", USAGE); return 0; }
  int off = 1, uc = 1;                                   // DEF_USECLASSIC = true
  if (!std::strcmp(argv[1], "-nc")) { off = 2; uc = 0; }
  if (argc < off + 4) { std::fprintf(stderr, "%s\n", USAGE); return 0; }
  char* end;
  long long np = std::strtoll(argv[off], &end, 10);
  if (end == argv[off] || *end) np = 100000;             // DEF_NPHOTON on parse failure
  double r = std::strtod(argv[off + 1], &end);
  double radius2 = (end == argv[off + 1] || *end) ? 0.1 * 0.1 : r * r;
  ppm_camera cam;
  ppm_scene* sc = nullptr;
  if (!cli_load_camera(argv[off + 2], &cam) || !cli_load_scene(argv[off + 3], &sc)) return 1;
  ppm_ctx* ctx = nullptr;
  if (!cli_engine(&ctx, sc)) return 1;
  CLI_CK(ctx, ppm_camera_set(ctx, &cam));
  CLI_CK(ctx, ppm_render_pass(ctx, cli_seed(), cli_pass(), np, radius2, uc));
  std::vector<double> img((size_t)cam.xreso * cam.yreso * 3);
  CLI_CK(ctx, ppm_pass_image_read(ctx, img.data()));
  if (ppm_write_image(nullptr, &cam, img.data(), 1) != PPM_OK) return 1;   // ppmpa always prints radiances
  ppm_destroy(ctx);
  ppm_scene_free(sc);
  return 0;
}
