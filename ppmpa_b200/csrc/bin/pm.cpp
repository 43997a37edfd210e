// pm -- photon tracer: writes the photon map of one pass to stdout.
// Same argv and stdout as the reference's src/bin/pm.rs:16-75:
//   pm <scene file> [<#photon>]
// stdout: "<#photon>", "<power>", then "<Red|Green|Blue> px py pz dx dy dz" per record.
#include "cli_common.h"

static const char* USAGE = "Usage: pm [-c|-h] <scene file> [<#photon>] (output photon map to stdout)";

int main(int argc, char** argv) {
  if (argc < 2 || !std::strcmp(argv[1], "-h")) { std::printf("%s\n", USAGE); return 0; }
  long long np = 100000;
  if (argc == 3) {
    char* end;
    long long v = std::strtoll(argv[2], &end, 10);
    if (end != argv[2] && !*end) np = v;
  }
  const int uc = 1;                                      // DEF_USECLASSIC
  ppm_scene* sc = nullptr;
  if (!cli_load_scene(argv[1], &sc)) return 1;
  ppm_ctx* ctx = nullptr;
  if (!cli_engine(&ctx, sc)) return 1;
  double power;
  std::vector<int64_t> ns((size_t)ppm_scene_nlights(sc));
  if (ppm_photon_budget(ppm_scene_lights(sc), ppm_scene_nlights(sc), np, &power, ns.data()) != PPM_OK) return 1;
  uint64_t n = 0;
  CLI_CK(ctx, ppm_trace_photons(ctx, cli_seed(), cli_pass(), uc, ns.data(), power, &n));
  std::vector<ppm_photon> ph((size_t)(n ? n : 1));
  CLI_CK(ctx, ppm_photons_export(ctx, ph.data(), n, nullptr));
  if (ppm_write_photon_dump(nullptr, np, power, ph.data(), n) != PPM_OK) return 1;
  ppm_destroy(ctx);
  ppm_scene_free(sc);
  return 0;
}
