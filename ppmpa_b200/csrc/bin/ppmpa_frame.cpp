// ppmpa_frame -- a whole progressive frame in one process: what util/iterator.rb:90-117 (N `ppmpa` processes, one
// per pass, radius schedule :34-38) followed by util/averager2.rb:49-110 (sum the pass images, divide, write) do
// offline.  Passes are sharded round-robin over the GPUs of the box (one host thread and one engine context per GPU);
// every GPU accumulates its own passes on the device and ONE NCCL sum-reduce (ppm_accum_reduce, C ABI only: no Python,
// no MPI) combines the frame on GPU 0.
//   ppmpa_frame [-nc] [-g <#gpus>] [-r <checkpoint>] <#pass> <#photon> <radius> <camera file> <scene file> <out.ppm|out.exr>
// -r: resume.  The reference's passes are files, so a job that is killed or run in instalments keeps what it rendered
// (iterator.rb:96-117 skips nothing, but averager2.rb sums whatever pass files exist); here the sums live on the GPU, so
// the accumulator is loaded from <checkpoint> when that file exists -- the <#pass> passes of this run then continue the
// job: pass ids, Philox streams and the radius schedule go on from the number of passes it holds -- and is written back
// (atomically) when the run ends.  Four runs of 250 passes give the frame of one run of 1000 (up to summation order).
// Environment: PPM_SEED (default: OS entropy), PPM_DEVICE = first GPU (default 0).
#include "cli_common.h"

#include <atomic>
#include <chrono>
#include <thread>

static const char* USAGE = "Usage: ppmpa_frame [-nc] [-g <#gpus>] [-r <checkpoint>] <#pass> <#photon> <radius> <camera file> <scene file> <out.ppm|out.exr>";

int main(int argc, char** argv) {
  int a = 1, uc = 1, ngpu = 1;
  const char* ckpt = nullptr;
  while (a < argc && argv[a][0] == '-' && argv[a][1]) {
    if (!std::strcmp(argv[a], "-nc")) { uc = 0; ++a; }
    else if (!std::strcmp(argv[a], "-g") && a + 1 < argc) { ngpu = std::atoi(argv[a + 1]); a += 2; }
    else if (!std::strcmp(argv[a], "-r") && a + 1 < argc) { ckpt = argv[a + 1]; a += 2; }
    else { std::fprintf(stderr, "%s\n", USAGE); return 0; }
  }
  if (argc - a < 6 || ngpu < 1) { std::fprintf(stderr, "%s\n", USAGE); return 0; }
  const int npass = std::atoi(argv[a]);
  const long long nphoton = std::atoll(argv[a + 1]);
  const double r0 = std::atof(argv[a + 2]);
  const char* out = argv[a + 5];
  if (npass < 1 || nphoton < 1 || !(r0 > 0.0)) { std::fprintf(stderr, "%s\n", USAGE); return 1; }
  ppm_camera cam;
  ppm_scene* sc = nullptr;
  if (!cli_load_camera(argv[a + 3], &cam) || !cli_load_scene(argv[a + 4], &sc)) return 1;
  const uint64_t seed = cli_seed();
  const int dev0 = cli_device();

  std::vector<ppm_ctx*> ctx((size_t)ngpu, nullptr);
  for (int g = 0; g < ngpu; ++g) {
    int rc = ppm_create(dev0 + g, &ctx[(size_t)g]);
    if (rc != PPM_OK) { std::fprintf(stderr, "GPU %d: no usable CUDA device (error %d); this engine has no CPU path\n", dev0 + g, rc); return 1; }
    CLI_CK(ctx[(size_t)g], ppm_scene_set(ctx[(size_t)g], ppm_scene_prims(sc), ppm_scene_nprims(sc), ppm_scene_materials(sc),
                                        ppm_scene_nmaterials(sc), ppm_scene_lights(sc), ppm_scene_nlights(sc)));
    CLI_CK(ctx[(size_t)g], ppm_camera_set(ctx[(size_t)g], &cam));
    CLI_CK(ctx[(size_t)g], ppm_accum_reset(ctx[(size_t)g]));
  }
  // resume: GPU 0 starts from the checkpoint's sums, the run continues at the pass id it holds
  uint32_t done = 0;
  if (ckpt) {
    FILE* f = std::fopen(ckpt, "rb");
    if (f) {
      std::fclose(f);
      CLI_CK(ctx[0], ppm_accum_load(ctx[0], ckpt, &done));
      std::fprintf(stderr, "resuming after %u passes (%s)\n", done, ckpt);
    }
  }
  std::vector<double> radius((size_t)done + (size_t)npass);
  ppm_radius_schedule(r0, (int32_t)radius.size(), radius.data());
  char uid[128];
  if (ngpu > 1 && ppm_comm_unique_id(uid) != PPM_OK) { std::fprintf(stderr, "NCCL is not available (libnccl.so.2)\n"); return 1; }

  std::atomic<int> failed{0};
  auto run_ranks = [&](auto&& fn) {
    std::vector<std::thread> th;
    for (int g = 1; g < ngpu; ++g) th.emplace_back(fn, g);
    fn(0);
    for (auto& t : th) t.join();
  };
  // ncclCommInitRank blocks until every rank has joined: one thread per rank
  const auto t_init = std::chrono::steady_clock::now();
  if (ngpu > 1)
    run_ranks([&](int g) {
      if (ppm_comm_init(ctx[(size_t)g], ngpu, g, uid) != PPM_OK) { std::fprintf(stderr, "GPU %d: %s\n", g, ppm_last_error(ctx[(size_t)g])); failed = 1; }
    });
  if (failed) return 1;
  const auto t0 = std::chrono::steady_clock::now();
  run_ranks([&](int g) {
    ppm_ctx* c = ctx[(size_t)g];
    std::vector<double> r2;
    for (int p = g; p < npass; p += ngpu) r2.push_back(radius[(size_t)done + (size_t)p] * radius[(size_t)done + (size_t)p]);
    // this rank's passes: ids done + g, done + g + ngpu, ... (Philox stream AND radius index = the global pass id)
    if (!r2.empty() && ppm_render_passes(c, seed, done + (uint32_t)g, (uint32_t)ngpu, (int32_t)r2.size(), nphoton, r2.data(), uc) != PPM_OK) {
      std::fprintf(stderr, "GPU %d: %s\n", g, ppm_last_error(c)); failed = 1;
    }
    // the collective must be entered by every rank, failed or not, or the others would wait forever
    if (ngpu > 1 && ppm_accum_reduce(c, nullptr, 0) != PPM_OK) { std::fprintf(stderr, "GPU %d: %s\n", g, ppm_last_error(c)); failed = 1; }
  });
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  const double init_s = std::chrono::duration<double>(t0 - t_init).count();
  if (failed) return 1;

  std::vector<double> sum((size_t)cam.xreso * cam.yreso * 3);
  uint32_t n = 0;
  CLI_CK(ctx[0], ppm_accum_read(ctx[0], sum.data(), &n));
  std::fprintf(stderr, "%u passes x %lld photons at %dx%d on %d GPU(s): %.3f s (%.2f ms/pass, incl. calibration, buffers, graph capture%s); "
                       "communicator set-up %.3f s\n", n, nphoton, cam.xreso, cam.yreso, ngpu, secs, 1e3 * secs / (double)npass,
               ngpu > 1 ? " and the frame reduce" : "", init_s);
  const size_t len = std::strlen(out);
  const bool exr = len > 4 && !std::strcmp(out + len - 4, ".exr");
  int rc = exr ? ppm_write_mean_exr(out, &cam, sum.data(), n) : ppm_write_mean_ppm(out, &cam, sum.data(), n);
  if (rc != PPM_OK) { std::fprintf(stderr, "cannot write %s\n", out); return 1; }
  if (ckpt) CLI_CK(ctx[0], ppm_accum_save(ctx[0], ckpt));
  for (ppm_ctx* c : ctx) ppm_destroy(c);
  ppm_scene_free(sc);
  return (int)(n != done + (uint32_t)npass);
}
