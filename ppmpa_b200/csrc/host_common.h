// host_common.h -- shared declarations of the host-side (CPU) part of the C ABI.
#ifndef PPM_HOST_COMMON_H_
#define PPM_HOST_COMMON_H_

#include "../../include/ppm.h"

#include <string>
#include <vector>

// Host-side owner of a parsed / built-in scene (opaque in ppm.h).
struct ppm_scene {
  std::vector<ppm_prim> prims;
  std::vector<ppm_material> mats;
  std::vector<ppm_light> lights;
  std::vector<std::string> prim_names, mat_names;
};

namespace ppmhost {
double dot3(const double a[3], const double b[3]);
void cross3(const double a[3], const double b[3], double o[3]);
bool normalize3(const double a[3], double o[3]);
// Rust `{}` (exp_form = false) / `{:e}` (true) formatting of an f64
std::string fmt_f64(double v, bool exp_form);
}  // namespace ppmhost

#endif
