// host_io.cpp -- text / image formats of the reference's CLI protocols.
//
//   photon dump : pm.rs:44-45,65-74 (writer)  <->  photonmap.rs:31-74 (reader)
//   pass image  : camera.rs:77-90 header + camera.rs:208-214 pixel lines
//                 (ppmpa.rs:40-45, rt.rs:48-59)
//   mean image  : util/averager2.rb:84-110 (P3 PPM) and :154-218 (float32 EXR)
//
// Rust prints f64 with `{}` / `{:e}` = shortest digits that round-trip, never
// an exponent for `{}`, no '+' / zero padding in the exponent for `{:e}`.
#include "host_common.h"

#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace ppmhost {

// shortest round-trip decimal digits of |v| and the decimal exponent of the
// first digit (v = 0.d1d2d3... * 10^(e10+1), i.e. d1.d2d3 * 10^e10)
static void shortest_digits(double av, std::string* digits, int* e10) {
  char buf[64];
  auto r = std::to_chars(buf, buf + sizeof buf, av, std::chars_format::scientific);
  std::string s(buf, r.ptr);
  size_t epos = s.find('e');
  std::string mant = s.substr(0, epos);
  *e10 = std::atoi(s.c_str() + epos + 1);
  digits->clear();
  for (char c : mant) if (c != '.') digits->push_back(c);
  while (digits->size() > 1 && digits->back() == '0') digits->pop_back();
}

std::string fmt_f64(double v, bool exp_form) {
  if (std::isnan(v)) return "NaN";
  std::string out;
  if (std::signbit(v)) out = "-";
  double av = std::fabs(v);
  if (std::isinf(av)) return out + "inf";
  if (av == 0.0) return out + (exp_form ? "0e0" : "0");
  std::string d;
  int e;
  shortest_digits(av, &d, &e);
  if (exp_form) {
    out += d[0];
    if (d.size() > 1) { out += '.'; out.append(d, 1, std::string::npos); }
    out += 'e';
    out += std::to_string(e);
    return out;
  }
  int nd = (int)d.size();
  if (e < 0) {                      // 0.000ddd
    out += "0.";
    out.append((size_t)(-e - 1), '0');
    out += d;
  } else if (e + 1 >= nd) {         // ddd000
    out += d;
    out.append((size_t)(e + 1 - nd), '0');
  } else {                          // dd.ddd
    out.append(d, 0, (size_t)e + 1);
    out += '.';
    out.append(d, (size_t)e + 1, std::string::npos);
  }
  return out;
}

// Ruby Float#to_s (used by averager2.rb's "## max radiance = #{@hd2}")
static std::string fmt_ruby_float(double v) {
  if (std::isnan(v)) return "NaN";
  if (std::isinf(v)) return v < 0 ? "-Infinity" : "Infinity";
  double av = std::fabs(v);
  std::string sign = std::signbit(v) ? "-" : "";
  if (av == 0.0) return sign + "0.0";
  std::string d;
  int e;
  shortest_digits(av, &d, &e);
  if (e >= -4 && e < 16) {
    std::string s = fmt_f64(av, false);
    if (s.find('.') == std::string::npos) s += ".0";
    return sign + s;
  }
  std::string s = sign;
  s += d[0];
  s += '.';
  if (d.size() > 1) s.append(d, 1, std::string::npos); else s += '0';
  char eb[16];
  std::snprintf(eb, sizeof eb, "e%c%02d", e < 0 ? '-' : '+', std::abs(e));
  return s + eb;
}

}  // namespace ppmhost

using namespace ppmhost;

namespace {

const char* WL_NAME[3] = {"Red", "Green", "Blue"};

struct Out {
  FILE* f; bool own;
  explicit Out(const char* path, const char* mode = "w") : f(nullptr), own(false) {
    if (path) { f = std::fopen(path, mode); own = true; } else { f = stdout; }
  }
  ~Out() { if (f) { if (own) std::fclose(f); else std::fflush(f); } }
};

// camera.rs:77-90
std::vector<std::string> pnm_header(const ppm_camera* cam) {
  std::string ss = cam->shut_speed < 1.0 ? "1/" + fmt_f64(1.0 / cam->shut_speed, false) : fmt_f64(cam->shut_speed, false);
  std::vector<std::string> h;
  h.push_back("P3");
  h.push_back("## max radiance = " + fmt_f64(cam->max_radiance, false));
  h.push_back("## image parameters = " + ss + ", F" + fmt_f64(cam->f_number, false) + ", ISO" + fmt_f64(cam->iso_sens, false));
  h.push_back(std::to_string(cam->xreso) + " " + std::to_string(cam->yreso));
  h.push_back("255");
  return h;
}

// averager2.rb:84-94
int averager_clip(double c, uint32_t nfile, double max_radiance) {
  double c2 = c / (double)nfile / max_radiance;
  double r = std::pow(c2 > 1.0 ? 1.0 : c2, 1.0 / 2.2) * 255.0;
  if (!(r == r)) return 0;
  return (int)r;
}

void put_i32(std::string& s, int32_t v) { s.append((const char*)&v, 4); }
void put_f32(std::string& s, float v) { s.append((const char*)&v, 4); }
void put_u64(std::string& s, uint64_t v) { s.append((const char*)&v, 8); }
void put_str(std::string& s, const char* z) { s.append(z); s.push_back('\0'); }

}  // namespace

extern "C" {

int ppm_format_f64(double v, int exp_form, char* buf, size_t buflen) {
  if (!buf || buflen == 0) return PPM_ERR_ARG;
  std::string s = fmt_f64(v, exp_form != 0);
  if (s.size() + 1 > buflen) return PPM_ERR_CAPACITY;
  std::memcpy(buf, s.c_str(), s.size() + 1);
  return PPM_OK;
}

void ppm_radiance_to_rgb(double max_radiance, const double rad[3], int32_t rgb[3]) {
  for (int i = 0; i < 3; ++i) {
    double d2 = rad[i] / max_radiance;
    double r2 = d2 > 1.0 ? 1.0 : d2;
    rgb[i] = (int32_t)std::floor(std::pow(r2, 1.0 / 2.2) * 255.0);
  }
}

int ppm_write_photon_dump(const char* path, int64_t nphoton, double power, const ppm_photon* ph, uint64_t n) {
  if (n && !ph) return PPM_ERR_ARG;
  Out o(path);
  if (!o.f) return PPM_ERR_IO;
  std::fprintf(o.f, "%lld\n", (long long)nphoton);
  std::fprintf(o.f, "%s\n", fmt_f64(power, false).c_str());
  std::string line;
  for (uint64_t i = 0; i < n; ++i) {
    int wl = ph[i].wl;
    if (wl < 0 || wl > 2) return PPM_ERR_ARG;
    line = WL_NAME[wl];
    for (int k = 0; k < 3; ++k) { line += ' '; line += fmt_f64(ph[i].pos[k], false); }
    for (int k = 0; k < 3; ++k) { line += ' '; line += fmt_f64(ph[i].dir[k], false); }
    line += '\n';
    if (std::fwrite(line.data(), 1, line.size(), o.f) != line.size()) return PPM_ERR_IO;
  }
  return PPM_OK;
}

int ppm_read_photon_dump(const char* path, ppm_photon** out, uint64_t* n, double* power) {
  if (!out || !n || !power) return PPM_ERR_ARG;
  FILE* f = path ? std::fopen(path, "r") : stdin;
  if (!f) return PPM_ERR_IO;
  std::vector<ppm_photon> v;
  char* line = nullptr;
  size_t cap = 0;
  ssize_t len;
  int lineno = 0;
  double pw = 1.0;
  int rc = PPM_OK;
  while ((len = getline(&line, &cap, f)) >= 0) {
    ++lineno;
    if (lineno == 1) continue;                       // "#photon" line is ignored, photonmap.rs:35
    if (lineno == 2) {                               // power, default 1.0 on parse failure, :37-45
      char* end = nullptr;
      double p = std::strtod(line, &end);
      if (end != line) pw = p;
      continue;
    }
    while (len > 0 && (line[len - 1] == '\n' || line[len - 1] == '\r')) line[--len] = '\0';
    if (len == 0) continue;
    ppm_photon ph;
    std::memset(&ph, 0, sizeof ph);
    char* sp = std::strchr(line, ' ');
    if (!sp) { rc = PPM_ERR_PARSE; break; }
    *sp = '\0';
    ph.wl = !std::strcmp(line, "Green") ? PPM_WL_GREEN : (!std::strcmp(line, "Blue") ? PPM_WL_BLUE : PPM_WL_RED);
    double e[6];
    char* p = sp + 1;
    bool ok = true;
    for (int k = 0; k < 6; ++k) {
      char* end = nullptr;
      e[k] = std::strtod(p, &end);
      if (end == p) { ok = false; break; }
      p = end;
    }
    if (!ok) { rc = PPM_ERR_PARSE; break; }
    ph.pos[0] = e[0]; ph.pos[1] = e[1]; ph.pos[2] = e[2];
    if (!normalize3(e + 3, ph.dir)) { rc = PPM_ERR_PARSE; break; }   // Ray::new_from_elem, geometry.rs:46-54
    v.push_back(ph);
  }
  std::free(line);
  if (path) std::fclose(f);
  if (rc != PPM_OK) return rc;
  ppm_photon* buf = (ppm_photon*)std::malloc(sizeof(ppm_photon) * (v.size() ? v.size() : 1));
  if (!buf) return PPM_ERR_IO;
  if (!v.empty()) std::memcpy(buf, v.data(), sizeof(ppm_photon) * v.size());
  *out = buf; *n = v.size(); *power = pw;
  return PPM_OK;
}

void ppm_free(void* p) { std::free(p); }

int ppm_write_image(const char* path, const ppm_camera* cam, const double* rgb3, int progressive) {
  if (!cam || !rgb3) return PPM_ERR_ARG;
  Out o(path);
  if (!o.f) return PPM_ERR_IO;
  for (auto& l : pnm_header(cam)) std::fprintf(o.f, "%s\n", l.c_str());
  size_t npix = (size_t)cam->xreso * (size_t)cam->yreso;
  std::string buf;
  buf.reserve(1 << 20);
  for (size_t i = 0; i < npix; ++i) {
    const double* c = rgb3 + i * 3;
    if (progressive) {
      buf += fmt_f64(c[0], true); buf += ' '; buf += fmt_f64(c[1], true); buf += ' '; buf += fmt_f64(c[2], true);
    } else {
      int32_t rgb[3];
      ppm_radiance_to_rgb(cam->max_radiance, c, rgb);
      buf += std::to_string(rgb[0]); buf += ' '; buf += std::to_string(rgb[1]); buf += ' '; buf += std::to_string(rgb[2]);
    }
    buf += '\n';
    if (buf.size() > (1 << 20) - 256) {
      if (std::fwrite(buf.data(), 1, buf.size(), o.f) != buf.size()) return PPM_ERR_IO;
      buf.clear();
    }
  }
  if (!buf.empty() && std::fwrite(buf.data(), 1, buf.size(), o.f) != buf.size()) return PPM_ERR_IO;
  return PPM_OK;
}

int ppm_write_mean_ppm(const char* path, const ppm_camera* cam, const double* sum, uint32_t n_pass) {
  if (!cam || !sum || n_pass == 0) return PPM_ERR_ARG;
  Out o(path);
  if (!o.f) return PPM_ERR_IO;
  // output_ppm, averager2.rb:96-110 (note: no "image parameters" line)
  std::fprintf(o.f, "P3\n## max radiance = %s\n%d %d\n255\n", fmt_ruby_float(cam->max_radiance).c_str(), cam->xreso, cam->yreso);
  size_t npix = (size_t)cam->xreso * (size_t)cam->yreso;
  for (size_t i = 0; i < npix; ++i) {
    const double* c = sum + i * 3;
    if (!(c[0] > 0 || c[1] > 0 || c[2] > 0)) { std::fputs("0 0 0\n", o.f); continue; }   // @count[i] == 0
    std::fprintf(o.f, "%d %d %d\n", averager_clip(c[0], n_pass, cam->max_radiance),
                 averager_clip(c[1], n_pass, cam->max_radiance), averager_clip(c[2], n_pass, cam->max_radiance));
  }
  return PPM_OK;
}

int ppm_write_mean_exr(const char* path, const ppm_camera* cam, const double* sum, uint32_t n_pass) {
  if (!cam || !sum || n_pass == 0) return PPM_ERR_ARG;
  Out o(path, "wb");
  if (!o.f) return PPM_ERR_IO;
  const int32_t FLOAT = 2;
  int32_t xlen = cam->xreso, ylen = cam->yreso;
  // header_exr, averager2.rb:154-176
  std::string ch;
  for (const char* name : {"B", "G", "R"}) {
    put_str(ch, name); put_i32(ch, FLOAT); put_i32(ch, 0); put_i32(ch, 1); put_i32(ch, 1);
  }
  put_str(ch, "");
  std::string h;
  put_i32(h, 20000630); put_i32(h, 2);
  put_str(h, "channels"); put_str(h, "chlist"); put_i32(h, (int32_t)ch.size()); h += ch;
  put_str(h, "compression"); put_str(h, "compression"); put_i32(h, 1); h.push_back((char)0);
  put_str(h, "dataWindow"); put_str(h, "box2i"); put_i32(h, 16);
  put_i32(h, 0); put_i32(h, 0); put_i32(h, xlen - 1); put_i32(h, ylen - 1);
  put_str(h, "displayWindow"); put_str(h, "box2i"); put_i32(h, 16);
  put_i32(h, 0); put_i32(h, 0); put_i32(h, xlen - 1); put_i32(h, ylen - 1);
  put_str(h, "lineOrder"); put_str(h, "lineOrder"); put_i32(h, 1); h.push_back((char)0);
  put_str(h, "pixelAspectRatio"); put_str(h, "float"); put_i32(h, 4); put_f32(h, 1.0f);
  put_str(h, "screenWindowCenter"); put_str(h, "v2f"); put_i32(h, 8); put_f32(h, 0.0f); put_f32(h, 0.0f);
  put_str(h, "screenWindowWidth"); put_str(h, "float"); put_i32(h, 4); put_f32(h, 1.0f);
  put_str(h, "");
  // output_exr, averager2.rb:178-218
  double cmag = (double)n_pass * cam->max_radiance;
  uint64_t offset = h.size() + 8ull * (uint64_t)ylen;
  uint64_t datalen = 4ull * (uint64_t)xlen * 3ull;
  uint64_t linelen = 8ull + datalen;
  std::string tab;
  for (int32_t y = 0; y < ylen; ++y) put_u64(tab, offset + linelen * (uint64_t)y);
  if (std::fwrite(h.data(), 1, h.size(), o.f) != h.size()) return PPM_ERR_IO;
  if (std::fwrite(tab.data(), 1, tab.size(), o.f) != tab.size()) return PPM_ERR_IO;
  std::string line;
  for (int32_t y = 0; y < ylen; ++y) {
    line.clear();
    put_i32(line, y); put_i32(line, (int32_t)datalen);
    for (int chn = 2; chn >= 0; --chn)          // B, G, R planes
      for (int32_t x = 0; x < xlen; ++x) {
        const double* c = sum + ((size_t)y * xlen + x) * 3;
        bool lit = c[0] > 0 || c[1] > 0 || c[2] > 0;
        put_f32(line, lit ? (float)(c[chn] / cmag) : 0.0f);
      }
    if (std::fwrite(line.data(), 1, line.size(), o.f) != line.size()) return PPM_ERR_IO;
  }
  return PPM_OK;
}

}  // extern "C"
