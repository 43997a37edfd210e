// host_parse.cpp -- parsers for example/<name>.scene and example/<name>.scr.
//
// The reference ships these files but never reads them (read_scene /
// read_camera ignore their argument, scene.rs:20, camera.rs:108), so there is
// no reference parser to mirror.  The grammar implemented here is the one
// derived from the example files and doc/ebnf-camera.txt:
//
//   scene : line-oriented YAML subset.  `#` starts a comment.  Sections at
//           column 0 (`light:`, `material:`, `vertex:`, `object:`; `air:` and
//           `ambient_light:` are accepted and ignored).  Items start with
//           `- key : value`, continuation lines are `key : value`.  Vectors
//           are `[ x, y, z ]`.  Object order in the file = object index.
//   camera: `key : value` lines; legacy keys (`xresolution`) and EBNF keys
//           (`x_resolution`) are both accepted; missing keys keep the defaults
//           of camera.rs:109-128.
//
// The mapping file value -> model value follows the hard-coded equivalents in
// scene.rs (checked against example/ex-11.9.scene, which describes the same
// room): light colour is normalised (scene.rs:24), plain dist = -(normal .
// position) (scene.rs:360-383), polygons go through Shape::new_polygon,
// `smoothness` is passed as new_simple's 5th (roughness) argument.
#include "host_common.h"

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>

using namespace ppmhost;

namespace {

struct ParseError {
  std::string msg;
};

std::string trim(const std::string& s) {
  size_t b = 0, e = s.size();
  while (b < e && std::isspace((unsigned char)s[b])) ++b;
  while (e > b && std::isspace((unsigned char)s[e - 1])) --e;
  return s.substr(b, e - b);
}
std::string strip_comment(const std::string& s) {
  size_t p = s.find('#');
  return p == std::string::npos ? s : s.substr(0, p);
}
bool parse_double(const std::string& s, double* out) {
  std::string t = trim(s);
  if (t.empty()) return false;
  char* end = nullptr;
  double v = std::strtod(t.c_str(), &end);
  if (end == t.c_str() || *end != '\0') return false;
  *out = v;
  return true;
}
bool parse_vec3(const std::string& s, double out[3]) {
  std::string t = trim(s);
  if (t.size() < 2 || t.front() != '[' || t.back() != ']') return false;
  t = t.substr(1, t.size() - 2);
  std::stringstream ss(t);
  std::string tok;
  int n = 0;
  while (std::getline(ss, tok, ',')) {
    if (n >= 3 || !parse_double(tok, &out[n])) return false;
    ++n;
  }
  return n == 3;
}
bool parse_yesno(const std::string& s, int* out) {
  std::string t = trim(s);
  for (auto& c : t) c = (char)std::tolower((unsigned char)c);
  if (t == "yes" || t == "true") { *out = 1; return true; }
  if (t == "no" || t == "false") { *out = 0; return true; }
  return false;
}

typedef std::map<std::string, std::string> Item;
struct Section {
  std::vector<Item> items;
  std::vector<int> lines;
};

const std::string& need(const Item& it, const char* key, const char* what, int line) {
  auto f = it.find(key);
  if (f == it.end()) {
    ParseError e; e.msg = std::string(what) + " item at line " + std::to_string(line) + ": missing key '" + key + "'";
    throw e;
  }
  return f->second;
}
void need_vec(const Item& it, const char* key, const char* what, int line, double out[3]) {
  if (!parse_vec3(need(it, key, what, line), out)) {
    ParseError e; e.msg = std::string(what) + " item at line " + std::to_string(line) + ": key '" + key + "' is not a vector";
    throw e;
  }
}
double need_num(const Item& it, const char* key, const char* what, int line) {
  double v;
  if (!parse_double(need(it, key, what, line), &v)) {
    ParseError e; e.msg = std::string(what) + " item at line " + std::to_string(line) + ": key '" + key + "' is not a number";
    throw e;
  }
  return v;
}
void opt_vec(const Item& it, const char* key, double out[3]) {
  auto f = it.find(key);
  if (f != it.end()) parse_vec3(f->second, out);
}
double opt_num(const Item& it, const char* key, double dflt) {
  auto f = it.find(key);
  double v;
  if (f != it.end() && parse_double(f->second, &v)) return v;
  return dflt;
}

void parse_scene_text(std::istream& in, ppm_scene* sc) {
  std::map<std::string, Section> sections;
  Section* cur = nullptr;
  std::string raw;
  int lineno = 0;
  while (std::getline(in, raw)) {
    ++lineno;
    std::string line = strip_comment(raw);
    if (trim(line).empty()) continue;
    bool col0 = !std::isspace((unsigned char)line[0]);
    std::string t = trim(line);
    if (col0) {
      // section header `name:`
      size_t c = t.find(':');
      if (c == std::string::npos) { ParseError e; e.msg = "line " + std::to_string(lineno) + ": expected a section header"; throw e; }
      cur = &sections[trim(t.substr(0, c))];
      continue;
    }
    if (!cur) { ParseError e; e.msg = "line " + std::to_string(lineno) + ": entry outside any section"; throw e; }
    bool new_item = false;
    if (t[0] == '-') { new_item = true; t = trim(t.substr(1)); }
    size_t c = t.find(':');
    if (c == std::string::npos) { ParseError e; e.msg = "line " + std::to_string(lineno) + ": expected 'key : value'"; throw e; }
    std::string key = trim(t.substr(0, c)), val = trim(t.substr(c + 1));
    if (new_item || cur->items.empty()) { cur->items.push_back(Item()); cur->lines.push_back(lineno); }
    cur->items.back()[key] = val;
  }

  // vertices: `- name: [x,y,z]`
  std::map<std::string, std::vector<double>> verts;
  for (auto& it : sections["vertex"].items)
    for (auto& kv : it) {
      double v[3];
      if (!parse_vec3(kv.second, v)) { ParseError e; e.msg = "vertex '" + kv.first + "' is not a vector"; throw e; }
      verts[kv.first] = {v[0], v[1], v[2]};
    }

  // materials
  std::map<std::string, int32_t> mat_index;
  {
    Section& s = sections["material"];
    for (size_t i = 0; i < s.items.size(); ++i) {
      const Item& it = s.items[i];
      int ln = s.lines[i];
      double em[3] = {0, 0, 0}, refl[3] = {0, 0, 0}, tr[3] = {0, 0, 0}, spec[3] = {0, 0, 0}, ior[3] = {0, 0, 0};
      opt_vec(it, "emittance", em); opt_vec(it, "reflectance", refl); opt_vec(it, "transmittance", tr);
      opt_vec(it, "specularrefl", spec); opt_vec(it, "ior", ior);
      double diff = opt_num(it, "diffuseness", 0.0), metal = opt_num(it, "metalness", 0.0), smooth = opt_num(it, "smoothness", 0.0);
      ppm_material m;
      ppm_material_simple(&m, em, tr, ior, refl, spec, diff, metal, smooth);
      std::string name = need(it, "name", "material", ln);
      // All 253 materials of the reference's example files say 0.0 and the reference never reads a scene file
      // (scene.rs:20), so a non-zero value has no defined meaning: it is passed on as new_simple's roughness, loudly.
      if (smooth != 0.0)
        std::fprintf(stderr, "ppm scene: material '%s' (line %d): smoothness %g is passed as Surface::new_simple's roughness; "
                             "the reference defines no mapping for a non-zero value\n", name.c_str(), ln, smooth);
      mat_index[name] = (int32_t)sc->mats.size();
      sc->mats.push_back(m);
      sc->mat_names.push_back(name);
    }
  }

  // lights
  {
    Section& s = sections["light"];
    for (size_t i = 0; i < s.items.size(); ++i) {
      const Item& it = s.items[i];
      int ln = s.lines[i];
      std::string type = need(it, "type", "light", ln);
      ppm_light l;
      std::memset(&l, 0, sizeof l);
      double col[3];
      need_vec(it, "color", "light", ln, col);
      ppm_color_normalize(col, l.color);
      l.flux = it.count("flux") ? need_num(it, "flux", "light", ln) : need_num(it, "power", "light", ln);
      need_vec(it, "position", "light", ln, l.pos);
      if (type == "point") {
        l.type = PPM_LIGHT_POINT;
      } else if (type == "parallelogram" || type == "sun") {
        l.type = type == "sun" ? PPM_LIGHT_SUN : PPM_LIGHT_PARALLELOGRAM;
        need_vec(it, "dir1", "light", ln, l.dir1);
        need_vec(it, "dir2", "light", ln, l.dir2);
        double c[3];
        cross3(l.dir1, l.dir2, c);
        if (!normalize3(c, l.nvec)) { ParseError e; e.msg = "light at line " + std::to_string(ln) + ": dir1 x dir2 is zero"; throw e; }
        if (l.type == PPM_LIGHT_SUN) {
          double ld[3];
          need_vec(it, "ldir", "light", ln, ld);
          if (!normalize3(ld, l.dir)) { ParseError e; e.msg = "light at line " + std::to_string(ln) + ": ldir is zero"; throw e; }
        }
      } else {
        ParseError e; e.msg = "light at line " + std::to_string(ln) + ": unknown type '" + type + "'"; throw e;
      }
      sc->lights.push_back(l);
    }
  }

  // objects
  {
    Section& s = sections["object"];
    auto vertex = [&](const Item& it, const char* key, int ln, double out[3]) {
      const std::string& v = need(it, key, "object", ln);
      if (parse_vec3(v, out)) return;
      auto f = verts.find(v);
      if (f == verts.end()) { ParseError e; e.msg = "object at line " + std::to_string(ln) + ": unknown vertex '" + v + "'"; throw e; }
      out[0] = f->second[0]; out[1] = f->second[1]; out[2] = f->second[2];
    };
    for (size_t i = 0; i < s.items.size(); ++i) {
      const Item& it = s.items[i];
      int ln = s.lines[i];
      std::string type = need(it, "type", "object", ln);
      std::string mname = need(it, "material", "object", ln);
      auto mf = mat_index.find(mname);
      if (mf == mat_index.end()) { ParseError e; e.msg = "object at line " + std::to_string(ln) + ": unknown material '" + mname + "'"; throw e; }
      ppm_prim p;
      if (type == "plain") {
        double n[3], pos[3];
        need_vec(it, "normal", "object", ln, n);
        need_vec(it, "position", "object", ln, pos);
        // dist = -(normal . position); 0 - x so that a zero dot product gives +0.0
        // like the literal `dist: 0.0` of scene.rs:361
        double d = dot3(n, pos);
        ppm_prim_plain(&p, n, d == 0.0 ? 0.0 : -d, mf->second);
      } else if (type == "sphere") {
        double c[3];
        need_vec(it, "center", "object", ln, c);
        ppm_prim_sphere(&p, c, need_num(it, "radius", "object", ln), mf->second);
      } else if (type == "polygon" || type == "parallelogram") {
        double p0[3], p1[3], p2[3];
        vertex(it, "pos1", ln, p0); vertex(it, "pos2", ln, p1); vertex(it, "pos3", ln, p2);
        if (ppm_prim_polygon(&p, p0, p1, p2, type == "parallelogram", mf->second) != PPM_OK) {
          ParseError e; e.msg = "object at line " + std::to_string(ln) + ": degenerate " + type; throw e;
        }
      } else {
        ParseError e; e.msg = "object at line " + std::to_string(ln) + ": unknown type '" + type + "'"; throw e;
      }
      sc->prims.push_back(p);
      auto nf = it.find("name");
      sc->prim_names.push_back(nf == it.end() ? std::string() : nf->second);
    }
  }
  if (sc->lights.empty()) { ParseError e; e.msg = "scene has no light"; throw e; }
  if (sc->prims.empty()) { ParseError e; e.msg = "scene has no object"; throw e; }
}

void set_err(char* err, size_t errlen, const std::string& m) {
  if (err && errlen) { std::snprintf(err, errlen, "%s", m.c_str()); }
}

}  // namespace

extern "C" {

int ppm_scene_load(const char* path, ppm_scene** out, char* err, size_t errlen) {
  if (!path || !out) return PPM_ERR_ARG;
  std::ifstream f(path);
  if (!f) { set_err(err, errlen, std::string("cannot open ") + path); return PPM_ERR_IO; }
  ppm_scene* sc = new ppm_scene();
  try {
    parse_scene_text(f, sc);
  } catch (const ParseError& e) {
    set_err(err, errlen, std::string(path) + ": " + e.msg);
    delete sc;
    return PPM_ERR_PARSE;
  } catch (...) {
    set_err(err, errlen, std::string(path) + ": parse failure");
    delete sc;
    return PPM_ERR_PARSE;
  }
  *out = sc;
  return PPM_OK;
}

int ppm_camera_load(const char* path, ppm_camera* out, char* err, size_t errlen) {
  if (!path || !out) return PPM_ERR_ARG;
  std::ifstream f(path);
  if (!f) { set_err(err, errlen, std::string("cannot open ") + path); return PPM_ERR_IO; }
  ppm_camera c;
  ppm_camera_default(&c);
  std::string raw;
  int lineno = 0;
  while (std::getline(f, raw)) {
    ++lineno;
    std::string line = trim(strip_comment(raw));
    if (line.empty()) continue;
    size_t p = line.find(':');
    if (p == std::string::npos) { set_err(err, errlen, std::string(path) + ": line " + std::to_string(lineno) + ": expected 'key : value'"); return PPM_ERR_PARSE; }
    std::string key = trim(line.substr(0, p)), val = trim(line.substr(p + 1));
    std::string k;  // canonical key: drop underscores so both dialects match
    for (char ch : key) if (ch != '_') k.push_back((char)std::tolower((unsigned char)ch));
    bool ok = true;
    double d;
    int b;
    if (k == "xresolution") { ok = parse_double(val, &d); c.xreso = (int32_t)d; }
    else if (k == "yresolution") { ok = parse_double(val, &d); c.yreso = (int32_t)d; }
    else if (k == "progressive") { ok = parse_yesno(val, &b); c.progressive = b; }
    else if (k == "antialias") { ok = parse_yesno(val, &b); c.antialias = b; }
    else if (k == "useclassic") { ok = parse_yesno(val, &b); c.use_classic = b; }
    else if (k == "blur") { ok = parse_yesno(val, &b); c.blur = b; }
    else if (k == "estimateradius") { ok = parse_double(val, &d); c.radius = d * d; }
    else if (k == "maxradiance") { ok = parse_double(val, &c.max_radiance); }
    else if (k == "isosensitivity") { ok = parse_double(val, &c.iso_sens); }
    else if (k == "shutterspeed") { ok = parse_double(val, &c.shut_speed); }
    else if (k == "focallength") { ok = parse_double(val, &d); c.focal_len = d / 1000.0; }
    else if (k == "fnumber") { ok = parse_double(val, &c.f_number); }
    else if (k == "focus") { ok = parse_double(val, &c.focus); }
    else if (k == "ambient") { ok = parse_vec3(val, c.ambient); }
    else if (k == "eyeposition") { ok = parse_vec3(val, c.eye_pos); }
    else if (k == "targetposition") { ok = parse_vec3(val, c.target_pos); }
    else if (k == "upperdirection") { ok = parse_vec3(val, c.upper_dir); }
    else if (k == "photonfilter") {
      std::string v = val;
      for (auto& ch : v) ch = (char)std::tolower((unsigned char)ch);
      if (v == "none") c.pfilter = PPM_FILTER_NONE;
      else if (v == "cone") c.pfilter = PPM_FILTER_CONE;
      else if (v == "gauss") c.pfilter = PPM_FILTER_GAUSS;
      else ok = false;
    }
    else if (k == "samplephoton") { ok = parse_double(val, &d); c.n_sample_photon = (int32_t)d; }
    else if (k == "nphoton") { /* legacy key, commented out in camera.rs:24; the bins take it from argv */ }
    else { /* unknown keys are ignored */ }
    if (!ok) { set_err(err, errlen, std::string(path) + ": line " + std::to_string(lineno) + ": bad value for '" + key + "'"); return PPM_ERR_PARSE; }
  }
  if (ppm_camera_finalize(&c) != PPM_OK) { set_err(err, errlen, std::string(path) + ": degenerate camera"); return PPM_ERR_PARSE; }
  *out = c;
  return PPM_OK;
}

}  // extern "C"
