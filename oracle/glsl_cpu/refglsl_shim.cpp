// oracle/glsl_cpu/refglsl_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// C entry points of oracle/_ref/librefglsl.so: a glUniform* / glBindTexture / glDispatchCompute shaped API over the
// reference's shader programs compiled for the CPU (see glsl_emu.h).  The caller (tests) plays the role of the
// reference's renderer classes: it sets the uniforms by name exactly as their Update() / CreateRenderingPass() do.
#include "refglsl_runtime.h"
#include <atomic>
#include <cstring>
#include <string>

namespace glsl {
static thread_local UniformTable* g_current = nullptr;
UniformTable*& current_uniforms() { return g_current; }
static std::atomic<long> g_faults{0};
static std::string g_first_fault;
void unbound_sampler(const char* what) {
  if (g_faults.fetch_add(1) == 0) g_first_fault = what;
}
const UniformValue& lookup_uniform(const char* name) {
  static const UniformValue zero;
  UniformTable* t = current_uniforms();
  if (!t) return zero;
  t->declared[name] = 1;
  auto it = t->values.find(name);
  if (it == t->values.end()) { t->unset.push_back(name); return zero; }
  return it->second;
}
}  // namespace glsl

namespace refglsl {
static std::map<std::string, DispatchFn>& registry() { static std::map<std::string, DispatchFn> r; return r; }
void register_program(const char* name, DispatchFn fn) { registry()[name] = fn; }
struct Program {
  DispatchFn fn;
  glsl::UniformTable table;
};
}  // namespace refglsl

extern "C" {
void* rg_program_create(const char* name) {
  auto it = refglsl::registry().find(name);
  if (it == refglsl::registry().end()) return nullptr;
  refglsl::Program* p = new refglsl::Program();
  p->fn = it->second;
  return p;
}
void rg_program_destroy(void* p) { delete (refglsl::Program*)p; }
int rg_program_names(char* buf, int cap) {
  std::string s;
  for (auto& kv : refglsl::registry()) s += kv.first + " ";
  std::snprintf(buf, cap, "%s", s.c_str());
  return (int)refglsl::registry().size();
}
void rg_set_f(void* p, const char* name, const float* v, int n) {
  glsl::UniformValue& u = ((refglsl::Program*)p)->table.values[name];
  u.f.assign(v, v + n);
}
void rg_set_i(void* p, const char* name, const int* v, int n) {
  glsl::UniformValue& u = ((refglsl::Program*)p)->table.values[name];
  u.i.assign(v, v + n);
}
// A texture borrows the caller's arrays (fp32 values already rounded to the internal format, x fastest, channels
// interleaved); dims 1 / 2 / 3; whd = 3 ints per level.
void* rg_texture_create(int dims, int channels, int nlevels, const int* whd, const float* const* level_data) {
  glsl::Texture* t = new glsl::Texture();
  t->dims = dims; t->channels = channels;
  for (int l = 0; l < nlevels; ++l) {
    orc::Tex3D L; L.w = whd[3 * l]; L.h = whd[3 * l + 1]; L.d = whd[3 * l + 2]; L.c = channels; L.data = level_data[l];
    t->mip.levels.push_back(L);
  }
  return t;
}
void rg_texture_destroy(void* t) { delete (glsl::Texture*)t; }
void rg_set_texture(void* p, const char* name, void* tex) { ((refglsl::Program*)p)->table.values[name].tex = (glsl::Texture*)tex; }
void* rg_image_create(float* data, int w, int h, int d, int channels, int half_storage) {
  glsl::Image* i = new glsl::Image();
  i->data = data; i->w = w; i->h = h; i->d = d; i->channels = channels; i->half_storage = half_storage != 0;
  return i;
}
void rg_image_destroy(void* i) { delete (glsl::Image*)i; }
void rg_set_image(void* p, const char* name, void* img) { ((refglsl::Program*)p)->table.values[name].img = (glsl::Image*)img; }

// glDispatchCompute: returns the number of sampler faults (unbound sampler used, texelFetch out of range) seen.
long rg_dispatch(void* p, const int* groups3, const int* local3) {
  refglsl::Program* P = (refglsl::Program*)p;
  glsl::g_faults = 0; glsl::g_first_fault.clear();
  P->table.unset.clear();
  P->fn(&P->table, groups3, local3);
  return glsl::g_faults.load();
}
int rg_first_fault(char* buf, int cap) { std::snprintf(buf, cap, "%s", glsl::g_first_fault.c_str()); return (int)glsl::g_first_fault.size(); }
// uniforms the linked shaders declare but the host never set (they read as 0, as in GL); valid after a dispatch
int rg_unset_uniforms(void* p, char* buf, int cap) {
  refglsl::Program* P = (refglsl::Program*)p;
  std::string s;
  for (auto& n : P->table.unset) s += n + " ";
  std::snprintf(buf, cap, "%s", s.c_str());
  return (int)P->table.unset.size();
}
// uniforms the host set that no linked shader declares (glGetUniformLocation would return -1); valid after a dispatch
int rg_unknown_uniforms(void* p, char* buf, int cap) {
  refglsl::Program* P = (refglsl::Program*)p;
  std::string s; int n = 0;
  for (auto& kv : P->table.values) if (!P->table.declared.count(kv.first)) { s += kv.first + " "; ++n; }
  std::snprintf(buf, cap, "%s", s.c_str());
  return n;
}
}
