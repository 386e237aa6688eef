// oracle/glsl_cpu/glsl_emu.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A small GLSL 4.30 compute-shader execution environment for the CPU: enough of the language's types, swizzles,
// built-ins, samplers and images to compile the REFERENCE's own .comp sources (rewritten mechanically by
// gen_shader_cpp.py: qualifiers only, no logic) as member functions of a C++ struct and to run one invocation per
// pixel.  The shader logic that runs is the reference's, from /root/reference; what is ours is this environment:
//   * fp32 arithmetic, one rounding per GLSL operator, no contraction (compiled with -ffp-contract=off);
//   * samplers follow the OpenGL 4.3 rules of SURVEY.md A.1 through the oracle's own sampler (oracle_common.h:
//     GL_LINEAR / clamp-to-edge, fused lerp, mip levels), so that a comparison with the oracle isolates shader logic;
//   * imageStore rounds to the image's internal format (fp16 for rgba16f / rg16f), imageLoad returns the stored value.
// Built into oracle/_ref/librefglsl.so by `make -C oracle refglsl`; only tests load it.
#pragma once
#include "../oracle_common.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

namespace glsl {

typedef unsigned int uint;

// ---------------------------------------------------------------- swizzle proxies
// A proxy aliases the N scalars of its owner (it lives in a union with them) and selects the components I...
template <class V, class T, int N, int... I>
struct swz {
  T d[N];
  operator V() const { return V(d[I]...); }
  swz& operator=(const V& v) { assign(v, 0, I...); return *this; }
  swz& operator=(const swz& o) { V v = o; return *this = v; }
  template <int... J> swz& operator=(const swz<V, T, N, J...>& o) { V v = o; return *this = v; }
  swz& operator+=(const V& v) { return *this = V(*this) + v; }
  swz& operator-=(const V& v) { return *this = V(*this) - v; }
  swz& operator*=(const V& v) { return *this = V(*this) * v; }
  swz& operator*=(T s) { return *this = V(*this) * s; }
  swz& operator/=(T s) { return *this = V(*this) / s; }
 private:
  template <class... R> void assign(const V& v, int k, int i, R... rest) { d[i] = v[k]; assign(v, k + 1, rest...); }
  void assign(const V&, int) {}
};

struct ivec2; struct ivec3; struct uvec2; struct uvec3;

struct vec2 {
  union {
    struct { float x, y; };
    struct { float r, g; };
    swz<vec2, float, 2, 0, 1> xy, rg;
    swz<vec2, float, 2, 1, 0> yx;
  };
  vec2() : x(0), y(0) {}
  explicit vec2(float s) : x(s), y(s) {}
  vec2(float a, float b) : x(a), y(b) {}
  vec2(const vec2& o) : x(o.x), y(o.y) {}
  explicit vec2(const ivec2& v);
  vec2& operator=(const vec2& o) { x = o.x; y = o.y; return *this; }
  float& operator[](int i) { return (&x)[i]; }
  const float& operator[](int i) const { return (&x)[i]; }
};

struct vec3 {
  union {
    struct { float x, y, z; };
    struct { float r, g, b; };
    swz<vec3, float, 3, 0, 1, 2> xyz, rgb;
    swz<vec2, float, 3, 0, 1> xy, rg;
    swz<vec2, float, 3, 1, 2> yz;
    swz<vec2, float, 3, 0, 2> xz;
  };
  vec3() : x(0), y(0), z(0) {}
  explicit vec3(float s) : x(s), y(s), z(s) {}
  vec3(float a, float b, float c) : x(a), y(b), z(c) {}
  vec3(const vec2& v, float c) : x(v.x), y(v.y), z(c) {}
  vec3(float a, const vec2& v) : x(a), y(v.x), z(v.y) {}
  vec3(const vec3& o) : x(o.x), y(o.y), z(o.z) {}
  vec3(const ivec3& v);                                   // GLSL converts ivec3 -> vec3 implicitly
  vec3& operator=(const vec3& o) { x = o.x; y = o.y; z = o.z; return *this; }
  float& operator[](int i) { return (&x)[i]; }
  const float& operator[](int i) const { return (&x)[i]; }
};

struct vec4 {
  union {
    struct { float x, y, z, w; };
    struct { float r, g, b, a; };
    swz<vec4, float, 4, 0, 1, 2, 3> xyzw, rgba;
    swz<vec3, float, 4, 0, 1, 2> xyz, rgb;
    swz<vec2, float, 4, 0, 1> xy, rg;
    swz<vec2, float, 4, 2, 3> zw, ba;
  };
  vec4() : x(0), y(0), z(0), w(0) {}
  explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
  vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
  vec4(const vec3& v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
  vec4(const vec2& v, float c, float d) : x(v.x), y(v.y), z(c), w(d) {}
  vec4(const vec4& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
  vec4& operator=(const vec4& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
  float& operator[](int i) { return (&x)[i]; }
  const float& operator[](int i) const { return (&x)[i]; }
};

struct ivec2 {
  union { struct { int x, y; }; swz<ivec2, int, 2, 0, 1> xy; };
  ivec2() : x(0), y(0) {}
  explicit ivec2(int s) : x(s), y(s) {}
  ivec2(int a, int b) : x(a), y(b) {}
  ivec2(const ivec2& o) : x(o.x), y(o.y) {}
  explicit ivec2(const uvec2& v);
  explicit ivec2(const vec2& v) : x((int)v.x), y((int)v.y) {}
  ivec2& operator=(const ivec2& o) { x = o.x; y = o.y; return *this; }
  int& operator[](int i) { return (&x)[i]; }
  const int& operator[](int i) const { return (&x)[i]; }
};
struct uvec2 {
  uint x, y;
  uvec2() : x(0), y(0) {}
  uvec2(uint a, uint b) : x(a), y(b) {}
  uint operator[](int i) const { return (&x)[i]; }
};
struct ivec3 {
  union { struct { int x, y, z; }; swz<ivec3, int, 3, 0, 1, 2> xyz; swz<ivec2, int, 3, 0, 1> xy; };
  ivec3() : x(0), y(0), z(0) {}
  explicit ivec3(int s) : x(s), y(s), z(s) {}
  ivec3(int a, int b, int c) : x(a), y(b), z(c) {}
  ivec3(const ivec3& o) : x(o.x), y(o.y), z(o.z) {}
  explicit ivec3(const vec3& v) : x((int)v.x), y((int)v.y), z((int)v.z) {}     // float -> int truncates toward zero
  explicit ivec3(const uvec3& v);
  ivec3& operator=(const ivec3& o) { x = o.x; y = o.y; z = o.z; return *this; }
  int& operator[](int i) { return (&x)[i]; }
  const int& operator[](int i) const { return (&x)[i]; }
};
struct uvec3 {
  union { struct { uint x, y, z; }; swz<uvec2, uint, 3, 0, 1> xy; swz<uvec3, uint, 3, 0, 1, 2> xyz; };
  uvec3() : x(0), y(0), z(0) {}
  uvec3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
  uvec3(const uvec3& o) : x(o.x), y(o.y), z(o.z) {}
  uvec3& operator=(const uvec3& o) { x = o.x; y = o.y; z = o.z; return *this; }
  uint operator[](int i) const { return (&x)[i]; }
};
inline ivec3::ivec3(const uvec3& v) : x((int)v.x), y((int)v.y), z((int)v.z) {}
inline vec2::vec2(const ivec2& v) : x((float)v.x), y((float)v.y) {}
inline vec3::vec3(const ivec3& v) : x((float)v.x), y((float)v.y), z((float)v.z) {}
inline ivec2::ivec2(const uvec2& v) : x((int)v.x), y((int)v.y) {}

// ---------------------------------------------------------------- operators (component-wise, one rounding each)
#define GLSL_VEC_OPS(V, N)                                                                                              \
  inline V operator+(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] + b[i]; return r; }         \
  inline V operator-(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] - b[i]; return r; }         \
  inline V operator*(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] * b[i]; return r; }         \
  inline V operator/(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] / b[i]; return r; }         \
  inline V operator+(const V& a, float s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] + s; return r; }               \
  inline V operator-(const V& a, float s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] - s; return r; }               \
  inline V operator*(const V& a, float s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] * s; return r; }               \
  inline V operator/(const V& a, float s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] / s; return r; }               \
  inline V operator+(float s, const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = s + a[i]; return r; }               \
  inline V operator-(float s, const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = s - a[i]; return r; }               \
  inline V operator*(float s, const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = s * a[i]; return r; }               \
  inline V operator/(float s, const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = s / a[i]; return r; }               \
  inline V operator-(const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = -a[i]; return r; }                           \
  inline V& operator+=(V& a, const V& b) { a = a + b; return a; }                                                       \
  inline V& operator-=(V& a, const V& b) { a = a - b; return a; }                                                       \
  inline V& operator*=(V& a, const V& b) { a = a * b; return a; }                                                       \
  inline V& operator/=(V& a, const V& b) { a = a / b; return a; }                                                       \
  inline V& operator+=(V& a, float s) { a = a + s; return a; }                                                          \
  inline V& operator-=(V& a, float s) { a = a - s; return a; }                                                          \
  inline V& operator*=(V& a, float s) { a = a * s; return a; }                                                          \
  inline V& operator/=(V& a, float s) { a = a / s; return a; }                                                          \
  inline bool operator==(const V& a, const V& b) { for (int i = 0; i < N; ++i) if (!(a[i] == b[i])) return false; return true; } \
  inline bool operator!=(const V& a, const V& b) { return !(a == b); }                                                  \
  inline V min(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = b[i] < a[i] ? b[i] : a[i]; return r; } \
  inline V max(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] < b[i] ? b[i] : a[i]; return r; } \
  inline V min(const V& a, float b) { V r; for (int i = 0; i < N; ++i) r[i] = b < a[i] ? b : a[i]; return r; }          \
  inline V max(const V& a, float b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] < b ? b : a[i]; return r; }          \
  inline V clamp(const V& a, const V& lo, const V& hi) { return min(max(a, lo), hi); }                                  \
  inline V clamp(const V& a, float lo, float hi) { return min(max(a, lo), hi); }                                        \
  inline V abs(const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = std::fabs(a[i]); return r; }                       \
  inline V floor(const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = std::floor(a[i]); return r; }                    \
  inline V ceil(const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = std::ceil(a[i]); return r; }                      \
  inline V exp(const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = std::exp(a[i]); return r; }                        \
  inline V mix(const V& a, const V& b, float t) { return a * (1.0f - t) + b * t; }
GLSL_VEC_OPS(vec2, 2)
GLSL_VEC_OPS(vec3, 3)
GLSL_VEC_OPS(vec4, 4)
#undef GLSL_VEC_OPS

// dot / length / normalize: sums left to right as written out by a scalarising compiler; normalize = v * inversesqrt(dot)
// (the same spelling as glm, include/glm/detail/func_geometric.inl:257-266, and as the oracle's orc::normalize).
inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(const vec4& a, const vec4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline float length(const vec2& a) { return std::sqrt(dot(a, a)); }
inline float length(const vec3& a) { return std::sqrt(dot(a, a)); }
inline float distance(const vec3& a, const vec3& b) { return length(a - b); }
inline vec2 normalize(const vec2& a) { return a * (1.0f / std::sqrt(dot(a, a))); }
inline vec3 normalize(const vec3& a) { return a * (1.0f / std::sqrt(dot(a, a))); }
inline vec3 cross(const vec3& a, const vec3& b) {
  return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}

// scalar built-ins (fp32).  min / max follow the GLSL definitions (y < x ? y : x and x < y ? y : x).
inline float min(float a, float b) { return b < a ? b : a; }
inline float max(float a, float b) { return a < b ? b : a; }
inline int min(int a, int b) { return b < a ? b : a; }
inline int max(int a, int b) { return a < b ? b : a; }
inline float min(int a, float b) { return min((float)a, b); }
inline float max(int a, float b) { return max((float)a, b); }
inline float min(float a, int b) { return min(a, (float)b); }
inline float max(float a, int b) { return max(a, (float)b); }
inline float clamp(float a, float lo, float hi) { return min(max(a, lo), hi); }
inline int clamp(int a, int lo, int hi) { return min(max(a, lo), hi); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline float abs(float a) { return std::fabs(a); }
inline float exp(float a) { return std::exp(a); }
inline float exp2(float a) { return std::exp2(a); }
inline float log(float a) { return std::log(a); }
inline float log2(float a) { return std::log2(a); }
inline float sqrt(float a) { return std::sqrt(a); }
inline float inversesqrt(float a) { return 1.0f / std::sqrt(a); }
inline float pow(float a, float b) { return std::pow(a, b); }
inline float sin(float a) { return std::sin(a); }
inline float cos(float a) { return std::cos(a); }
inline float tan(float a) { return std::tan(a); }
inline float acos(float a) { return std::acos(a); }
inline float floor(float a) { return std::floor(a); }
inline float ceil(float a) { return std::ceil(a); }
inline float fract(float a) { return a - std::floor(a); }
inline float sign(float a) { return a > 0.0f ? 1.0f : (a < 0.0f ? -1.0f : 0.0f); }
inline float radians(float d) { return d * 0.017453292519943295f; }

inline ivec3 operator+(const ivec3& a, int s) { return ivec3(a.x + s, a.y + s, a.z + s); }
inline ivec3 operator-(const ivec3& a, int s) { return ivec3(a.x - s, a.y - s, a.z - s); }
inline ivec3 operator+(const ivec3& a, const ivec3& b) { return ivec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline ivec3 operator-(const ivec3& a, const ivec3& b) { return ivec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline ivec3 min(const ivec3& a, const ivec3& b) { return ivec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline ivec3 max(const ivec3& a, const ivec3& b) { return ivec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline ivec3 clamp(const ivec3& a, const ivec3& lo, const ivec3& hi) { return min(max(a, lo), hi); }
inline ivec2 operator+(const ivec2& a, const ivec2& b) { return ivec2(a.x + b.x, a.y + b.y); }
inline ivec2 operator-(const ivec2& a, const ivec2& b) { return ivec2(a.x - b.x, a.y - b.y); }
inline ivec2 operator*(const ivec2& a, int s) { return ivec2(a.x * s, a.y * s); }
inline ivec2 operator/(const ivec2& a, int s) { return ivec2(a.x / s, a.y / s); }
inline ivec2 operator+(const ivec2& a, int s) { return ivec2(a.x + s, a.y + s); }

// ---------------------------------------------------------------- matrices (column major, like GLSL and glm)
struct mat4 {
  float m[16];
  mat4() { for (int i = 0; i < 16; ++i) m[i] = 0.0f; }
};
struct mat3 {
  float c[3][3];                                   // c[column][row]
  mat3() { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) c[i][j] = 0.0f; }
  explicit mat3(const mat4& M) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) c[i][j] = M.m[4 * i + j]; }
};
// row vector * matrix: component i = dot(v, column i) (GLSL 4.30 section 5.10)
inline vec3 operator*(const vec3& v, const mat3& M) {
  return vec3(v.x * M.c[0][0] + v.y * M.c[0][1] + v.z * M.c[0][2],
              v.x * M.c[1][0] + v.y * M.c[1][1] + v.z * M.c[1][2],
              v.x * M.c[2][0] + v.y * M.c[2][1] + v.z * M.c[2][2]);
}
inline vec3 operator*(const mat3& M, const vec3& v) {
  return vec3(M.c[0][0] * v.x + M.c[1][0] * v.y + M.c[2][0] * v.z,
              M.c[0][1] * v.x + M.c[1][1] * v.y + M.c[2][1] * v.z,
              M.c[0][2] * v.x + M.c[1][2] * v.y + M.c[2][2] * v.z);
}
inline vec4 operator*(const mat4& M, const vec4& v) {
  vec4 r;
  for (int j = 0; j < 4; ++j) r[j] = M.m[j] * v.x + M.m[4 + j] * v.y + M.m[8 + j] * v.z + M.m[12 + j] * v.w;
  return r;
}

// ---------------------------------------------------------------- textures and images
struct Texture {                 // what a sampler is bound to
  orc::Tex3DMip mip;             // levels[0] is the base; 1-D textures use levels[0] with h = d = 1
  int channels = 1;
  int dims = 3;
};
struct sampler3D { const Texture* t = nullptr; };
struct sampler2D { const Texture* t = nullptr; };
struct sampler1D { const Texture* t = nullptr; };

// Missing components read as (0, 0, 0, 1) (OpenGL 4.3 table 8.15 texture base formats RED / RG / RGB).
inline vec4 fill4(const float* v, int c) { return vec4(v[0], c > 1 ? v[1] : 0.0f, c > 2 ? v[2] : 0.0f, c > 3 ? v[3] : 1.0f); }
void unbound_sampler(const char* what);

inline vec4 texture(const sampler3D& s, const vec3& p) {
  if (!s.t) { unbound_sampler("sampler3D"); return vec4(0, 0, 0, 1); }
  float v[4] = {0, 0, 0, 1};
  for (int ch = 0; ch < s.t->channels; ++ch) v[ch] = orc::tex3d(s.t->mip.levels[0], orc::v3(p.x, p.y, p.z), ch);
  return fill4(v, s.t->channels);
}
inline vec4 textureLod(const sampler3D& s, const vec3& p, float lod) {
  if (!s.t) { unbound_sampler("sampler3D"); return vec4(0, 0, 0, 1); }
  float v[4] = {0, 0, 0, 1};
  for (int ch = 0; ch < s.t->channels; ++ch) v[ch] = s.t->mip.lod(orc::v3(p.x, p.y, p.z), lod, ch);
  return fill4(v, s.t->channels);
}
inline vec4 texelFetch(const sampler3D& s, const ivec3& p, int level) {
  if (!s.t) { unbound_sampler("sampler3D"); return vec4(0, 0, 0, 1); }
  const orc::Tex3D& L = s.t->mip.levels[level];
  float v[4] = {0, 0, 0, 1};
  // out-of-range texelFetch is undefined in GL; the on-path shaders clamp before they fetch (checked here)
  if (p.x < 0 || p.y < 0 || p.z < 0 || p.x >= L.w || p.y >= L.h || p.z >= L.d) { unbound_sampler("texelFetch out of range"); return fill4(v, 1); }
  for (int ch = 0; ch < s.t->channels; ++ch) v[ch] = L.at(p.x, p.y, p.z, ch);
  return fill4(v, s.t->channels);
}
inline vec4 texture(const sampler1D& s, float p) {
  if (!s.t) { unbound_sampler("sampler1D"); return vec4(0, 0, 0, 1); }
  const orc::Tex3D& L = s.t->mip.levels[0];
  int i0, i1; float f;
  orc::lin_coord(p, L.w, &i0, &i1, &f);
  float v[4] = {0, 0, 0, 1};
  for (int ch = 0; ch < s.t->channels; ++ch) v[ch] = orc::lerp(L.at(i0, 0, 0, ch), L.at(i1, 0, 0, ch), f);
  return fill4(v, s.t->channels);
}
inline vec4 texelFetch(const sampler1D& s, int i, int level) {
  if (!s.t) { unbound_sampler("sampler1D"); return vec4(0, 0, 0, 1); }
  const orc::Tex3D& L = s.t->mip.levels[level];
  float v[4] = {0, 0, 0, 1};
  if (i < 0 || i >= L.w) { unbound_sampler("texelFetch out of range"); return fill4(v, 1); }
  for (int ch = 0; ch < s.t->channels; ++ch) v[ch] = L.at(i, 0, 0, ch);
  return fill4(v, s.t->channels);
}
inline vec4 texture(const sampler2D& s, const vec2& p) {
  if (!s.t) { unbound_sampler("sampler2D"); return vec4(0, 0, 0, 1); }
  const orc::Tex3D& L = s.t->mip.levels[0];
  int x0, x1, y0, y1; float fx, fy;
  orc::lin_coord(p.x, L.w, &x0, &x1, &fx);
  orc::lin_coord(p.y, L.h, &y0, &y1, &fy);
  float v[4] = {0, 0, 0, 1};
  for (int ch = 0; ch < s.t->channels; ++ch)
    v[ch] = orc::lerp(orc::lerp(L.at(x0, y0, 0, ch), L.at(x1, y0, 0, ch), fx), orc::lerp(L.at(x0, y1, 0, ch), L.at(x1, y1, 0, ch), fx), fy);
  return fill4(v, s.t->channels);
}

// texelFetch outside a 2-D texture is undefined in GL 4.3 without robust buffer access; the frame filters rely on it at the
// image borders.  Zero is returned, the convention the oracle (oracle_frame.cpp) and the product's kernels document.
inline vec4 texelFetch(const sampler2D& s, const ivec2& p, int level) {
  if (!s.t) { unbound_sampler("sampler2D"); return vec4(0, 0, 0, 1); }
  const orc::Tex3D& L = s.t->mip.levels[level];
  if (p.x < 0 || p.y < 0 || p.x >= L.w || p.y >= L.h) return vec4(0, 0, 0, 0);
  float v[4] = {0, 0, 0, 1};
  for (int ch = 0; ch < s.t->channels; ++ch) v[ch] = L.at(p.x, p.y, 0, ch);
  return fill4(v, s.t->channels);
}

struct Image {                    // what an image unit is bound to (one level of a 2-D or 3-D texture)
  float* data = nullptr;          // w * h * d * channels floats, x fastest; row 0 = bottom (GL image coordinates)
  int w = 0, h = 0, d = 1, channels = 4;
  bool half_storage = true;       // rgba16f / rg16f / r16f
};
struct image2D { Image* i = nullptr; };
inline ivec2 imageSize(const image2D& im) { return im.i ? ivec2(im.i->w, im.i->h) : ivec2(0, 0); }
inline void imageStore(const image2D& im, const ivec2& p, const vec4& v) {
  if (!im.i || p.x < 0 || p.y < 0 || p.x >= im.i->w || p.y >= im.i->h) return;      // out-of-bounds stores are dropped
  float* o = im.i->data + ((size_t)p.y * im.i->w + p.x) * im.i->channels;
  for (int c = 0; c < im.i->channels; ++c) o[c] = im.i->half_storage ? orc::round_f16(v[c]) : v[c];
}
inline vec4 imageLoad(const image2D& im, const ivec2& p) {
  float v[4] = {0, 0, 0, 1};
  if (!im.i || p.x < 0 || p.y < 0 || p.x >= im.i->w || p.y >= im.i->h) return vec4(0, 0, 0, 0);
  const float* o = im.i->data + ((size_t)p.y * im.i->w + p.x) * im.i->channels;
  for (int c = 0; c < im.i->channels; ++c) v[c] = o[c];
  return fill4(v, im.i->channels);
}

struct image3D { Image* i = nullptr; };
inline ivec3 imageSize(const image3D& im) { return im.i ? ivec3(im.i->w, im.i->h, im.i->d) : ivec3(0, 0, 0); }
inline bool inside(const Image* i, const ivec3& p) { return i && p.x >= 0 && p.y >= 0 && p.z >= 0 && p.x < i->w && p.y < i->h && p.z < i->d; }
inline void imageStore(const image3D& im, const ivec3& p, const vec4& v) {
  if (!inside(im.i, p)) return;
  float* o = im.i->data + (((size_t)p.z * im.i->h + p.y) * im.i->w + p.x) * im.i->channels;
  for (int c = 0; c < im.i->channels; ++c) o[c] = im.i->half_storage ? orc::round_f16(v[c]) : v[c];
}
inline vec4 imageLoad(const image3D& im, const ivec3& p) {
  float v[4] = {0, 0, 0, 1};
  if (!inside(im.i, p)) return vec4(0, 0, 0, 0);
  const float* o = im.i->data + (((size_t)p.z * im.i->h + p.y) * im.i->w + p.x) * im.i->channels;
  for (int c = 0; c < im.i->channels; ++c) v[c] = o[c];
  return fill4(v, im.i->channels);
}

// ---------------------------------------------------------------- uniforms: looked up by name when a program object is
// instantiated (the values the host set with glUniform* before the dispatch; unset uniforms read as 0 like in GL)
struct UniformValue {
  std::vector<float> f;           // floats (float, vecN, matN and arrays of them, flattened)
  std::vector<int> i;             // ints (int and arrays)
  const Texture* tex = nullptr;
  Image* img = nullptr;
};
struct UniformTable {
  std::map<std::string, UniformValue> values;
  std::map<std::string, int> declared;     // name -> 1 when some shader of the program declares it
  std::vector<std::string> unset;           // declared but never set by the host
};
UniformTable*& current_uniforms();

template <class T> struct uniform_reader;
template <> struct uniform_reader<float> { static float get(const UniformValue& v, int k) { return (size_t)k < v.f.size() ? v.f[k] : ((size_t)k < v.i.size() ? (float)v.i[k] : 0.0f); } };
template <> struct uniform_reader<int> { static int get(const UniformValue& v, int k) { return (size_t)k < v.i.size() ? v.i[k] : ((size_t)k < v.f.size() ? (int)v.f[k] : 0); } };
template <> struct uniform_reader<vec2> { static vec2 get(const UniformValue& v, int k) { return (size_t)(2 * k + 1) < v.f.size() ? vec2(v.f[2 * k], v.f[2 * k + 1]) : vec2(); } };
template <> struct uniform_reader<vec3> { static vec3 get(const UniformValue& v, int k) { return (size_t)(3 * k + 2) < v.f.size() ? vec3(v.f[3 * k], v.f[3 * k + 1], v.f[3 * k + 2]) : vec3(); } };
template <> struct uniform_reader<vec4> { static vec4 get(const UniformValue& v, int k) { return (size_t)(4 * k + 3) < v.f.size() ? vec4(v.f[4 * k], v.f[4 * k + 1], v.f[4 * k + 2], v.f[4 * k + 3]) : vec4(); } };
template <> struct uniform_reader<mat4> { static mat4 get(const UniformValue& v, int k) { mat4 m; if ((size_t)(16 * k + 15) < v.f.size()) for (int i = 0; i < 16; ++i) m.m[i] = v.f[16 * k + i]; return m; } };
template <> struct uniform_reader<sampler3D> { static sampler3D get(const UniformValue& v, int) { sampler3D s; s.t = v.tex; return s; } };
template <> struct uniform_reader<sampler2D> { static sampler2D get(const UniformValue& v, int) { sampler2D s; s.t = v.tex; return s; } };
template <> struct uniform_reader<sampler1D> { static sampler1D get(const UniformValue& v, int) { sampler1D s; s.t = v.tex; return s; } };
template <> struct uniform_reader<image2D> { static image2D get(const UniformValue& v, int) { image2D s; s.i = v.img; return s; } };
template <> struct uniform_reader<image3D> { static image3D get(const UniformValue& v, int) { image3D s; s.i = v.img; return s; } };

const UniformValue& lookup_uniform(const char* name);
template <class T> inline T U(const char* name) { return uniform_reader<T>::get(lookup_uniform(name), 0); }
template <class T, int N> struct arr {
  T v[N];
  T& operator[](int i) { return v[i]; }
  const T& operator[](int i) const { return v[i]; }
};
template <class T, int N> inline arr<T, N> UA(const char* name) {
  arr<T, N> a;
  const UniformValue& u = lookup_uniform(name);
  for (int k = 0; k < N; ++k) a.v[k] = uniform_reader<T>::get(u, k);
  return a;
}

// one invocation's built-in inputs
struct Invocation {
  uvec3 gl_GlobalInvocationID;
};

}  // namespace glsl
