// oracle/oracle_common.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the OpenGL 4.3 sampler semantics and of the small GLSL helpers that
// every on-path shader of lquatrin/cpp_volume_rendering shares.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load anything built from oracle/.
//
// PARITY STATUS: the reference ships no tests, golden images or known-answer vectors and no GL exists in
// this container (F4).  What pins this restatement: (1) the pieces of the reference that compile here
// (SummedAreaTable3D, TransferFunction1D, ConeGaussianSampler, gradient generators, VCT pre-passes, colour difference,
// PVM/DDS reader)
// are built into oracle/_ref/libref.so and must agree bit for bit (tests/test_oracle_ref.py, test_dos.py,
// test_gradient.py, test_eval_harness.py, test_pvm_dds.py); (2) the reference's own GLSL compute shaders are compiled
// for the CPU (oracle/glsl_cpu -> oracle/_ref/librefglsl.so) and every marcher / light cache / pyramid / filter of this
// oracle must produce the same fp16 values (tests/test_refglsl.py).  Still "parity unpinned": whatever GL leaves to the
// driver (filter precision, exp/pow).
//
// Citations are relative to /root/reference.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector>

namespace orc {

// ---------------------------------------------------------------------------------------------
// fp16 storage rounding (GL_R16F / GL_RGBA16F uploads and imageStore).
// Round-to-nearest-even, IEEE binary16 with subnormals, overflow -> inf.
// ---------------------------------------------------------------------------------------------
static inline uint16_t f32_to_f16_bits(float f) {
  uint32_t x; std::memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t ax = x & 0x7fffffffu;
  if (ax >= 0x7f800000u) {                       // inf / nan
    return (uint16_t)(sign | 0x7c00u | ((ax > 0x7f800000u) ? 0x200u : 0u));
  }
  if (ax >= 0x477ff000u) {                       // >= 65520 rounds to inf
    return (uint16_t)(sign | 0x7c00u);
  }
  if (ax < 0x38800000u) {                        // subnormal half or zero (|f| < 2^-14)
    if (ax < 0x33000000u) return (uint16_t)sign; // < 2^-25 -> 0 (2^-25 itself ties to even = 0)
    int e = (int)(ax >> 23);                     // biased exponent (>= 102)
    uint32_t m = (ax & 0x7fffffu) | 0x800000u;   // 24-bit significand
    int shift = 126 - e;                         // bits to drop so that unit = 2^-24
    uint32_t q = m >> shift;
    uint32_t rem = m & ((1u << shift) - 1u);
    uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) q++;
    return (uint16_t)(sign | q);
  }
  uint32_t e = (ax >> 23) - 112u;                // rebias 127 -> 15
  uint32_t m = ax & 0x7fffffu;
  uint32_t q = (e << 10) | (m >> 13);
  uint32_t rem = m & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (q & 1u))) q++;  // carry may bump the exponent: correct
  return (uint16_t)(sign | q);
}

static inline float f16_bits_to_f32(uint16_t h) {
  uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
  uint32_t e = (h >> 10) & 0x1fu;
  uint32_t m = h & 0x3ffu;
  uint32_t x;
  if (e == 0) {
    if (m == 0) x = sign;
    else {                                       // subnormal: value = m * 2^-24
      float v = (float)m * 5.9604644775390625e-8f;
      std::memcpy(&x, &v, 4); x |= sign;
    }
  } else if (e == 31) {
    x = sign | 0x7f800000u | (m << 13);
  } else {
    x = sign | ((e + 112u) << 23) | (m << 13);
  }
  float f; std::memcpy(&f, &x, 4); return f;
}

static inline float round_f16(float f) { return f16_bits_to_f32(f32_to_f16_bits(f)); }

// ---------------------------------------------------------------------------------------------
// tiny vector type (GLSL vec3 semantics, fp32, evaluation order as written in the shaders)
// ---------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
static inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
static inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline V3 operator/(V3 a, V3 b) { return V3{a.x / b.x, a.y / b.y, a.z / b.z}; }
static inline V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
static inline V3 operator*(float s, V3 a) { return V3{a.x * s, a.y * s, a.z * s}; }
static inline V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
static inline V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) {
  return V3{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
// GLSL normalize(); glm spells it x * inversesqrt(dot(x,x)) (include/glm/detail/func_geometric.inl:257-266)
static inline V3 normalize(V3 a) { float r = 1.0f / std::sqrt(dot(a, a)); return a * r; }
static inline V3 vmin(V3 a, V3 b) { return V3{std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)}; }
static inline V3 vmax(V3 a, V3 b) { return V3{std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)}; }
static inline V3 vabs(V3 a) { return V3{std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}; }
static inline V3 vclamp(V3 a, V3 lo, V3 hi) { return vmin(vmax(a, lo), hi); }
static inline float length(V3 a) { return std::sqrt(dot(a, a)); }

struct V4 { float x, y, z, w; };

// ---------------------------------------------------------------------------------------------
// Camera uniforms exactly as the reference uploads them (rc1prenderer.cpp:91-101,
// ebsrenderer.cpp:206-219): eye, glm::lookAt matrix (column major), tan(fovy/2), aspect.
// ---------------------------------------------------------------------------------------------
struct Camera {
  float eye[3];
  float lookat[16];   // column-major, glm::lookAt(eye, center, up)
  float tan_fovy;     // (float)tan(DEGREE_TO_RADIANS(fovy)/2.0)
  float aspect;       // float(w)/float(h)
};

// glm::lookAt restated (include/glm/gtc/matrix_transform.inl:403-428), fp32.
static inline void look_at(const float eye[3], const float center[3], const float up[3], float out[16]) {
  V3 e = v3(eye[0], eye[1], eye[2]), c = v3(center[0], center[1], center[2]), u0 = v3(up[0], up[1], up[2]);
  V3 f = normalize(c - e);
  V3 s = normalize(cross(f, u0));
  V3 u = cross(s, f);
  for (int i = 0; i < 16; ++i) out[i] = 0.0f;
  out[15] = 1.0f;
  out[0] = s.x; out[4] = s.y; out[8] = s.z;
  out[1] = u.x; out[5] = u.y; out[9] = u.z;
  out[2] = -f.x; out[6] = -f.y; out[10] = -f.z;
  out[12] = -dot(s, e); out[13] = -dot(u, e); out[14] = dot(f, e);
}

// GLSL `v * mat3(M)`: row vector times matrix = dot products with the COLUMNS of M.
static inline V3 vec_times_mat3(V3 v, const float m[16]) {
  return V3{v.x * m[0] + v.y * m[1] + v.z * m[2],
            v.x * m[4] + v.y * m[5] + v.z * m[6],
            v.x * m[8] + v.y * m[9] + v.z * m[10]};
}

// Pixel -> camera ray direction (ray_marching_1p.comp:93-99; same in every lit shader).
static inline V3 pixel_ray_dir(const Camera& cam, int px, int py, int W, int H) {
  float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
  float vx = (fx / (float)W) * 2.0f - 1.0f;
  float vy = (fy / (float)H) * 2.0f - 1.0f;
  V3 d = vec_times_mat3(v3(vx * cam.tan_fovy * cam.aspect, vy * cam.tan_fovy, -1.0f), cam.lookat);
  return normalize(d);
}

// IntersectBox + RayAABBIntersection (ray_bbox_intersection.comp:18-52); dir is normalised again there.
static inline bool ray_aabb(V3 eye, V3 dir_in, V3 boxmin, V3 boxmax, V3* dir_out, float* tnear, float* tfar) {
  V3 dir = normalize(dir_in);
  V3 inv = V3{1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z};
  V3 tb0 = inv * (boxmin - eye);
  V3 tb1 = inv * (boxmax - eye);
  V3 tmn = vmin(tb0, tb1), tmx = vmax(tb0, tb1);
  float tn = std::fmax(std::fmax(tmn.x, tmn.y), tmn.z);
  float tf = std::fmin(std::fmin(tmx.x, tmx.y), tmx.z);
  bool hit = tf > tn;
  tn = std::fmax(tn, 0.0f);
  *dir_out = dir; *tnear = tn; *tfar = tf;
  return hit;
}

// ---------------------------------------------------------------------------------------------
// Textures.  All reference textures are GL_LINEAR + GL_CLAMP_TO_EDGE (libs/volvis_utils/utils.cpp:8-9,
// transferfunction1d.cpp:66,97; texture3d.cpp:54-59).  Texel values are stored already rounded to the
// internal format (fp16 for R16F/RGBA16F/RG16F, fp32 for the R32F SAT).
// OpenGL 4.3 spec section 8.14: u = s*N - 0.5, i0 = floor(u), weights frac(u), indices clamped.
// ---------------------------------------------------------------------------------------------
struct Tex3D {
  int w = 0, h = 0, d = 0, c = 1;     // c channels interleaved
  const float* data = nullptr;
  inline float at(int x, int y, int z, int ch = 0) const {
    return data[((size_t)x + (size_t)w * ((size_t)y + (size_t)h * (size_t)z)) * c + ch];
  }
};

static inline void lin_coord(float s, int n, int* i0, int* i1, float* f) {
  float u = s * (float)n - 0.5f;
  float fl = std::floor(u);
  *f = u - fl;
  int a = (int)fl, b = a + 1;
  *i0 = std::min(std::max(a, 0), n - 1);
  *i1 = std::min(std::max(b, 0), n - 1);
}

// GL_LINEAR blend.  The GL spec leaves the arithmetic to the implementation; this restatement fixes it as
// a + t*(b-a) with a fused multiply-add (llvmpipe's lp_build_lerp form), and the CUDA kernels use the same
// expression so that the fp32 SAT look-ups of rc1pextbsd (differences of large prefix sums) agree bit for bit.
static inline float lerp(float a, float b, float t) { return std::fmaf(t, b - a, a); }

static inline float tex3d(const Tex3D& t, V3 p, int ch = 0) {
  int x0, x1, y0, y1, z0, z1; float fx, fy, fz;
  lin_coord(p.x, t.w, &x0, &x1, &fx);
  lin_coord(p.y, t.h, &y0, &y1, &fy);
  lin_coord(p.z, t.d, &z0, &z1, &fz);
  float c00 = lerp(t.at(x0, y0, z0, ch), t.at(x1, y0, z0, ch), fx);
  float c10 = lerp(t.at(x0, y1, z0, ch), t.at(x1, y1, z0, ch), fx);
  float c01 = lerp(t.at(x0, y0, z1, ch), t.at(x1, y0, z1, ch), fx);
  float c11 = lerp(t.at(x0, y1, z1, ch), t.at(x1, y1, z1, ch), fx);
  return lerp(lerp(c00, c10, fy), lerp(c01, c11, fy), fz);
}

// 1-D RGBA texture (transfer function / section tables), n texels of 4 floats.
struct Tex1D { int n = 0; const float* data = nullptr; };
static inline V4 tex1d(const Tex1D& t, float s) {
  int i0, i1; float f; lin_coord(s, t.n, &i0, &i1, &f);
  const float* a = t.data + 4 * (size_t)i0; const float* b = t.data + 4 * (size_t)i1;
  return V4{lerp(a[0], b[0], f), lerp(a[1], b[1], f), lerp(a[2], b[2], f), lerp(a[3], b[3], f)};
}
static inline V4 texel1d(const Tex1D& t, int i) {
  const float* a = t.data + 4 * (size_t)i; return V4{a[0], a[1], a[2], a[3]};
}

// Mip-mapped 3-D texture, GL_LINEAR_MIPMAP_LINEAR, textureLod: level sizes max(1,N>>l),
// lod clamped to [0,maxLevel], trilinear inside each level, linear between floor(lod) and floor(lod)+1.
struct Tex3DMip {
  std::vector<Tex3D> levels;
  inline float lod(V3 p, float l, int ch = 0) const {
    int maxl = (int)levels.size() - 1;
    if (!(l > 0.0f)) return tex3d(levels[0], p, ch);
    if (l >= (float)maxl) return tex3d(levels[maxl], p, ch);
    int l0 = (int)std::floor(l); float f = l - (float)l0;
    float a = tex3d(levels[l0], p, ch);
    if (f == 0.0f) return a;
    float b = tex3d(levels[l0 + 1], p, ch);
    return lerp(a, b, f);
  }
};

// Volume texel values as the reference uploads them: half(float(double(v)/255.0)) resp. /65535.0
// (structuredgridvolume.cpp:121-151, utils.cpp:20-56 with USE_16F_INTERNAL_FORMAT utils.h:15).
static inline void volume_to_r16f(const void* vox, size_t n, int bytes_per_voxel, float* out) {
  if (bytes_per_voxel == 1) {
    float lut[256];
    for (int v = 0; v < 256; ++v) lut[v] = round_f16((float)((double)v / (256.0 - 1.0)));
    const uint8_t* p = (const uint8_t*)vox;
    for (size_t i = 0; i < n; ++i) out[i] = lut[p[i]];
  } else {
    std::vector<float> lut(65536);
    for (int v = 0; v < 65536; ++v) lut[v] = round_f16((float)((double)v / (65536.0 - 1.0)));
    const uint16_t* p = (const uint16_t*)vox;
    for (size_t i = 0; i < n; ++i) out[i] = lut[p[i]];
  }
}

}  // namespace orc

// ---------------------------------------------------------------------------------------------
// Lighting uniforms shared by the lit renderers (same POD layout as vrb_lighting in include/vrb200.h).
// apply_phong = the ApplyPhongShading / ApplyGradientPhongShading uniform, i.e.
// (m_apply_gradient_shading && GetCurrentGradientTexture()) ? 1 : 0 (rc1prenderer.cpp:112, dosrcrenderer.cpp:221,
// ebsrenderer.cpp:221, vctrenderer.cpp:211, crtgtrenderer.cpp:217-218).
// ---------------------------------------------------------------------------------------------
struct Lighting {
  float ka, kd, ks, shininess;
  float ispecular[3], light_pos[3], light_forward[3], light_up[3], light_right[3];
  float spot_angle_deg;
  int apply_phong;
};

namespace orc {
// TexVolumeGradient: the RGB16F gradient texture bound by every renderer (oracle_gradient.cpp: orc_set_gradient).
const Tex3D* gradient_texture();

// The part of the Blinn-Phong branch every shader shares (ray_marching_1p.comp:50-70, rc1pdosct/ray_bbox_marching.comp:
// 629-643, ebs_ray_bbox_marching.comp:526-538, vct_ray_bbox_marching.comp:165-177, gt_ray_marching.comp:279-291):
// returns false when the sampled gradient is exactly zero (the shaders then leave the colour untouched).
static inline bool phong_terms(const Tex3D& grad, V3 Tpos, V3 G, V3 light_pos, V3 eye, float shininess, float* dot_diff, float* spec) {
  V3 s = Tpos / G;
  V3 n = v3(tex3d(grad, s, 0), tex3d(grad, s, 1), tex3d(grad, s, 2));
  if (n.x == 0.0f && n.y == 0.0f && n.z == 0.0f) return false;
  V3 Wpos = Tpos - (G * 0.5f);
  n = normalize(n);
  V3 light_direction = normalize(light_pos - Wpos);
  V3 eye_direction = normalize(eye - Wpos);
  V3 halfway_vector = normalize(eye_direction + light_direction);
  *dot_diff = std::fmax(0.0f, dot(n, light_direction));
  float dot_spec = std::fmax(0.0f, dot(halfway_vector, n));
  *spec = std::pow(dot_spec, shininess);
  return true;
}
}  // namespace orc
