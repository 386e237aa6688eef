// oracle/oracle_dos.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// CPU restatement of the directional-occlusion / cone-shadow renderer (Campagnolo & Celes 2019):
//   - ConeGaussianSampler (rc1pdosct/conegaussiansampler.cpp:212-498) -- pinned against the reference's own
//     conegaussiansampler.cpp compiled into oracle/_ref (tests/test_oracle_ref.py);
//   - ExtinctionCoefficientVolume (rc1pdosct/extcoefvolumegenerator.cpp:92-408) with glslextgen/gen_extcoefvol_*.comp,
//     gen_extcoefvol_*_mmlevel.comp and backtotau.comp;
//   - rc1pdosct/ray_bbox_marching.comp (whole file; CONSIDER_BORDERS defined, the other switches off), uniforms as
//     uploaded by dosrcrenderer.cpp:134-247,805-985.
// Pinned against the reference's own GLSL run on the CPU (pyramid shaders, marcher, light cache: tests/test_refglsl.py);
// see oracle_common.h for what stays unpinned.
#include "oracle_common.h"
#include <omp.h>
#include <cstdio>

using namespace orc;

namespace {
const double kPi = 3.14159265358979323846264338327950288;

// RodriguesRotation(glm::vec3, float, glm::vec3) (libs/math_utils/utils.cpp:149-156): the float overload is the one
// overload resolution picks for the sampler's (vec3, double, vec3) calls (glm's vec3->dvec3 constructor is explicit).
V3 rodrigues(V3 v, float teta, V3 k) {
  float c = std::cos(teta), s = std::sin(teta);
  V3 r = v * c + cross(k, v) * s + k * dot(k, v) * (1.0f - c);
  return normalize(r);
}
}  // namespace

extern "C" {

struct ConeSamplerParams {       // ConeGaussianSampler members (conegaussiansampler.cpp:30-47 + renderer set-up)
  float cone_half_angle;         // degrees
  float initial_step;            // 3.0
  int max_packing;               // 0: 1 ray, 1: 3 rays, 2: 7 rays
  float covered_distance;
  float d_sigma;                 // integration half step multiplier, 1.25
  float r_sigma;                 // sigma limit multiplier, 2.0
  float ui_weight;
};
struct ConeSamplerOut {
  int n_sections;
  int counts[3];                 // gaussian_samples_1 / _3 / _7
  float ray_axes[10][3];         // 3-ray axes then 7-ray axes, as uploaded to *ConeRayAxes[10]
  float ray3_adj_weight, ray7_adj_weight;
};

#define D_HEMISPHERE_CONE_DIV_3 (1.0 + (2.0 / std::sqrt(3.0)))
#define D_HEMISPHERE_CONE_DIV_7 3.010000

// sections_out: cap x 4 floats [interval distance, mip level, d_integral, amplitude] = the GL_FLOAT client array of
// GetConeSectionsInfoTex (:179-205), BEFORE the RGBA16F rounding.  Returns 0, or -1 if cap is too small, -2 on the
// invariant violations the reference exit()s on (:337-349).
int orc_cone_sampler_compute(const ConeSamplerParams* P, double min_sg_gaussian, float* sections_out, int cap, ConeSamplerOut* out) {
  struct Sec { int n; double pos, radius, sigma, d_integral, amplitude, mip; };
  struct Itv { double pos, dist; };
  std::vector<Sec> secs; std::vector<Itv> itvs;
  const double half_angle = (double)P->cone_half_angle;   // GetConeHalfAngle() returns float, promoted in the double expressions
  // 3 axis rays (:221-233)
  V3 ray3[3], ray7[7];
  {
    double adj_angle = half_angle / D_HEMISPHERE_CONE_DIV_3;
    double t1 = (half_angle - adj_angle) * kPi / 180.0;
    ray3[0] = rodrigues(v3(0, 0, 1), (float)t1, v3(0, 1, 0));
    ray3[0] = rodrigues(ray3[0], (float)(30.0 * kPi / 180.0), v3(0, 0, 1));
    double angle_t = 120.0 * kPi / 180.0;
    ray3[1] = rodrigues(ray3[0], (float)angle_t, v3(0, 0, 1));
    ray3[2] = rodrigues(ray3[1], (float)angle_t, v3(0, 0, 1));
  }
  {  // 7 axis rays (:237-252)
    ray7[0] = v3(0, 0, 1);
    double adj_angle = half_angle / D_HEMISPHERE_CONE_DIV_7;
    double t1 = (half_angle - adj_angle) * kPi / 180.0;
    ray7[1] = rodrigues(ray7[0], (float)t1, v3(0, 1, 0));
    double angle_t = 60.0 * kPi / 180.0;
    for (int i = 2; i < 7; ++i) ray7[i] = rodrigues(ray7[i - 1], (float)angle_t, v3(0, 0, 1));
  }
  // AddGaussianSampleStep{,With3,With7} (:418-498)
  int n_gaussians = 1;
  auto add_step = [&](double curr_pos, double sg) -> bool {
    static const double div[3] = {1.0, D_HEMISPHERE_CONE_DIV_3, D_HEMISPHERE_CONE_DIV_7};
    static const int cnt[3] = {1, 3, 7};
    int stage = n_gaussians > 3 ? 2 : (n_gaussians > 1 ? 1 : 0);
    for (;; ++stage) {
      n_gaussians = cnt[stage];
      double rad = (stage == 0 ? half_angle : (half_angle / div[stage])) * kPi / 180.0;
      double cone_radius = curr_pos * std::tan(rad);
      if (cone_radius > (double)P->r_sigma * sg) {
        if (stage < 2 && P->max_packing > stage) continue;   // try more gaussians
        return false;                                        // caller doubles sigma
      }
      secs.push_back(Sec{cnt[stage], curr_pos, cone_radius, sg, 0, 0, 0});
      return true;
    }
  };
  // ComputeConeIntegrationSteps (:254-285)
  double curr_pos = (double)P->initial_step;
  double sigma = min_sg_gaussian;
  while (!add_step(curr_pos, sigma)) sigma *= 2.0;
  const double dsg = (double)P->d_sigma;
  while (curr_pos < (double)P->covered_distance) {
    double si = dsg * sigma;
    while (!add_step(curr_pos + si + (dsg * sigma), sigma)) sigma *= 2.0;
    si += dsg * sigma;
    itvs.push_back(Itv{curr_pos, si});
    curr_pos += si;
  }
  // ComputeAdditionalInfo (:318-386)
  for (size_t i = 0; i + 1 < secs.size(); ++i)
    if (!(secs[i].n <= secs[i + 1].n)) return -2;
  if (!(secs.size() == itvs.size() + 1)) return -2;
  itvs.push_back(Itv{itvs.back().pos + itvs.back().dist, 0.0});
  int c1 = 0, c3 = 0, c7 = 0;
  auto gaussian_eval = [](double x, double sig) { return (1.0 / (std::sqrt(2.0 * kPi) * sig)) * std::exp(-(x * x) / (2.0 * sig * sig)); };
  auto integrate = [&](double sdev, double cone_radius) {   // IntegrateGaussian (:393-414)
    double t = (2.0 * cone_radius) / 0.05;
    int nt = (int)std::ceil(t);
    double segment = (2.0 * cone_radius) / double(nt);
    double s0 = -cone_radius + segment * 0.5;
    double S = 0.0;
    for (int i = 0; i < nt; i++) S += gaussian_eval(s0 + segment * double(i), sdev) * segment;
    return S;
  };
  for (size_t i = 0; i < secs.size(); ++i) {
    Sec& s = secs[i];
    if (s.n == 1) c1++; else if (s.n == 3) c3++; else if (s.n == 7) c7++;
    if (i == 0) s.d_integral = s.sigma * std::sqrt(2.0 * kPi) * 0.5;
    else s.d_integral = itvs[i - 1].dist * 0.5;
    double pr = integrate(s.sigma, s.radius);
    double Ac = kPi * s.radius * s.radius;
    double Ig = s.sigma * std::sqrt(2.0 * kPi);
    s.amplitude = ((pr * pr) * (Ig * Ig)) / Ac;
    s.mip = std::log2(s.sigma / min_sg_gaussian);
  }
  out->n_sections = (int)secs.size();
  out->counts[0] = c1; out->counts[1] = c3; out->counts[2] = c7;
  for (int i = 0; i < 3; ++i) { out->ray_axes[i][0] = ray3[i].x; out->ray_axes[i][1] = ray3[i].y; out->ray_axes[i][2] = ray3[i].z; }
  for (int i = 0; i < 7; ++i) { out->ray_axes[3 + i][0] = ray7[i].x; out->ray_axes[3 + i][1] = ray7[i].y; out->ray_axes[3 + i][2] = ray7[i].z; }
  out->ray3_adj_weight = (float)(double)dot(v3(0, 0, 1), ray3[0]);
  out->ray7_adj_weight = (float)(double)dot(v3(0, 0, 1), ray7[1]);
  if ((int)secs.size() > cap) return -1;
  for (size_t i = 0; i < secs.size(); ++i) {
    sections_out[4 * i + 0] = (float)itvs[i].dist;
    sections_out[4 * i + 1] = (float)secs[i].mip;
    sections_out[4 * i + 2] = (float)secs[i].d_integral;
    sections_out[4 * i + 3] = (float)secs[i].amplitude;
  }
  return 0;
}

// -------------------------------------------------------------------------------------------------------------------
// Extinction-coefficient pyramid.  Level sizes follow the GL mip chain (max(1, N>>l)); every level is stored R16F.
// levels_out: concatenated levels (level 0 first), each x fastest, already fp16-rounded, holding EXTINCTION (after
// backtotau).  level_dims: n_levels x 3.  Returns the number of levels written, or -1 when cap_floats is too small.
// -------------------------------------------------------------------------------------------------------------------
int orc_extcoef_levels(int rw, int rh, int rd) {
  int m = std::max(rw, std::max(rh, rd)), n = 1;
  while (m > 1) { m >>= 1; ++n; }
  return n;
}

// rw <= 0: the "same size" build (GenerateExtinctionCoefficientVolumeSameSize, extcoefvolumegenerator.cpp:92-228, taken when no
// custom resolution is set): base resolution = the volume's, and gen_extcoefvol_samesize.comp places its taps with the volume's
// own VoxelSize = grid_size / resolution as the HOST holds it (voxel_scale), not with a quotient recomputed from the grid size.
int orc_extcoef_build_ex(const float* vol_r16f, int vw, int vh, int vd, const float grid_size[3], const float* voxel_scale, const float* tf_rgba, int tf_n,
                         float S0, int rw, int rh, int rd, float* levels_out, size_t cap_floats, int* level_dims);
int orc_extcoef_build(const float* vol_r16f, int vw, int vh, int vd, const float grid_size[3], const float* tf_rgba, int tf_n,
                      float S0, int rw, int rh, int rd, float* levels_out, size_t cap_floats, int* level_dims) {
  return orc_extcoef_build_ex(vol_r16f, vw, vh, vd, grid_size, nullptr, tf_rgba, tf_n, S0, rw, rh, rd, levels_out, cap_floats, level_dims);
}
int orc_extcoef_build_ex(const float* vol_r16f, int vw, int vh, int vd, const float grid_size[3], const float* voxel_scale, const float* tf_rgba, int tf_n,
                         float S0, int rw, int rh, int rd, float* levels_out, size_t cap_floats, int* level_dims) {
  const bool same_size = rw <= 0 || rh <= 0 || rd <= 0;
  if (same_size) { rw = vw; rh = vh; rd = vd; if (!voxel_scale) return -2; }
  Tex3D vol; vol.w = vw; vol.h = vh; vol.d = vd; vol.c = 1; vol.data = vol_r16f;
  Tex1D tf; tf.n = tf_n; tf.data = tf_rgba;
  const V3 G = v3(grid_size[0], grid_size[1], grid_size[2]);
  const int nlev = orc_extcoef_levels(rw, rh, rd);
  size_t total = 0;
  std::vector<size_t> off(nlev);
  for (int l = 0; l < nlev; ++l) {
    int w = std::max(1, rw >> l), h = std::max(1, rh >> l), d = std::max(1, rd >> l);
    level_dims[3 * l] = w; level_dims[3 * l + 1] = h; level_dims[3 * l + 2] = d;
    off[l] = total; total += (size_t)w * h * d;
  }
  if (total > cap_floats) return -1;
  auto outside = [](V3 p) { return p.x < 0.0f || p.y < 0.0f || p.z < 0.0f || p.x > 1.0f || p.y > 1.0f || p.z > 1.0f; };
  // level 0 (gen_extcoefvol_anysize.comp / _samesize.comp): opacity of TF(volume) under a 7^3 Gaussian of sigma S0
  {
    const int w = level_dims[0], h = level_dims[1], d = level_dims[2];
    const V3 voxel = same_size ? v3(voxel_scale[0], voxel_scale[1], voxel_scale[2])      // VoxelSize (gen_extcoefvol_samesize.comp:45)
                               : G / v3((float)rw, (float)rh, (float)rd);               // base_level_voxel_sizes (extcoefvolumegenerator.cpp:233)
    float* out = levels_out + off[0];
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int z = 0; z < d; ++z)
      for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
          float sum_wkck = 0.0f, sum_wk = 0.0f;
          V3 grid_pos = (v3((float)x, (float)y, (float)z) + v3(0.5f, 0.5f, 0.5f)) * voxel;
          for (int ptx = -3; ptx < 4; ptx++)
            for (int pty = -3; pty < 4; pty++)
              for (int ptz = -3; ptz < 4; ptz++) {
                float fx = float(ptx) * S0, fy = float(pty) * S0, fz = float(ptz) * S0;
                float wk = (S0 * S0 * S0) * std::exp(-(fx * fx + fy * fy + fz * fz) / (2.0f * S0 * S0));
                V3 tp = (grid_pos + v3(fx, fy, fz)) / G;
                float ck = 0.0f;
                if (!outside(tp)) ck = tex1d(tf, tex3d(vol, tp)).w;
                sum_wkck += wk * ck;
                sum_wk += wk;
              }
          out[(size_t)x + (size_t)w * ((size_t)y + (size_t)h * z)] = round_f16(sum_wkck / sum_wk);
        }
  }
  // levels 1.. (gen_extcoefvol_*_mmlevel.comp): same filter with sigma Si = S0 * 2^i over level i-1 (still opacity)
  for (int l = 1; l < nlev; ++l) {
    const int w = level_dims[3 * l], h = level_dims[3 * l + 1], d = level_dims[3 * l + 2];
    Tex3D prev; prev.w = level_dims[3 * (l - 1)]; prev.h = level_dims[3 * (l - 1) + 1]; prev.d = level_dims[3 * (l - 1) + 2];
    prev.c = 1; prev.data = levels_out + off[l - 1];
    const float Si = S0 * std::pow(2.0f, (float)l);
    const V3 voxel = G / v3((float)w, (float)h, (float)d);
    float* out = levels_out + off[l];
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int z = 0; z < d; ++z)
      for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
          float sum_wkck = 0.0f, sum_wk = 0.0f;
          V3 grid_pos = (v3((float)x, (float)y, (float)z) + v3(0.5f, 0.5f, 0.5f)) * voxel;
          for (int ptx = -3; ptx < 4; ptx++)
            for (int pty = -3; pty < 4; pty++)
              for (int ptz = -3; ptz < 4; ptz++) {
                float fx = float(ptx) * Si, fy = float(pty) * Si, fz = float(ptz) * Si;
                float wk = (Si * Si * Si) * std::exp(-(fx * fx + fy * fy + fz * fz) / (2.0f * Si * Si));
                V3 tp = (grid_pos + v3(fx, fy, fz)) / G;
                float ck = 0.0f;
                if (!outside(tp)) ck = tex3d(prev, tp);
                sum_wkck += wk * ck;
                sum_wk += wk;
              }
          out[(size_t)x + (size_t)w * ((size_t)y + (size_t)h * z)] = round_f16(sum_wkck / sum_wk);
        }
  }
  // backtotau.comp on every level, after all levels exist (extcoefvolumegenerator.cpp:217-219,369-408)
  for (size_t i = 0; i < total; ++i) levels_out[i] = round_f16(-1.0f * std::log(1.0f - levels_out[i]));
  return nlev;
}

// -------------------------------------------------------------------------------------------------------------------
struct DosCone {                  // one sampler's uniforms
  const float* sections;          // n x 4, fp16-ROUNDED texel values (texelFetch of an RGBA16F texture)
  int n_sections;
  int counts[3];
  float initial_step, ray7_adj_weight, ui_weight;
  float axes[10][3];
};
struct DosParams {
  float step_size;
  int apply_occlusion, apply_shadow, type_of_shadow;
  float spot_cos;                 // SpotLightMaxAngle uniform = cos(pi * angle / 180) (dosrcrenderer.cpp:159)
  int count_samples;
};

namespace {
struct Dos {
  Tex3DMip pyr; V3 VSS;
  const DosCone* occ; const DosCone* sdw;
  DosParams P; Lighting L; V3 eye;

  float GetGaussianExtinction(V3 tp, float mip) const {   // ray_bbox_marching.comp:92-112
    float rg = pyr.lod(tp / VSS, mip);
    if (tp.x < 0.0f || tp.x > VSS.x || tp.y < 0.0f || tp.y > VSS.y || tp.z < 0.0f || tp.z > VSS.z) {
      float sg = std::pow(2.0f, mip);
      V3 c = vclamp(tp, v3(0, 0, 0), VSS) - tp;
      float dist = c.x * c.x + c.y * c.y + c.z * c.z;
      rg = rg * std::exp(-(dist) / (2.0f * sg * sg));
    }
    return rg;
  }
  // Cone{1,3,7}Ray{Occlusion,Shadow}: identical structure, parameterised by the sampler block.
  // (k,u,v) are the names at the CALL of Cone1Ray*; `swap_uv` reproduces Cone1RayShadow's (k, v, u) parameter list
  // (ray_bbox_marching.comp:481 vs :561), which hands u and v swapped to the 3- and 7-ray stages.
  float cone(const DosCone& C, V3 pos0, V3 k, V3 u, V3 v, bool swap_uv) const {
    if (swap_uv) std::swap(u, v);
    float rays[7], last[7];
    float track = C.initial_step;
    rays[0] = 0.0f; last[0] = 0.0f;
    auto sec = [&](int id) { const float* s = C.sections + 4 * (size_t)id; return V4{s[0], s[1], s[2], s[3]}; };
    for (int i = 0; i < C.counts[0]; ++i) {
      V4 si = sec(i);
      V3 pos = pos0 + k * track;
      float amptau = GetGaussianExtinction(pos, si.y) * si.w;
      rays[0] += (last[0] + amptau) * si.z * C.ui_weight;
      last[0] = amptau;
      track += si.x;
    }
    if (!(C.counts[1] + C.counts[2] > 0)) return std::exp(-rays[0]);
    // 1 -> 3
    rays[2] = rays[0]; rays[1] = rays[0];
    last[2] = last[0]; last[1] = last[0];
    V3 vk3[3];
    for (int i = 0; i < 3; ++i) vk3[i] = k * C.axes[i][2] + u * C.axes[i][1] + v * C.axes[i][0];
    for (int s = 0; s < C.counts[1]; ++s) {
      V4 si = sec(C.counts[0] + s);
      for (int i = 0; i < 3; ++i) {
        V3 pos = pos0 + vk3[i] * track;
        float amptau = GetGaussianExtinction(pos, si.y) * si.w;
        rays[i] += (last[i] + amptau) * si.z * C.ui_weight;
        last[i] = amptau;
      }
      track += si.x;
    }
    if (!(C.counts[2] > 0)) return (std::exp(-rays[0]) + std::exp(-rays[1]) + std::exp(-rays[2])) / 3.0f;
    // 3 -> 7
    rays[6] = rays[5] = rays[2];
    rays[4] = rays[3] = rays[1];
    float avg = (rays[2] + rays[1] + rays[0]) / 3.0f;
    rays[2] = rays[1] = rays[0];
    rays[0] = avg;
    last[6] = last[5] = last[2];
    last[4] = last[3] = last[1];
    float avgt = (last[2] + last[1] + last[0]) / 3.0f;
    last[2] = last[1] = last[0];
    last[0] = avgt;
    V3 vk7[7];
    for (int i = 0; i < 7; ++i) vk7[i] = k * C.axes[3 + i][2] + u * C.axes[3 + i][1] + v * C.axes[3 + i][0];
    for (int s = 0; s < C.counts[2]; ++s) {
      V4 si = sec(C.counts[0] + C.counts[1] + s);
      for (int i = 0; i < 7; ++i) {
        V3 pos = pos0 + vk7[i] * track;
        float amptau = si.w * GetGaussianExtinction(pos, si.y);
        rays[i] += (last[i] + amptau) * si.z * C.ui_weight;
        last[i] = amptau;
      }
      track += si.x;
    }
    return (std::exp(-rays[0]) + (std::exp(-rays[1]) + std::exp(-rays[2]) + std::exp(-rays[3]) + std::exp(-rays[4]) +
                                  std::exp(-rays[5]) + std::exp(-rays[6])) * C.ray7_adj_weight) / (1.0f + C.ray7_adj_weight * 6.0f);
  }
  float Occlusion(V3 pos0, V3 up, V3 right, V3 realpos) const {       // :324-333
    V3 k = normalize(eye - realpos);
    return cone(*occ, pos0, k, up, right, false);
  }
  float Shadow(V3 pos0) const {                                        // :533-562
    V3 lp = v3(L.light_pos[0], L.light_pos[1], L.light_pos[2]);
    V3 fwd = v3(L.light_forward[0], L.light_forward[1], L.light_forward[2]);
    V3 upv = v3(L.light_up[0], L.light_up[1], L.light_up[2]);
    V3 rgt = v3(L.light_right[0], L.light_right[1], L.light_right[2]);
    V3 k = v3(0, 0, 0), u = v3(0, 0, 0), v = v3(0, 0, 0);
    if (P.type_of_shadow == 0 || P.type_of_shadow == 1) {
      V3 cone_vec = normalize(lp - (pos0 - (VSS / 2.0f)));
      k = cone_vec;
      u = normalize(cross(k, rgt));
      v = normalize(cross(k, u));
      if (P.type_of_shadow == 1 && dot(cone_vec, fwd) < P.spot_cos) return 0.0f;
    } else if (P.type_of_shadow == 2) {
      k = fwd; v = upv; u = rgt;
    }
    return cone(*sdw, pos0, k, u, v, true);
  }
};
}  // namespace

// pyramid: concatenated fp16-rounded extinction levels from orc_extcoef_build.
int orc_dos_render(const float* vol_r16f, int vw, int vh, int vd, const float voxel_scale[3], const float* pyramid,
                   const int* level_dims, int n_levels, const float* tf_rgbt, int tf_n, const Camera* cam, const Lighting* light,
                   const DosCone* occ, const DosCone* sdw, const DosParams* prm, int W, int H, float* out_rgba, uint32_t* out_nsamples) {
  Dos Dd;
  Tex3D vol; vol.w = vw; vol.h = vh; vol.d = vd; vol.c = 1; vol.data = vol_r16f;
  Tex1D tf; tf.n = tf_n; tf.data = tf_rgbt;
  size_t off = 0;
  for (int l = 0; l < n_levels; ++l) {
    Tex3D t; t.w = level_dims[3 * l]; t.h = level_dims[3 * l + 1]; t.d = level_dims[3 * l + 2]; t.c = 1; t.data = pyramid + off;
    off += (size_t)t.w * t.h * t.d;
    Dd.pyr.levels.push_back(t);
  }
  Dd.VSS = v3((float)vw * voxel_scale[0], (float)vh * voxel_scale[1], (float)vd * voxel_scale[2]);
  Dd.occ = occ; Dd.sdw = sdw; Dd.P = *prm; Dd.L = *light;
  Dd.eye = v3(cam->eye[0], cam->eye[1], cam->eye[2]);
  const V3 G = Dd.VSS;
  const V3 InvG = v3(1.0f, 1.0f, 1.0f) / G;
  const Tex3D* grad = (light->apply_phong == 1) ? gradient_texture() : nullptr;
  if (light->apply_phong == 1 && !grad) return -2;
#pragma omp parallel for schedule(dynamic, 1)
  for (int py = 0; py < H; ++py) {
    for (int px = 0; px < W; ++px) {
      float* o = out_rgba + 4 * ((size_t)py * W + px);
      o[0] = o[1] = o[2] = o[3] = 0.0f;
      uint32_t ns = 0;
      V3 cdir = pixel_ray_dir(*cam, px, py, W, H);          // camera_dir (normalised once in main, :673-674)
      V3 dir; float tnear, tfar;
      bool inbox = ray_aabb(Dd.eye, cdir, -G * 0.5f, G * 0.5f, &dir, &tnear, &tfar);
      if (inbox) {
        V3 v_right = normalize(cross(cdir, v3(0, 1, 0)));   // :682-684 (uses camera_dir, not r.Dir)
        V3 v_up = normalize(cross(-cdir, v_right));
        float D = std::fabs(tfar - tnear);
        float dr = 0, dg = 0, db = 0, da = 0;
        V3 wd = Dd.eye + dir * tnear;
        wd = wd + (G * 0.5f);
        for (float s = 0.0f; s < D;) {
          float h = std::fmin(prm->step_size, D - s);
          V3 tx = wd + dir * (s + h * 0.5f);
          float density = tex3d(vol, tx * InvG);
          V4 src = tex1d(tf, density);
          ++ns;
          if (src.w > 0.0f) {
            // ShadeSample (:607-656)
            float ka = 0.0f, kd = 0.0f, ks = 0.0f, IOcc = 0.0f, ISdw = 0.0f;
            if (prm->apply_occlusion == 1) { ka = light->ka; IOcc = Dd.Occlusion(tx, v_up, v_right, tx - (G * 0.5f)); }
            if (prm->apply_shadow == 1) { kd = light->kd; ks = light->ks; ISdw = Dd.Shadow(tx); }
            float cr, cg, cb;
            if (grad) {                                      // ApplyPhongShading == 1 (:629-648)
              cr = src.x; cg = src.y; cb = src.z;             // a zero gradient leaves L.rgb = clr.rgb
              float dot_diff, spec;
              if (phong_terms(*grad, tx, G, v3(light->light_pos[0], light->light_pos[1], light->light_pos[2]), Dd.eye, light->shininess, &dot_diff, &spec)) {
                float f = ((1.0f / (ka + kd)) * (IOcc * ka + ISdw * kd * dot_diff));
                float sp = (ISdw * ks * spec);
                cr = src.x * f + light->ispecular[0] * sp;
                cg = src.y * f + light->ispecular[1] * sp;
                cb = src.z * f + light->ispecular[2] * sp;
              }
            } else {
              float kk = (1.0f / (ka + kd));
              cr = kk * (src.x * IOcc * ka + src.x * ISdw * kd);
              cg = kk * (src.y * IOcc * ka + src.y * ISdw * kd);
              cb = kk * (src.z * IOcc * ka + src.z * ISdw * kd);
            }
            float a = 1.0f - std::exp(-src.w * h);
            float om = 1.0f - da;
            dr = dr + om * (cr * a); dg = dg + om * (cg * a); db = db + om * (cb * a); da = da + om * a;
            if (da > 0.99f) break;
          }
          s = s + h;
        }
        o[0] = round_f16(dr); o[1] = round_f16(dg); o[2] = round_f16(db); o[3] = round_f16(da);
      }
      if (out_nsamples) out_nsamples[(size_t)py * W + px] = ns;
    }
  }
  return 0;
}

// K6: rc1pdosct/lightcachecomputation.comp main (:523-546), dispatched by PreComputeLightCache (dosrcrenderer.cpp:555-657).
// One (Iocc, Ishadow) pair per light-cache voxel, stored rg16f.  The cone functions are those of the marcher; only the
// occlusion frame differs (:280-293): v_right = normalize(cross(-cone_vec, EyeCamUp)), v_up = normalize(cross(cone_vec,
// v_right)).  out_rg: rw*rh*rd*2 floats, fp16-rounded (x fastest).
int orc_dos_light_cache(int vw, int vh, int vd, const float voxel_scale[3], const float* pyramid, const int* level_dims,
                        int n_levels, const float eye[3], const float eye_up[3], const Lighting* light, const DosCone* occ,
                        const DosCone* sdw, const DosParams* prm, int rw, int rh, int rd, float* out_rg) {
  Dos Dd;
  size_t off = 0;
  for (int l = 0; l < n_levels; ++l) {
    Tex3D t; t.w = level_dims[3 * l]; t.h = level_dims[3 * l + 1]; t.d = level_dims[3 * l + 2]; t.c = 1; t.data = pyramid + off;
    off += (size_t)t.w * t.h * t.d;
    Dd.pyr.levels.push_back(t);
  }
  Dd.VSS = v3((float)vw * voxel_scale[0], (float)vh * voxel_scale[1], (float)vd * voxel_scale[2]);
  Dd.occ = occ; Dd.sdw = sdw; Dd.P = *prm; Dd.L = *light;
  Dd.eye = v3(eye[0], eye[1], eye[2]);
  const V3 up = v3(eye_up[0], eye_up[1], eye_up[2]);
  // VolumeScales * (VolumeDimensions / LightCacheDimensions)
  const V3 cell = v3(voxel_scale[0] * ((float)vw / (float)rw), voxel_scale[1] * ((float)vh / (float)rh), voxel_scale[2] * ((float)vd / (float)rd));
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
  for (int z = 0; z < rd; ++z)
    for (int y = 0; y < rh; ++y)
      for (int x = 0; x < rw; ++x) {
        float Idao = 1.0f, Idcs = 1.0f;
        V3 tex_pos = v3(((float)x + 0.5f) * cell.x, ((float)y + 0.5f) * cell.y, ((float)z + 0.5f) * cell.z);
        V3 realpos = tex_pos - (Dd.VSS * 0.5f);
        if (prm->apply_occlusion == 1) {
          V3 cone_vec = normalize(Dd.eye - realpos);
          V3 v_right = normalize(cross(-cone_vec, up));
          V3 v_up = normalize(cross(cone_vec, v_right));
          Idao = Dd.cone(*occ, tex_pos, cone_vec, v_up, v_right, false);
        }
        if (prm->apply_shadow == 1) Idcs = Dd.Shadow(tex_pos);
        float* o = out_rg + 2 * ((size_t)x + (size_t)rw * ((size_t)y + (size_t)rh * (size_t)z));
        o[0] = round_f16(Idao); o[1] = round_f16(Idcs);
      }
  return 0;
}

// K7: _common_shaders/obj_ray_marching.comp, active #else branch (:210-333): the primary march with the shading read from
// the light cache (rg16f, GL_LINEAR, clamp-to-edge), with the gradient Blinn-Phong branch (:236-257) when
// light->apply_phong == 1.  Samples are composited only when src.a > 0 AND (ApplyOcclusion || ApplyShadow) (:312).
int orc_obj_march_lit(const float* vol_r16f, int vw, int vh, int vd, const float voxel_scale[3], const float* tf_rgbt, int tf_n,
                      const Camera* cam, const Lighting* light, int apply_occlusion, int apply_shadow, float step_size,
                      const float* cache_rg, int rw, int rh, int rd, int W, int H, float* out_rgba, uint32_t* out_nsamples) {
  const float Kambient = light->ka, Kdiffuse = light->kd;
  const Tex3D* grad = light->apply_phong == 1 ? gradient_texture() : nullptr;      // ApplyPhongShading (:236-257)
  if (light->apply_phong == 1 && !grad) return -2;
  Tex3D vol; vol.w = vw; vol.h = vh; vol.d = vd; vol.c = 1; vol.data = vol_r16f;
  Tex3D lc; lc.w = rw; lc.h = rh; lc.d = rd; lc.c = 2; lc.data = cache_rg;
  Tex1D tf; tf.n = tf_n; tf.data = tf_rgbt;
  const V3 G = v3((float)vw * voxel_scale[0], (float)vh * voxel_scale[1], (float)vd * voxel_scale[2]);
  const V3 InvG = v3(1.0f, 1.0f, 1.0f) / G;
  const V3 eye = v3(cam->eye[0], cam->eye[1], cam->eye[2]);
  const bool Shade = apply_occlusion == 1 || apply_shadow == 1;
#pragma omp parallel for schedule(dynamic, 1)
  for (int py = 0; py < H; ++py) {
    for (int px = 0; px < W; ++px) {
      float* o = out_rgba + 4 * ((size_t)py * W + px);
      o[0] = o[1] = o[2] = o[3] = 0.0f;
      uint32_t ns = 0;
      V3 cdir = pixel_ray_dir(*cam, px, py, W, H);
      V3 dir; float tnear, tfar;
      bool inbox = ray_aabb(eye, cdir, -G * 0.5f, G * 0.5f, &dir, &tnear, &tfar);
      if (inbox) {
        float D = std::fabs(tfar - tnear);
        float cr = 0, cg = 0, cb = 0, ca = 0;
        V3 wd = eye + dir * tnear;
        wd = wd + (G * 0.5f);
        for (float s = 0.0f; s < D;) {
          float h = std::fmin(step_size, D - s);
          V3 tx = wd + dir * (s + h * 0.5f);
          float density = tex3d(vol, tx * InvG);
          V4 src = tex1d(tf, density);
          ++ns;
          if (src.w > 0.0f && Shade) {
            V3 lp = tx / G;
            float Ia = tex3d(lc, lp, 0), Is = tex3d(lc, lp, 1);
            float ka = 0.0f, kd = 0.0f, ks = 0.0f, IOcc = 0.0f, ISdw = 0.0f;
            if (apply_occlusion == 1) { ka = Kambient; IOcc = Ia; }
            if (apply_shadow == 1) { kd = Kdiffuse; ks = light->ks; ISdw = Is; }
            float r, g, b;
            if (grad) {
              r = src.x; g = src.y; b = src.z;                   // a zero gradient leaves L = clr (:241)
              float dot_diff, spec;
              if (phong_terms(*grad, tx, G, v3(light->light_pos[0], light->light_pos[1], light->light_pos[2]), eye, light->shininess, &dot_diff, &spec)) {
                float kk = (1.0f / (ka + kd));
                r = kk * (src.x * IOcc * ka + ISdw * (src.x * kd * dot_diff)) + ISdw * (ks * light->ispecular[0] * spec);
                g = kk * (src.y * IOcc * ka + ISdw * (src.y * kd * dot_diff)) + ISdw * (ks * light->ispecular[1] * spec);
                b = kk * (src.z * IOcc * ka + ISdw * (src.z * kd * dot_diff)) + ISdw * (ks * light->ispecular[2] * spec);
              }
            } else {
              float kk = (1.0f / (ka + kd));
              r = kk * (src.x * IOcc * ka + src.x * ISdw * kd);
              g = kk * (src.y * IOcc * ka + src.y * ISdw * kd);
              b = kk * (src.z * IOcc * ka + src.z * ISdw * kd);
            }
            float a = 1.0f - std::exp(-src.w * h);
            float om = 1.0f - ca;
            cr = cr + om * (r * a); cg = cg + om * (g * a); cb = cb + om * (b * a); ca = ca + om * a;
            if (ca > 0.99f) break;
          }
          s = s + h;
        }
        o[0] = round_f16(cr); o[1] = round_f16(cg); o[2] = round_f16(cb); o[3] = round_f16(ca);
      }
      if (out_nsamples) out_nsamples[(size_t)py * W + px] = ns;
    }
  }
  return 0;
}

int orc_obj_march(const float* vol_r16f, int vw, int vh, int vd, const float voxel_scale[3], const float* tf_rgbt, int tf_n,
                  const Camera* cam, float Kambient, float Kdiffuse, int apply_occlusion, int apply_shadow, float step_size,
                  const float* cache_rg, int rw, int rh, int rd, int W, int H, float* out_rgba, uint32_t* out_nsamples) {
  Lighting l;
  std::memset(&l, 0, sizeof(l));
  l.ka = Kambient; l.kd = Kdiffuse;
  return orc_obj_march_lit(vol_r16f, vw, vh, vd, voxel_scale, tf_rgbt, tf_n, cam, &l, apply_occlusion, apply_shadow, step_size, cache_rg, rw, rh, rd,
                           W, H, out_rgba, out_nsamples);
}

}  // extern "C"
