// oracle/oracle_ebs.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// CPU restatement of the extinction-based shading renderer (Schlegel et al. 2011):
//   - RC1PExtinctionBasedShading::GenerateExtinctionSAT3DTex (rc1pextbsd/ebsrenderer.cpp:624-723) with
//     vis::SummedAreaTable3D<double>::BuildSAT (libs/vis_utils/summedareatable.h:218-278), same recurrence order
//     (pinned bit-for-bit against the reference header compiled into oracle/_ref, tests/test_oracle_ref.py);
//   - rc1pextbsd/ebs_ray_bbox_marching.comp (whole file), uniforms as uploaded by ebsrenderer.cpp:125-247,556-583.
// Pinned against the reference's own GLSL run on the CPU (marcher and light cache: tests/test_refglsl.py); see oracle_common.h.
#include "oracle_common.h"
#include <omp.h>

using namespace orc;

extern "C" {

// SAT over the zero-bordered (W+2)(H+2)(D+2) grid; ext_lut[v] = tf->GetExtN(v/max) per voxel value.
// out_f32 = float(double SAT).  Layout x + w*y + w*h*z.
void orc_sat_build(const void* vox, int vw, int vh, int vd, int bpv, const float* ext_lut, float* out_f32, double* out_f64) {
  const int w = vw + 2, h = vh + 2, d = vd + 2;
  const size_t n = (size_t)w * h * d;
  std::vector<double> S(n, 0.0);
  auto at = [&](int x, int y, int z) -> double& { return S[(size_t)x + (size_t)w * y + (size_t)w * h * z]; };
  auto get = [&](int x, int y, int z) -> double {   // SummedAreaTable3D::GetValue (summedareatable.h:204-213)
    if (x < 0 || y < 0 || z < 0) return 0.0;
    if (x >= w) x = w - 1;
    if (y >= h) y = h - 1;
    if (z >= d) z = d - 1;
    return S[(size_t)x + (size_t)w * y + (size_t)w * h * z];
  };
  for (int z = 1; z < d - 1; ++z)
    for (int y = 1; y < h - 1; ++y)
      for (int x = 1; x < w - 1; ++x) {
        size_t id = (size_t)(x - 1) + (size_t)vw * (y - 1) + (size_t)vw * vh * (z - 1);
        at(x, y, z) = bpv == 1 ? (double)ext_lut[((const uint8_t*)vox)[id]] : (double)ext_lut[((const uint16_t*)vox)[id]];
      }
  // BuildSAT, steps 1-4 in the reference's order
  for (int x = 1; x < w; x++) at(x, 0, 0) = get(x - 1, 0, 0) + get(x, 0, 0);
  for (int y = 1; y < h; y++) at(0, y, 0) = get(0, y - 1, 0) + get(0, y, 0);
  for (int z = 1; z < d; z++) at(0, 0, z) = get(0, 0, z - 1) + get(0, 0, z);
  for (int x = 1; x < w; x++)
    for (int z = 1; z < d; z++) at(x, 0, z) = get(x - 1, 0, z) + get(x, 0, z - 1) - get(x - 1, 0, z - 1) + get(x, 0, z);
  for (int x = 1; x < w; x++)
    for (int y = 1; y < h; y++) at(x, y, 0) = get(x - 1, y, 0) + get(x, y - 1, 0) - get(x - 1, y - 1, 0) + get(x, y, 0);
  for (int y = 1; y < h; y++)
    for (int z = 1; z < d; z++) at(0, y, z) = get(0, y - 1, z) + get(0, y, z - 1) - get(0, y - 1, z - 1) + get(0, y, z);
  // step 4: x outer, z inner in the reference; the value at (x,y,z) depends only on smaller indices, so any order
  // that respects the dependences gives identical doubles.  z outer here for cache friendliness.
  for (int z = 1; z < d; z++)
    for (int y = 1; y < h; y++)
      for (int x = 1; x < w; x++) {
        double val = get(x, y, z) + get(x - 1, y - 1, z - 1) + get(x, y, z - 1) + get(x, y - 1, z) + get(x - 1, y, z) -
                     get(x - 1, y - 1, z) - get(x, y - 1, z - 1) - get(x - 1, y, z - 1);
        at(x, y, z) = val;
      }
  if (out_f32) for (size_t i = 0; i < n; ++i) out_f32[i] = (float)S[i];
  if (out_f64) std::memcpy(out_f64, S.data(), n * sizeof(double));
}

// Integer SAT (bit-exact mode): inclusive 3-D prefix sum of lut[v] over the UNbordered grid, u64.
void orc_sat_build_u64(const void* vox, int w, int h, int d, int bpv, const uint32_t* lut, uint64_t* out) {
  auto idx = [&](int x, int y, int z) { return (size_t)x + (size_t)w * y + (size_t)w * h * z; };
  for (int z = 0; z < d; ++z)
    for (int y = 0; y < h; ++y) {
      uint64_t run = 0;
      for (int x = 0; x < w; ++x) {
        size_t i = idx(x, y, z);
        run += bpv == 1 ? lut[((const uint8_t*)vox)[i]] : lut[((const uint16_t*)vox)[i]];
        out[i] = run;
      }
    }
  for (int z = 0; z < d; ++z)
    for (int y = 1; y < h; ++y)
      for (int x = 0; x < w; ++x) out[idx(x, y, z)] += out[idx(x, y - 1, z)];
  for (int z = 1; z < d; ++z)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) out[idx(x, y, z)] += out[idx(x, y, z - 1)];
}

struct EbsParams {
  float step_size;
  int apply_occlusion, apply_shadow;
  int amb_occ_shells; float amb_occ_radius;
  float sdw_cone_angle_rad, sdw_sample_interval, sdw_initial_step, sdw_ui_weight, sdw_cone_max_distance;
  int type_of_shadow;
  int count_samples;
};

namespace {
struct Ebs {
  Tex3D vol, sat; Tex1D tf;
  V3 VS, VSS;                 // VolumeScales, VolumeScaledSizes
  V3 MinSAT, MaxSAT, MinVol, MaxVol, inv_vol_scaled;
  EbsParams P; Lighting L;
  V3 eye;
  const Tex3D* grad = nullptr;   // TexVolumeGradient when ApplyPhongShading == 1

  float GetSummed3Density(float x, float y, float z) const { return tex3d(sat, v3(x, y, z) * inv_vol_scaled); }
  float EvaluateSAT3D(V3 p1, V3 p2) const {
    float V1 = GetSummed3Density(p2.x, p2.y, p2.z);
    float V2 = GetSummed3Density(p1.x, p2.y, p2.z);
    float V3_ = GetSummed3Density(p2.x, p2.y, p1.z);
    float V4_ = GetSummed3Density(p1.x, p2.y, p1.z);
    float V5 = GetSummed3Density(p2.x, p1.y, p2.z);
    float V6 = GetSummed3Density(p1.x, p1.y, p2.z);
    float V7 = GetSummed3Density(p2.x, p1.y, p1.z);
    float V8 = GetSummed3Density(p1.x, p1.y, p1.z);
    return (V1 - V2 - V3_ + V4_ - V5 + V6 + V7 - V8);
  }
  float EvaluateAmbientOcclusionSAT3D(V3 p1, V3 p2) const {
    p1 = vclamp(p1 + VS, MinSAT, MaxSAT);
    p2 = vclamp(p2 + VS, MinSAT, MaxSAT);
    return EvaluateSAT3D(p1, p2);
  }
  float ExtinctionAmbientOcclusion(V3 tx) const {   // ebs_ray_bbox_marching.comp:112-146
    float SAT_Sh0 = EvaluateAmbientOcclusionSAT3D(tx - P.amb_occ_radius * VS, tx + P.amb_occ_radius * VS);
    float rsh0 = P.amb_occ_radius;
    float tSh0 = SAT_Sh0 * (1.0f / (rsh0 * rsh0));
    float SAT_Shi = SAT_Sh0, tshi = tSh0;
    int ith = 1;
    while (ith < P.amb_occ_shells) {
      float r1 = P.amb_occ_radius * (float)(ith + 1);
      float S1 = EvaluateAmbientOcclusionSAT3D(tx - r1 * VS, tx + r1 * VS);
      float t1 = tshi + (S1 - SAT_Shi) * (1.0f / (r1 * r1));
      SAT_Shi = S1; tshi = t1; ith = ith + 1;
    }
    float rshi = P.amb_occ_radius * (float)P.amb_occ_shells;
    float W_A = 1.0f / (rshi * rshi);
    float Stau = W_A * tshi;
    return std::exp(-(Stau));
  }
  float EvaluateShadowSAT3D(V3 p1, V3 p2) const {   // :148-188, non-texelFetch branch
    float volquery = ((std::fabs(p1.x - p2.x) / VS.x)) * ((std::fabs(p1.y - p2.y) / VS.y)) * ((std::fabs(p1.z - p2.z) / VS.z));
    p1 = vclamp(p1 + VS, MinSAT, MaxSAT);
    p2 = vclamp(p2 + VS, MinSAT, MaxSAT);
    return ((EvaluateSAT3D(p1, p2) / volquery)) * P.sdw_ui_weight;
  }
  // One implementation for the three dominant axes: a = dominant axis index, (b, c) = the two lateral axes in the
  // order the shader treats them.  The rotation formulas differ per axis in the shader (:190-430) and are kept verbatim
  // below through the `rot` lambdas.
  static float comp(V3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
  static V3 set3(int ia, float a, int ib, float b, int ic, float c) {
    float r[3]; r[ia] = a; r[ib] = b; r[ic] = c; return v3(r[0], r[1], r[2]);
  }
  float ConeZAxis(V3 pos, V3 cv) const {
    float Stau = 0.0f;
    float signal = 1.0f; if (cv.z < 0) signal = -1.0f;
    V3 proj_y = normalize(v3(0.0f, cv.y, cv.z));
    V3 proj_x = normalize(v3(cv.x, 0.0f, cv.z));
    float ra = P.sdw_cone_angle_rad;
    float p_cs = std::cos(ra), p_sn = std::sin(ra), n_cs = std::cos(-ra), n_sn = std::sin(-ra);
    V3 pj_x1 = normalize(v3(proj_x.x * n_cs - proj_x.z * n_sn, 0.0f, proj_x.x * n_sn + proj_x.z * n_cs));
    V3 pj_x2 = normalize(v3(proj_x.x * p_cs - proj_x.z * p_sn, 0.0f, proj_x.x * p_sn + proj_x.z * p_cs));
    V3 pj_y1 = normalize(v3(0.0f, proj_y.y * n_cs - proj_y.z * n_sn, proj_y.y * n_sn + proj_y.z * n_cs));
    V3 pj_y2 = normalize(v3(0.0f, proj_y.y * p_cs - proj_y.z * p_sn, proj_y.y * p_sn + proj_y.z * p_cs));
    float si = P.sdw_sample_interval * signal * VS.z;
    float z_pos = P.sdw_initial_step * signal * VS.z;
    while ((z_pos / cv.z) < P.sdw_cone_max_distance &&
           (pos.z + (z_pos + si) > MinVol.z && pos.z + (z_pos + si) < MaxVol.z)) {
      float z_mean = std::fabs(z_pos + si * 0.5f);
      float p_x1 = pj_x1.x * (z_mean / std::fabs(pj_x1.z));
      float p_x2 = pj_x2.x * (z_mean / std::fabs(pj_x2.z));
      float p_y1 = pj_y1.y * (z_mean / std::fabs(pj_y1.z));
      float p_y2 = pj_y2.y * (z_mean / std::fabs(pj_y2.z));
      float x1 = std::fmin(p_x1, p_x2), x2 = std::fmax(p_x1, p_x2);
      float y1 = std::fmin(p_y1, p_y2), y2 = std::fmax(p_y1, p_y2);
      float xdiff = std::fabs(x2 - x1), ydiff = std::fabs(y2 - y1);
      float xs = (std::ceil(xdiff / VS.x) - (xdiff / VS.x)) * 0.5f;
      float ys = (std::ceil(ydiff / VS.y) - (ydiff / VS.y)) * 0.5f;
      x1 = x1 - xs * VS.x; x2 = x2 + xs * VS.x;
      y1 = y1 - ys * VS.y; y2 = y2 + ys * VS.y;
      float z1 = std::fmin(z_pos, z_pos + si), z2 = std::fmax(z_pos, z_pos + si);
      Stau += EvaluateShadowSAT3D(pos + v3(x1, y1, z1), pos + v3(x2, y2, z2));
      z_pos = z_pos + si;
    }
    return Stau;
  }
  float ConeYAxis(V3 pos, V3 cv) const {
    float Stau = 0.0f;
    float signal = 1.0f; if (cv.y < 0) signal = -1.0f;
    V3 proj_x = normalize(v3(cv.x, cv.y, 0.0f));
    V3 proj_z = normalize(v3(0.0f, cv.y, cv.z));
    float ra = P.sdw_cone_angle_rad;
    float p_cs = std::cos(ra), p_sn = std::sin(ra), n_cs = std::cos(-ra), n_sn = std::sin(-ra);
    V3 pj_x1 = normalize(v3(proj_x.x * n_cs - proj_x.y * n_sn, proj_x.x * n_sn + proj_x.y * n_cs, 0.0f));
    V3 pj_x2 = normalize(v3(proj_x.x * p_cs - proj_x.y * p_sn, proj_x.x * p_sn + proj_x.y * p_cs, 0.0f));
    V3 pj_z1 = normalize(v3(0.0f, proj_z.z * n_sn + proj_z.y * n_cs, proj_z.z * n_cs - proj_z.y * n_sn));
    V3 pj_z2 = normalize(v3(0.0f, proj_z.z * p_sn + proj_z.y * p_cs, proj_z.z * p_cs - proj_z.y * p_sn));
    float si = P.sdw_sample_interval * signal * VS.y;
    float y_pos = P.sdw_initial_step * signal * VS.y;
    while ((y_pos / cv.y) < P.sdw_cone_max_distance &&
           (pos.y + (y_pos + si) > MinVol.y && pos.y + (y_pos + si) < MaxVol.y)) {
      float y_mean = std::fabs(y_pos + si * 0.5f);
      float p_x1 = pj_x1.x * (y_mean / std::fabs(pj_x1.y));
      float p_x2 = pj_x2.x * (y_mean / std::fabs(pj_x2.y));
      float p_z1 = pj_z1.z * (y_mean / std::fabs(pj_z1.y));
      float p_z2 = pj_z2.z * (y_mean / std::fabs(pj_z2.y));
      float x1 = std::fmin(p_x1, p_x2), x2 = std::fmax(p_x1, p_x2);
      float z1 = std::fmin(p_z1, p_z2), z2 = std::fmax(p_z1, p_z2);
      float xdiff = std::fabs(x2 - x1), zdiff = std::fabs(z2 - z1);
      float xs = (std::ceil(xdiff / VS.x) - (xdiff / VS.x)) * 0.5f;
      float zs = (std::ceil(zdiff / VS.z) - (zdiff / VS.z)) * 0.5f;
      x1 = x1 - xs * VS.x; x2 = x2 + xs * VS.x;
      z1 = z1 - zs * VS.z; z2 = z2 + zs * VS.z;
      float y1 = std::fmin(y_pos, y_pos + si), y2 = std::fmax(y_pos, y_pos + si);
      Stau += EvaluateShadowSAT3D(pos + v3(x1, y1, z1), pos + v3(x2, y2, z2));
      y_pos = y_pos + si;
    }
    return Stau;
  }
  float ConeXAxis(V3 pos, V3 cv) const {
    float Stau = 0.0f;
    float signal = 1.0f; if (cv.x < 0) signal = -1.0f;
    V3 proj_y = normalize(v3(cv.x, cv.y, 0.0f));
    V3 proj_z = normalize(v3(cv.x, 0.0f, cv.z));
    float ra = P.sdw_cone_angle_rad;
    float p_cs = std::cos(ra), p_sn = std::sin(ra), n_cs = std::cos(-ra), n_sn = std::sin(-ra);
    V3 pj_y1 = normalize(v3(proj_y.y * n_sn + proj_y.x * n_cs, proj_y.y * n_cs - proj_y.x * n_sn, 0.0f));
    V3 pj_y2 = normalize(v3(proj_y.y * p_sn + proj_y.x * p_cs, proj_y.y * p_cs - proj_y.x * p_sn, 0.0f));
    V3 pj_z1 = normalize(v3(proj_z.z * n_sn + proj_z.x * n_cs, 0.0f, proj_z.z * n_cs - proj_z.x * n_sn));
    V3 pj_z2 = normalize(v3(proj_z.z * p_sn + proj_z.x * p_cs, 0.0f, proj_z.z * p_cs - proj_z.x * p_sn));
    float si = P.sdw_sample_interval * signal * VS.x;
    float x_pos = P.sdw_initial_step * signal * VS.x;
    while ((x_pos / cv.x) < P.sdw_cone_max_distance &&
           (pos.x + (x_pos + si) > MinVol.x && pos.x + (x_pos + si) < MaxVol.x)) {
      float x_mean = std::fabs(x_pos + si * 0.5f);
      float p_y1 = pj_y1.y * (x_mean / std::fabs(pj_y1.x));
      float p_y2 = pj_y2.y * (x_mean / std::fabs(pj_y2.x));
      float p_z1 = pj_z1.z * (x_mean / std::fabs(pj_z1.x));
      float p_z2 = pj_z2.z * (x_mean / std::fabs(pj_z2.x));
      float y1 = std::fmin(p_y1, p_y2), y2 = std::fmax(p_y1, p_y2);
      float z1 = std::fmin(p_z1, p_z2), z2 = std::fmax(p_z1, p_z2);
      float ydiff = std::fabs(y2 - y1), zdiff = std::fabs(z2 - z1);
      float ys = (std::ceil(ydiff / VS.y) - (ydiff / VS.y)) * 0.5f;
      float zs = (std::ceil(zdiff / VS.z) - (zdiff / VS.z)) * 0.5f;
      y1 = y1 - ys * VS.y; y2 = y2 + ys * VS.y;
      z1 = z1 - zs * VS.z; z2 = z2 + zs * VS.z;
      float x1 = std::fmin(x_pos, x_pos + si), x2 = std::fmax(x_pos, x_pos + si);
      Stau += EvaluateShadowSAT3D(pos + v3(x1, y1, z1), pos + v3(x2, y2, z2));
      x_pos = x_pos + si;
    }
    return Stau;
  }
  float ExtinctionDirectionalShadows(V3 tx) const {   // :432-456
    V3 realpos = tx - (VSS * 0.5f);
    V3 cone_vec = v3(0, 0, 0);
    if (P.type_of_shadow == 0) cone_vec = normalize(v3(L.light_pos[0], L.light_pos[1], L.light_pos[2]) - realpos);
    else if (P.type_of_shadow == 1) cone_vec = normalize(v3(L.light_forward[0], L.light_forward[1], L.light_forward[2]));
    V3 a = vabs(cone_vec);
    float Stau = 0.0f;
    if (a.z > a.x && a.z > a.y) Stau = ConeZAxis(tx, cone_vec);
    else if (a.y > a.x) Stau = ConeYAxis(tx, cone_vec);
    else Stau = ConeXAxis(tx, cone_vec);
    return std::exp(-Stau);
  }
  V4 ShadeSample(V4 clr, V3 tx) const {   // :500-551
    float ka = 0.0f, kd = 0.0f, ks = 0.0f;
    float IOcc = 0.0f;
    if (P.apply_occlusion == 1) { ka = L.ka; IOcc = ExtinctionAmbientOcclusion(tx); }
    float ISdw = 0.0f;
    if (P.apply_shadow == 1) { kd = L.kd; ks = L.ks; ISdw = ExtinctionDirectionalShadows(tx); }
    V4 o = clr;
    if (grad) {                            // ApplyPhongShading == 1 (:524-544); a zero gradient leaves L = clr
      float dot_diff, spec;
      if (phong_terms(*grad, tx, VSS, v3(L.light_pos[0], L.light_pos[1], L.light_pos[2]), eye, L.shininess, &dot_diff, &spec)) {
        float k = (1.0f / (ka + kd));
        o.x = k * (clr.x * IOcc * ka + ISdw * (clr.x * kd * dot_diff)) + ISdw * (ks * L.ispecular[0] * spec);
        o.y = k * (clr.y * IOcc * ka + ISdw * (clr.y * kd * dot_diff)) + ISdw * (ks * L.ispecular[1] * spec);
        o.z = k * (clr.z * IOcc * ka + ISdw * (clr.z * kd * dot_diff)) + ISdw * (ks * L.ispecular[2] * spec);
      }
      return o;
    }
    float k = (1.0f / (ka + kd));
    o.x = k * (clr.x * IOcc * ka + clr.x * ISdw * kd);
    o.y = k * (clr.y * IOcc * ka + clr.y * ISdw * kd);
    o.z = k * (clr.z * IOcc * ka + clr.z * ISdw * kd);
    return o;
  }
};
}  // namespace

// sat_f32: (vw+2)(vh+2)(vd+2) floats.  Other arguments as orc_rc1pass_render.  voxel_scale = VolumeScales.
int orc_ebs_render(const float* vol_r16f, int vw, int vh, int vd, const float voxel_scale[3], const float* sat_f32,
                   const float* tf_rgbt, int tf_n, const Camera* cam, const Lighting* light, const EbsParams* prm,
                   int W, int H, float* out_rgba, uint32_t* out_nsamples) {
  Ebs E;
  E.vol.w = vw; E.vol.h = vh; E.vol.d = vd; E.vol.c = 1; E.vol.data = vol_r16f;
  E.sat.w = vw + 2; E.sat.h = vh + 2; E.sat.d = vd + 2; E.sat.c = 1; E.sat.data = sat_f32;
  E.tf.n = tf_n; E.tf.data = tf_rgbt;
  E.VS = v3(voxel_scale[0], voxel_scale[1], voxel_scale[2]);
  E.VSS = v3((float)vw, (float)vh, (float)vd) * E.VS;      // vol_resolution * vol_voxelsize (ebsrenderer.cpp:558-566)
  E.MinSAT = E.VS * 0.5f; E.MaxSAT = E.VSS + E.VS * 1.5f;
  E.MinVol = E.VS * 0.5f; E.MaxVol = E.VSS - E.VS * 0.5f;
  E.inv_vol_scaled = v3(1.0f, 1.0f, 1.0f) / (E.VSS + E.VS * 2.0f);
  E.P = *prm; E.L = *light;
  E.grad = (light->apply_phong == 1) ? gradient_texture() : nullptr;
  if (light->apply_phong == 1 && !E.grad) return -2;
  E.eye = v3(cam->eye[0], cam->eye[1], cam->eye[2]);
  const V3 G = E.VSS;
  const V3 InvG = v3(1.0f, 1.0f, 1.0f) / G;
#pragma omp parallel for schedule(dynamic, 1)
  for (int py = 0; py < H; ++py) {
    for (int px = 0; px < W; ++px) {
      float* o = out_rgba + 4 * ((size_t)py * W + px);
      o[0] = o[1] = o[2] = o[3] = 0.0f;
      uint32_t ns = 0;
      // camera_dir = normalize(vec3(...) * mat3(ViewMatrix)) then normalised again in RayAABBIntersection (:553-571)
      V3 cdir = pixel_ray_dir(*cam, px, py, W, H);
      V3 dir; float tnear, tfar;
      bool inbox = ray_aabb(E.eye, cdir, -G * 0.5f, G * 0.5f, &dir, &tnear, &tfar);
      if (inbox) {
        float D = std::fabs(tfar - tnear);
        float dr = 0, dg = 0, db = 0, da = 0;
        V3 wd = E.eye + dir * tnear;
        wd = wd + (G * 0.5f);
        for (float s = 0.0f; s < D;) {
          float h = std::fmin(prm->step_size, D - s);
          V3 tx = wd + dir * (s + h * 0.5f);
          float density = tex3d(E.vol, tx * InvG);
          V4 src = tex1d(E.tf, density);
          ++ns;
          if (src.w > 0.0f) {
            src = E.ShadeSample(src, tx);
            float a = 1.0f - std::exp(-src.w * h);
            float r = src.x * a, g = src.y * a, b = src.z * a;
            float om = 1.0f - da;
            dr = dr + om * r; dg = dg + om * g; db = db + om * b; da = da + om * a;
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

// K9: rc1pextbsd/lightcachecomputation.comp main (:443-469), dispatched by PreComputeLightCache (ebsrenderer.cpp:441-555).
// ExtinctionAmbientOcclusion / ExtinctionDirectionalShadows are the marcher's functions (CONE_RESCALE_CEIL_INTERVAL is
// the active variant in both files).  out_rg: rw*rh*rd*2 floats, fp16-rounded.
int orc_ebs_light_cache(int vw, int vh, int vd, const float voxel_scale[3], const float* sat_f32, const Lighting* light,
                        const EbsParams* prm, int rw, int rh, int rd, float* out_rg) {
  Ebs E;
  E.vol.w = vw; E.vol.h = vh; E.vol.d = vd; E.vol.c = 1; E.vol.data = nullptr;
  E.sat.w = vw + 2; E.sat.h = vh + 2; E.sat.d = vd + 2; E.sat.c = 1; E.sat.data = sat_f32;
  E.VS = v3(voxel_scale[0], voxel_scale[1], voxel_scale[2]);
  E.VSS = v3((float)vw, (float)vh, (float)vd) * E.VS;
  E.MinSAT = E.VS * 0.5f; E.MaxSAT = E.VSS + E.VS * 1.5f;
  E.MinVol = E.VS * 0.5f; E.MaxVol = E.VSS - E.VS * 0.5f;
  E.inv_vol_scaled = v3(1.0f, 1.0f, 1.0f) / (E.VSS + E.VS * 2.0f);
  E.P = *prm; E.L = *light;
  const V3 cell = v3(voxel_scale[0] * ((float)vw / (float)rw), voxel_scale[1] * ((float)vh / (float)rh), voxel_scale[2] * ((float)vd / (float)rd));
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
  for (int z = 0; z < rd; ++z)
    for (int y = 0; y < rh; ++y)
      for (int x = 0; x < rw; ++x) {
        V3 tex_pos = v3(((float)x + 0.5f) * cell.x, ((float)y + 0.5f) * cell.y, ((float)z + 0.5f) * cell.z);
        float Iao = 1.0f, Ids = 1.0f;
        if (prm->apply_occlusion == 1) Iao = E.ExtinctionAmbientOcclusion(tex_pos);
        if (prm->apply_shadow == 1) Ids = E.ExtinctionDirectionalShadows(tex_pos);
        float* o = out_rg + 2 * ((size_t)x + (size_t)rw * ((size_t)y + (size_t)rh * (size_t)z));
        o[0] = round_f16(Iao); o[1] = round_f16(Ids);
      }
  return 0;
}

}  // extern "C"
