// oracle/oracle_vct.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// CPU restatement of the voxel-cone-traced shadow renderer (Shih et al. 2016, single GPU version):
//   - VCTPreProcessing::PreProcessSuperVoxels (rc1pvctsg/preprocessingstages.cpp:35-145): mean / stddev "super voxel"
//     pyramid, 2x box down-sampling in double, uploaded per level as RG16F;
//   - OpacityGaussianEvaluation / PreProcessPreIntegrationTable (:147-202): R16F look-up table [density][stddev];
//   - rc1pvctsg/vct_ray_bbox_marching.comp (EvaluationVoxelConeTracing :97-144, ShadeSample :146-189, main :191-264),
//     uniforms as uploaded by vctrenderer.cpp:124-237.
// Marcher and light cache: pinned against the reference's own GLSL run on the CPU (tests/test_refglsl.py).  Pre-passes: pinned
// bit for bit against preprocessingstages.cpp compiled in place (oracle/_ref/libref.so, tests/test_oracle_ref.py).
#include "oracle_common.h"
#include <omp.h>

using namespace orc;

extern "C" {

// Number of levels PreProcessSuperVoxels creates: halve (floor) until one dimension reaches 0 (:62-66,113-119).
int orc_vct_levels(int w, int h, int d) {
  int n = 1;
  w /= 2; h /= 2; d /= 2;
  while ((long long)w * h * d >= 1) { ++n; w /= 2; h /= 2; d /= 2; }
  return n;
}

// levels_out: concatenated levels, each w*h*d x (mean, stddev) floats, fp16-ROUNDED (RG16F).  level_dims: n x 3.
// Returns maximum_standard_deviation (double, over all levels) through *max_stddev.
int orc_vct_supervoxels(const void* vox, int vw, int vh, int vd, int bpv, float* levels_out, size_t cap_floats, int* level_dims,
                        double* max_stddev) {
  const int nlev = orc_vct_levels(vw, vh, vd);
  std::vector<std::vector<double>> mean(nlev);
  int w = vw, h = vh, d = vd;
  size_t total = 0;
  std::vector<size_t> off(nlev);
  for (int l = 0; l < nlev; ++l) {
    level_dims[3 * l] = w; level_dims[3 * l + 1] = h; level_dims[3 * l + 2] = d;
    off[l] = total; total += (size_t)w * h * d * 2;
    w /= 2; h /= 2; d /= 2;
  }
  if (total > cap_floats) return -1;
  // level 0: mean = GetNormalizedSample * 255.0 for EVERY storage type, stddev 0
  {
    const size_t n = (size_t)vw * vh * vd;
    mean[0].resize(n);
    const double maxv = bpv == 1 ? (256.0 - 1.0) : (65536.0 - 1.0);
    for (size_t i = 0; i < n; ++i) {
      double v = bpv == 1 ? (double)((const uint8_t*)vox)[i] : (double)((const uint16_t*)vox)[i];
      mean[0][i] = (v / maxv) * 255.0;
      levels_out[off[0] + 2 * i] = round_f16((float)mean[0][i]);
      levels_out[off[0] + 2 * i + 1] = 0.0f;
    }
  }
  double mx = 0.0;
  for (int l = 1; l < nlev; ++l) {
    const int pw = level_dims[3 * (l - 1)], ph = level_dims[3 * (l - 1) + 1];
    const int cw = level_dims[3 * l], ch = level_dims[3 * l + 1], cd = level_dims[3 * l + 2];
    mean[l].resize((size_t)cw * ch * cd);
    const std::vector<double>& P = mean[l - 1];
    auto pm = [&](int x, int y, int z) { return P[(size_t)x + (size_t)y * pw + (size_t)z * pw * ph]; };
    for (int iz = 0; iz < cd; ++iz)
      for (int iy = 0; iy < ch; ++iy)
        for (int ix = 0; ix < cw; ++ix) {
          int lw = ix * 2, lh = iy * 2, ld = iz * 2;
          double vm0 = pm(lw, lh, ld), vm1 = pm(lw, lh, ld + 1), vm2 = pm(lw, lh + 1, ld), vm3 = pm(lw, lh + 1, ld + 1);
          double vm4 = pm(lw + 1, lh, ld), vm5 = pm(lw + 1, lh, ld + 1), vm6 = pm(lw + 1, lh + 1, ld), vm7 = pm(lw + 1, lh + 1, ld + 1);
          double vmn = (vm0 + vm1 + vm2 + vm3 + vm4 + vm5 + vm6 + vm7) / 8.0;
          double vstdd = std::sqrt((std::pow(vm0 - vmn, 2.0) + std::pow(vm1 - vmn, 2.0) + std::pow(vm2 - vmn, 2.0) + std::pow(vm3 - vmn, 2.0) +
                                    std::pow(vm4 - vmn, 2.0) + std::pow(vm5 - vmn, 2.0) + std::pow(vm6 - vmn, 2.0) + std::pow(vm7 - vmn, 2.0)) / 8.0);
          size_t i = (size_t)ix + (size_t)iy * cw + (size_t)iz * cw * ch;
          mean[l][i] = vmn;
          levels_out[off[l] + 2 * i] = round_f16((float)vmn);
          levels_out[off[l] + 2 * i + 1] = round_f16((float)vstdd);
          mx = std::max(vstdd, mx);
        }
  }
  *max_stddev = mx;
  return nlev;
}

// Pre-integration table.  opc_by_density[i] = tf->GetOpc(i, dens_val) for i in [0, int(dens_val)] (one extra entry for
// the stddev == 0 row, which evaluates GetOpc(mean) at iw <= w-1).  Rows [row0, row1) of the h = ceil(max_stddev)
// rows are computed (all when row1 < 0).  out: w x (row1-row0) floats, fp16-rounded, x (density) fastest.
void orc_vct_preintegration(const float* opc_by_density, int dens_val, double max_stddev, int row0, int row1, float* out) {
  const int w = (int)std::ceil((double)dens_val);
  const int h = (int)std::ceil(max_stddev);
  if (row1 < 0) { row0 = 0; row1 = h; }
  const double s2pi = std::sqrt(2.0 * 3.14159265358979323846264338327950288);
#pragma omp parallel for schedule(dynamic, 16) collapse(2)
  for (int ih = row0; ih < row1; ++ih)
    for (int iw = 0; iw < w; ++iw) {
      double mean = iw, stddev = ih, SumG = 0.0, SumW = 0.0;
      if (std::fabs(stddev) > 0.0001) {
        for (int i = 0; i < dens_val; i++) {
          double nf = 1.0 / (stddev * s2pi);
          double W = nf * std::exp(-(((double)i - mean) * ((double)i - mean)) / (2.0 * stddev * stddev));
          SumG += W * (double)opc_by_density[i];
          SumW += W;
        }
        SumG = SumG / SumW;
      } else {
        SumG = (double)opc_by_density[iw];
      }
      out[(size_t)iw + (size_t)(ih - row0) * w] = round_f16((float)SumG);
    }
}

struct VctParams {
  float step_size;
  int apply_occlusion, apply_shadow;
  float tan_cone_apex_angle, cone_step_size, cone_step_increase_rate, cone_initial_step;
  float opacity_correction_factor; int apply_opacity_correction;
  int cone_number_of_samples;
  float volume_max_density, volume_max_stddev;
  int count_samples;
};

// 2-D R16F texture, GL_LINEAR, clamp to edge
static float tex2d(const float* t, int w, int h, float sx, float sy) {
  int x0, x1, y0, y1; float fx, fy;
  lin_coord(sx, w, &x0, &x1, &fx);
  lin_coord(sy, h, &y0, &y1, &fy);
  float a = lerp(t[(size_t)x0 + (size_t)w * y0], t[(size_t)x1 + (size_t)w * y0], fx);
  float b = lerp(t[(size_t)x0 + (size_t)w * y1], t[(size_t)x1 + (size_t)w * y1], fx);
  return lerp(a, b, fy);
}

int orc_vct_render(const float* vol_r16f, int vw, int vh, int vd, const float voxel_scale[3], const float* sv_levels,
                   const int* level_dims, int n_levels, const float* lut, int lut_w, int lut_h, const float* tf_rgbt, int tf_n,
                   const Camera* cam, const Lighting* light, const VctParams* prm, int W, int H, float* out_rgba, uint32_t* out_nsamples) {
  Tex3D vol; vol.w = vw; vol.h = vh; vol.d = vd; vol.c = 1; vol.data = vol_r16f;
  Tex1D tf; tf.n = tf_n; tf.data = tf_rgbt;
  Tex3DMip sv;
  size_t off = 0;
  for (int l = 0; l < n_levels; ++l) {
    Tex3D t; t.w = level_dims[3 * l]; t.h = level_dims[3 * l + 1]; t.d = level_dims[3 * l + 2]; t.c = 2; t.data = sv_levels + off;
    off += (size_t)t.w * t.h * t.d * 2;
    sv.levels.push_back(t);
  }
  const V3 VSS = v3((float)vw * voxel_scale[0], (float)vh * voxel_scale[1], (float)vd * voxel_scale[2]);
  const V3 eye = v3(cam->eye[0], cam->eye[1], cam->eye[2]);
  const V3 lpos = v3(light->light_pos[0], light->light_pos[1], light->light_pos[2]);
  const V3 InvG = v3(1.0f, 1.0f, 1.0f) / VSS;
  const Tex3D* grad = (light->apply_phong == 1) ? gradient_texture() : nullptr;
  if (light->apply_phong == 1 && !grad) return -2;
  const float corr_fact = (float)prm->apply_opacity_correction * prm->opacity_correction_factor;
  auto cone = [&](V3 tex_pos) -> float {     // EvaluationVoxelConeTracing (:97-144)
    float Tvd = 1.0f;
    V3 realpos = tex_pos - (VSS * 0.5f);
    V3 cone_vec = normalize(lpos - realpos);
    float apex_distance = prm->cone_initial_step;
    float step_size = prm->cone_step_size;
    const float DXbase = 1.0f;
    for (int is = 0; is < prm->cone_number_of_samples; ++is) {
      float xl_x = (apex_distance + step_size * 0.5f);
      float mm_level = std::log2((2.0f * xl_x * prm->tan_cone_apex_angle) / DXbase);
      V3 wpos = (tex_pos + cone_vec * xl_x);
      if (wpos.x < 0 || wpos.x > VSS.x || wpos.y < 0 || wpos.y > VSS.y || wpos.z < 0 || wpos.z > VSS.z) break;
      V3 p = (tex_pos + cone_vec * xl_x) / VSS;
      float g_m = sv.lod(p, mm_level, 0), g_s = sv.lod(p, mm_level, 1);
      float opacity = tex2d(lut, lut_w, lut_h, (g_m + 0.5f) / prm->volume_max_density, (g_s + 0.5f) / prm->volume_max_stddev);
      opacity = 1.0f - std::pow(1.0f - opacity, step_size * corr_fact);
      Tvd *= (1.0f - opacity);
      apex_distance = apex_distance + step_size;
      step_size = step_size * prm->cone_step_increase_rate;
    }
    return Tvd;
  };
#pragma omp parallel for schedule(dynamic, 1)
  for (int py = 0; py < H; ++py) {
    for (int px = 0; px < W; ++px) {
      float* o = out_rgba + 4 * ((size_t)py * W + px);
      o[0] = o[1] = o[2] = o[3] = 0.0f;
      uint32_t ns = 0;
      V3 cdir = pixel_ray_dir(*cam, px, py, W, H);
      V3 dir; float tnear, tfar;
      bool inbox = ray_aabb(eye, cdir, -VSS * 0.5f, VSS * 0.5f, &dir, &tnear, &tfar);
      if (inbox) {
        float D = std::fabs(tfar - tnear);
        float dr = 0, dg = 0, db = 0, da = 0;
        V3 wd = eye + dir * tnear;
        wd = wd + (VSS * 0.5f);
        for (float s = 0.0f; s < D;) {
          float h = std::fmin(prm->step_size, D - s);
          V3 tx = wd + dir * (s + h * 0.5f);
          float density = tex3d(vol, tx * InvG);
          V4 src = tex1d(tf, density);
          ++ns;
          if (src.w > 0.0f) {
            float ka = 0.0f, kd = 0.0f, ks = 0.0f, Ivd = 0.0f;
            if (prm->apply_occlusion == 1) ka = light->ka;
            if (prm->apply_shadow == 1) { kd = light->kd; ks = light->ks; Ivd = cone(tx); }
            float cr, cg, cb;
            if (grad) {                                      // ApplyPhongShading == 1 (:163-182)
              cr = src.x; cg = src.y; cb = src.z;
              float dot_diff, spec;
              if (phong_terms(*grad, tx, VSS, v3(light->light_pos[0], light->light_pos[1], light->light_pos[2]), eye, light->shininess, &dot_diff, &spec)) {
                float kk = (1.0f / (ka + kd));
                cr = kk * (src.x * ka + Ivd * (src.x * kd * dot_diff)) + Ivd * (ks * light->ispecular[0] * spec);
                cg = kk * (src.y * ka + Ivd * (src.y * kd * dot_diff)) + Ivd * (ks * light->ispecular[1] * spec);
                cb = kk * (src.z * ka + Ivd * (src.z * kd * dot_diff)) + Ivd * (ks * light->ispecular[2] * spec);
              }
            } else {
              float kk = (1.0f / (ka + kd));
              cr = kk * (src.x * ka + src.x * Ivd * kd);
              cg = kk * (src.y * ka + src.y * Ivd * kd);
              cb = kk * (src.z * ka + src.z * Ivd * kd);
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

// K13: rc1pvctsg/lightcachecomputation.comp main (:85-118), dispatched by PreComputeLightCache (vctrenderer.cpp:393-515).
// Its EvaluationVoxelConeTracing (:45-83) differs from the marcher's: CUT_WHEN_AWAY_FROM_VOLUME is NOT defined (:3), so the
// cone keeps stepping outside the volume (clamp-to-edge samples).  Iao is always 1.0.  out_rg: fp16-rounded pairs.
int orc_vct_light_cache(int vw, int vh, int vd, const float voxel_scale[3], const float* sv_levels, const int* level_dims, int n_levels,
                        const float* lut, int lut_w, int lut_h, const Lighting* light, const VctParams* prm, int rw, int rh, int rd,
                        float* out_rg) {
  Tex3DMip sv;
  size_t off = 0;
  for (int l = 0; l < n_levels; ++l) {
    Tex3D t; t.w = level_dims[3 * l]; t.h = level_dims[3 * l + 1]; t.d = level_dims[3 * l + 2]; t.c = 2; t.data = sv_levels + off;
    off += (size_t)t.w * t.h * t.d * 2;
    sv.levels.push_back(t);
  }
  const V3 VSS = v3((float)vw * voxel_scale[0], (float)vh * voxel_scale[1], (float)vd * voxel_scale[2]);
  const V3 lpos = v3(light->light_pos[0], light->light_pos[1], light->light_pos[2]);
  const float corr_fact = (float)prm->apply_opacity_correction * prm->opacity_correction_factor;
  const V3 cell = v3(voxel_scale[0] * ((float)vw / (float)rw), voxel_scale[1] * ((float)vh / (float)rh), voxel_scale[2] * ((float)vd / (float)rd));
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
  for (int z = 0; z < rd; ++z)
    for (int y = 0; y < rh; ++y)
      for (int x = 0; x < rw; ++x) {
        V3 tex_pos = v3(((float)x + 0.5f) * cell.x, ((float)y + 0.5f) * cell.y, ((float)z + 0.5f) * cell.z);
        V3 realpos = tex_pos - (VSS * 0.5f);
        float Iao = 1.0f, Ivd = 1.0f;
        if (prm->apply_shadow == 1) {
          float Tvd = 1.0f;
          V3 cone_vec = normalize(lpos - realpos);
          float apex_distance = prm->cone_initial_step;
          float step_size = prm->cone_step_size;
          const float DXbase = 1.0f;
          for (int is = 0; is < prm->cone_number_of_samples; ++is) {
            float xl_x = (apex_distance + step_size * 0.5f);
            float mm_level = std::log2((2.0f * xl_x * prm->tan_cone_apex_angle) / DXbase);
            V3 p = (tex_pos + cone_vec * xl_x) / VSS;
            float g_m = sv.lod(p, mm_level, 0), g_s = sv.lod(p, mm_level, 1);
            float opacity = tex2d(lut, lut_w, lut_h, (g_m + 0.5f) / prm->volume_max_density, (g_s + 0.5f) / prm->volume_max_stddev);
            opacity = 1.0f - std::pow(1.0f - opacity, step_size * corr_fact);
            Tvd *= (1.0f - opacity);
            apex_distance = apex_distance + step_size;
            step_size = step_size * prm->cone_step_increase_rate;
          }
          Ivd = Tvd;
        }
        float* o = out_rg + 2 * ((size_t)x + (size_t)rw * ((size_t)y + (size_t)rh * (size_t)z));
        o[0] = round_f16(Iao); o[1] = round_f16(Ivd);
      }
  return 0;
}

}  // extern "C"
