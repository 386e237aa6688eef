// oracle/oracle_iso.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// CPU restatement of cppvolrend/structured/rc1pisoadapt/ray_marching_1p_iso_adapt.comp (main :91-173, ShadeBlinnPhong
// :48-88) with the uniforms of RayCasting1PassIsoAdapt::Update (rc1pisoadaptrenderer.cpp:113-165).
// Pinned against the reference's own GLSL run on the CPU (tests/test_refglsl.py); see oracle_common.h.
#include "oracle_common.h"

using namespace orc;

struct IsoParams {
  float isovalue, step_size_small, step_size_large, step_size_range;
  float color[4];
  int count_samples;
};

extern "C" {

int orc_iso_render(const float* vol_r16f, int vw, int vh, int vd, const float grid_size[3], const Camera* cam, const Lighting* light,
                   const IsoParams* P, int W, int H, float* out_rgba, uint32_t* out_nsamples) {
  const Tex3D* grad = (light && light->apply_phong == 1) ? gradient_texture() : nullptr;
  if (light && light->apply_phong == 1 && !grad) return -2;
  Tex3D vol; vol.w = vw; vol.h = vh; vol.d = vd; vol.c = 1; vol.data = vol_r16f;
  V3 G = v3(grid_size[0], grid_size[1], grid_size[2]);
  V3 eye = v3(cam->eye[0], cam->eye[1], cam->eye[2]);
#pragma omp parallel for schedule(dynamic, 4)
  for (int py = 0; py < H; ++py) {
    for (int px = 0; px < W; ++px) {
      float* o = out_rgba + 4 * ((size_t)py * W + px);
      o[0] = o[1] = o[2] = o[3] = 0.0f;
      uint32_t ns = 0;
      V3 cdir = pixel_ray_dir(*cam, px, py, W, H);
      V3 dir; float tnear, tfar;
      if (ray_aabb(eye, cdir, -G * 0.5f, G * 0.5f, &dir, &tnear, &tfar)) {
        float D = std::fabs(tfar - tnear);
        float dr = 0, dg = 0, db = 0, da = 0;
        V3 tex_pos = (eye + dir * tnear) + (G * 0.5f);
        float prevDensity = tex3d(vol, tex_pos / G);
        for (float s = 0.0f; s < D;) {
          float cur = (std::fabs(prevDensity - P->isovalue) < P->step_size_range) ? P->step_size_small : P->step_size_large;
          float h = std::fmin(cur, D - s);
          V3 sp = tex_pos + dir * (s + h);
          float density = tex3d(vol, sp / G);
          ++ns;
          if ((prevDensity <= P->isovalue && P->isovalue < density) || (prevDensity >= P->isovalue && P->isovalue > density)) {
            float t = (P->isovalue - prevDensity) / (density - prevDensity);
            sp = tex_pos + dir * (s + t * h);
            float cr = P->color[0], cg = P->color[1], cb = P->color[2], ca = P->color[3];
            if (grad) {
              float dot_diff, spec;
              if (phong_terms(*grad, sp, G, v3(light->light_pos[0], light->light_pos[1], light->light_pos[2]), eye, light->shininess, &dot_diff, &spec)) {
                float kad = light->ka + light->kd * dot_diff;
                cr = cr * kad + light->ispecular[0] * light->ks * spec;
                cg = cg * kad + light->ispecular[1] * light->ks * spec;
                cb = cb * kad + light->ispecular[2] * light->ks * spec;
              }
            }
            float om = 1.0f - da;
            dr = dr + om * (cr * ca); dg = dg + om * (cg * ca); db = db + om * (cb * ca); da = da + om * ca;
            if (da > 0.99f) break;
          }
          prevDensity = density;
          s = s + h;
        }
        o[0] = round_f16(dr); o[1] = round_f16(dg); o[2] = round_f16(db); o[3] = round_f16(da);
      }
      if (out_nsamples) out_nsamples[(size_t)py * W + px] = ns;
    }
  }
  return 0;
}

}  // extern "C"
