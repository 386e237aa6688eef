// oracle/oracle_rc1pass.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// CPU restatement of cppvolrend/structured/rc1pass/ray_marching_1p.comp:85-179 (main) with
// _common_shaders/ray_bbox_intersection.comp:18-52, uniforms as uploaded by rc1prenderer.cpp:72-138,231-262.
// Gradient Blinn-Phong (ShadeBlinnPhong, ray_marching_1p.comp:48-81; off by default, datamanager.cpp:27) is restated in
// orc_rc1pass_render_lit: it needs the lighting uniforms and the gradient texture bound with orc_set_gradient.
// Pinned against the reference's own GLSL run on the CPU (tests/test_refglsl.py); see oracle_common.h.
#include "oracle_common.h"
#include <omp.h>

using namespace orc;

extern "C" {

// vol_r16f: W*H*D fp16-rounded texel values (orc_volume_to_r16f); tf_rgbt: tf_n x 4 fp16-rounded texels
// (orc_tf_texture_rgbt).  grid_size = resolution * voxel scale (VolumeGridSize).  out_rgba: W*H*4 floats, row 0 =
// bottom (GL image coords), fp16-rounded like imageStore into rgba16f; pixels whose ray misses stay 0
// (renderoutputframe.cpp:187-190).  out_nsamples (optional): loop iterations executed per pixel.
int orc_rc1pass_render_lit(const float* vol_r16f, int vw, int vh, int vd, const float grid_size[3],
                           const float* tf_rgbt, int tf_n, const Camera* cam, float step_size,
                           int W, int H, float* out_rgba, uint32_t* out_nsamples, const Lighting* light) {
  const Tex3D* grad = (light && light->apply_phong == 1) ? gradient_texture() : nullptr;
  if (light && light->apply_phong == 1 && !grad) return -2;        // ApplyGradientPhongShading needs TexVolumeGradient
  Tex3D vol; vol.w = vw; vol.h = vh; vol.d = vd; vol.c = 1; vol.data = vol_r16f;
  Tex1D tf; tf.n = tf_n; tf.data = tf_rgbt;
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
      bool inbox = ray_aabb(eye, cdir, -G * 0.5f, G * 0.5f, &dir, &tnear, &tfar);
      if (inbox) {
        float D = std::fabs(tfar - tnear);
        float dr = 0, dg = 0, db = 0, da = 0;
        V3 wld = eye + dir * tnear;
        V3 tex_pos = wld + (G * 0.5f);
        for (float s = 0.0f; s < D;) {
          float h = std::fmin(step_size, D - s);
          V3 sp = tex_pos + dir * (s + h * 0.5f);
          float density = tex3d(vol, sp / G);
          V4 src = tex1d(tf, density);
          ++ns;
          if (src.w > 0.0f) {
            if (grad) {                                     // ShadeBlinnPhong (:48-81)
              float dot_diff, spec;
              if (phong_terms(*grad, sp, G, v3(light->light_pos[0], light->light_pos[1], light->light_pos[2]), eye, light->shininess, &dot_diff, &spec)) {
                float kad = light->ka + light->kd * dot_diff;
                src.x = src.x * kad + light->ispecular[0] * light->ks * spec;
                src.y = src.y * kad + light->ispecular[1] * light->ks * spec;
                src.z = src.z * kad + light->ispecular[2] * light->ks * spec;
              }
            }
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

int orc_rc1pass_render(const float* vol_r16f, int vw, int vh, int vd, const float grid_size[3],
                       const float* tf_rgbt, int tf_n, const Camera* cam, float step_size,
                       int W, int H, float* out_rgba, uint32_t* out_nsamples) {
  return orc_rc1pass_render_lit(vol_r16f, vw, vh, vd, grid_size, tf_rgbt, tf_n, cam, step_size, W, H, out_rgba, out_nsamples, nullptr);
}

int orc_num_threads(void) { return omp_get_max_threads(); }
// the bench's reference arm runs under torchrun, which exports OMP_NUM_THREADS=1: the arm sets the thread count itself
void orc_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

}  // extern "C"
