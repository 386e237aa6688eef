// oracle/oracle_gt.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// CPU restatement of the cone "ground truth" renderer: rc1pcrtgt/gt_ray_marching.comp (main :371-483,
// ConeOcclusionEvaluationRayCasting :109-169, ConeShadowsEvaluationRayCasting :171-252, ShadeSample :254-302) driven
// to convergence the way RC1PConeLightGroundTruthSteps::RedrawFrameTexture does (crtgtrenderer.cpp:272-325): one
// primary sample per dispatch, running colour round-tripped through an rgba16f image and the ray parameter through
// an rg16f image between dispatches.  The converged image is what this restatement returns: the per-dispatch fp16
// round trips of colour and s are applied after every sample.  Ray-direction tables are explicit inputs (the
// reference draws them from an implementation-defined std::default_random_engine, crtgtrenderer.cpp:131-187).
// Pinned against gt_ray_marching.comp run on the CPU and re-dispatched to convergence (tests/test_refglsl.py); see oracle_common.h.
#include "oracle_common.h"
#include <omp.h>

using namespace orc;

extern "C" {

struct GtParams {
  float step_size;
  float light_ray_initial_gap, light_ray_step_size;
  int apply_occlusion, occ_num_rays; float occ_cone_distance;
  int apply_shadow, sdw_num_rays; float sdw_cone_distance; int shadow_type;
  int count_samples;
};

namespace {
struct Gt {
  Tex3D vol; Tex1D tf; V3 G; GtParams P; Lighting L;
  const float* occ_rays; const float* sdw_rays;     // n x 3, fp16-rounded (RGB16F texelFetch)

  float cone(const float* table, int nrays, float dist_eval, V3 tx, V3 v_right, V3 v_up, V3 v_dir, uint64_t* nsteps) const {
    float S = 0.0f, Sw = 0.0f;
    for (int rayid = 0; rayid < nrays; ++rayid) {
      V3 c = v3(table[3 * rayid], table[3 * rayid + 1], table[3 * rayid + 2]);
      V3 w = normalize(v_right * c.x + v_up * c.y + v_dir * c.z);
      float Vt = 1.0f;
      float s = P.light_ray_initial_gap;
      float density0 = tex3d(vol, (tx + s * w) / G);
      float st0 = tex1d(tf, density0).w;
      while (s < dist_eval) {
        float h = std::fmin(P.light_ray_step_size, dist_eval - s);
        V3 at = tx + (s + h) * w;
        if (at.x < 0.0f || at.x > G.x || at.y < 0.0f || at.y > G.y || at.z < 0.0f || at.z > G.z) break;
        float density1 = tex3d(vol, at / G);
        float st1 = tex1d(tf, density1).w;
        Vt *= std::exp(-((st0 + st1) * 0.5f) * h);
        if (nsteps) ++*nsteps;
        if ((1 - Vt) > 0.99f) break;
        st0 = st1;
        s = s + h;
      }
      float rw = dot(v_dir, w);
      S += Vt * rw;
      Sw += rw;
    }
    return (S / Sw);
  }
  float Occlusion(V3 tx, V3 v_up, V3 v_right, V3 ray_dir, uint64_t* n) const {
    V3 v_dir = normalize(-ray_dir);
    return cone(occ_rays, P.occ_num_rays, P.occ_cone_distance, tx, v_right, v_up, v_dir, n);
  }
  float Shadow(V3 tx, uint64_t* n) const {
    V3 Wpos = tx - (G * 0.5f);
    V3 lp = v3(L.light_pos[0], L.light_pos[1], L.light_pos[2]);
    V3 fwd = v3(L.light_forward[0], L.light_forward[1], L.light_forward[2]);
    V3 upv = v3(L.light_up[0], L.light_up[1], L.light_up[2]);
    V3 rgt = v3(L.light_right[0], L.light_right[1], L.light_right[2]);
    V3 v_dir = v3(0, 0, 0), v_up = v_dir, v_right = v_dir;
    if (P.shadow_type == 0 || P.shadow_type == 1) {
      v_dir = normalize(lp - Wpos);
      v_up = normalize(cross(v_dir, rgt));
      v_right = normalize(cross(v_dir, v_up));
      // reference quirk: a cosine is compared with 30.0, so a spot light is always dark
      if (P.shadow_type == 1 && dot(v_dir, fwd) < 30.0f) return 0.0f;
    } else if (P.shadow_type == 2) {
      v_dir = fwd; v_up = upv; v_right = rgt;
    }
    return cone(sdw_rays, P.sdw_num_rays, P.sdw_cone_distance, tx, v_right, v_up, v_dir, n);
  }
};
}  // namespace

// light_forward in `light` is the LightCamForward UNIFORM of this shader, which the host sets to
// -GetBlinnPhongLightSourceCameraForward() (crtgtrenderer.cpp:226-236).
int orc_gt_render(const float* vol_r16f, int vw, int vh, int vd, const float grid_size[3], const float* tf_rgbt, int tf_n,
                  const Camera* cam, const Lighting* light, const GtParams* prm, const float* occ_rays, const float* sdw_rays,
                  int W, int H, float* out_rgba, uint32_t* out_nsamples, uint64_t* out_secondary_steps) {
  Gt g;
  g.vol.w = vw; g.vol.h = vh; g.vol.d = vd; g.vol.c = 1; g.vol.data = vol_r16f;
  g.tf.n = tf_n; g.tf.data = tf_rgbt;
  g.G = v3(grid_size[0], grid_size[1], grid_size[2]);
  g.P = *prm; g.L = *light; g.occ_rays = occ_rays; g.sdw_rays = sdw_rays;
  const V3 G = g.G;
  const V3 eye = v3(cam->eye[0], cam->eye[1], cam->eye[2]);
  const Tex3D* grad = (light->apply_phong == 1) ? gradient_texture() : nullptr;
  if (light->apply_phong == 1 && !grad) return -2;
  uint64_t total_steps = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total_steps)
  for (int py = 0; py < H; ++py) {
    for (int px = 0; px < W; ++px) {
      float* o = out_rgba + 4 * ((size_t)py * W + px);
      o[0] = o[1] = o[2] = o[3] = 0.0f;
      uint32_t ns = 0;
      uint64_t nsec = 0;
      V3 cdir = pixel_ray_dir(*cam, px, py, W, H);
      V3 dir; float tnear, tfar;
      bool inbox = ray_aabb(eye, cdir, -G * 0.5f, G * 0.5f, &dir, &tnear, &tfar);
      if (inbox) {
        V3 v_right = normalize(cross(cdir, v3(0, 1, 0)));
        V3 v_up = normalize(cross(-cdir, v_right));
        float D = tfar - tnear;                        // no abs() in this shader (:398)
        float cr = 0, cg = 0, cb = 0, ca = 0;          // state after imageStore/imageLoad round trips (fp16)
        V3 wld = eye + dir * tnear;
        V3 tex_pos = wld + (G * 0.5f);
        float s = 0.0f;                                // ifrag.x, an rg16f texel
        while (s < D) {                                // one iteration == one dispatch
          float h = std::fmin(prm->step_size, D - s);
          V3 sp = tex_pos + dir * (s + h * 0.5f);
          float density = tex3d(g.vol, sp / G);
          V4 src = tex1d(g.tf, density);
          ++ns;
          bool done = false;
          if (src.w > 0.0f) {
            // ShadeSample (:254-302); v_dir argument is camera_dir
            float ka = 0.0f, kd = 0.0f, ks = 0.0f, IOcc = 0.0f, ISdw = 0.0f;
            if (prm->apply_occlusion == 1) { ka = light->ka; IOcc = g.Occlusion(sp, v_up, v_right, cdir, &nsec); }
            if (prm->apply_shadow == 1) { kd = light->kd; ks = light->ks; ISdw = g.Shadow(sp, &nsec); }
            float r, gg, b;
            if (grad) {                                      // ApplyGradientPhongShading == 1 (:277-295), specular colour vec3(1)
              r = src.x; gg = src.y; b = src.z;
              float dot_diff, spec;
              if (phong_terms(*grad, sp, G, v3(light->light_pos[0], light->light_pos[1], light->light_pos[2]), eye, light->shininess, &dot_diff, &spec)) {
                float f = ((1.0f / (ka + kd)) * (IOcc * ka + ISdw * kd * dot_diff));
                float sc = (ISdw * ks * spec);
                r = src.x * f + 1.0f * sc; gg = src.y * f + 1.0f * sc; b = src.z * f + 1.0f * sc;
              }
            } else {
              float kk = (1.0f / (ka + kd));
              r = kk * (src.x * IOcc * ka + src.x * ISdw * kd);
              gg = kk * (src.y * IOcc * ka + src.y * ISdw * kd);
              b = kk * (src.z * IOcc * ka + src.z * ISdw * kd);
            }
            float a = 1.0f - std::exp(-src.w * h);
            float om = 1.0f - ca;
            cr = cr + om * (r * a); cg = cg + om * (gg * a); cb = cb + om * (b * a); ca = ca + om * a;
            if (ca > 0.99f) done = true;               // state <- (s, 1); colour stored; finished
          }
          // imageStore(OutputFrag) at the end of every dispatch: rgba16f
          cr = round_f16(cr); cg = round_f16(cg); cb = round_f16(cb); ca = round_f16(ca);
          if (done) break;
          s = s + h;
          if (!(s < D)) break;                         // state <- (s, 1)
          s = round_f16(s);                            // state <- (s, 0): rg16f round trip before the next dispatch
        }
        o[0] = cr; o[1] = cg; o[2] = cb; o[3] = ca;
      }
      total_steps += nsec;
      if (out_nsamples) out_nsamples[(size_t)py * W + px] = ns;
    }
  }
  if (out_secondary_steps) *out_secondary_steps = total_steps;
  return 0;
}

// RedrawCube (crtgtrenderer.cpp:327-338) = rc1pcrtgt/vol_intersection.comp:64-110: what the reference shows instead of the
// ground-truth frame while "Show Generated Frame Texture" is off (its default, and after every parameter or camera change):
// the entry point of each ray on the volume's bounding box, coloured by the face it lies on (blue z, green y, red x; ties go to
// z, then y).  Pixels whose ray misses stay cleared.
int orc_gt_cube_render(const float grid_size[3], const Camera* cam, int W, int H, float* out_rgba) {
  const V3 G = v3(grid_size[0], grid_size[1], grid_size[2]);
  const V3 eye = v3(cam->eye[0], cam->eye[1], cam->eye[2]);
#pragma omp parallel for schedule(static)
  for (int py = 0; py < H; ++py)
    for (int px = 0; px < W; ++px) {
      float* o = out_rgba + 4 * ((size_t)py * W + px);
      o[0] = o[1] = o[2] = o[3] = 0.0f;
      V3 cdir = pixel_ray_dir(*cam, px, py, W, H);
      V3 dir; float tnear, tfar;
      if (!ray_aabb(eye, cdir, -G * 0.5f, G * 0.5f, &dir, &tnear, &tfar)) continue;
      V3 wld = eye + dir * tnear;
      V3 c = vabs(wld) / (G * 0.5f);
      if (c.z >= c.x && c.z >= c.y) { o[2] = 1.0f; o[3] = 1.0f; }
      else if (c.y >= c.x && c.y >= c.z) { o[1] = 1.0f; o[3] = 1.0f; }
      else if (c.x >= c.y && c.x >= c.z) { o[0] = 1.0f; o[3] = 1.0f; }
    }
  return 0;
}

}  // extern "C"
