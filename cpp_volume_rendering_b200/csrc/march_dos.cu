// march_dos.cu -- directional ambient occlusion + cone shadows by cone tracing a Gaussian extinction pyramid
// (Campagnolo & Celes 2019), sm_100a.  Replaces the dispatch of rc1pdosct/ray_bbox_marching.comp made by
// RC1PConeTracingDirOcclusionShading::Redraw; uniforms as uploaded by dosrcrenderer.cpp:134-247,805-985.
// CONSIDER_BORDERS is defined in the shader, the other switches (USE_EARLY_TERMINATION, ALWAYS_SPLIT_CONES,
// USE_FALLOFF_FUNCTION) are off, and so they are here.
//
// Per shaded sample the work is ~18 (AO: 1 + 3x17 taps) and ~159 (shadow) trilinear taps of an fp16 pyramid that is
// 4.8 MB in total: the kernel lives in L1/L2, one thread per ray, 8x4 ray tile per warp.  Compiled with -fmad=false:
// operation order follows the shader / oracle, only the GL_LINEAR blends are explicit fmaf().
#include "vrb_internal.cuh"
#include <vector>

#include "march_dos_common.cuh"
#define DOS_HW 0
namespace dos_exact {
#include "march_dos_body.cuh"
}

static int upload_cone(vrb_ctx* c, int which, const vrb_cone_sampler* s) {
  VRB_REQUIRE(s->sections && s->n_sections >= 1 && s->n_sections <= 65536, VRB_ERR_INVALID, "vrb_dos_set_cones: bad section table");
  VRB_REQUIRE(s->integration_samples[0] >= 0 && s->integration_samples[1] >= 0 && s->integration_samples[2] >= 0 &&
              s->integration_samples[0] + s->integration_samples[1] + s->integration_samples[2] <= s->n_sections,
              VRB_ERR_INVALID, "vrb_dos_set_cones: integration sample counts exceed the section table");
  std::vector<float4> h((size_t)s->n_sections);
  for (int i = 0; i < s->n_sections; ++i) {
    h[i].x = __half2float(__float2half_rn(s->sections[4 * i + 0]));
    h[i].y = __half2float(__float2half_rn(s->sections[4 * i + 1]));
    h[i].z = __half2float(__float2half_rn(s->sections[4 * i + 2]));
    h[i].w = __half2float(__float2half_rn(s->sections[4 * i + 3]));
  }
  if (c->d_cone_sections[which]) { VRB_CUDA(cudaFree(c->d_cone_sections[which])); c->d_cone_sections[which] = nullptr; }
  VRB_CUDA(cudaMalloc(&c->d_cone_sections[which], h.size() * sizeof(float4)));
  VRB_CUDA(cudaMemcpyAsync(c->d_cone_sections[which], h.data(), h.size() * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
  VRB_CUDA(cudaStreamSynchronize(c->stream));
  ConeView& v = c->cone[which];
  v.sections = c->d_cone_sections[which];
  for (int i = 0; i < 3; ++i) v.counts[i] = s->integration_samples[i];
  v.initial_step = s->initial_step; v.ray7_adj_weight = s->ray7_adj_weight; v.ui_weight = s->ui_weight;
  for (int i = 0; i < 10; ++i) for (int j = 0; j < 3; ++j) v.axes[i][j] = s->ray_axes[i][j];
  return VRB_OK;
}

extern "C" int vrb_dos_set_cones(vrb_ctx* c, const vrb_cone_sampler* occ, const vrb_cone_sampler* sdw) {
  VRB_REQUIRE(c && occ && sdw, VRB_ERR_INVALID, "vrb_dos_set_cones: NULL argument");
  VRB_CUDA(cudaSetDevice(c->device));
  int rc = upload_cone(c, 0, occ);
  if (rc != VRB_OK) return rc;
  rc = upload_cone(c, 1, sdw);
  if (rc != VRB_OK) return rc;
  c->cones_set = true;
  return VRB_OK;
}

static d3 hd3(const float* p) { d3 r; r.x = p[0]; r.y = p[1]; r.z = p[2]; return r; }

static void dos_fill_const(vrb_ctx* c, const float eye[3], const vrb_lighting* light, const vrb_dos_params* p, DosConst& C) {
  memset(&C, 0, sizeof(C));
  for (int l = 0; l < c->pyr_levels; ++l) { C.lev[l].tex = c->d_pyr[l]; C.lev[l].w = c->pyr_dims[l][0]; C.lev[l].h = c->pyr_dims[l][1]; C.lev[l].d = c->pyr_dims[l][2]; }
  C.n_levels = c->pyr_levels;
  C.VSS.x = (float)c->vw * c->scale[0]; C.VSS.y = (float)c->vh * c->scale[1]; C.VSS.z = (float)c->vd * c->scale[2];
  C.occ = c->cone[0]; C.sdw = c->cone[1];
  C.P = *p;
  C.ka = light->ka; C.kd = light->kd;
  C.eye = hd3(eye); C.light_pos = hd3(light->light_pos); C.light_fwd = hd3(light->light_forward);
  C.light_up = hd3(light->light_up); C.light_right = hd3(light->light_right);
  C.inv_VSS.x = 1.0f / C.VSS.x; C.inv_VSS.y = 1.0f / C.VSS.y; C.inv_VSS.z = 1.0f / C.VSS.z;
}

extern "C" int vrb_dos_render(vrb_ctx* c, const vrb_camera* cam, const vrb_lighting* light, const vrb_dos_params* p) {
  VRB_REQUIRE(c && cam && light && p, VRB_ERR_INVALID, "vrb_dos_render: NULL argument");
  VRB_REQUIRE(c->d_vol, VRB_ERR_STATE, "vrb_dos_render: no volume uploaded");
  VRB_REQUIRE(c->d_tf_rgbt, VRB_ERR_STATE, "vrb_dos_render: no transfer function uploaded");
  VRB_REQUIRE(c->pyr_levels > 0, VRB_ERR_STATE, "vrb_dos_render: no extinction pyramid (vrb_extcoef_build)");
  VRB_REQUIRE(c->cones_set, VRB_ERR_STATE, "vrb_dos_render: no cone samplers (vrb_dos_set_cones)");
  VRB_REQUIRE(c->d_frame, VRB_ERR_STATE, "vrb_dos_render: no frame (vrb_frame_resize)");
  VRB_REQUIRE(p->step_size > 0.0f, VRB_ERR_INVALID, "vrb_dos_render: step_size %g", p->step_size);
  VRB_CUDA(cudaSetDevice(c->device));
  DosConst C;
  dos_fill_const(c, cam->eye, light, p, C);
  { int rc = vrb_make_phong_view(c, light, &C.ph, "vrb_dos_render"); if (rc != VRB_OK) return rc; }
  if (!c->d_frame_target) VRB_CUDA(cudaMemsetAsync(c->d_frame, 0, (size_t)c->fw * c->fh * 4 * sizeof(__half), c->stream));
  if (p->count_samples) { int rc = vrb_counters_reset(c); if (rc != VRB_OK) return rc; }
  int rc = VRB_OK;
  if (c->filter_mode == VRB_FILTER_HARDWARE) {
    rc = vrb_vol_tex3d_prepare(c);
    if (rc == VRB_OK) rc = vrb_pyr_tex_prepare(c);
    if (rc != VRB_OK) return rc;
    C.pyr_tex = c->pyr_tex;
    rc = vrb_dos_launch_hw(c, cam, C, p->count_samples);
  } else {
    rc = dos_exact::dos_launch(c, cam, C, p->count_samples);
  }
  if (rc != VRB_OK) return rc;
  c->launches++;
  if (p->count_samples) return vrb_counters_fetch(c);
  return VRB_OK;
}

// PreComputeLightCache (dosrcrenderer.cpp:555-657): dispatch of rc1pdosct/lightcachecomputation.comp over the cache voxels.
extern "C" int vrb_dos_light_cache_build(vrb_ctx* c, const float eye[3], const float eye_up[3], const vrb_lighting* light,
                                         const vrb_dos_params* p, int rw, int rh, int rd) {
  VRB_REQUIRE(c && eye && eye_up && light && p, VRB_ERR_INVALID, "vrb_dos_light_cache_build: NULL argument");
  VRB_REQUIRE(c->d_vol, VRB_ERR_STATE, "vrb_dos_light_cache_build: no volume uploaded");
  VRB_REQUIRE(c->pyr_levels > 0, VRB_ERR_STATE, "vrb_dos_light_cache_build: no extinction pyramid (vrb_extcoef_build)");
  VRB_REQUIRE(c->cones_set, VRB_ERR_STATE, "vrb_dos_light_cache_build: no cone samplers (vrb_dos_set_cones)");
  VRB_REQUIRE(rw >= 1 && rh >= 1 && rd >= 1 && rw <= 1024 && rh <= 1024 && rd <= 1024, VRB_ERR_INVALID,
              "vrb_dos_light_cache_build: bad resolution %dx%dx%d", rw, rh, rd);
  VRB_CUDA(cudaSetDevice(c->device));
  int rc = vrb_light_cache_alloc(c, rw, rh, rd);
  if (rc != VRB_OK) return rc;
  DosConst C;
  dos_fill_const(c, eye, light, p, C);
  if (c->filter_mode == VRB_FILTER_HARDWARE) {
    rc = vrb_pyr_tex_prepare(c);
    if (rc != VRB_OK) return rc;
    C.pyr_tex = c->pyr_tex;
    rc = vrb_dos_light_cache_launch_hw(c, C, eye_up, rw, rh, rd);
  } else {
    rc = dos_exact::dos_light_cache_launch(c, C, eye_up, rw, rh, rd);
  }
  if (rc != VRB_OK) return rc;
  vrb_light_cache_finish(c);
  c->launches += 2;
  return VRB_OK;
}
