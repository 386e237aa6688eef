// march_vct.cu -- voxel-cone-traced shadows over the mean/stddev super-voxel pyramid (Shih et al. 2016), sm_100a.
// Replaces the dispatch of rc1pvctsg/vct_ray_bbox_marching.comp (EvaluationVoxelConeTracing :97-144, ShadeSample
// :146-189, main :191-264) made by RC1PVoxelConeTracingSGPU::Redraw; uniforms as uploaded by vctrenderer.cpp:124-237.
// Per shaded sample: up to ConeNumberOfSamples (50) cone steps, each a fractional-lod fetch of the RG16F pyramid (two
// levels x 8 half2 taps), a bilinear R16F LUT fetch, log2 and pow.  Compiled with -fmad=false (oracle operation order).
#include "vrb_internal.cuh"
#include "march_list.cuh"
#include <cstdlib>
#include <cstring>

#include "march_vct_common.cuh"
#define VCT_HW 0
namespace vct_exact {
#include "march_vct_body.cuh"
}

static void vct_fill_const(vrb_ctx* c, const vrb_lighting* light, const vrb_vct_params* p, VctConst& C) {
  memset(&C, 0, sizeof(C));
  for (int l = 0; l < c->sv_levels; ++l) {
    SvLevel& L = C.lev[l];
    L.tex = c->d_sv[l]; L.w = c->sv_dims[l][0]; L.h = c->sv_dims[l][1]; L.d = c->sv_dims[l][2];
    L.gw = c->sv_gdims[l][0]; L.gh = c->sv_gdims[l][1]; L.gd = c->sv_gdims[l][2];
    L.ox = c->sv_off[l][0]; L.oy = c->sv_off[l][1]; L.oz = c->sv_off[l][2];
  }
  C.n_levels = c->sv_levels;
  C.lut = c->d_preint; C.lut_w = c->preint_w; C.lut_h = c->preint_h;
  C.VSS.x = (float)c->vw * c->scale[0]; C.VSS.y = (float)c->vh * c->scale[1]; C.VSS.z = (float)c->vd * c->scale[2];
  C.light_pos.x = light->light_pos[0]; C.light_pos.y = light->light_pos[1]; C.light_pos.z = light->light_pos[2];
  C.P = *p; C.ka = light->ka; C.kd = light->kd;
  C.corr_fact = (float)p->apply_opacity_correction * p->opacity_correction_factor;     // const float corr_fact (:96)
  C.inv_VSS.x = 1.0f / C.VSS.x; C.inv_VSS.y = 1.0f / C.VSS.y; C.inv_VSS.z = 1.0f / C.VSS.z;
  C.sv_scale = C.inv_VSS;                                  // sv_bias = 0: fmaf(w, s, 0) == w * s
}

extern "C" int vrb_vct_render(vrb_ctx* c, const vrb_camera* cam, const vrb_lighting* light, const vrb_vct_params* p) {
  VRB_REQUIRE(c && cam && light && p, VRB_ERR_INVALID, "vrb_vct_render: NULL argument");
  VRB_REQUIRE(c->d_vol, VRB_ERR_STATE, "vrb_vct_render: no volume uploaded");
  VRB_REQUIRE(c->d_tf_rgbt, VRB_ERR_STATE, "vrb_vct_render: no transfer function uploaded");
  VRB_REQUIRE(c->sv_levels > 0 && c->d_preint, VRB_ERR_STATE, "vrb_vct_render: no super-voxel pyramid / LUT (vrb_vct_build)");
  VRB_REQUIRE(c->d_frame, VRB_ERR_STATE, "vrb_vct_render: no frame (vrb_frame_resize)");
  VRB_REQUIRE(p->step_size > 0.0f, VRB_ERR_INVALID, "vrb_vct_render: step_size %g", p->step_size);
  VRB_REQUIRE(p->cone_number_of_samples >= 0 && p->cone_number_of_samples <= 100000, VRB_ERR_INVALID, "vrb_vct_render: cone_number_of_samples");
  VRB_CUDA(cudaSetDevice(c->device));
  VctConst C;
  vct_fill_const(c, light, p, C);
  { int rc = vrb_make_phong_view(c, light, &C.ph, "vrb_vct_render"); if (rc != VRB_OK) return rc; }
  if (!c->d_frame_target) VRB_CUDA(cudaMemsetAsync(c->d_frame, 0, (size_t)c->fw * c->fh * 4 * sizeof(__half), c->stream));
  if (p->count_samples) { int rc = vrb_counters_reset(c); if (rc != VRB_OK) return rc; }
  int rc = VRB_OK;
  if (c->filter_mode == VRB_FILTER_HARDWARE) {
    rc = vrb_vol_tex3d_prepare(c);
    if (rc == VRB_OK) rc = vrb_sv_tex_prepare(c);
    if (rc != VRB_OK) return rc;
    C.sv_tex = c->sv_tex; C.lut_tex = c->preint_tex;
    rc = vrb_vct_launch_hw(c, cam, C, p->count_samples);
  } else {
    // default: deferred frame (march_list.cu -> k_vct_shade -> composite); VRB_VCT_KERNEL=ray: round 1's one-thread-per-ray k_vct
    const char* kern = getenv("VRB_VCT_KERNEL");
    if (kern && !strcmp(kern, "ray")) rc = vct_exact::vct_launch(c, cam, C, p->count_samples);
    else rc = vct_exact::vct_deferred_launch(c, cam, C, p->count_samples);
  }
  if (rc != VRB_OK) return rc;
  c->launches++;
  if (p->count_samples) return vrb_counters_fetch(c);
  return VRB_OK;
}

// Sort-last brick of the VCT renderer.  The context holds a WINDOW of the volume (vrb_volume_upload of owned + ghost voxels)
// and the window's pyramid (vrb_sv_build_brick) with the LUT built for the deviation range of the whole volume
// (vrb_preint_build).  The ghost layers must cover the cone reach (cpp_volume_rendering_b200/dist.py: vct_halo).
extern "C" int vrb_vct_render_brick(vrb_ctx* c, const vrb_camera* cam, const vrb_lighting* light, const vrb_vct_params* p, const vrb_brick* b,
                                    int mode, const void* const* front_alphas, int n_front) {
  VRB_REQUIRE(c && cam && light && p && b, VRB_ERR_INVALID, "vrb_vct_render_brick: NULL argument");
  VRB_REQUIRE(mode >= VRB_BRICK_SEGMENT && mode <= VRB_BRICK_EXACT, VRB_ERR_INVALID, "vrb_vct_render_brick: mode %d", mode);
  VRB_REQUIRE(n_front >= 0 && n_front <= VRB_VCT_MAX_FRONT && (n_front == 0 || front_alphas), VRB_ERR_INVALID, "vrb_vct_render_brick: bad front list");
  VRB_REQUIRE(light->apply_phong != 1, VRB_ERR_UNSUPPORTED, "vrb_vct_render_brick: gradient Blinn-Phong is not available on bricks");
  VRB_REQUIRE(mode == VRB_BRICK_EXACT || n_front == 0, VRB_ERR_INVALID, "vrb_vct_render_brick: a front list only makes sense in VRB_BRICK_EXACT mode");
  VRB_REQUIRE(c->d_vol && c->d_tf_rgbt && c->d_frame, VRB_ERR_STATE, "vrb_vct_render_brick: volume / transfer function / frame missing");
  VRB_REQUIRE(mode == VRB_BRICK_ALPHA || (c->sv_levels > 0 && c->d_preint), VRB_ERR_STATE, "vrb_vct_render_brick: no super-voxel pyramid / LUT (vrb_sv_build_brick, vrb_preint_build)");
  VRB_REQUIRE(p->step_size > 0.0f, VRB_ERR_INVALID, "vrb_vct_render_brick: step_size %g", p->step_size);
  VRB_REQUIRE(p->cone_number_of_samples >= 0 && p->cone_number_of_samples <= 100000, VRB_ERR_INVALID, "vrb_vct_render_brick: cone_number_of_samples");
  { int rc = vrb_brick_check(c, b, "vrb_vct_render_brick"); if (rc != VRB_OK) return rc; }
  if (mode != VRB_BRICK_ALPHA)
    for (int a = 0; a < 3; ++a)
      VRB_REQUIRE(c->sv_gdims[0][a] == b->global_dims[a] && c->sv_off[0][a] == b->origin[a] - b->ghost_lo[a], VRB_ERR_STATE,
                  "vrb_vct_render_brick: the pyramid was not built for this brick (axis %d)", a);
  VRB_CUDA(cudaSetDevice(c->device));
  { int rc = vrb_partial_alloc(c); if (rc != VRB_OK) return rc; }
  VctConst C;
  vct_fill_const(c, light, p, C);
  VctBrick B;
  C.VSS.x = (float)b->global_dims[0] * c->scale[0]; C.VSS.y = (float)b->global_dims[1] * c->scale[1]; C.VSS.z = (float)b->global_dims[2] * c->scale[2];
  C.inv_VSS.x = 1.0f / C.VSS.x; C.inv_VSS.y = 1.0f / C.VSS.y; C.inv_VSS.z = 1.0f / C.VSS.z;
  B.kx = (float)b->global_dims[0] / C.VSS.x; B.ky = (float)b->global_dims[1] / C.VSS.y; B.kz = (float)b->global_dims[2] / C.VSS.z;
  B.offx = (float)(b->origin[0] - b->ghost_lo[0]); B.offy = (float)(b->origin[1] - b->ghost_lo[1]); B.offz = (float)(b->origin[2] - b->ghost_lo[2]);
  for (int a = 0; a < 3; ++a) { B.lo[a] = b->origin[a]; B.hi[a] = b->origin[a] + b->owned[a]; }
  B.nx = b->global_dims[0]; B.ny = b->global_dims[1]; B.nz = b->global_dims[2];
  // hardware mode: one affine map serves every level because the window is aligned to the coarsest level built
  C.sv_scale.x = B.kx / (float)c->vw; C.sv_scale.y = B.ky / (float)c->vh; C.sv_scale.z = B.kz / (float)c->vd;
  C.sv_bias.x = -B.offx / (float)c->vw; C.sv_bias.y = -B.offy / (float)c->vh; C.sv_bias.z = -B.offz / (float)c->vd;
  VctFront front; front.n = n_front;
  for (int i = 0; i < n_front; ++i) { VRB_REQUIRE(front_alphas[i], VRB_ERR_INVALID, "vrb_vct_render_brick: front alpha %d is NULL", i); front.p[i] = (const float*)front_alphas[i]; }
  if (p->count_samples) { int rc = vrb_counters_reset(c); if (rc != VRB_OK) return rc; }
  int rc = VRB_OK;
  if (c->filter_mode == VRB_FILTER_HARDWARE) {
    rc = vrb_vol_tex3d_prepare(c);
    if (rc == VRB_OK && mode != VRB_BRICK_ALPHA) rc = vrb_sv_tex_prepare(c);
    if (rc != VRB_OK) return rc;
    C.sv_tex = c->sv_tex; C.lut_tex = c->preint_tex;
    rc = vrb_vct_brick_launch_hw(c, cam, C, B, front, mode, p->count_samples);
  } else {
    rc = vct_exact::vct_brick_launch(c, cam, C, B, front, mode, p->count_samples);
  }
  if (rc != VRB_OK) return rc;
  c->launches++;
  if (p->count_samples) return vrb_counters_fetch(c);
  return VRB_OK;
}

// PreComputeLightCache (vctrenderer.cpp:393-515): dispatch of rc1pvctsg/lightcachecomputation.comp over the cache voxels.
extern "C" int vrb_vct_light_cache_build(vrb_ctx* c, const vrb_lighting* light, const vrb_vct_params* p, int rw, int rh, int rd) {
  VRB_REQUIRE(c && light && p, VRB_ERR_INVALID, "vrb_vct_light_cache_build: NULL argument");
  VRB_REQUIRE(c->d_vol, VRB_ERR_STATE, "vrb_vct_light_cache_build: no volume uploaded");
  VRB_REQUIRE(c->sv_levels > 0 && c->d_preint, VRB_ERR_STATE, "vrb_vct_light_cache_build: no super-voxel pyramid / LUT (vrb_vct_build)");
  VRB_REQUIRE(p->cone_number_of_samples >= 0 && p->cone_number_of_samples <= 100000, VRB_ERR_INVALID, "vrb_vct_light_cache_build: cone_number_of_samples");
  VRB_REQUIRE(rw >= 1 && rh >= 1 && rd >= 1 && rw <= 1024 && rh <= 1024 && rd <= 1024, VRB_ERR_INVALID,
              "vrb_vct_light_cache_build: bad resolution %dx%dx%d", rw, rh, rd);
  VRB_CUDA(cudaSetDevice(c->device));
  int rc = vrb_light_cache_alloc(c, rw, rh, rd);
  if (rc != VRB_OK) return rc;
  VctConst C;
  vct_fill_const(c, light, p, C);
  if (c->filter_mode == VRB_FILTER_HARDWARE) {
    rc = vrb_sv_tex_prepare(c);
    if (rc != VRB_OK) return rc;
    C.sv_tex = c->sv_tex; C.lut_tex = c->preint_tex;
    rc = vrb_vct_light_cache_launch_hw(c, C, rw, rh, rd);
  } else {
    rc = vct_exact::vct_light_cache_launch(c, C, rw, rh, rd);
  }
  if (rc != VRB_OK) return rc;
  vrb_light_cache_finish(c);
  c->launches += 2;
  return VRB_OK;
}
