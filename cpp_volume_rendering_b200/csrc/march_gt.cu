// march_gt.cu -- cone "ground truth" renderer (many occlusion / shadow rays per primary sample), sm_100a.
// Replaces the progressive dispatch loop of RC1PConeLightGroundTruthSteps::RedrawFrameTexture (crtgtrenderer.cpp:272-325,
// one primary sample per glDispatchCompute + two full-image glGetTexImage read-backs per iteration) and the shader
// rc1pcrtgt/gt_ray_marching.comp by ONE kernel that marches every ray to convergence.  The reference's running colour
// and ray parameter live in rgba16f / rg16f images between dispatches; those fp16 round trips are part of its result
// and are applied here after every primary sample.
//
// This is the compute-bound path: per shaded sample N_occ + N_sdw secondary rays of up to dist/step trapezoid steps,
// each one trilinear volume tap + one TF extinction lookup + one exp.  One thread per pixel, 8x4 pixels per warp: the
// 32 lanes trace the SAME secondary-ray index at the same time from neighbouring origins, so their taps share lines.
#include "vrb_internal.cuh"
#include <vector>

#include "march_gt_common.cuh"
// secondary rays marched together by one thread (measured on B200 at cfg4: 1396 / 1491 / 1674 ms for 1 / 2 / 4 rays in flight: the software-filtered kernel is issue-bound, more rays only add registers)
#ifndef GT_ILP
#define GT_ILP 1
#endif
#define GT_HW 0
namespace gt_exact {
#include "march_gt_body.cuh"
}
#include "march_list.cuh"
#include "march_gt_deferred.cuh"
#include <cstdlib>
#include <cstring>

// march (march_list.cu, fp16 state round trips on) -> one lane per secondary ray -> composite
static int gt_deferred_launch(vrb_ctx* c, const vrb_camera* cam, const GtConst& C, int count_samples) {
  ListFrame f;
  int rc = vrb_list_march(c, cam, C.P.step_size, VRB_LIST_GT, count_samples, &f);
  if (rc != VRB_OK) return rc;
  if (f.n_entries) {
    const int nt = (C.P.apply_occlusion == 1 ? C.P.occ_num_rays : 0) + (C.P.apply_shadow == 1 ? C.P.sdw_num_rays : 0);
    const CamView cv = make_cam_view(cam);
    const VolView vol = c->vol_view();
    // "task": one secondary ray per lane with refill (needs the per-warp result buffers to hold a group); "entry": one list
    // entry per lane, 32 parallel rays per warp (VRB_GT_SHADE=entry, VRB_GT_ILP = rays in flight per lane).  Measured at config 4 on
    // B200: task 531 ms, entry 635 / 838 / 1175 ms with 1 / 2 / 4 rays in flight (round 1, one thread per primary ray: 1370 ms)
    const char* mode = getenv("VRB_GT_SHADE");
    const bool by_task = nt <= GT_TASK_CAP && nt > 0 && !(mode && !strcmp(mode, "entry"));
    VrbKernelTimer timer(c, "k_gt_shade");
    if (by_task) {
      const int group = std::max(1, std::min(GT_MAX_GROUP, GT_TASK_CAP / nt));
      const unsigned n_groups = (f.n_entries + (unsigned)group - 1u) / (unsigned)group;
      const size_t smem = (size_t)((c->tf_n + 2 + 3) & ~3) * sizeof(float) + GT_WARPS * sizeof(gt_deferred::GroupSmem);
      const unsigned blocks = std::max(1u, std::min((n_groups + GT_WARPS - 1) / GT_WARPS, 148u * 12u));
#define VRB_GT_SHADE(PH, Q) gt_deferred::k_gt_shade<PH, Q><<<blocks, GT_WARPS * 32, smem, c->stream>>>(vol, c->d_vol_quad, c->d_tf_rgbt, c->tf_n, \
          c->frame_view(), cv, C, f.L, f.n_entries, group, count_samples, c->d_counter)
      if (C.ph.grad) { if (c->d_vol_quad) VRB_GT_SHADE(true, true); else VRB_GT_SHADE(true, false); }
      else           { if (c->d_vol_quad) VRB_GT_SHADE(false, true); else VRB_GT_SHADE(false, false); }
#undef VRB_GT_SHADE
    } else {
      const size_t smem = (size_t)(c->tf_n + 2) * sizeof(float);
      const unsigned blocks = (f.n_entries + 127u) / 128u;
      const char* ie = getenv("VRB_GT_ILP");
      const int ilp = ie ? atoi(ie) : 2;
#define VRB_GT_SHADE_E(PH, Q, I) gt_deferred::k_gt_shade_e<PH, Q, I><<<blocks, 128, smem, c->stream>>>(vol, c->d_vol_quad, c->d_tf_rgbt, c->tf_n, \
          c->frame_view(), cv, C, f.L, f.n_entries, count_samples, c->d_counter)
#define VRB_GT_SHADE_E2(PH, Q) do { if (ilp >= 4) VRB_GT_SHADE_E(PH, Q, 4); else if (ilp >= 2) VRB_GT_SHADE_E(PH, Q, 2); else VRB_GT_SHADE_E(PH, Q, 1); } while (0)
      if (C.ph.grad) { if (c->d_vol_quad) VRB_GT_SHADE_E2(true, true); else VRB_GT_SHADE_E2(true, false); }
      else           { if (c->d_vol_quad) VRB_GT_SHADE_E2(false, true); else VRB_GT_SHADE_E2(false, false); }
#undef VRB_GT_SHADE_E2
#undef VRB_GT_SHADE_E
    }
    VRB_CUDA(cudaGetLastError());
    c->launches++;
  }
  return vrb_list_composite(c, cam, VRB_LIST_GT, f);
}

static int upload_rays(vrb_ctx* c, int which, const float* rays, int n) {
  std::vector<float> h((size_t)std::max(n, 1) * 3, 0.0f);
  for (int i = 0; i < n * 3; ++i) h[i] = __half2float(__float2half_rn(rays[i]));     // GL_RGB16F
  if (c->d_gt_rays[which]) { VRB_CUDA(cudaFree(c->d_gt_rays[which])); c->d_gt_rays[which] = nullptr; }
  VRB_CUDA(cudaMalloc(&c->d_gt_rays[which], h.size() * sizeof(float)));
  VRB_CUDA(cudaMemcpyAsync(c->d_gt_rays[which], h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  VRB_CUDA(cudaStreamSynchronize(c->stream));
  c->gt_nrays[which] = n;
  return VRB_OK;
}

extern "C" int vrb_gt_set_rays(vrb_ctx* c, const float* occ_rays, int n_occ, const float* sdw_rays, int n_sdw) {
  VRB_REQUIRE(c, VRB_ERR_INVALID, "vrb_gt_set_rays: ctx is NULL");
  VRB_REQUIRE(n_occ >= 0 && n_sdw >= 0 && n_occ <= (1 << 20) && n_sdw <= (1 << 20), VRB_ERR_INVALID, "vrb_gt_set_rays: bad ray counts");
  VRB_REQUIRE((n_occ == 0 || occ_rays) && (n_sdw == 0 || sdw_rays), VRB_ERR_INVALID, "vrb_gt_set_rays: NULL table");
  VRB_CUDA(cudaSetDevice(c->device));
  int rc = upload_rays(c, 0, occ_rays, n_occ);
  if (rc != VRB_OK) return rc;
  return upload_rays(c, 1, sdw_rays, n_sdw);
}

static g3 hg3(const float* p) { g3 r; r.x = p[0]; r.y = p[1]; r.z = p[2]; return r; }

extern "C" int vrb_gt_render(vrb_ctx* c, const vrb_camera* cam, const vrb_lighting* light, const vrb_gt_params* p) {
  VRB_REQUIRE(c && cam && light && p, VRB_ERR_INVALID, "vrb_gt_render: NULL argument");
  VRB_REQUIRE(c->d_vol, VRB_ERR_STATE, "vrb_gt_render: no volume uploaded");
  VRB_REQUIRE(c->d_tf_rgbt, VRB_ERR_STATE, "vrb_gt_render: no transfer function uploaded");
  VRB_REQUIRE(c->d_frame, VRB_ERR_STATE, "vrb_gt_render: no frame (vrb_frame_resize)");
  VRB_REQUIRE(c->tf_n + 2 <= 2048, VRB_ERR_UNSUPPORTED, "vrb_gt_render: transfer functions above 2046 texels are not supported by this kernel");
  VRB_REQUIRE(p->step_size > 0.0f && p->light_ray_step_size > 0.0f, VRB_ERR_INVALID, "vrb_gt_render: step sizes must be positive");
  VRB_REQUIRE(!p->apply_occlusion || (c->d_gt_rays[0] && c->gt_nrays[0] == p->occ_num_rays && p->occ_num_rays > 0), VRB_ERR_STATE,
              "vrb_gt_render: occlusion ray table missing or of different size (vrb_gt_set_rays)");
  VRB_REQUIRE(!p->apply_shadow || (c->d_gt_rays[1] && c->gt_nrays[1] == p->sdw_num_rays && p->sdw_num_rays > 0), VRB_ERR_STATE,
              "vrb_gt_render: shadow ray table missing or of different size (vrb_gt_set_rays)");
  VRB_CUDA(cudaSetDevice(c->device));
  GtConst C;
  C.G.x = (float)c->vw * c->scale[0]; C.G.y = (float)c->vh * c->scale[1]; C.G.z = (float)c->vd * c->scale[2];
  C.P = *p; C.ka = light->ka; C.kd = light->kd;
  C.light_pos = hg3(light->light_pos); C.light_fwd = hg3(light->light_forward); C.light_up = hg3(light->light_up); C.light_right = hg3(light->light_right);
  C.occ_rays = c->d_gt_rays[0]; C.sdw_rays = c->d_gt_rays[1];
  { int rc = vrb_make_phong_view(c, light, &C.ph, "vrb_gt_render"); if (rc != VRB_OK) return rc; }
  if (!c->d_frame_target) VRB_CUDA(cudaMemsetAsync(c->d_frame, 0, (size_t)c->fw * c->fh * 4 * sizeof(__half), c->stream));
  if (p->count_samples) { int rc = vrb_counters_reset(c); if (rc != VRB_OK) return rc; }
  int rc = VRB_OK;
  if (c->filter_mode == VRB_FILTER_HARDWARE) {
    rc = vrb_vol_tex3d_prepare(c);
    if (rc != VRB_OK) return rc;
    rc = vrb_gt_launch_hw(c, cam, C, p->count_samples);
  } else {
    const int nt = (p->apply_occlusion == 1 ? p->occ_num_rays : 0) + (p->apply_shadow == 1 ? p->sdw_num_rays : 0);
    const char* kern = getenv("VRB_GT_KERNEL");
    // VRB_GT_KERNEL=ray: one thread per primary ray (round 1's k_gt), kept for A/B runs
    (void)nt;
    if (kern && !strcmp(kern, "ray")) rc = gt_exact::gt_launch(c, cam, C, p->count_samples);
    else rc = gt_deferred_launch(c, cam, C, p->count_samples);
  }
  if (rc != VRB_OK) return rc;
  c->launches++;
  if (p->count_samples) return vrb_counters_fetch(c);
  return VRB_OK;
}

// RedrawCube (crtgtrenderer.cpp:327-338) = rc1pcrtgt/vol_intersection.comp:64-110: the frame the reference shows while
// "Show Generated Frame Texture" is off (its default, and after every camera or parameter change): the point where each ray
// enters the volume's bounding box, coloured by the face it lies on (ties go to z, then y).  Same ray set-up as every marcher.
__global__ void __launch_bounds__(64) k_gt_cube(FrameView fr, CamView cam, PartView part, float gx, float gy, float gz) {
  int px, py;
  vrb_cta_origin(part, fr.w, 8, 8, px, py);
  px += threadIdx.x; py += threadIdx.y;
  if (px >= fr.w || py >= fr.h || !vrb_owns_pixel(part, px, py, fr.w)) return;
  Ray r = vrb_make_ray(cam, px, py, fr.w, fr.h, gx, gy, gz);
  if (!r.hit) { if (fr.zero_miss) vrb_store_pixel(fr, px, py, 0.f, 0.f, 0.f, 0.f); return; }
  const float wx = __fadd_rn(r.ox, __fmul_rn(r.dx, r.tnear)), wy = __fadd_rn(r.oy, __fmul_rn(r.dy, r.tnear)), wz = __fadd_rn(r.oz, __fmul_rn(r.dz, r.tnear));
  const float cx = __fdiv_rn(fabsf(wx), __fmul_rn(gx, 0.5f)), cy = __fdiv_rn(fabsf(wy), __fmul_rn(gy, 0.5f)), cz = __fdiv_rn(fabsf(wz), __fmul_rn(gz, 0.5f));
  float R = 0.f, G = 0.f, B = 0.f, A = 0.f;
  if (cz >= cx && cz >= cy) { B = 1.f; A = 1.f; }
  else if (cy >= cx && cy >= cz) { G = 1.f; A = 1.f; }
  else if (cx >= cy && cx >= cz) { R = 1.f; A = 1.f; }
  vrb_store_pixel(fr, px, py, R, G, B, A);
}

extern "C" int vrb_gt_cube_render(vrb_ctx* c, const vrb_camera* cam) {
  VRB_REQUIRE(c && cam, VRB_ERR_INVALID, "vrb_gt_cube_render: NULL argument");
  VRB_REQUIRE(c->d_vol, VRB_ERR_STATE, "vrb_gt_cube_render: no volume uploaded");
  VRB_REQUIRE(c->d_frame, VRB_ERR_STATE, "vrb_gt_cube_render: no frame (vrb_frame_resize)");
  VRB_CUDA(cudaSetDevice(c->device));
  if (!c->d_frame_target) VRB_CUDA(cudaMemsetAsync(c->d_frame, 0, (size_t)c->fw * c->fh * 4 * sizeof(__half), c->stream));
  PartView part;
  dim3 block(8, 8), grid = vrb_make_grid(c, 8, 8, &part);
  k_gt_cube<<<grid, block, 0, c->stream>>>(c->frame_view(), make_cam_view(cam), part, (float)c->vw * c->scale[0], (float)c->vh * c->scale[1], (float)c->vd * c->scale[2]);
  VRB_CUDA(cudaGetLastError());
  c->launches++;
  return VRB_OK;
}
