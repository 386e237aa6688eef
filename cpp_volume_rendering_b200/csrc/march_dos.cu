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
#include "march_list.cuh"
#include "march_dos_compact.cuh"
#include "march_dos_deferred.cuh"
#include <cstdlib>

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
  c->h_cone_sections[which] = h;
  c->cones_gen++;
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

// ---- k_dos_compact: host side -------------------------------------------------------------------------------------------
// The pyramid level textureLod would pick for a section (dos_texture_lod): lod <= 0 -> 0, lod >= last -> last, an integer in
// between -> that level.  A fractional lod blends two levels: not handled by k_dos_compact (returns -1).
static int dos_section_level(float lod, int n_levels) {
  const int maxl = n_levels - 1;
  if (!(lod > 0.0f)) return 0;
  if (lod >= (float)maxl) return maxl;
  const float fl = floorf(lod);
  if (lod - fl != 0.0f) return -1;
  return (int)fl;
}

static bool is_pow2f(float v) { int e; return v > 0.0f && std::isfinite(v) && frexpf(v, &e) == 0.5f; }

// Builds / refreshes d_dos_packed and fills F.  *ok = false when this scene needs the general kernel (k_dos).
static int dos_compact_prepare(vrb_ctx* c, const DosConst& C, DosFast& F, bool* ok, bool* pow2) {
  *ok = false;
  memset(&F, 0, sizeof(F));
  const char* env = getenv("VRB_DOS_KERNEL");
  if (env && !strcmp(env, "ray")) return VRB_OK;
  const int n0 = (int)c->h_cone_sections[0].size(), n1 = (int)c->h_cone_sections[1].size();
  if (n0 + n1 > 1024 || n0 > 65535 || n1 > 65535) return VRB_OK;
  if ((size_t)(c->tf_n + 2 <= 1026 ? c->tf_n + 2 : 0) * sizeof(float4) + (size_t)(n0 + n1) * sizeof(float4) > 40 * 1024) return VRB_OK;
  for (int l = 0; l < c->pyr_levels; ++l)
    if ((long long)(c->pyr_dims[l][0] + 2) * (c->pyr_dims[l][1] + 2) * (c->pyr_dims[l][2] + 2) >= (1LL << 31)) return VRB_OK;
  if (c->pyr_levels > 255) return VRB_OK;
  const unsigned long long sig = c->cones_gen * 64ull + (unsigned long long)c->pyr_levels;
  if (!c->d_dos_packed || c->dos_packed_sig != sig) {
    std::vector<float4> pk((size_t)(n0 + n1));
    for (int which = 0; which < 2; ++which) {
      const std::vector<float4>& h = c->h_cone_sections[which];
      const int n = (int)h.size(), base = which ? n0 : 0;
      std::vector<int> lvl(n), mip(n);
      for (int i = 0; i < n; ++i) {
        lvl[i] = dos_section_level(h[i].y, c->pyr_levels);
        if (lvl[i] < 0 || h[i].y != floorf(h[i].y) || h[i].y < -60.0f || h[i].y > 60.0f) { c->dos_packed_sig = 0; return VRB_OK; }
        mip[i] = (int)h[i].y;
      }
      for (int i = 0; i < n; ++i) {
        int end = i + 1;
        while (end < n && lvl[end] == lvl[i] && mip[end] == mip[i]) ++end;
        const int bits = lvl[i] | ((mip[i] + 64) << 8) | (end << 16);
        float w; memcpy(&w, &bits, 4);
        pk[base + i] = make_float4(h[i].x, h[i].z, h[i].w, w);
      }
    }
    if (c->d_dos_packed) { VRB_CUDA(cudaFree(c->d_dos_packed)); c->d_dos_packed = nullptr; }
    VRB_CUDA(cudaMalloc(&c->d_dos_packed, std::max<size_t>(1, pk.size()) * sizeof(float4)));
    VRB_CUDA(cudaMemcpyAsync(c->d_dos_packed, pk.data(), pk.size() * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    VRB_CUDA(cudaStreamSynchronize(c->stream));
    c->dos_packed_sig = sig; c->dos_packed_n[0] = n0; c->dos_packed_n[1] = n1;
  }
  int rc = vrb_pyr_quads_prepare(c);
  if (rc != VRB_OK) return rc;
  *pow2 = is_pow2f(C.VSS.x) && is_pow2f(C.VSS.y) && is_pow2f(C.VSS.z);
  for (int l = 0; l < c->pyr_levels; ++l)
    for (int k = 0; k < 3; ++k) *pow2 = *pow2 && (c->pyr_dims[l][k] & (c->pyr_dims[l][k] - 1)) == 0;
  { const char* e = getenv("VRB_DOS_POW2"); if (e && !strcmp(e, "0")) *pow2 = false; }
  for (int l = 0; l < c->pyr_levels; ++l) {
    DosLevelQ& L = F.lev[l];
    L.q = c->d_pyr_quad[l];
    L.w = c->pyr_dims[l][0]; L.h = c->pyr_dims[l][1]; L.d = c->pyr_dims[l][2];
    L.pw = L.w + 2; L.pslice = (L.w + 2) * (L.h + 2);
    L.fw = (float)L.w; L.fh = (float)L.h; L.fd = (float)L.d;
    if (*pow2) { L.fw *= 1.0f / C.VSS.x; L.fh *= 1.0f / C.VSS.y; L.fd *= 1.0f / C.VSS.z; }   // exact: powers of two
  }
  F.packed = c->d_dos_packed; F.n_occ = n0; F.n_sdw = n1;
  F.inv_vss.x = 1.0f / C.VSS.x; F.inv_vss.y = 1.0f / C.VSS.y; F.inv_vss.z = 1.0f / C.VSS.z;
  *ok = true;
  return VRB_OK;
}

template <bool PHONG, bool HAS7>
static void dos_compact_launch2(vrb_ctx* c, dim3 grid, size_t smem, const vrb_camera* cam, const PartView& part, const DosConst& C, const DosFast& F, bool pow2) {
  if (pow2) dos_compact::k_dos_compact<PHONG, HAS7, true><<<grid, dim3(8, 8), smem, c->stream>>>(c->vol_view(), c->d_tf_rgbt, c->tf_n, c->frame_view(), make_cam_view(cam), part, C, F, c->d_counter);
  else      dos_compact::k_dos_compact<PHONG, HAS7, false><<<grid, dim3(8, 8), smem, c->stream>>>(c->vol_view(), c->d_tf_rgbt, c->tf_n, c->frame_view(), make_cam_view(cam), part, C, F, c->d_counter);
}

static int dos_compact_launch(vrb_ctx* c, const vrb_camera* cam, const DosConst& C, DosFast& F, int count_samples, bool pow2) {
  PartView part;
  dim3 grid = vrb_make_grid(c, 8, 8, &part);
  F.count = count_samples;
  const size_t smem = ((size_t)(c->tf_n + 2 <= 1026 ? c->tf_n + 2 : 0) + (size_t)(F.n_occ + F.n_sdw)) * sizeof(float4);
  const bool has7 = C.occ.counts[2] > 0 || C.sdw.counts[2] > 0;
  if (C.ph.grad) { if (has7) dos_compact_launch2<true, true>(c, grid, smem, cam, part, C, F, pow2); else dos_compact_launch2<true, false>(c, grid, smem, cam, part, C, F, pow2); }
  else           { if (has7) dos_compact_launch2<false, true>(c, grid, smem, cam, part, C, F, pow2); else dos_compact_launch2<false, false>(c, grid, smem, cam, part, C, F, pow2); }
  VRB_CUDA(cudaGetLastError());
  return VRB_OK;
}

template <bool PHONG, bool HAS7>
static void dos_shade_launch2(vrb_ctx* c, unsigned n, size_t smem, const vrb_camera* cam, const DosConst& C, const DosFast& F, const ShadeListView& L, bool pow2) {
  const unsigned blocks = (n + 127u) / 128u;
  if (pow2) dos_deferred::k_dos_shade<PHONG, HAS7, true><<<blocks, 128, smem, c->stream>>>(c->vol_view(), c->frame_view(), make_cam_view(cam), C, F, L, n, c->d_counter);
  else      dos_deferred::k_dos_shade<PHONG, HAS7, false><<<blocks, 128, smem, c->stream>>>(c->vol_view(), c->frame_view(), make_cam_view(cam), C, F, L, n, c->d_counter);
}

// march (march_list.cu) -> shade -> composite (march_list.cu)
static int dos_deferred_launch(vrb_ctx* c, const vrb_camera* cam, const DosConst& C, DosFast& F, int count_samples, bool pow2) {
  F.count = count_samples;
  ListFrame f;
  int rc = vrb_list_march(c, cam, C.P.step_size, 0, count_samples, &f);
  if (rc != VRB_OK) return rc;
  const unsigned n = f.n_entries;
  const ShadeListView& L = f.L;
  if (n) {
    VrbKernelTimer timer(c, "k_dos_shade");
    const size_t smem = (size_t)(F.n_occ + F.n_sdw) * sizeof(float4);
    const bool has7 = C.occ.counts[2] > 0 || C.sdw.counts[2] > 0;
    if (C.ph.grad) { if (has7) dos_shade_launch2<true, true>(c, n, smem, cam, C, F, L, pow2); else dos_shade_launch2<true, false>(c, n, smem, cam, C, F, L, pow2); }
    else           { if (has7) dos_shade_launch2<false, true>(c, n, smem, cam, C, F, L, pow2); else dos_shade_launch2<false, false>(c, n, smem, cam, C, F, L, pow2); }
    VRB_CUDA(cudaGetLastError());
    c->launches++;
  }
  return vrb_list_composite(c, cam, 0, f);
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
    DosFast F; bool compact = false, pow2 = false;
    rc = dos_compact_prepare(c, C, F, &compact, &pow2);
    if (rc != VRB_OK) return rc;
    const char* kern = getenv("VRB_DOS_KERNEL");
    if (!compact) rc = dos_exact::dos_launch(c, cam, C, p->count_samples);
    else if (kern && !strcmp(kern, "compact")) rc = dos_compact_launch(c, cam, C, F, p->count_samples, pow2);
    else rc = dos_deferred_launch(c, cam, C, F, p->count_samples, pow2);
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
