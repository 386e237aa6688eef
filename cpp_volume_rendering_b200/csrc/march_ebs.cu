// march_ebs.cu -- extinction-based shading marcher (Schlegel et al. 2011) for sm_100a.
// Replaces the dispatch of rc1pextbsd/ebs_ray_bbox_marching.comp made by RC1PExtinctionBasedShading::Redraw
// (ebsrenderer.cpp:249-260); uniforms as uploaded by ebsrenderer.cpp:125-247.
//
// Numerics: the SAT queries are fp32 differences of prefix sums as large as 1e7-1e8, so the reference RESULT contains
// the cancellation noise of its exact sequence of fp32 operations.  This file is therefore compiled with -fmad=false
// (see Makefile) and written in the shader's operation order; the only fused operations are the explicit fmaf() of
// the GL_LINEAR blend (vrb_lerp), which the oracle spells the same way.  cos/sin of the cone angle depend on uniforms
// only and are evaluated once on the host.
#include "vrb_internal.cuh"
#include "march_list.cuh"
#include <cmath>
#include <cstdlib>
#include <cstring>

struct f3 { float x, y, z; };
__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ f3 operator*(float s, f3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ f3 clamp3(f3 a, f3 lo, f3 hi) {
  return mk3(fminf(fmaxf(a.x, lo.x), hi.x), fminf(fmaxf(a.y, lo.y), hi.y), fminf(fmaxf(a.z, lo.z), hi.z));
}
__device__ __forceinline__ f3 norm3(f3 a) {
  float d = a.x * a.x + a.y * a.y + a.z * a.z;
  float r = 1.0f / sqrtf(d);
  return mk3(a.x * r, a.y * r, a.z * r);
}

struct EbsConst {
  f3 VS, VSS, MinSAT, MaxSAT, MinVol, MaxVol, inv_vol_scaled, InvG;
  float p_cs, p_sn, n_cs, n_sn;
  const float* sat; const void* sat_packed; int sw, sh, sd; long long sslice;
  cudaTextureObject_t sat_tex; int atlas_tiles_x, atlas_tile_w, atlas_tile_h;
  PhongView ph;               // ApplyPhongShading (ebs_ray_bbox_marching.comp:524-544); handled by k_ebs (one thread per ray)
  vrb_ebs_params P;
  float ka, kd;
  f3 light_pos, light_fwd;
};

// The body is instantiated once per SAT layout; the layout is chosen at vrb_sat_build time (vrb_ctx::sat_pack).
#define EBS_MIN_BLOCKS 8       // <= 128 registers: 16 warps / SM
#define EBS_PACK 1
namespace ebs_pack1 {
#include "march_ebs_body.cuh"
}
#undef EBS_PACK
#define EBS_PACK 2
namespace ebs_pack2 {
#include "march_ebs_body.cuh"
}
#undef EBS_PACK
#define EBS_PACK 4
namespace ebs_pack4 {
#include "march_ebs_body.cuh"
}
#undef EBS_PACK
#define EBS_PACK 8
namespace ebs_pack8 {
#include "march_ebs_body.cuh"
}
#undef EBS_PACK

static f3 h3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }

static void ebs_fill_const(vrb_ctx* c, const vrb_lighting* light, const vrb_ebs_params* p, EbsConst& E) {
  E.VS = h3(c->scale[0], c->scale[1], c->scale[2]);
  E.VSS = h3((float)c->vw * E.VS.x, (float)c->vh * E.VS.y, (float)c->vd * E.VS.z);
  E.MinSAT = h3(E.VS.x * 0.5f, E.VS.y * 0.5f, E.VS.z * 0.5f);
  E.MaxSAT = h3(E.VSS.x + E.VS.x * 1.5f, E.VSS.y + E.VS.y * 1.5f, E.VSS.z + E.VS.z * 1.5f);
  E.MinVol = E.MinSAT;
  E.MaxVol = h3(E.VSS.x - E.VS.x * 0.5f, E.VSS.y - E.VS.y * 0.5f, E.VSS.z - E.VS.z * 0.5f);
  E.inv_vol_scaled = h3(1.0f / (E.VSS.x + E.VS.x * 2.0f), 1.0f / (E.VSS.y + E.VS.y * 2.0f), 1.0f / (E.VSS.z + E.VS.z * 2.0f));
  E.InvG = h3(1.0f / E.VSS.x, 1.0f / E.VSS.y, 1.0f / E.VSS.z);
  E.p_cs = cosf(p->sdw_cone_angle_rad); E.p_sn = sinf(p->sdw_cone_angle_rad);
  E.n_cs = cosf(-p->sdw_cone_angle_rad); E.n_sn = sinf(-p->sdw_cone_angle_rad);
  E.sat = c->d_sat; E.sw = c->sat_w; E.sh = c->sat_h; E.sd = c->sat_d; E.sslice = (long long)c->sat_w * c->sat_h;
  E.P = *p;
  E.ka = light->ka; E.kd = light->kd;
  E.light_pos = h3(light->light_pos[0], light->light_pos[1], light->light_pos[2]);
  E.light_fwd = h3(light->light_forward[0], light->light_forward[1], light->light_forward[2]);

  memset(&E.ph, 0, sizeof(E.ph));
  E.sat_packed = c->d_sat_packed;
  E.sat_tex = c->sat_tex; E.atlas_tiles_x = c->atlas_tiles_x; E.atlas_tile_w = c->sat_w + 2; E.atlas_tile_h = c->sat_h + 2;
}

extern "C" int vrb_ebs_render(vrb_ctx* c, const vrb_camera* cam, const vrb_lighting* light, const vrb_ebs_params* p) {
  VRB_REQUIRE(c && cam && light && p, VRB_ERR_INVALID, "vrb_ebs_render: NULL argument");
  VRB_REQUIRE(c->d_vol, VRB_ERR_STATE, "vrb_ebs_render: no volume uploaded");
  VRB_REQUIRE(c->d_tf_rgbt, VRB_ERR_STATE, "vrb_ebs_render: no transfer function uploaded");
  VRB_REQUIRE(c->d_sat, VRB_ERR_STATE, "vrb_ebs_render: no SAT (vrb_sat_build)");
  VRB_REQUIRE(c->d_frame, VRB_ERR_STATE, "vrb_ebs_render: no frame (vrb_frame_resize)");
  VRB_REQUIRE(p->step_size > 0.0f, VRB_ERR_INVALID, "vrb_ebs_render: step_size %g", p->step_size);
  VRB_REQUIRE(p->amb_occ_shells >= 1 && p->amb_occ_shells <= 4096, VRB_ERR_INVALID, "vrb_ebs_render: amb_occ_shells %d", p->amb_occ_shells);
  VRB_REQUIRE(p->sdw_sample_interval > 0.0f, VRB_ERR_INVALID, "vrb_ebs_render: sdw_sample_interval %g", p->sdw_sample_interval);
  VRB_CUDA(cudaSetDevice(c->device));
  EbsConst E;
  ebs_fill_const(c, light, p, E);
  { int rc = vrb_make_phong_view(c, light, &E.ph, "vrb_ebs_render"); if (rc != VRB_OK) return rc; }
  if (!c->d_frame_target) VRB_CUDA(cudaMemsetAsync(c->d_frame, 0, (size_t)c->fw * c->fh * 4 * sizeof(__half), c->stream));
  if (p->count_samples) { int rc = vrb_counters_reset(c); if (rc != VRB_OK) return rc; }
  const int pack = (c->sat_pack == 8 && c->sat_tex) ? 8 : (c->d_sat_packed ? c->sat_pack : 1);
  // default: M lanes per ray with the longest-first CTA order (k_ebs_coop).  VRB_EBS_KERNEL=deferred: the three-kernel frame
  // the other lit renderers use (march_list.cu -> k_ebs_shade -> composite); =ray: one thread per ray.  Measured on B200 at
  // config 2, 1 / 8 GPUs: coop 8.19 / 1.66 ms, deferred 9.8 / 1.68 ms.  Both sit on the texture pipe (16 tex2Dgather per SAT
  // box query); the fused kernel keeps it busier (L1TEX 91 % against 79 %), and with ~10 visible samples per ray there is no
  // long march for a deferred frame to take out of the way, so the fused kernel stays the default here.
  const char* kern = getenv("VRB_EBS_KERNEL");
  if (kern && !strcmp(kern, "deferred")) {
    ListFrame f;
    int rc = vrb_list_march(c, cam, p->step_size, 0, p->count_samples, &f);
    if (rc != VRB_OK) return rc;
    if (f.n_entries) {
      const unsigned blocks = (f.n_entries + 127u) / 128u;
      const CamView cv = make_cam_view(cam);
      VrbKernelTimer timer(c, "k_ebs_shade");
#define VRB_EBS_SHADE(NS) do { if (p->count_samples) NS::k_ebs_shade<true><<<blocks, 128, 0, c->stream>>>(c->vol_view(), cv, E, f.L, f.n_entries, c->d_counter); \
                               else NS::k_ebs_shade<false><<<blocks, 128, 0, c->stream>>>(c->vol_view(), cv, E, f.L, f.n_entries, c->d_counter); } while (0)
      if (pack == 8) VRB_EBS_SHADE(ebs_pack8); else if (pack == 4) VRB_EBS_SHADE(ebs_pack4); else if (pack == 2) VRB_EBS_SHADE(ebs_pack2); else VRB_EBS_SHADE(ebs_pack1);
#undef VRB_EBS_SHADE
      VRB_CUDA(cudaGetLastError());
      c->launches++;
    }
    rc = vrb_list_composite(c, cam, 0, f);
    if (rc != VRB_OK) return rc;
    if (p->count_samples) return vrb_counters_fetch(c);
    return VRB_OK;
  }
  PartView part;
  dim3 block(8, 8), grid = vrb_make_grid(c, 8, 8, &part);
  size_t smem = (c->tf_n + 2 <= 1026) ? (size_t)(c->tf_n + 2) * sizeof(float4) : 0;
#define VRB_EBS_LAUNCH(NS)                                                                                                          \
  do {                                                                                                                              \
    if (p->count_samples) NS::k_ebs<true><<<grid, block, smem, c->stream>>>(c->vol_view(), c->d_tf_rgbt, c->tf_n, c->frame_view(),  \
                                                                           make_cam_view(cam), part, E, c->d_counter);          \
    else NS::k_ebs<false><<<grid, block, smem, c->stream>>>(c->vol_view(), c->d_tf_rgbt, c->tf_n, c->frame_view(),                  \
                                                            make_cam_view(cam), part, E, c->d_counter);                         \
  } while (0)
  // lanes per ray (see k_ebs_coop): 1 = one thread per ray
  static const int lanes_env = getenv("VRB_EBS_LANES") ? atoi(getenv("VRB_EBS_LANES")) : 0;
  // default: 4 lanes per ray.  Measured on B200 at config 2 with the longest-first CTA order, full frame / one eighth of
  // it (a rank of an 8-GPU sort-first run): 4 lanes 8.22 / 1.71 ms, 8 lanes 10.73 / 1.37, 16 lanes 11.56 / 1.50.  More
  // lanes speculate more samples per round (more work) but shorten every ray, which is what bounds a rank that owns a
  // small share of the frame: 8 lanes from 6 ranks up.  (Handing rays with many shaded samples to a warp-per-ray
  // continuation kernel was tried too: such rays are too rare at this config to matter, no gain.)
  const int lanes_auto = c->part.nranks >= 6 ? 8 : 4;
  const int lanes_req = lanes_env ? lanes_env : lanes_auto;
  // the gradient Blinn-Phong branch lives in the one-thread-per-ray kernel only
  const int lanes = (E.ph.grad || (kern && !strcmp(kern, "ray"))) ? 1 : ((lanes_req == 2 || lanes_req == 4 || lanes_req == 8 || lanes_req == 16) ? lanes_req : 1);
#define VRB_EBS_LAUNCH_COOP(NS, M)                                                                                                  \
  do {                                                                                                                              \
    const int tw = (M >= 16) ? 2 : (M >= 4) ? 4 : 8, th = (64 / M) / tw;                                                            \
    PartView part; dim3 g2 = vrb_make_grid(c, tw, th, &part);                                                                       \
    const unsigned n_ctas = g2.x * g2.y;                                                                                            \
    const unsigned long long sig = ((unsigned long long)c->fw << 40) ^ ((unsigned long long)c->fh << 24) ^ ((unsigned long long)M << 16) ^ \
                                   ((unsigned long long)part.nranks << 8) ^ (unsigned long long)part.rank ^                         \
                                   ((unsigned long long)part.tile_w << 52) ^ ((unsigned long long)part.compact << 63);              \
    const unsigned int* order = nullptr; unsigned int* cost = nullptr;                                                              \
    { int rc = vrb_cta_order_prepare(c, n_ctas, sig, &order, &cost); if (rc != VRB_OK) return rc; }                                 \
    if (p->count_samples) NS::k_ebs_coop<true, M><<<n_ctas, 64, smem, c->stream>>>(c->vol_view(), c->d_tf_rgbt, c->tf_n,            \
                                                      c->frame_view(), make_cam_view(cam), part, E, c->d_counter, order, cost);     \
    else NS::k_ebs_coop<false, M><<<n_ctas, 64, smem, c->stream>>>(c->vol_view(), c->d_tf_rgbt, c->tf_n, c->frame_view(),           \
                                                              make_cam_view(cam), part, E, c->d_counter, order, cost);             \
  } while (0)
  VrbKernelTimer timer(c, lanes > 1 ? "k_ebs_coop" : "k_ebs");
  if (lanes > 1 && (pack == 8 || pack == 1)) {
    if (pack == 8) { if (lanes == 2) VRB_EBS_LAUNCH_COOP(ebs_pack8, 2); else if (lanes == 4) VRB_EBS_LAUNCH_COOP(ebs_pack8, 4); else if (lanes == 8) VRB_EBS_LAUNCH_COOP(ebs_pack8, 8); else VRB_EBS_LAUNCH_COOP(ebs_pack8, 16); }
    else           { if (lanes == 2) VRB_EBS_LAUNCH_COOP(ebs_pack1, 2); else if (lanes == 4) VRB_EBS_LAUNCH_COOP(ebs_pack1, 4); else if (lanes == 8) VRB_EBS_LAUNCH_COOP(ebs_pack1, 8); else VRB_EBS_LAUNCH_COOP(ebs_pack1, 16); }
  } else
  if (pack == 8) VRB_EBS_LAUNCH(ebs_pack8);
  else if (pack == 4) VRB_EBS_LAUNCH(ebs_pack4);
  else if (pack == 2) VRB_EBS_LAUNCH(ebs_pack2);
  else VRB_EBS_LAUNCH(ebs_pack1);
#undef VRB_EBS_LAUNCH
  VRB_CUDA(cudaGetLastError());
  c->launches++;
  if (p->count_samples) return vrb_counters_fetch(c);
  return VRB_OK;
}

// PreComputeLightCache (ebsrenderer.cpp:441-555): dispatch of rc1pextbsd/lightcachecomputation.comp over the cache voxels.
extern "C" int vrb_ebs_light_cache_build(vrb_ctx* c, const vrb_lighting* light, const vrb_ebs_params* p, int rw, int rh, int rd) {
  VRB_REQUIRE(c && light && p, VRB_ERR_INVALID, "vrb_ebs_light_cache_build: NULL argument");
  VRB_REQUIRE(c->d_vol, VRB_ERR_STATE, "vrb_ebs_light_cache_build: no volume uploaded");
  VRB_REQUIRE(c->d_sat, VRB_ERR_STATE, "vrb_ebs_light_cache_build: no SAT (vrb_sat_build)");
  VRB_REQUIRE(p->amb_occ_shells >= 1 && p->amb_occ_shells <= 4096, VRB_ERR_INVALID, "vrb_ebs_light_cache_build: amb_occ_shells %d", p->amb_occ_shells);
  VRB_REQUIRE(p->sdw_sample_interval > 0.0f, VRB_ERR_INVALID, "vrb_ebs_light_cache_build: sdw_sample_interval %g", p->sdw_sample_interval);
  VRB_REQUIRE(rw >= 1 && rh >= 1 && rd >= 1 && rw <= 1024 && rh <= 1024 && rd <= 1024, VRB_ERR_INVALID,
              "vrb_ebs_light_cache_build: bad resolution %dx%dx%d", rw, rh, rd);
  VRB_CUDA(cudaSetDevice(c->device));
  int rc = vrb_light_cache_alloc(c, rw, rh, rd);
  if (rc != VRB_OK) return rc;
  EbsConst E;
  ebs_fill_const(c, light, p, E);
  f3 cell = h3(c->scale[0] * ((float)c->vw / (float)rw), c->scale[1] * ((float)c->vh / (float)rh), c->scale[2] * ((float)c->vd / (float)rd));
  const int n = rw * rh * rd;
  if (c->sat_pack == 8 && c->sat_tex) ebs_pack8::k_ebs_light_cache<<<(n + 63) / 64, 64, 0, c->stream>>>(E, cell, rw, rh, rd, c->d_light_cache);
  else                                ebs_pack1::k_ebs_light_cache<<<(n + 63) / 64, 64, 0, c->stream>>>(E, cell, rw, rh, rd, c->d_light_cache);
  VRB_CUDA(cudaGetLastError());
  vrb_light_cache_finish(c);
  c->launches += 2;
  return VRB_OK;
}
