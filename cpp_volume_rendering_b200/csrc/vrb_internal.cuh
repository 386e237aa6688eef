// vrb_internal.cuh -- shared device/host internals of libvrb200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <string>
#include <vector>
#include "../../include/vrb200.h"
#include "shade_list.cuh"

// ---------------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------------
void vrb_set_error(const char* fmt, ...);
#define VRB_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      vrb_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return VRB_ERR_CUDA;                                                                      \
    }                                                                                           \
  } while (0)
#define VRB_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      vrb_set_error(__VA_ARGS__);      \
      return (code);                   \
    }                                  \
  } while (0)

// ---------------------------------------------------------------------------------------------------------
// device-side views
// ---------------------------------------------------------------------------------------------------------
// Volume as the reference's R16F 3-D texture, GL_LINEAR + CLAMP_TO_EDGE (libs/volvis_utils/utils.cpp:8-9,44-47).
// Layout in HBM: fp16 texels, x fastest, padded by ONE replicated texel on every side so that clamp-to-edge needs
// no index clamping in the hot loop: texel (x,y,z) lives at (x+1) + pw*((y+1) + ph*(z+1)).
struct VolView {
  const __half* tex;      // padded fp16 texels
  int w, h, d;            // resolution
  int pw, ph, pd;         // padded dims (w+2, ...)
  long long slice;        // pw*ph
  float gx, gy, gz;       // VolumeGridSize = resolution * voxel scale
  float sx, sy, sz;       // voxel scale
  // optional texture-unit path: the same padded texels as a 2-D fp16 cudaArray atlas, one (w+2)x(h+2) tile per padded
  // z slice, sampled with tex2Dgather (exact texels, no hardware filtering).  0 = not available (volume too large).
  cudaTextureObject_t atlas;
  int atlas_tiles_x;
  // hardware-filtered path (filter mode VRB_FILTER_HARDWARE): the unpadded texels as a 3-D fp16 cudaArray with
  // linear filtering and clamp addressing, i.e. the texture unit's own GL_LINEAR (9-bit blend weights).  0 = not built.
  cudaTextureObject_t tex3d;
};

// 1-D RGBA16F texture with clamp-to-edge (transfer function, cone section tables), as fp16-rounded float4 texels
// padded by one replicated texel at both ends: texel i at index i+1.
struct TfView {
  const float4* tex;      // n+2 entries
  int n;
};

struct FrameView {
  __half* rgba;           // W*H*4 halves, row 0 = bottom (the context's frame, or the target set by vrb_frame_set_target)
  int w, h;
  int zero_miss;          // != 0: rays that miss store zeros (no per-render clear: the buffer is shared with other writers)
};

struct CamView {
  float ex, ey, ez;
  float m[9];             // columns of mat3(lookAt): m[3*c + r]
  float tan_fovy, aspect;
};

// Uniforms of the gradient Blinn-Phong branch every shader carries (ShadeBlinnPhong, ray_marching_1p.comp:48-81, and the
// ApplyPhongShading branches of the lit shaders).  grad == nullptr: ApplyPhongShading / ApplyGradientPhongShading == 0.
struct PhongView {
  const uint2* grad;        // TexVolumeGradient: padded half4 texels (x, y, z, 0), indexed like the volume (gradient.cu)
  float lx, ly, lz;         // LightSourcePosition / WorldLightingPos
  float ka, kd, ks, shininess;
  float isx, isy, isz;      // BlinnPhongIspecular / Ispecular
};

// occupancy cells of empty_space.cu as the marchers see them (flags == nullptr: no skipping)
struct CellView { const unsigned char* flags; int cw, ch, cd; };

struct PartView { int rank, nranks, tile_w, tile_h; int compact; };   // compact: 1-D grid over the owned tiles only (vrb_make_grid)

// One level of a mip pyramid (or the volume itself): padded fp16 texels, texel (x,y,z) at (x+1,y+1,z+1).
struct LevelView { const __half* tex; int w, h, d; };
#define VRB_MAX_LEVELS 16

// Uniform block of one ConeGaussianSampler as the DOS shader sees it (ray_bbox_marching.comp:21-37).
struct ConeView {
  const float4* sections;   // texelFetch of the RGBA16F section table: [interval, mip level, d_integral, amplitude]
  int counts[3];            // *ConeIntegrationSamples
  float initial_step, ray7_adj_weight, ui_weight;
  float axes[10][3];        // *ConeRayAxes
};

// ---------------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------------
struct vrb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  uint64_t launches = 0;
  uint64_t last_samples = 0, last_aux = 0;
  // vrb_ctx_set_kernel_timing: CUDA events around the dominant kernel of every render call (VrbKernelTimer below)
  bool time_kernels = false;
  cudaEvent_t ev_kern[2] = {nullptr, nullptr};
  bool kern_timed = false;
  const char* kern_name = "";
  float last_prepass_ms = 0.f;   // device time of the kernels of the last pre-pass (SAT scans), CUDA events
  unsigned long long* d_counter = nullptr;   // device counters: [0] primary samples, [1] secondary work items
  PartView part{0, 1, 64, 64, 0};

  // volume
  int vw = 0, vh = 0, vd = 0, bpv = 0;
  float scale[3] = {1, 1, 1};
  void* d_raw = nullptr;        // raw voxels (u8/u16), kept for the pre-passes (SAT fill, super-voxel pyramid)
  __half* d_vol = nullptr;      // padded fp16 texels
  cudaArray_t vol_array = nullptr;          // gather atlas of the same texels (env VRB_VOL_GATHER=0 disables)
  cudaTextureObject_t vol_tex = 0;
  int vol_atlas_tiles_x = 0;
  cudaArray_t vol_array3d = nullptr;        // hardware-filtered 3-D texture of the same texels (built on first use)
  cudaTextureObject_t vol_tex3d = 0;
  int filter_mode = 0;                      // VRB_FILTER_EXACT / VRB_FILTER_HARDWARE (vrb_ctx_set_filter)

  // longest-first CTA schedule of the EBS marcher: durations of the previous frame's CTAs (cta_order.cu)
  unsigned int* d_cta_cost = nullptr;     // written by the running frame (max over the CTA's warps of clock ticks >> 6)
  unsigned int* d_cta_keys = nullptr;     // previous frame's costs (sort input), sorted keys (scratch)
  unsigned int* d_cta_keys_sorted = nullptr;
  unsigned int* d_cta_iota = nullptr;
  unsigned int* d_cta_order = nullptr;    // logical CTA ids, heaviest first
  void* d_cta_sort_tmp = nullptr; size_t cta_sort_tmp_bytes = 0;
  unsigned int cta_n = 0;
  unsigned long long cta_sig = 0;         // frame size / partition / tile shape the costs belong to
  bool cta_cost_valid = false;

  // empty-space cells (rc1pass skip_empty): min/max of the (8+1)^3 padded texels a sample with floor index in the
  // cell can touch, and one byte per cell = "some sample in this cell can have alpha != 0 under the current TF"
  __half2* d_cell_mm = nullptr;
  unsigned char* d_cell_flags = nullptr;
  int* d_tf_nz = nullptr;       // prefix count of padded TF texels with alpha != 0
  int cell_dims[3] = {0, 0, 0};
  bool cell_mm_valid = false, cell_flags_valid = false;
  float cell_empty_fraction = 0.f;   // share of cells flagged empty under the current TF (vrb_cells_prepare)
  unsigned* d_cell_count = nullptr;

  // the padded fp16 volume again as 2x2 texel quads (march_list.cu): entry (x,y,z) = texels (x,y) (x+1,y) (x,y+1) (x+1,y+1)
  // of slice z as four halves, so a trilinear footprint is two 8-byte loads; built on first use, dropped with the volume
  uint2* d_vol_quad = nullptr;
  bool vol_quad_tried = false;

  // gradient texture (gradient.cu): padded half4 texels, nullptr = none (the reference's default, datamanager.cpp:27)
  void* d_grad = nullptr;
  int grad_mode = 0;

  // transfer function
  int tf_n = 0;
  float4* d_tf_rgbt = nullptr;  // n+2, .w = extinction
  float4* d_tf_rgba = nullptr;  // n+2, .w = opacity

  // frame
  int fw = 0, fh = 0;
  __half* d_frame = nullptr;
  __half* d_frame_target = nullptr;   // vrb_frame_set_target: where the marchers store (local extra buffer or peer memory)
  __half* d_frame_extra[2] = {nullptr, nullptr};   // vrb_frame_extra: context-owned buffers for double-buffered targets
  // pipelined read-back (vrb_frame_read_rgba32f_async): fp32 staging x2, copy stream, events
  cudaStream_t copy_stream = nullptr;
  float* d_stage[2] = {nullptr, nullptr};
  size_t stage_px = 0;
  cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  bool stage_busy[2] = {false, false};
  unsigned stage_next = 0;
  __half* d_filtered = nullptr;   // pixel multi-scaling (frame_filters.cu): the W x H "filtered screen output"
  int flw = 0, flh = 0;
  void* d_partial = nullptr;    // sort-last: premultiplied fp32 RGBA of this brick's ray segments (float4 per pixel)
  void* d_brick_alpha = nullptr; // sort-last, exact two-pass mode: opacity of this brick's segment (float per pixel)
  size_t partial_px = 0;

  // SAT (rc1pextbsd)
  float* d_sat = nullptr;       // (vw+2)(vh+2)(vd+2) fp32
  int sat_w = 0, sat_h = 0, sat_d = 0;
  double* d_sat_slab64 = nullptr; int sat_slab_lo = 0, sat_slab_hi = 0;   // sharded build: this rank's z-slab in fp64 (vrb_sat_build_slab)
  void* d_sat_packed = nullptr; // same texels with their +x (pack 2: float2) or +x,+y,+xy (pack 4: float4) neighbours
  int sat_order = 0;            // VRB_SAT_ORDER_REFERENCE / VRB_SAT_ORDER_SCAN (vrb_sat_set_order, env VRB_SAT_ORDER=scan)
  int sat_pack = 8;             // layout the marcher samples: 1 linear, 2 x-pairs, 4 xy-quads, 8 texture-gather atlas (env VRB_SAT_PACK)
  cudaArray_t sat_array = nullptr;            // pack 8: 2-D atlas of (w+2)x(h+2) tiles, one per z slice
  cudaTextureObject_t sat_tex = 0;
  int atlas_tiles_x = 0;

  // extinction-coefficient pyramid + cone section tables (rc1pdosct)
  __half* d_pyr[VRB_MAX_LEVELS] = {};
  int pyr_levels = 0;
  int pyr_dims[VRB_MAX_LEVELS][3] = {};
  cudaMipmappedArray_t pyr_mip = nullptr;    // VRB_FILTER_HARDWARE: the same levels as a mipmapped 3-D texture
  cudaTextureObject_t pyr_tex = 0;
  float4* d_cone_sections[2] = {nullptr, nullptr};   // [0] occlusion, [1] shadow
  ConeView cone[2] = {};
  bool cones_set = false;
  // k_dos_compact (march_dos.cu): every level again as 2x2 texel quads (one 8-byte load = the four x/y neighbours of a
  // trilinear footprint in one z slice), and the section tables with their pyramid level resolved on the host
  uint2* d_pyr_quad[VRB_MAX_LEVELS] = {};
  bool pyr_quad_valid = false;
  std::vector<float4> h_cone_sections[2];             // fp16-rounded section texels as uploaded (host copy)
  float4* d_dos_packed = nullptr;                     // [occlusion sections][shadow sections]: interval, d_integral, amplitude, packed ints
  int dos_packed_n[2] = {0, 0};
  unsigned long long cones_gen = 0, dos_packed_sig = 0;   // generation of the cone tables / what d_dos_packed was built from

  // deferred shading list of the lit marchers (shade_list.cuh / shade_list.cu)
  float4* d_sl_a = nullptr; float4* d_sl_b = nullptr; unsigned* d_sl_next = nullptr;
  unsigned* d_sl_head = nullptr; unsigned* d_sl_counters = nullptr;
  unsigned* h_sl_counters = nullptr;                  // pinned, 2 entries
  unsigned sl_capacity = 0, sl_heads = 0;
  unsigned sl_last_entries = 0, sl_last_chunks = 0;   // of the last frame (diagnostics: vrb_last_shade_list)

  // object-space light cache (PreIlluminationStructuredVolume): RG16F (Iocc, Ishadow), padded by one replicated texel
  __half2* d_light_cache = nullptr;
  int lc_dims[3] = {0, 0, 0};

  // secondary-ray direction tables (rc1pcrtgt): [0] occlusion, [1] shadow; n x 3 fp16-rounded floats
  float* d_gt_rays[2] = {nullptr, nullptr};
  int gt_nrays[2] = {0, 0};

  // super-voxel mean/stddev pyramid + pre-integration LUT (rc1pvctsg)
  __half2* d_sv[VRB_MAX_LEVELS] = {};        // RG16F levels, padded by one replicated texel
  int sv_levels = 0;
  int sv_dims[VRB_MAX_LEVELS][3] = {};
  // brick builds (vrb_sv_build_brick): the levels are windows of the WHOLE volume's pyramid -- global level dims and the
  // window's first texel; whole-volume builds have sv_gdims == sv_dims and zero offsets
  int sv_gdims[VRB_MAX_LEVELS][3] = {};
  int sv_off[VRB_MAX_LEVELS][3] = {};
  double* d_sv_top_means = nullptr;          // fp64 means of the last level built (brick builds: input of the cross-brick top levels)
  __half* d_preint = nullptr;                // R16F 2-D LUT, padded
  int preint_w = 0, preint_h = 0;
  float sv_max_stddev = 0.0f;
  cudaMipmappedArray_t sv_mip = nullptr;     // VRB_FILTER_HARDWARE: the same pyramid / LUT as textures
  cudaTextureObject_t sv_tex = 0;
  cudaArray_t preint_array = nullptr;
  cudaTextureObject_t preint_tex = 0;

  VolView vol_view() const {
    VolView v;
    v.tex = d_vol; v.w = vw; v.h = vh; v.d = vd; v.pw = vw + 2; v.ph = vh + 2; v.pd = vd + 2;
    v.slice = (long long)v.pw * v.ph;
    v.sx = scale[0]; v.sy = scale[1]; v.sz = scale[2];
    v.gx = (float)vw * scale[0]; v.gy = (float)vh * scale[1]; v.gz = (float)vd * scale[2];
    v.atlas = vol_tex; v.atlas_tiles_x = vol_atlas_tiles_x;
    v.tex3d = (filter_mode == 1) ? vol_tex3d : 0;
    return v;
  }
  FrameView frame_view() const { return FrameView{d_frame_target ? d_frame_target : d_frame, fw, fh, d_frame_target ? 1 : 0}; }
  __half* frame_ptr() const { return d_frame_target ? d_frame_target : d_frame; }
};

// Launch grid of a marcher whose CTA covers TW x TH pixels, and the partition view to pass to it (see vrb_cta_origin).
// Brackets the launch of a render call's dominant kernel with events on the launching stream when the caller asked
// for it (bench.py's roofline: that kernel's own duration, not the frame's).  No cost otherwise.
struct VrbKernelTimer {
  vrb_ctx* c;
  VrbKernelTimer(vrb_ctx* ctx, const char* name) : c(ctx) {
    if (c->time_kernels) { cudaEventRecord(c->ev_kern[0], c->stream); c->kern_name = name; }
  }
  ~VrbKernelTimer() { if (c->time_kernels) { cudaEventRecord(c->ev_kern[1], c->stream); c->kern_timed = true; } }
};

static inline dim3 vrb_make_grid(const vrb_ctx* c, int TW, int TH, PartView* pv) {
  *pv = c->part;
  pv->compact = 0;
  if (c->part.nranks > 1 && c->part.tile_w % TW == 0 && c->part.tile_h % TH == 0) {
    const int tiles_x = (c->fw + c->part.tile_w - 1) / c->part.tile_w, tiles_y = (c->fh + c->part.tile_h - 1) / c->part.tile_h;
    const int ntiles = tiles_x * tiles_y;
    const int owned = (ntiles - c->part.rank + c->part.nranks - 1) / c->part.nranks;
    pv->compact = 1;
    return dim3((unsigned)(owned > 0 ? owned : 1) * (unsigned)((c->part.tile_w / TW) * (c->part.tile_h / TH)));
  }
  return dim3((unsigned)((c->fw + TW - 1) / TW), (unsigned)((c->fh + TH - 1) / TH));
}

// Sort-last bricks: may a brick start its ray loop at sample k0 with s = k0 * step instead of adding `step` k0 times?
// Only when every multiple k * step (k up to the longest ray's sample count) is exactly representable in fp32: then the
// shader's repeated `s = s + h` produces exactly those multiples, and the jump is bit-identical ("a proven-
// equal closed form").  step = m * 2^e with m odd: needs m * kmax < 2^24.  True for the default 0.5 (m = 1).
static inline bool vrb_step_multiples_exact(float step, float longest_ray) {
  if (!(step > 0.0f) || !std::isfinite(step) || !std::isfinite(longest_ray)) return false;
  int e = 0;
  double m = std::frexp((double)step, &e);             // step = m * 2^e, m in [0.5, 1)
  for (int i = 0; i < 24 && m != std::floor(m); ++i) m *= 2.0;
  if (m != std::floor(m)) return false;
  const double kmax = std::ceil((double)longest_ray / (double)step) + 4.0;
  return m * kmax < 16777216.0;
}

// The part of a brick's ray loop that can contain owned samples: the ray (position = w + dir * t, t in [0, D]) against the
// owned cells [lo, hi) enlarged by 1.5 voxels per side (cells are found from rounded positions; the enlargement dwarfs the
// rounding).  Returns false when no sample can be owned; else the first sample index to visit and the ray parameter after
// which nothing can be owned (two steps of slack on both sides; the loop's own ownership test stays in place).
__device__ __forceinline__ bool vrb_brick_span(float wx, float wy, float wz, float dx, float dy, float dz, float D, float step,
                                               const int lo[3], const int hi[3], float kx, float ky, float kz, int& k0, float& s_end) {
  float t0 = 0.0f, t1 = D;
  const float w[3] = {wx, wy, wz}, d[3] = {dx, dy, dz}, k[3] = {kx, ky, kz};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float lw = ((float)lo[a] - 1.5f) / k[a], hw = ((float)hi[a] + 1.5f) / k[a];
    if (fabsf(d[a]) > 1e-20f) {
      const float ta = (lw - w[a]) / d[a], tb = (hw - w[a]) / d[a];
      t0 = fmaxf(t0, fminf(ta, tb)); t1 = fminf(t1, fmaxf(ta, tb));
    } else if (w[a] < lw || w[a] > hw) t1 = -1.0f;
  }
  if (!(t1 >= t0)) return false;
  k0 = max(0, (int)floorf(t0 / step) - 2);
  s_end = t1 + 2.0f * step;
  return true;
}

void vrb_free_pyramid(vrb_ctx* c);    // extcoef_pyramid.cu
int vrb_pyr_quads_prepare(vrb_ctx* c); // extcoef_pyramid.cu: build d_pyr_quad[] from d_pyr[] if stale
void vrb_free_vct(vrb_ctx* c);        // vct_prepass.cu
void vrb_free_gradient(vrb_ctx* c);   // gradient.cu
void vrb_free_filtered(vrb_ctx* c);   // frame_filters.cu
// PhongView of a render call: grad = nullptr unless light->apply_phong (then the gradient texture must exist)
int vrb_make_phong_view(const vrb_ctx* c, const vrb_lighting* light, PhongView* out, const char* who);   // gradient.cu
int vrb_partial_alloc(vrb_ctx* c);    // sort_last.cu: (re)allocate the fp32 partial frame + segment opacity of a brick context
int vrb_brick_check(const vrb_ctx* c, const vrb_brick* b, const char* who);   // sort_last.cu: brick description vs uploaded array
int vrb_vol_tex3d_prepare(vrb_ctx* c); // vrb_api.cu: build the hardware-filtered volume texture if the filter mode asks for it
int vrb_light_cache_alloc(vrb_ctx* c, int rw, int rh, int rd);   // march_obj.cu: (re)allocate the padded RG16F cache
void vrb_light_cache_finish(vrb_ctx* c);                          // march_obj.cu: replicate the border texels
void vrb_free_light_cache(vrb_ctx* c);
// cta_order.cu: returns the schedule for this launch (nullptr = none yet: natural order) and the buffer this frame's
// CTA durations go to (nullptr = feature off, VRB_CTA_ORDER=0)
int vrb_cta_order_prepare(vrb_ctx* c, unsigned n_ctas, unsigned long long sig, const unsigned int** order, unsigned int** cost);
void vrb_free_cta_order(vrb_ctx* c);
// shade_list.cu: vrb_sl_begin (re)allocates for `n_warps` marching warps and zeroes the counters; vrb_sl_counts waits for
// the march kernel and returns {entries, chunks}; *overflow = the list was too small: it has been enlarged, march again
int vrb_sl_begin(vrb_ctx* c, unsigned n_lanes, ShadeListView* out);
int vrb_sl_counts(vrb_ctx* c, unsigned* entries, bool* overflow);
void vrb_free_shade_list(vrb_ctx* c);
void vrb_free_cells(vrb_ctx* c);      // empty_space.cu
void vrb_free_vol_quads(vrb_ctx* c);  // march_list.cu
int vrb_vol_quads_prepare(vrb_ctx* c); // march_list.cu: builds d_vol_quad when it fits (it may stay nullptr: callers fall back to d_vol)
void vrb_build_quads(const __half* lev, uint2* quad, int pw, int ph, int pd, cudaStream_t stream);   // extcoef_pyramid.cu
int vrb_cells_prepare(vrb_ctx* c);    // empty_space.cu: (re)build what is stale; VRB_OK or error
void vrb_free_sat_atlas(vrb_ctx* c);  // sat_scan.cu
void vrb_free_vol_atlas(vrb_ctx* c);  // vrb_api.cu

// counters of the *_render(count_samples=1) variants
static inline int vrb_counters_reset(vrb_ctx* c) {
  VRB_CUDA(cudaMemsetAsync(c->d_counter, 0, 2 * sizeof(unsigned long long), c->stream));
  return VRB_OK;
}
static inline int vrb_counters_fetch(vrb_ctx* c) {
  unsigned long long n[2] = {0, 0};
  VRB_CUDA(cudaMemcpyAsync(n, c->d_counter, sizeof(n), cudaMemcpyDeviceToHost, c->stream));
  VRB_CUDA(cudaStreamSynchronize(c->stream));
  c->last_samples = n[0]; c->last_aux = n[1];
  return VRB_OK;
}

static inline CamView make_cam_view(const vrb_camera* c) {
  CamView v;
  v.ex = c->eye[0]; v.ey = c->eye[1]; v.ez = c->eye[2];
  for (int col = 0; col < 3; ++col)
    for (int r = 0; r < 3; ++r) v.m[3 * col + r] = c->lookat[4 * col + r];
  v.tan_fovy = c->tan_fovy; v.aspect = c->aspect;
  return v;
}

// ---------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// GL_LINEAR weights: a + t*(b-a), one FADD + one FFMA.
__device__ __forceinline__ float vrb_lerp(float a, float b, float t) { return fmaf(t, b - a, a); }

// floor() of a non-negative float < 2^22 without the (quarter-rate) conversion pipe: FADD with round-down against
// 2^23 leaves floor(x) in the low mantissa bits.  Returns floor as int, *fl as float.
__device__ __forceinline__ int vrb_floor_pos(float x, float* fl) {
  float t = __fadd_rd(x, 8388608.0f);
  *fl = t - 8388608.0f;
  return __float_as_int(t) - 0x4B000000;
}

struct Ray {
  float ox, oy, oz, dx, dy, dz;
  float tnear, tfar;
  bool hit;
};

// Pixel -> ray -> AABB (ray_marching_1p.comp:93-104, ray_bbox_intersection.comp:18-52).  Every operation is an
// explicitly rounded IEEE fp32 op in the order the shader writes them (no FMA contraction), so the ray interval
// and hence the sample count do not depend on compiler fusion decisions.
__device__ __forceinline__ void vrb_normalize3(float& x, float& y, float& z) {
  float d = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
  float r = __fdiv_rn(1.0f, __fsqrt_rn(d));
  x = __fmul_rn(x, r); y = __fmul_rn(y, r); z = __fmul_rn(z, r);
}

__device__ __forceinline__ Ray vrb_make_ray(const CamView& cam, int px, int py, int W, int H,
                                            float gx, float gy, float gz) {
  Ray r;
  float fx = __fadd_rn((float)px, 0.5f), fy = __fadd_rn((float)py, 0.5f);
  float vx = __fadd_rn(__fmul_rn(__fdiv_rn(fx, (float)W), 2.0f), -1.0f);
  float vy = __fadd_rn(__fmul_rn(__fdiv_rn(fy, (float)H), 2.0f), -1.0f);
  float cx = __fmul_rn(__fmul_rn(vx, cam.tan_fovy), cam.aspect);
  float cy = __fmul_rn(vy, cam.tan_fovy);
  float cz = -1.0f;
  // v * mat3(M): dot with the columns
  float dx = __fadd_rn(__fadd_rn(__fmul_rn(cx, cam.m[0]), __fmul_rn(cy, cam.m[1])), __fmul_rn(cz, cam.m[2]));
  float dy = __fadd_rn(__fadd_rn(__fmul_rn(cx, cam.m[3]), __fmul_rn(cy, cam.m[4])), __fmul_rn(cz, cam.m[5]));
  float dz = __fadd_rn(__fadd_rn(__fmul_rn(cx, cam.m[6]), __fmul_rn(cy, cam.m[7])), __fmul_rn(cz, cam.m[8]));
  vrb_normalize3(dx, dy, dz);   // normalize() in main
  vrb_normalize3(dx, dy, dz);   // and again in RayAABBIntersection
  float ix = __fdiv_rn(1.0f, dx), iy = __fdiv_rn(1.0f, dy), iz = __fdiv_rn(1.0f, dz);
  float hx = __fmul_rn(gx, 0.5f), hy = __fmul_rn(gy, 0.5f), hz = __fmul_rn(gz, 0.5f);
  float ax = __fmul_rn(ix, __fadd_rn(-hx, -cam.ex)), bx = __fmul_rn(ix, __fadd_rn(hx, -cam.ex));
  float ay = __fmul_rn(iy, __fadd_rn(-hy, -cam.ey)), by = __fmul_rn(iy, __fadd_rn(hy, -cam.ey));
  float az = __fmul_rn(iz, __fadd_rn(-hz, -cam.ez)), bz = __fmul_rn(iz, __fadd_rn(hz, -cam.ez));
  float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
  float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
  r.hit = tf > tn;
  r.tnear = fmaxf(tn, 0.0f);
  r.tfar = tf;
  r.ox = cam.ex; r.oy = cam.ey; r.oz = cam.ez;
  r.dx = dx; r.dy = dy; r.dz = dz;
  return r;
}

// Trilinear fetch from the padded fp16 volume at TEXTURE-SPACE position (px,py,pz) in [0,G] (world units, origin at
// the box corner).  kx = N/G per axis.  up = p*k + 0.5 is the padded continuous index (u + 1).
// vrb_volume_coords: padded floor indices and blend fractions of a sample (shared by the fetch and the empty-space
// cells of empty_space.cu); vrb_fetch_volume: the eight taps and the blend.
__device__ __forceinline__ void vrb_volume_coords(const VolView& v, float kx, float ky, float kz, float px, float py, float pz,
                                                  int& ix, int& iy, int& iz, float& fx, float& fy, float& fz) {
  float ux = fmaf(px, kx, 0.5f), uy = fmaf(py, ky, 0.5f), uz = fmaf(pz, kz, 0.5f);
  // positions are inside the box up to rounding: clamp so that indices stay inside the padded array
  ux = fminf(fmaxf(ux, 0.0f), (float)v.w + 0.999f);
  uy = fminf(fmaxf(uy, 0.0f), (float)v.h + 0.999f);
  uz = fminf(fmaxf(uz, 0.0f), (float)v.d + 0.999f);
  float flx, fly, flz;
  ix = vrb_floor_pos(ux, &flx); iy = vrb_floor_pos(uy, &fly); iz = vrb_floor_pos(uz, &flz);
  fx = ux - flx; fy = uy - fly; fz = uz - flz;
}

__device__ __forceinline__ float vrb_fetch_volume(const VolView& v, int ix, int iy, int iz, float fx, float fy, float fz) {
  if (v.atlas) {
    // two gathers (z slices iz, iz+1) instead of eight 16-bit loads; gather order x=(i0,j1) y=(i1,j1) z=(i1,j0) w=(i0,j0)
    const float gx = (float)(ix + 1), gy = (float)(iy + 1);
    const int t0 = iz % v.atlas_tiles_x, u0 = iz / v.atlas_tiles_x;
    const int t1 = (iz + 1) % v.atlas_tiles_x, u1 = (iz + 1) / v.atlas_tiles_x;
    float4 a = tex2Dgather<float4>(v.atlas, gx + (float)(t0 * v.pw), gy + (float)(u0 * v.ph), 0);
    float4 c = tex2Dgather<float4>(v.atlas, gx + (float)(t1 * v.pw), gy + (float)(u1 * v.ph), 0);
    float c00 = vrb_lerp(a.w, a.z, fx), c10 = vrb_lerp(a.x, a.y, fx);
    float c01 = vrb_lerp(c.w, c.z, fx), c11 = vrb_lerp(c.x, c.y, fx);
    return vrb_lerp(vrb_lerp(c00, c10, fy), vrb_lerp(c01, c11, fy), fz);
  }
  const __half* p = v.tex + ((long long)iz * v.slice + (long long)iy * v.pw + ix);
  const __half* q = p + v.slice;
  float c000 = __half2float(__ldg(p)),          c100 = __half2float(__ldg(p + 1));
  float c010 = __half2float(__ldg(p + v.pw)),   c110 = __half2float(__ldg(p + v.pw + 1));
  float c001 = __half2float(__ldg(q)),          c101 = __half2float(__ldg(q + 1));
  float c011 = __half2float(__ldg(q + v.pw)),   c111 = __half2float(__ldg(q + v.pw + 1));
  float c00 = vrb_lerp(c000, c100, fx), c10 = vrb_lerp(c010, c110, fx);
  float c01 = vrb_lerp(c001, c101, fx), c11 = vrb_lerp(c011, c111, fx);
  return vrb_lerp(vrb_lerp(c00, c10, fy), vrb_lerp(c01, c11, fy), fz);
}

// The same fetch from the 2x2 quad copy of the padded volume (vrb_ctx::d_vol_quad, march_list.cu): entry (x,y,z) holds the
// padded texels (x,y) (x+1,y) (x,y+1) (x+1,y+1) of slice z, so the footprint is two 8-byte loads; same texels, same blend order.
__device__ __forceinline__ float vrb_h2f_lo(unsigned v) { return __half2float(__ushort_as_half((unsigned short)(v & 0xffffu))); }
__device__ __forceinline__ float vrb_h2f_hi(unsigned v) { return __half2float(__ushort_as_half((unsigned short)(v >> 16))); }
__device__ __forceinline__ float vrb_fetch_volume_quad(const VolView& v, const uint2* __restrict__ vq, int ix, int iy, int iz, float fx, float fy, float fz) {
  const uint2* p = vq + ((long long)iz * v.slice + (long long)iy * v.pw + ix);
  const uint2 a = __ldg(p), b = __ldg(p + v.slice);
  const float c00 = vrb_lerp(vrb_h2f_lo(a.x), vrb_h2f_hi(a.x), fx), c10 = vrb_lerp(vrb_h2f_lo(a.y), vrb_h2f_hi(a.y), fx);
  const float c01 = vrb_lerp(vrb_h2f_lo(b.x), vrb_h2f_hi(b.x), fx), c11 = vrb_lerp(vrb_h2f_lo(b.y), vrb_h2f_hi(b.y), fx);
  return vrb_lerp(vrb_lerp(c00, c10, fy), vrb_lerp(c01, c11, fy), fz);
}

__device__ __forceinline__ float vrb_sample_volume(const VolView& v, float kx, float ky, float kz,
                                                   float px, float py, float pz) {
  int ix, iy, iz; float fx, fy, fz;
  vrb_volume_coords(v, kx, ky, kz, px, py, pz, ix, iy, iz, fx, fy, fz);
  return vrb_fetch_volume(v, ix, iy, iz, fx, fy, fz);
}

// texture(TexVolumeGradient, Tpos / VolumeGridSize).xyz: the gradient texture has the volume's resolution and sampler
// state, so the taps and weights are those of the density sample at the same position.
__device__ __forceinline__ void vrb_fetch_gradient(const VolView& v, const uint2* __restrict__ grad, int ix, int iy, int iz,
                                                   float fx, float fy, float fz, float& gx, float& gy, float& gz) {
  const uint2* p = grad + ((long long)iz * v.slice + (long long)iy * v.pw + ix);
  const uint2* q = p + v.slice;
  float c[8][3];
  const uint2 t[8] = {__ldg(p), __ldg(p + 1), __ldg(p + v.pw), __ldg(p + v.pw + 1), __ldg(q), __ldg(q + 1), __ldg(q + v.pw), __ldg(q + v.pw + 1)};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    unsigned int lo = t[k].x, hi = t[k].y;
    float2 a = __half22float2(*reinterpret_cast<__half2*>(&lo));
    c[k][0] = a.x; c[k][1] = a.y; c[k][2] = __half2float(__ushort_as_half((unsigned short)(hi & 0xffffu)));
  }
  float o[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float c00 = vrb_lerp(c[0][ch], c[1][ch], fx), c10 = vrb_lerp(c[2][ch], c[3][ch], fx);
    float c01 = vrb_lerp(c[4][ch], c[5][ch], fx), c11 = vrb_lerp(c[6][ch], c[7][ch], fx);
    o[ch] = vrb_lerp(vrb_lerp(c00, c10, fy), vrb_lerp(c01, c11, fy), fz);
  }
  gx = o[0]; gy = o[1]; gz = o[2];
}

// The part of the Blinn-Phong branch all five shaders share: false when the sampled gradient is exactly zero (the
// shaders then leave the colour untouched), else max(0, N.L) and pow(max(0, H.N), shininess).  (px,py,pz) is the
// texture-space position (Tpos), eye the camera position (CameraEye).  Operation order as in the shaders; the marcher
// translation units are compiled with -fmad=false.
__device__ __forceinline__ bool vrb_phong_terms(const VolView& v, const PhongView& ph, float kx, float ky, float kz,
                                                float px, float py, float pz, float ex, float ey, float ez,
                                                float& dot_diff, float& spec) {
  int ix, iy, iz; float fx, fy, fz;
  vrb_volume_coords(v, kx, ky, kz, px, py, pz, ix, iy, iz, fx, fy, fz);
  float nx, ny, nz;
  vrb_fetch_gradient(v, ph.grad, ix, iy, iz, fx, fy, fz, nx, ny, nz);
  if (nx == 0.0f && ny == 0.0f && nz == 0.0f) return false;
  const float wx = px - (v.gx * 0.5f), wy = py - (v.gy * 0.5f), wz = pz - (v.gz * 0.5f);       // Wpos
  float r = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
  nx = nx * r; ny = ny * r; nz = nz * r;
  float lx = ph.lx - wx, ly = ph.ly - wy, lz = ph.lz - wz;
  r = 1.0f / sqrtf(lx * lx + ly * ly + lz * lz);
  lx = lx * r; ly = ly * r; lz = lz * r;
  float vx = ex - wx, vy = ey - wy, vz = ez - wz;
  r = 1.0f / sqrtf(vx * vx + vy * vy + vz * vz);
  vx = vx * r; vy = vy * r; vz = vz * r;
  float hx = vx + lx, hy = vy + ly, hz = vz + lz;
  r = 1.0f / sqrtf(hx * hx + hy * hy + hz * hz);
  hx = hx * r; hy = hy * r; hz = hz * r;
  dot_diff = fmaxf(0.0f, nx * lx + ny * ly + nz * lz);
  const float dot_spec = fmaxf(0.0f, hx * nx + hy * ny + hz * nz);
  spec = powf(dot_spec, ph.shininess);
  return true;
}

// 1-D RGBA lookup with clamp-to-edge from a padded table of n+2 float4 (shared or global memory).
__device__ __forceinline__ float4 vrb_sample_tf(const float4* __restrict__ tf, int n, float s) {
  float up = fmaf(s, (float)n, 0.5f);                    // u + 1
  up = fminf(fmaxf(up, 0.0f), (float)n + 0.5f);
  float fl; int i = vrb_floor_pos(up, &fl);
  float f = up - fl;
  float4 a = tf[i], b = tf[i + 1];
  return make_float4(vrb_lerp(a.x, b.x, f), vrb_lerp(a.y, b.y, f), vrb_lerp(a.z, b.z, f), vrb_lerp(a.w, b.w, f));
}

__device__ __forceinline__ void vrb_store_pixel(const FrameView& fr, int px, int py, float r, float g, float b, float a) {
  __half2 lo = __floats2half2_rn(r, g), hi = __floats2half2_rn(b, a);
  uint2 pk;
  pk.x = *reinterpret_cast<unsigned int*>(&lo);
  pk.y = *reinterpret_cast<unsigned int*>(&hi);
  reinterpret_cast<uint2*>(fr.rgba)[(size_t)py * fr.w + px] = pk;
}

// CTA rows are issued from the image centre outwards: the rays through the middle of the volume are the longest, and
// the hardware launches CTAs in blockIdx order, so the expensive tiles start first and the cheap border tiles fill the
// tail of the launch (k = 0,1,2,3.. -> mid, mid+1, mid-1, mid+2 ..).
__device__ __forceinline__ int vrb_center_out_row(int k, int n) {
  int mid = (n - 1) >> 1;
  int up = n - 1 - mid;                    // rows above mid
  int r = (k & 1) ? mid + ((k + 1) >> 1) : mid - (k >> 1);
  // once one side is exhausted the remaining rows come in order from the other side
  if (k > 2 * mid && mid <= up) r = k;                         // lower side exhausted: rows k..n-1 ascend
  if (k > 2 * up && up < mid) r = n - 1 - k;                   // upper side exhausted: rows descend to 0
  return r;
}

// Sort-first ownership: tiles are dealt round-robin in row-major order, with every tile row rotated by three tiles
// against the row above.  Without the rotation a frame whose tile count per row is a multiple of the rank count (1920 / 16
// = 120 tiles, 8 ranks) gives every rank whole tile COLUMNS, and the ranks' loads differ systematically across the image
// (measured: 1.23 .. 1.74 ms per rank at config 2); rotated rows put a rank's tiles on diagonals.
__device__ __host__ __forceinline__ int vrb_skew_tile(int tx, int ty, int tiles_x) { return (tx + 3 * ty) % tiles_x; }
__device__ __host__ __forceinline__ int vrb_unskew_tile(int txs, int ty, int tiles_x) { int v = (txs - 3 * ty) % tiles_x; return v < 0 ? v + tiles_x : v; }

// Pixel origin of this CTA's TW x TH pixel tile.  Whole frame: 2-D grid, rows centre-out.  Sort-first partition whose
// tiles are multiples of the CTA tile (vrb_make_grid sets pt.compact): 1-D grid over the OWNED tiles only, so that a
// rank launches no CTA for pixels of other ranks; the owned tiles are visited centre-out too.
__device__ __forceinline__ void vrb_cta_origin(const PartView& pt, int W, int TW, int TH, int& px0, int& py0) {
  if (pt.compact) {
    const int cpx = pt.tile_w / TW, cpt = cpx * (pt.tile_h / TH);
    const int owned = gridDim.x / cpt;
    const int k = vrb_center_out_row(blockIdx.x / cpt, owned), sub = blockIdx.x % cpt;
    const int tiles_x = (W + pt.tile_w - 1) / pt.tile_w;
    const int t = pt.rank + k * pt.nranks;
    const int ty = t / tiles_x;
    px0 = vrb_unskew_tile(t % tiles_x, ty, tiles_x) * pt.tile_w + (sub % cpx) * TW;
    py0 = ty * pt.tile_h + (sub / cpx) * TH;
  } else {
    px0 = blockIdx.x * TW;
    py0 = vrb_center_out_row(blockIdx.y, gridDim.y) * TH;
  }
}

// Same mapping from a LINEAR logical CTA id (1-D launch).  `ordered`: the id comes from a longest-first schedule
// (CtaOrder), so no centre-out permutation is applied on top.
__device__ __forceinline__ void vrb_cta_origin_linear(const PartView& pt, int W, int H, int TW, int TH, unsigned id, unsigned n_ctas,
                                                      bool ordered, int& px0, int& py0) {
  if (pt.compact) {
    const int cpx = pt.tile_w / TW, cpt = cpx * (pt.tile_h / TH);
    const int owned = (int)(n_ctas / (unsigned)cpt);
    const int kk = (int)(id / (unsigned)cpt), sub = (int)(id % (unsigned)cpt);
    const int k = ordered ? kk : vrb_center_out_row(kk, owned);
    const int tiles_x = (W + pt.tile_w - 1) / pt.tile_w;
    const int t = pt.rank + k * pt.nranks;
    const int ty = t / tiles_x;
    px0 = vrb_unskew_tile(t % tiles_x, ty, tiles_x) * pt.tile_w + (sub % cpx) * TW;
    py0 = ty * pt.tile_h + (sub / cpx) * TH;
  } else {
    const int gx = (W + TW - 1) / TW, gy = (H + TH - 1) / TH;
    const int bx = (int)(id % (unsigned)gx), by = (int)(id / (unsigned)gx);
    px0 = bx * TW;
    py0 = (ordered ? by : vrb_center_out_row(by, gy)) * TH;
  }
}

// Does this context render pixel (px,py)?  (sort-first tile interleave)
__device__ __forceinline__ bool vrb_owns_pixel(const PartView& pt, int px, int py, int W) {
  if (pt.nranks <= 1) return true;
  const int tiles_x = (W + pt.tile_w - 1) / pt.tile_w;
  const int ty = py / pt.tile_h;
  const int t = ty * tiles_x + vrb_skew_tile(px / pt.tile_w, ty, tiles_x);
  return (t % pt.nranks) == pt.rank;
}

#endif  // __CUDACC__
