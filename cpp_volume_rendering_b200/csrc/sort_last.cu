// sort_last.cu -- sort-last partition of the single-pass ray caster over bricks + ordered RGBA compositing.
// The reference has no multi-GPU path at all and cannot even load a 2048^3 volume (F11); this is the new
// code of SURVEY.md section 8e.  Design:
//   * every GPU holds one brick (owned voxels + one ghost layer on interior faces) and walks the ray of the WHOLE
//     volume: same AABB, same s_k sequence, same sample positions as the single-GPU kernel (march_rc1pass.cu).  A
//     sample belongs to the brick whose owned region contains the sample's voxel cell, so every sample is composited
//     by exactly one brick and the per-brick segments concatenate along the ray.  Trilinear weights and texels are
//     bit-identical to the single-GPU ones (the local index is the global one minus an integer offset).
//   * the partial results are premultiplied fp32 RGBA; compositing is front-to-back "over" in visibility order with the
//     reference's 0.99 cut applied between segments.  One kernel per GPU reads its rows of ALL partial frames straight
//     from the peers' HBM through CUDA-IPC mappings (P2P loads over NVLink / NVSwitch): no staging copy, the transfer
//     IS the compositing pass.
#include "vrb_internal.cuh"
#include <cstdlib>
#include <cstring>

struct BrickView {
  float gx, gy, gz;          // VolumeGridSize of the WHOLE volume
  float kx, ky, kz;          // global N / G
  float offx, offy, offz;    // origin - ghost_lo: global voxel index of local texel 0
  int lo[3], hi[3];          // owned cells [lo, hi)
  int nx, ny, nz;            // global resolution
};

__device__ __forceinline__ float brick_sample(const VolView& v, const BrickView& b, float px, float py, float pz) {
  // global padded continuous index (u + 1), then shifted into the brick's array; the subtraction of an integer is exact
  float ux = fmaf(px, b.kx, 0.5f) - b.offx, uy = fmaf(py, b.ky, 0.5f) - b.offy, uz = fmaf(pz, b.kz, 0.5f) - b.offz;
  ux = fminf(fmaxf(ux, 0.0f), (float)v.w + 0.999f);
  uy = fminf(fmaxf(uy, 0.0f), (float)v.h + 0.999f);
  uz = fminf(fmaxf(uz, 0.0f), (float)v.d + 0.999f);
  float flx, fly, flz;
  int ix = vrb_floor_pos(ux, &flx), iy = vrb_floor_pos(uy, &fly), iz = vrb_floor_pos(uz, &flz);
  float fx = ux - flx, fy = uy - fly, fz = uz - flz;
  const __half* p = v.tex + ((long long)iz * v.slice + (long long)iy * v.pw + ix);
  const __half* q = p + v.slice;
  float c00 = vrb_lerp(__half2float(__ldg(p)), __half2float(__ldg(p + 1)), fx);
  float c10 = vrb_lerp(__half2float(__ldg(p + v.pw)), __half2float(__ldg(p + v.pw + 1)), fx);
  float c01 = vrb_lerp(__half2float(__ldg(q)), __half2float(__ldg(q + 1)), fx);
  float c11 = vrb_lerp(__half2float(__ldg(q + v.pw)), __half2float(__ldg(q + v.pw + 1)), fx);
  return vrb_lerp(vrb_lerp(c00, c10, fy), vrb_lerp(c01, c11, fy), fz);
}

#define VRB_MAX_PARTIALS 64
struct AlphaList { const float* p[VRB_MAX_PARTIALS]; int n; };

// MODE 0: independent segment (da starts at 0) -> float4 partial for the ordered "over" compositor.
// MODE 1: alpha-only pre-pass -> the segment's opacity (float per pixel).
// MODE 2: exact pass: da starts at the opacity accumulated by the bricks in FRONT of this one (their MODE-1 results,
//         read in visibility order, possibly from peer memory), so the reference's 0.99 cut falls on the same sample as
//         on one GPU; colours of all bricks then simply add and the final alpha is the maximum.
template <bool COUNT, int MODE>
__global__ void __launch_bounds__(64)
k_rc1pass_brick(VolView vol, BrickView B, const float4* __restrict__ tf_g, int tf_n, float4* __restrict__ partial, float* __restrict__ alpha_out,
                const __grid_constant__ AlphaList front, int W, int H, CamView cam, float step, int jump, unsigned long long* counter) {
  extern __shared__ float4 s_tf[];
  const float4* tf = tf_g;
  if (tf_n + 2 <= 1026) {
    for (int i = threadIdx.y * 8 + threadIdx.x; i < tf_n + 2; i += 64) s_tf[i] = tf_g[i];
    __syncthreads();
    tf = s_tf;
  }
  int px = blockIdx.x * 8 + threadIdx.x, py = vrb_center_out_row(blockIdx.y, gridDim.y) * 8 + threadIdx.y;
  unsigned int ns = 0;
  if (px < W && py < H) {
    float dr = 0.f, dg = 0.f, db = 0.f, da = 0.f;
    bool touched = false;
    if (MODE == 2) {
      for (int k = 0; k < front.n; ++k) { float a = front.p[k][(size_t)py * W + px]; da = fmaf(1.0f - da, a, da); }
    }
    Ray r = vrb_make_ray(cam, px, py, W, H, B.gx, B.gy, B.gz);
    if (r.hit && !(MODE == 2 && da > 0.99f)) {
      float D = fabsf(__fadd_rn(r.tfar, -r.tnear));
      float tx = __fadd_rn(__fadd_rn(r.ox, __fmul_rn(r.dx, r.tnear)), __fmul_rn(B.gx, 0.5f));
      float ty = __fadd_rn(__fadd_rn(r.oy, __fmul_rn(r.dy, r.tnear)), __fmul_rn(B.gy, 0.5f));
      float tz = __fadd_rn(__fadd_rn(r.oz, __fmul_rn(r.dz, r.tnear)), __fmul_rn(B.gz, 0.5f));
      // jump (host: vrb_step_multiples_exact): start at the first sample that can be owned, s = k0 * step, instead of
      // walking there -- the walk was the whole cost of a back brick's pass (ncu on the 1144^3 window: 2.6 ms, issue-bound,
      // DRAM 1 %)
      float s = 0.0f, s_end = 3.0e38f;
      bool any = true;
      if (jump) { int k0 = 0; any = vrb_brick_span(tx, ty, tz, r.dx, r.dy, r.dz, D, step, B.lo, B.hi, B.kx, B.ky, B.kz, k0, s_end); s = __fmul_rn((float)k0, step); }
      for (; any && s < D && s <= s_end;) {
        float h = fminf(step, __fadd_rn(D, -s));
        float t = __fadd_rn(s, __fmul_rn(h, 0.5f));
        float qx = fmaf(r.dx, t, tx), qy = fmaf(r.dy, t, ty), qz = fmaf(r.dz, t, tz);
        // voxel cell of the sample in the global grid (positions on the far faces belong to the last cell)
        int cx = min(max((int)floorf(qx * B.kx), 0), B.nx - 1);
        int cy = min(max((int)floorf(qy * B.ky), 0), B.ny - 1);
        int cz = min(max((int)floorf(qz * B.kz), 0), B.nz - 1);
        if (cx >= B.lo[0] && cx < B.hi[0] && cy >= B.lo[1] && cy < B.hi[1] && cz >= B.lo[2] && cz < B.hi[2]) {
          float density = brick_sample(vol, B, qx, qy, qz);
          float4 src = vrb_sample_tf(tf, tf_n, density);
          if (COUNT) ++ns;
          touched = true;
          if (src.w > 0.0f) {
            float a = 1.0f - __expf(-src.w * h);
            float om = (1.0f - da) * a;
            if (MODE != 1) { dr = fmaf(om, src.x, dr); dg = fmaf(om, src.y, dg); db = fmaf(om, src.z, db); }
            da = da + om;
            if (da > 0.99f) break;      // opaque: nothing behind this sample matters
          }
        }
        s = __fadd_rn(s, h);
      }
    }
    if (MODE == 1) alpha_out[(size_t)py * W + px] = da;
    else partial[(size_t)py * W + px] = make_float4(dr, dg, db, (MODE == 2 && !touched) ? 0.0f : da);
  }
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) ns += __shfl_xor_sync(0xffffffffu, ns, o);
    if (((threadIdx.y * 8 + threadIdx.x) & 31) == 0 && ns) atomicAdd(counter, (unsigned long long)ns);
  }
}

int vrb_brick_check(const vrb_ctx* c, const vrb_brick* b, const char* who) {
  const int dims[3] = {c->vw, c->vh, c->vd};
  for (int a = 0; a < 3; ++a) {
    VRB_REQUIRE(b->owned[a] > 0 && b->ghost_lo[a] >= 0 && b->ghost_hi[a] >= 0 && b->origin[a] >= 0 &&
                b->origin[a] + b->owned[a] <= b->global_dims[a], VRB_ERR_INVALID, "%s: bad brick on axis %d", who, a);
    VRB_REQUIRE(b->ghost_lo[a] <= b->origin[a] && b->origin[a] + b->owned[a] + b->ghost_hi[a] <= b->global_dims[a], VRB_ERR_INVALID,
                "%s: ghost layers reach outside the volume on axis %d", who, a);
    VRB_REQUIRE(b->ghost_lo[a] + b->owned[a] + b->ghost_hi[a] == dims[a], VRB_ERR_INVALID,
                "%s: uploaded array has %d voxels on axis %d, brick says %d + %d + %d", who, dims[a], a, b->ghost_lo[a], b->owned[a], b->ghost_hi[a]);
    VRB_REQUIRE((b->origin[a] == 0 || b->ghost_lo[a] >= 1) && (b->origin[a] + b->owned[a] == b->global_dims[a] || b->ghost_hi[a] >= 1),
                VRB_ERR_INVALID, "%s: interior faces need at least one ghost layer (axis %d)", who, a);
  }
  return VRB_OK;
}

int vrb_partial_alloc(vrb_ctx* c) {
  const size_t npx = (size_t)c->fw * c->fh;
  if (!c->d_partial || c->partial_px != npx) {
    if (c->d_partial) { VRB_CUDA(cudaStreamSynchronize(c->stream)); VRB_CUDA(cudaFree(c->d_partial)); VRB_CUDA(cudaFree(c->d_brick_alpha)); c->d_partial = nullptr; c->d_brick_alpha = nullptr; }
    VRB_CUDA(cudaMalloc(&c->d_partial, npx * sizeof(float4)));
    VRB_CUDA(cudaMalloc(&c->d_brick_alpha, npx * sizeof(float)));
    VRB_CUDA(cudaMemsetAsync(c->d_partial, 0, npx * sizeof(float4), c->stream));
    VRB_CUDA(cudaMemsetAsync(c->d_brick_alpha, 0, npx * sizeof(float), c->stream));
    c->partial_px = npx;
  }
  return VRB_OK;
}

static int render_brick(vrb_ctx* c, const vrb_camera* cam, const vrb_rc1pass_params* p, const vrb_brick* b, int mode,
                        const void* const* front_alphas, int n_front) {
  VRB_REQUIRE(c && cam && p && b, VRB_ERR_INVALID, "vrb_rc1pass_render_brick: NULL argument");
  VRB_REQUIRE(n_front >= 0 && n_front <= VRB_MAX_PARTIALS && (n_front == 0 || front_alphas), VRB_ERR_INVALID, "vrb_rc1pass_render_brick: bad front list");
  VRB_REQUIRE(c->d_vol && c->d_tf_rgbt && c->d_frame, VRB_ERR_STATE, "vrb_rc1pass_render_brick: volume / transfer function / frame missing");
  VRB_REQUIRE(p->step_size > 0.0f, VRB_ERR_INVALID, "vrb_rc1pass_render_brick: step_size %g", p->step_size);
  { int rc = vrb_brick_check(c, b, "vrb_rc1pass_render_brick"); if (rc != VRB_OK) return rc; }
  VRB_CUDA(cudaSetDevice(c->device));
  { int rc = vrb_partial_alloc(c); if (rc != VRB_OK) return rc; }
  AlphaList front; front.n = n_front;
  for (int i = 0; i < n_front; ++i) { VRB_REQUIRE(front_alphas[i], VRB_ERR_INVALID, "vrb_rc1pass_render_brick: front alpha %d is NULL", i); front.p[i] = (const float*)front_alphas[i]; }
  BrickView B;
  B.gx = (float)b->global_dims[0] * c->scale[0]; B.gy = (float)b->global_dims[1] * c->scale[1]; B.gz = (float)b->global_dims[2] * c->scale[2];
  B.kx = (float)b->global_dims[0] / B.gx; B.ky = (float)b->global_dims[1] / B.gy; B.kz = (float)b->global_dims[2] / B.gz;
  B.offx = (float)(b->origin[0] - b->ghost_lo[0]); B.offy = (float)(b->origin[1] - b->ghost_lo[1]); B.offz = (float)(b->origin[2] - b->ghost_lo[2]);
  for (int a = 0; a < 3; ++a) { B.lo[a] = b->origin[a]; B.hi[a] = b->origin[a] + b->owned[a]; }
  B.nx = b->global_dims[0]; B.ny = b->global_dims[1]; B.nz = b->global_dims[2];
  if (p->count_samples) { int rc = vrb_counters_reset(c); if (rc != VRB_OK) return rc; }
  dim3 block(8, 8), grid((c->fw + 7) / 8, (c->fh + 7) / 8);
  size_t smem = (c->tf_n + 2 <= 1026) ? (size_t)(c->tf_n + 2) * sizeof(float4) : 0;
  const float diag = sqrtf(B.gx * B.gx + B.gy * B.gy + B.gz * B.gz);
  static const bool no_jump = getenv("VRB_BRICK_JUMP") && !strcmp(getenv("VRB_BRICK_JUMP"), "0");
  const int jump = (!no_jump && vrb_step_multiples_exact(p->step_size, diag)) ? 1 : 0;
#define VRB_BRICK_LAUNCH(CNT, MD) k_rc1pass_brick<CNT, MD><<<grid, block, smem, c->stream>>>(c->vol_view(), B, c->d_tf_rgbt, c->tf_n, \
      (float4*)c->d_partial, (float*)c->d_brick_alpha, front, c->fw, c->fh, make_cam_view(cam), p->step_size, jump, c->d_counter)
  if (mode == 0) { if (p->count_samples) VRB_BRICK_LAUNCH(true, 0); else VRB_BRICK_LAUNCH(false, 0); }
  else if (mode == 1) { if (p->count_samples) VRB_BRICK_LAUNCH(true, 1); else VRB_BRICK_LAUNCH(false, 1); }
  else { if (p->count_samples) VRB_BRICK_LAUNCH(true, 2); else VRB_BRICK_LAUNCH(false, 2); }
#undef VRB_BRICK_LAUNCH
  VRB_CUDA(cudaGetLastError());
  c->launches++;
  if (p->count_samples) return vrb_counters_fetch(c);
  return VRB_OK;
}

extern "C" int vrb_rc1pass_render_brick(vrb_ctx* c, const vrb_camera* cam, const vrb_rc1pass_params* p, const vrb_brick* b) {
  return render_brick(c, cam, p, b, 0, nullptr, 0);
}
extern "C" int vrb_rc1pass_brick_alpha(vrb_ctx* c, const vrb_camera* cam, const vrb_rc1pass_params* p, const vrb_brick* b) {
  return render_brick(c, cam, p, b, 1, nullptr, 0);
}
extern "C" int vrb_rc1pass_render_brick_exact(vrb_ctx* c, const vrb_camera* cam, const vrb_rc1pass_params* p, const vrb_brick* b,
                                              const void* const* front_alphas, int n_front) {
  return render_brick(c, cam, p, b, 2, front_alphas, n_front);
}
extern "C" int vrb_brick_alpha_device_ptr(vrb_ctx* c, void** dev) {
  VRB_REQUIRE(c && dev, VRB_ERR_INVALID, "vrb_brick_alpha_device_ptr: NULL argument");
  VRB_REQUIRE(c->d_brick_alpha, VRB_ERR_STATE, "vrb_brick_alpha_device_ptr: no brick rendered yet");
  *dev = c->d_brick_alpha;
  return VRB_OK;
}

extern "C" int vrb_partial_device_ptr(vrb_ctx* c, void** dev) {
  VRB_REQUIRE(c && dev, VRB_ERR_INVALID, "vrb_partial_device_ptr: NULL argument");
  VRB_REQUIRE(c->d_partial, VRB_ERR_STATE, "vrb_partial_device_ptr: no partial frame (vrb_rc1pass_render_brick)");
  *dev = c->d_partial;
  return VRB_OK;
}

struct PartialList { const float4* p[VRB_MAX_PARTIALS]; int n; };

// rows [row0, row0 + rows): dst = over(partials in order), 0.99 cut between segments (ray_marching_1p.comp:167)
// exact two-pass partials (MODE 2): colours add, alpha is the maximum; the order does not matter
__global__ void __launch_bounds__(256)
k_composite_sum(const __grid_constant__ PartialList L, FrameView fr, int row0, int rows) {
  const long long n = (long long)rows * fr.w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long px = (long long)row0 * fr.w + i;
    float dr = 0.f, dg = 0.f, db = 0.f, da = 0.f;
    for (int k = 0; k < L.n; ++k) {
      float4 s = L.p[k][px];
      dr += s.x; dg += s.y; db += s.z; da = fmaxf(da, s.w);
    }
    __half2 lo = __floats2half2_rn(dr, dg), hi = __floats2half2_rn(db, da);
    uint2 pk; pk.x = *reinterpret_cast<unsigned int*>(&lo); pk.y = *reinterpret_cast<unsigned int*>(&hi);
    reinterpret_cast<uint2*>(fr.rgba)[px] = pk;
  }
}

__global__ void __launch_bounds__(256)
k_composite_ordered(const __grid_constant__ PartialList L, FrameView fr, int row0, int rows) {
  const long long n = (long long)rows * fr.w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long px = (long long)row0 * fr.w + i;
    float dr = 0.f, dg = 0.f, db = 0.f, da = 0.f;
    for (int k = 0; k < L.n; ++k) {
      float4 s = L.p[k][px];                        // local HBM or a peer's (P2P load)
      if (s.w > 0.0f || s.x > 0.0f || s.y > 0.0f || s.z > 0.0f) {
        float om = 1.0f - da;
        dr = fmaf(om, s.x, dr); dg = fmaf(om, s.y, dg); db = fmaf(om, s.z, db); da = fmaf(om, s.w, da);
        if (da > 0.99f) break;
      }
    }
    __half2 lo = __floats2half2_rn(dr, dg), hi = __floats2half2_rn(db, da);
    uint2 pk; pk.x = *reinterpret_cast<unsigned int*>(&lo); pk.y = *reinterpret_cast<unsigned int*>(&hi);
    reinterpret_cast<uint2*>(fr.rgba)[px] = pk;
  }
}

static int composite(vrb_ctx* c, const void* const* partials, int n, int row0, int rows, bool ordered);
extern "C" int vrb_composite_ordered(vrb_ctx* c, const void* const* partials, int n, int row0, int rows) { return composite(c, partials, n, row0, rows, true); }
extern "C" int vrb_composite_sum(vrb_ctx* c, const void* const* partials, int n, int row0, int rows) { return composite(c, partials, n, row0, rows, false); }
static int composite(vrb_ctx* c, const void* const* partials, int n, int row0, int rows, bool ordered) {
  VRB_REQUIRE(c && partials, VRB_ERR_INVALID, "vrb_composite_ordered: NULL argument");
  VRB_REQUIRE(n >= 1 && n <= VRB_MAX_PARTIALS, VRB_ERR_INVALID, "vrb_composite_ordered: n = %d (1..%d)", n, VRB_MAX_PARTIALS);
  VRB_REQUIRE(c->d_frame, VRB_ERR_STATE, "vrb_composite_ordered: no frame");
  VRB_REQUIRE(row0 >= 0 && rows >= 0 && row0 + rows <= c->fh, VRB_ERR_INVALID, "vrb_composite_ordered: rows [%d, %d) of %d", row0, row0 + rows, c->fh);
  if (rows == 0) return VRB_OK;
  VRB_CUDA(cudaSetDevice(c->device));
  PartialList L; L.n = n;
  for (int i = 0; i < n; ++i) { VRB_REQUIRE(partials[i], VRB_ERR_INVALID, "vrb_composite_ordered: partial %d is NULL", i); L.p[i] = (const float4*)partials[i]; }
  const long long npx = (long long)rows * c->fw;
  if (ordered) k_composite_ordered<<<(int)std::min<long long>((npx + 255) / 256, 148 * 16), 256, 0, c->stream>>>(L, c->frame_view(), row0, rows);
  else         k_composite_sum<<<(int)std::min<long long>((npx + 255) / 256, 148 * 16), 256, 0, c->stream>>>(L, c->frame_view(), row0, rows);
  VRB_CUDA(cudaGetLastError());
  c->launches++;
  return VRB_OK;
}

extern "C" int vrb_ipc_export(vrb_ctx* c, const void* dev_ptr, unsigned char handle[64]) {
  VRB_REQUIRE(c && dev_ptr && handle, VRB_ERR_INVALID, "vrb_ipc_export: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  VRB_CUDA(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  VRB_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
  memcpy(handle, &h, 64);
  return VRB_OK;
}
extern "C" int vrb_ipc_import(vrb_ctx* c, const unsigned char handle[64], void** dev_ptr) {
  VRB_REQUIRE(c && handle && dev_ptr, VRB_ERR_INVALID, "vrb_ipc_import: NULL argument");
  VRB_CUDA(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  VRB_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return VRB_OK;
}
extern "C" int vrb_ipc_close(vrb_ctx* c, void* dev_ptr) {
  VRB_REQUIRE(c && dev_ptr, VRB_ERR_INVALID, "vrb_ipc_close: NULL argument");
  VRB_CUDA(cudaSetDevice(c->device));
  VRB_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return VRB_OK;
}
