// march_list.cu -- the renderer-independent half of a deferred lit frame (shade_list.cuh, march_list.cuh), sm_100a.
//
// k_list_march is the primary-ray loop every lit shader of the reference shares (rc1pextbsd/ebs_ray_bbox_marching.comp:553-627,
// rc1pdosct/ray_bbox_marching.comp:658-734, rc1pvctsg/vct_ray_bbox_marching.comp:191-264, rc1pcrtgt/gt_ray_marching.comp:371-483)
// with ShadeSample taken out: the opacity a ray accumulates, hence where it stops, depends on the transfer function only.
// One lane per ray, 8x4 rays per warp.  What makes it fast:
//   * 2x2 texel QUADS of the padded fp16 volume (vrb_ctx::d_vol_quad): a trilinear footprint is two LDG.64 instead of
//     eight LDG.U16 -- the same texels, the same blend order, a quarter of the L1 requests;
//   * the opacity channel of the transfer function as its own float table in shared memory: a transparent sample costs
//     two LDS.32, the RGB texels are read only for samples that are appended;
//   * result-preserving empty-space skipping over the occupancy cells of empty_space.cu (switched on when at least 5 %
//     of the cells are empty under the current transfer function): a sample in an empty cell has alpha == 0 exactly, so
//     only its fetches are dropped; the ray parameter advances by the same fp32 additions and the loop counter counts it,
//     so pixels and loop counts are those of the plain loop.  Every sample looks up its own cell (one cached byte): no
//     per-lane fast-forward loop, the 32 rays of a warp stay converged through empty space;
//   * lanes never wait for each other to append (shade_list.cuh).
// k_list_composite walks a ray's entries in march order: dst += (1 - dst.a) * (rgb * a, a), operation by operation.
// Compiled with -fmad=false; every rounding that matters is spelled out anyway.
#include "march_list.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace {

struct RaySetup { float wx, wy, wz, dx, dy, dz, D; bool hit; };

// ray / AABB and entry point in texture space: wd = (eye + dir * tnear) + G / 2
template <bool GT>
__device__ __forceinline__ RaySetup ray_setup(const CamView& cam, const FrameView& fr, const VolView& vol, int px, int py) {
  RaySetup R;
  const Ray r = vrb_make_ray(cam, px, py, fr.w, fr.h, vol.gx, vol.gy, vol.gz);
  R.hit = r.hit;
  const float d = __fadd_rn(r.tfar, -r.tnear);
  R.D = GT ? d : fabsf(d);
  R.dx = r.dx; R.dy = r.dy; R.dz = r.dz;
  R.wx = __fadd_rn(__fadd_rn(r.ox, __fmul_rn(r.dx, r.tnear)), __fmul_rn(vol.gx, 0.5f));
  R.wy = __fadd_rn(__fadd_rn(r.oy, __fmul_rn(r.dy, r.tnear)), __fmul_rn(vol.gy, 0.5f));
  R.wz = __fadd_rn(__fadd_rn(r.oz, __fmul_rn(r.dz, r.tnear)), __fmul_rn(vol.gz, 0.5f));
  return R;
}

template <bool QUAD>
__device__ __forceinline__ float fetch_density(const VolView& v, const uint2* __restrict__ vq, int ix, int iy, int iz, float fx, float fy, float fz) {
  if (QUAD) return vrb_fetch_volume_quad(v, vq, ix, iy, iz, fx, fy, fz);
  VolView lin = v; lin.atlas = 0;
  return vrb_fetch_volume(lin, ix, iy, iz, fx, fy, fz);
}

__device__ __forceinline__ float round_h(float v) { return __half2float(__float2half_rn(v)); }

// Ray r (0..255) of a CTA's 16x16 pixel tile: 32 consecutive rays are an 8x4 patch (eight patches, 2 across, 4 down).
__device__ __forceinline__ void tile_pixel(int r, int& dx, int& dy) {
  const int sub = r >> 5, l = r & 31;
  dx = ((sub & 1) << 3) + (l & 7);
  dy = ((sub >> 1) << 2) + (l >> 3);
}

// threads of a march CTA: they share the 256 rays of one 16x16 pixel tile.  256 = one ray per thread to start with, the pool
// then balances early terminations.  Measured at config 3 on one B200: 64 threads 1.87 ms, 128 threads 1.03 ms, 256 threads
// 0.71 ms (a tile's serial work is the kernel's critical path; on an eighth of the frame, a rank of an 8-GPU run, even more so)
#define MARCH_THREADS 256
#define MARCH_THREADS_MAX 256

// MODE 0: lit shaders' loop, visible samples go to the shading list.  MODE 1: the same with rc1pcrtgt's fp16 state round
// trips.  MODE 2: rc1pass (ray_marching_1p.comp:85-179): no list, the sample is composited in place with k_rc1pass's own
// arithmetic (march_rc1pass.cu: fmaf positions, __expf, fmaf compositing) and the pixel is stored when the ray ends.
template <int MODE, bool SKIP, bool QUAD>
__global__ void __launch_bounds__(MARCH_THREADS_MAX)
k_list_march(VolView vol, const uint2* __restrict__ volq, const float4* __restrict__ tf_g, int tf_n, FrameView fr, CamView cam, PartView part,
             float step, ShadeListView L, CellView cells, int refill_min, int count, unsigned long long* counter) {
  constexpr bool GT = MODE == 1, DIRECT = MODE == 2;
  extern __shared__ float4 s_tf[];               // tf_n + 2 RGBA texels, then tf_n + 2 extinction floats
  const int tid = threadIdx.x;
  const bool tf_smem = tf_n + 2 <= 1026;
  float* s_tfw = reinterpret_cast<float*>(s_tf + (tf_n + 2));
  __shared__ unsigned long long s_chunk[MARCH_THREADS_MAX / 32];      // the warps' current chunks (shade_list.cuh)
  __shared__ unsigned s_next;                    // next ray of the tile nobody has taken yet
  if (tf_smem) {
    for (int i = tid; i < tf_n + 2; i += (int)blockDim.x) { const float4 t = tf_g[i]; s_tf[i] = t; s_tfw[i] = t.w; }
  }
  if (tid < MARCH_THREADS_MAX / 32) s_chunk[tid] = (unsigned long long)VRB_SL_CHUNK;
  if (tid == 0) s_next = 0u;
  __syncthreads();
  const float4* tf = tf_smem ? s_tf : tf_g;
  const unsigned lane = (unsigned)(tid & 31);
  unsigned long long* wc = &s_chunk[tid >> 5];
  int px0, py0;
  vrb_cta_origin(part, fr.w, 16, 16, px0, py0);
  const float kx = (float)vol.w / vol.gx, ky = (float)vol.h / vol.gy, kz = (float)vol.d / vol.gz;
  // ---- per-lane ray state.  A lane that finishes its ray takes the next ray of the tile: the rays of a tile stop after
  // very different numbers of steps (early termination behind dense noise), and a warp that waited for its longest ray
  // ran at 10 of 32 lanes (profiles/r2_v4_cfg3_k_list_march.txt)
  RaySetup R;
  R.hit = false; R.D = 0.0f; R.wx = R.wy = R.wz = R.dx = R.dy = R.dz = 0.0f;
  bool have = false, more = true;
  float s = 0.0f, da = 0.0f;
  float dr = 0.0f, dg = 0.0f, db = 0.0f;         // MODE 2: the running colour
  unsigned ns = 0, nw = 0, last = VRB_SL_NONE, first = VRB_SL_NONE;
  int pix = 0;
  // Empty-space state: samples with t < t_safe lie in a cell already found empty and are skipped without looking at the
  // volume.  On each axis the padded index u = p k + 0.5 (p = w + d t) reaches the cell face b at t = (b - 0.5 - w k) / (d k);
  // the sample loop computes u(t) with an fp32 error far below INDEX_EPS (|u| <= 4098, a handful of roundings: < 2e-3), and so
  // does this inversion, hence every sample with t < b * ia + ca, ia = 1 / (d k), ca = (-0.5 - w k) ia - INDEX_EPS |ia|, has its
  // floor index inside the cell whatever the ray direction (for a ray nearly parallel to a face the margin grows as 1 / |d|
  // and simply disables the shortcut near that face).  ia, ca are per-ray constants.  tests/test_skip_margin_cpu.py replays this
  // arithmetic in numpy fp32 on random rays: no skipped sample leaves its cell (a margin of 0 does produce violations, 1e-4
  // already none; INDEX_EPS is 78 times that).
  const float INDEX_EPS = 0.0078125f;
  float t_safe = -3.0e38f;
  float iax = 0.f, iay = 0.f, iaz = 0.f, cax = 3.0e38f, cay = 3.0e38f, caz = 3.0e38f;
  for (;;) {
    const unsigned idle = __ballot_sync(0xffffffffu, !have);
    if (more && (__popc(idle) >= refill_min || idle == 0xffffffffu)) {
      unsigned base = 0;
      if (lane == 0) base = atomicAdd(&s_next, (unsigned)__popc(idle));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base + (unsigned)__popc(idle) >= 256u) more = false;
      if (!have) {
        const unsigned r = base + (unsigned)__popc(idle & ((1u << lane) - 1u));
        if (r < 256u) {
          int dx, dy;
          tile_pixel((int)r, dx, dy);
          const int px = px0 + dx, py = py0 + dy;
          if (px < fr.w && py < fr.h && vrb_owns_pixel(part, px, py, fr.w)) {
            pix = py * fr.w + px;
            R = ray_setup<GT>(cam, fr, vol, px, py);
            if (R.hit) {
              have = true;
              s = 0.0f; da = 0.0f; first = last = VRB_SL_NONE;
              dr = dg = db = 0.0f;
              t_safe = -3.0e38f;
              if (SKIP) {
                iax = iay = iaz = 0.0f; cax = cay = caz = 3.0e38f;
                if (R.dx != 0.0f) { iax = 1.0f / (R.dx * kx); cax = (-0.5f - R.wx * kx) * iax - INDEX_EPS * fabsf(iax); }
                if (R.dy != 0.0f) { iay = 1.0f / (R.dy * ky); cay = (-0.5f - R.wy * ky) * iay - INDEX_EPS * fabsf(iay); }
                if (R.dz != 0.0f) { iaz = 1.0f / (R.dz * kz); caz = (-0.5f - R.wz * kz) * iaz - INDEX_EPS * fabsf(iaz); }
              }
            } else if (DIRECT) { if (fr.zero_miss) vrb_store_pixel(fr, px, py, 0.f, 0.f, 0.f, 0.f); }
            else L.head[pix] = VRB_SL_NONE;
          }
        }
      }
    }
    if (!__any_sync(0xffffffffu, have)) { if (more) continue; break; }
    if (!have) continue;
    // ---- one loop iteration of the shader's `for (s = 0; s < D;)` for this lane's ray
    const bool in_range = s < R.D;
    float h = 0.0f, t = 0.0f;
    if (in_range) { h = fminf(step, __fadd_rn(R.D, -s)); t = __fadd_rn(s, __fmul_rn(h, 0.5f)); }
    bool done = !in_range;
    if (in_range) {
      if (SKIP && t < t_safe) {
        // still inside a cell known to be empty: alpha == 0 exactly, only the loop bookkeeping remains.  Up to 8 such steps per
        // iteration of the warp, so that the lanes crossing empty space keep pace with the lanes that fetch
#pragma unroll 1
        for (int k = 0; k < 8; ++k) {
          ++ns;
          float s1 = __fadd_rn(s, h);
          if (GT && s1 < R.D) { s1 = round_h(s1); if (!(s1 > s)) { done = true; break; } }
          s = s1;
          if (!(s < R.D)) { done = true; break; }
          h = fminf(step, __fadd_rn(R.D, -s)); t = __fadd_rn(s, __fmul_rn(h, 0.5f));
          if (!(t < t_safe)) break;
        }
      } else {
        // sample position: the lit shaders round the product and the sum separately, k_rc1pass fuses them (as its oracle does)
        const float qx = DIRECT ? fmaf(R.dx, t, R.wx) : __fadd_rn(R.wx, __fmul_rn(R.dx, t));
        const float qy = DIRECT ? fmaf(R.dy, t, R.wy) : __fadd_rn(R.wy, __fmul_rn(R.dy, t));
        const float qz = DIRECT ? fmaf(R.dz, t, R.wz) : __fadd_rn(R.wz, __fmul_rn(R.dz, t));
        int ix, iy, iz; float fx, fy, fz;
        vrb_volume_coords(vol, kx, ky, kz, qx, qy, qz, ix, iy, iz, fx, fy, fz);
        bool fetch = true;
        if (SKIP) fetch = __ldg(cells.flags + (((iz >> 3) * cells.ch + (iy >> 3)) * cells.cw + (ix >> 3))) != 0;
        ++ns;
        if (!fetch) {
          // Ray parameter up to which the samples provably stay in this cell: per axis t_axis = b * ia + ca (see the ray
          // set-up above), b = the cell face the ray moves towards in padded index units
          const float bx = (float)((ix & ~7) + ((R.dx > 0.0f) ? 8 : 0)), by = (float)((iy & ~7) + ((R.dy > 0.0f) ? 8 : 0));
          const float bz = (float)((iz & ~7) + ((R.dz > 0.0f) ? 8 : 0));
          t_safe = fminf(fminf(fmaf(bx, iax, cax), fmaf(by, iay, cay)), fmaf(bz, iaz, caz));
        } else {
          const float density = fetch_density<QUAD>(vol, volq, ix, iy, iz, fx, fy, fz);
          // vrb_sample_tf's .w alone
          float up = fmaf(density, (float)tf_n, 0.5f);
          up = fminf(fmaxf(up, 0.0f), (float)tf_n + 0.5f);
          float fl; const int ti = vrb_floor_pos(up, &fl);
          const float tfrac = up - fl;
          const float tau = tf_smem ? vrb_lerp(s_tfw[ti], s_tfw[ti + 1], tfrac) : vrb_lerp(tf_g[ti].w, tf_g[ti + 1].w, tfrac);
          if (tau > 0.0f) {
            if (DIRECT) {
              // k_rc1pass's compositing, operation by operation (march_rc1pass.cu)
              const float4 t0 = tf[ti], t1 = tf[ti + 1];
              const float a = __fadd_rn(1.0f, -__expf(-__fmul_rn(tau, h)));
              const float om = __fmul_rn(__fadd_rn(1.0f, -da), a);
              dr = fmaf(om, vrb_lerp(t0.x, t1.x, tfrac), dr); dg = fmaf(om, vrb_lerp(t0.y, t1.y, tfrac), dg); db = fmaf(om, vrb_lerp(t0.z, t1.z, tfrac), db);
              da = __fadd_rn(da, om);
              done = da > 0.99f;
            } else {
            const float a = __fadd_rn(1.0f, -expf(-__fmul_rn(tau, h)));
            const unsigned e = sl_push(L, wc, lane);
            if (e != VRB_SL_NONE) {
              const float4 t0 = tf[ti], t1 = tf[ti + 1];
              L.a[e] = make_float4(qx, qy, qz, __int_as_float(pix));
              L.b[e] = make_float4(vrb_lerp(t0.x, t1.x, tfrac), vrb_lerp(t0.y, t1.y, tfrac), vrb_lerp(t0.z, t1.z, tfrac), a);
              L.next[e] = VRB_SL_NONE;
              if (last == VRB_SL_NONE) first = e; else L.next[last] = e;
              last = e;
              ++nw;
            }
            const float om = __fadd_rn(1.0f, -da);
            da = __fadd_rn(da, __fmul_rn(om, a));
            done = da > 0.99f;                        // tested on the fp32 value of this dispatch (gt_ray_marching.comp:449-455) ...
            if (GT) da = round_h(da);                 // ... then stored: OutputFrag is rgba16f, re-read by the next dispatch
            }
          }
        }
        if (!done) {
          // s = s + h; rc1pcrtgt: if (!(s < D)) stop, else the state image rounds s to fp16 (a ray whose s no longer grows stops)
          float s1 = __fadd_rn(s, h);
          if (GT && s1 < R.D) { s1 = round_h(s1); if (!(s1 > s)) done = true; }
          s = s1;
        }
      }
    }
    if (done) {
      if (DIRECT) vrb_store_pixel(fr, pix % fr.w, pix / fr.w, dr, dg, db, da);
      else L.head[pix] = first;
      have = false;
    }
  }
  __syncwarp();
  for (int o = 16; o > 0; o >>= 1) { ns += __shfl_xor_sync(0xffffffffu, ns, o); nw += __shfl_xor_sync(0xffffffffu, nw, o); }
  if (lane == 0) {
    if (count && ns) atomicAdd(counter, (unsigned long long)ns);
    if (!DIRECT && nw) atomicAdd(&L.counters[1], nw);
  }
}

// one lane per ray, same grid as k_list_march: front-to-back over the ray's entries in march order
template <bool GT>
__global__ void __launch_bounds__(64)
k_list_composite(FrameView fr, CamView cam, PartView part, float gx, float gy, float gz, ShadeListView L) {
  int px, py;
  vrb_cta_origin(part, fr.w, 8, 8, px, py);
  px += threadIdx.x; py += threadIdx.y;
  const bool mine = px < fr.w && py < fr.h && vrb_owns_pixel(part, px, py, fr.w);
  bool hit = false;
  if (mine) hit = vrb_make_ray(cam, px, py, fr.w, fr.h, gx, gy, gz).hit;
  float dr = 0.f, dg = 0.f, db = 0.f, da = 0.f;
  unsigned e = hit ? L.head[py * fr.w + px] : VRB_SL_NONE;
  while (e != VRB_SL_NONE) {
    const float4 o = L.b[e];
    e = L.next[e];
    const float om = __fadd_rn(1.0f, -da);
    dr = __fadd_rn(dr, __fmul_rn(om, __fmul_rn(o.x, o.w))); dg = __fadd_rn(dg, __fmul_rn(om, __fmul_rn(o.y, o.w)));
    db = __fadd_rn(db, __fmul_rn(om, __fmul_rn(o.z, o.w))); da = __fadd_rn(da, __fmul_rn(om, o.w));
    if (GT) { dr = round_h(dr); dg = round_h(dg); db = round_h(db); da = round_h(da); }
  }
  if (hit) vrb_store_pixel(fr, px, py, dr, dg, db, da);
  else if (mine && fr.zero_miss) vrb_store_pixel(fr, px, py, 0.f, 0.f, 0.f, 0.f);
}

}  // namespace

void vrb_free_vol_quads(vrb_ctx* c) {
  if (c->d_vol_quad) cudaFree(c->d_vol_quad);
  c->d_vol_quad = nullptr;
  c->vol_quad_tried = false;
}

// Builds the quad copy of the volume on first use.  It costs four times the fp16 volume: skipped (the marcher then reads
// the linear layout) when that is more than VRB_VOL_QUADS_MAX_GB (default 24) or than a third of the free device memory.
int vrb_vol_quads_prepare(vrb_ctx* c) {
  if (c->d_vol_quad || c->vol_quad_tried) return VRB_OK;
  c->vol_quad_tried = true;
  if (const char* e = getenv("VRB_VOL_QUADS")) if (!strcmp(e, "0")) return VRB_OK;
  const size_t np = (size_t)(c->vw + 2) * (c->vh + 2) * (c->vd + 2);
  const size_t bytes = np * sizeof(uint2);
  double max_gb = 24.0;
  if (const char* e = getenv("VRB_VOL_QUADS_MAX_GB")) max_gb = atof(e);
  size_t free_b = 0, total_b = 0;
  VRB_CUDA(cudaMemGetInfo(&free_b, &total_b));
  if ((double)bytes > max_gb * 1e9 || bytes > free_b / 3) return VRB_OK;
  if (cudaMalloc(&c->d_vol_quad, bytes) != cudaSuccess) { cudaGetLastError(); c->d_vol_quad = nullptr; return VRB_OK; }
  vrb_build_quads(c->d_vol, c->d_vol_quad, c->vw + 2, c->vh + 2, c->vd + 2, c->stream);
  VRB_CUDA(cudaGetLastError());
  c->launches++;
  return VRB_OK;
}

template <int MODE, bool SKIP>
static void march_launch(vrb_ctx* c, dim3 grid, size_t smem, const CamView& cv, const PartView& part, float step, const ShadeListView& L,
                         const CellView& cells, int count) {
  // Lanes idle before the warp takes new rays.  Long rays of very different lengths (volumes with empty space: skipping on)
  // want early refills (8); where every ray stops after a few samples (dense data) a warp keeps whole 8x4 patches (32):
  // the march is cheap there anyway, and the entries of a chunk stay neighbours, which the SAT gathers of k_ebs_shade
  // feel (config 2: 8.5 ms with whole patches, 9.8 ms with early refills)
  static const int refill_env = getenv("VRB_MARCH_REFILL") ? atoi(getenv("VRB_MARCH_REFILL")) : 0;
  const int refill_min = refill_env > 0 ? std::min(refill_env, 32) : (SKIP ? 8 : 32);
  static const int threads_env = getenv("VRB_MARCH_THREADS") ? atoi(getenv("VRB_MARCH_THREADS")) : 0;
  const int threads = (threads_env == 32 || threads_env == 64 || threads_env == 128 || threads_env == 256) ? threads_env
                      : MARCH_THREADS;
  if (c->d_vol_quad)
    k_list_march<MODE, SKIP, true><<<grid, threads, smem, c->stream>>>(c->vol_view(), c->d_vol_quad, c->d_tf_rgbt, c->tf_n, c->frame_view(), cv, part, step, L, cells, refill_min, count, c->d_counter);
  else
    k_list_march<MODE, SKIP, false><<<grid, threads, smem, c->stream>>>(c->vol_view(), nullptr, c->d_tf_rgbt, c->tf_n, c->frame_view(), cv, part, step, L, cells, refill_min, count, c->d_counter);
}

// march -> (the host learns the list size; a list that was too small is enlarged and the march repeated)
int vrb_list_march(vrb_ctx* c, const vrb_camera* cam, float step, int flags, int count_samples, ListFrame* out) {
  PartView part;
  const dim3 grid = vrb_make_grid(c, 16, 16, &part);      // a CTA shares the 256 rays of a 16x16 pixel tile
  const size_t smem = (c->tf_n + 2 <= 1026) ? (size_t)(c->tf_n + 2) * (sizeof(float4) + sizeof(float)) : 0;
  const CamView cv = make_cam_view(cam);
  int rc = vrb_vol_quads_prepare(c);
  if (rc != VRB_OK) return rc;
  CellView cells{nullptr, 0, 0, 0};
  bool skip = true;
  if (const char* e = getenv("VRB_LIST_SKIP")) skip = strcmp(e, "0") != 0;
  if (skip) {
    rc = vrb_cells_prepare(c);
    if (rc != VRB_OK) return rc;
    const char* e = getenv("VRB_LIST_SKIP");
    if (c->cell_empty_fraction >= 0.05f || (e && !strcmp(e, "force"))) {
      cells.flags = c->d_cell_flags; cells.cw = c->cell_dims[0]; cells.ch = c->cell_dims[1]; cells.cd = c->cell_dims[2];
    } else skip = false;
  }
  const bool gt = (flags & VRB_LIST_GT) != 0;
  static const bool trace = getenv("VRB_TRACE") != nullptr;     // stage times of the march on stderr (debugging aid)
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  if (trace) { for (auto& e : ev) cudaEventCreate(&e); cudaEventRecord(ev[0], c->stream); }
  for (int attempt = 0; ; ++attempt) {
    rc = vrb_sl_begin(c, (unsigned)c->fw * (unsigned)c->fh, &out->L);
    if (rc != VRB_OK) return rc;
    if (count_samples) { rc = vrb_counters_reset(c); if (rc != VRB_OK) return rc; }
    if (trace) cudaEventRecord(ev[1], c->stream);
    if (gt) { if (skip) march_launch<1, true>(c, grid, smem, cv, part, step, out->L, cells, count_samples); else march_launch<1, false>(c, grid, smem, cv, part, step, out->L, cells, count_samples); }
    else    { if (skip) march_launch<0, true>(c, grid, smem, cv, part, step, out->L, cells, count_samples); else march_launch<0, false>(c, grid, smem, cv, part, step, out->L, cells, count_samples); }
    VRB_CUDA(cudaGetLastError());
    c->launches++;
    if (trace) cudaEventRecord(ev[2], c->stream);
    bool overflow = false;
    rc = vrb_sl_counts(c, &out->n_entries, &overflow);
    if (rc != VRB_OK) return rc;
    if (trace) {
      float a = 0.f, b = 0.f;
      cudaEventElapsedTime(&a, ev[0], ev[1]); cudaEventElapsedTime(&b, ev[1], ev[2]);
      fprintf(stderr, "[vrb trace] list begin (memsets) %.3f ms, march %.3f ms, slots %u (capacity %u, written %u)%s\n", a, b, out->n_entries,
              c->sl_capacity, c->sl_last_chunks, overflow ? " OVERFLOW: marching again" : "");
    }
    if (!overflow) break;
    VRB_REQUIRE(attempt < 2, VRB_ERR_CUDA, "deferred frame: the shading list overflowed twice");
  }
  if (trace) for (auto& e : ev) cudaEventDestroy(e);
  return VRB_OK;
}

// rc1pass through the same persistent kernel (MODE 2): called by vrb_rc1pass_render in exact filter mode.  Counters are
// reset / fetched by the caller.
int vrb_list_rc1pass(vrb_ctx* c, const vrb_camera* cam, float step, int skip_requested, int count_samples) {
  PartView part;
  const dim3 grid = vrb_make_grid(c, 16, 16, &part);
  const size_t smem = (c->tf_n + 2 <= 1026) ? (size_t)(c->tf_n + 2) * (sizeof(float4) + sizeof(float)) : 0;
  const CamView cv = make_cam_view(cam);
  int rc = vrb_vol_quads_prepare(c);
  if (rc != VRB_OK) return rc;
  CellView cells{nullptr, 0, 0, 0};
  bool skip = true;
  if (const char* e = getenv("VRB_LIST_SKIP")) skip = strcmp(e, "0") != 0;
  if (skip || skip_requested) {
    rc = vrb_cells_prepare(c);
    if (rc != VRB_OK) return rc;
    if (skip_requested || c->cell_empty_fraction >= 0.05f) {
      cells.flags = c->d_cell_flags; cells.cw = c->cell_dims[0]; cells.ch = c->cell_dims[1]; cells.cd = c->cell_dims[2];
      skip = true;
    } else skip = false;
  }
  ShadeListView none{};
  if (skip) march_launch<2, true>(c, grid, smem, cv, part, step, none, cells, count_samples);
  else      march_launch<2, false>(c, grid, smem, cv, part, step, none, cells, count_samples);
  VRB_CUDA(cudaGetLastError());
  c->launches++;
  return VRB_OK;
}

int vrb_list_composite(vrb_ctx* c, const vrb_camera* cam, int flags, const ListFrame& f) {
  PartView part;
  const dim3 grid = vrb_make_grid(c, 8, 8, &part);
  const CamView cv = make_cam_view(cam);
  const VolView v = c->vol_view();
  if (flags & VRB_LIST_GT) k_list_composite<true><<<grid, dim3(8, 8), 0, c->stream>>>(c->frame_view(), cv, part, v.gx, v.gy, v.gz, f.L);
  else                     k_list_composite<false><<<grid, dim3(8, 8), 0, c->stream>>>(c->frame_view(), cv, part, v.gx, v.gy, v.gz, f.L);
  VRB_CUDA(cudaGetLastError());
  c->launches++;
  return VRB_OK;
}
