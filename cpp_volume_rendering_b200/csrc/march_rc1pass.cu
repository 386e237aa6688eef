// march_rc1pass.cu -- single-pass ray casting for sm_100a.
// Replaces the dispatch of rc1pass/ray_marching_1p.comp (main at :85-179) made by RayCasting1Pass::Redraw
// (rc1prenderer.cpp:140-151).  One thread per pixel, 8x4 pixel tile per warp (two warps per 8x8 CTA, the reference's
// work-group shape, ray_marching_1p.comp:38) so that the 32 rays of a warp walk the same few cache lines.
#include "vrb_internal.cuh"
#include "march_list.cuh"
#include <cstdlib>
#include <cstring>

#define TF_SMEM_MAX 1024   // transfer functions up to this many texels are staged in shared memory
extern "C" int vrb_rc1pass_render(vrb_ctx* c, const vrb_camera* cam, const vrb_rc1pass_params* p);

// SKIP: result-preserving empty-space skipping over the occupancy cells of empty_space.cu.  A sample whose cell is
// flagged empty has src.a == 0 exactly, so only its fetches are dropped; the ray parameter still advances by the same
// fp32 additions (pixels and loop counts are those of the plain loop).  After an empty sample the loop fast-forwards to
// just before the analytic exit of the cell: floor indices are monotone in t along a ray, so checking that the LAST
// fast-forwarded sample is still in the cell proves all of them were (otherwise the fast-forward is undone).

// QUAD: the eight taps of a sample come from the 2x2 quad copy of the volume (two LDG.64, vrb_fetch_volume_quad) instead of
// eight LDG.U16: same texels, same blend (config 1: 0.294 -> 0.223 ms).  Two other variants were measured at config 1 in round 2
// and not kept: the next sample's taps issued before this sample's TF lookup and compositing (two samples in flight:
// 0.241 ms, the wasted fetch and the extra registers cost more than the overlap gains) and the persistent kernel of the lit
// renderers with refill (VRB_RC1_KERNEL=list: 0.422 ms).
template <bool TF_SMEM, bool COUNT, bool SKIP, bool HW, bool QUAD>
__global__ void __launch_bounds__(64)
k_rc1pass(VolView vol, const uint2* __restrict__ volq, const float4* __restrict__ tf_g, int tf_n, FrameView fr, CamView cam, PartView part,
          float step, unsigned long long* counter, CellView cells) {
  extern __shared__ float4 s_tf[];
  const float4* tf = tf_g;
  if (TF_SMEM) {
    for (int i = threadIdx.y * 8 + threadIdx.x; i < tf_n + 2; i += 64) s_tf[i] = tf_g[i];
    __syncthreads();
    tf = s_tf;
  }
  int px, py;
  vrb_cta_origin(part, fr.w, 8, 8, px, py);
  px += threadIdx.x; py += threadIdx.y;
  unsigned int ns = 0;
  if (px < fr.w && py < fr.h && vrb_owns_pixel(part, px, py, fr.w)) {
    Ray r = vrb_make_ray(cam, px, py, fr.w, fr.h, vol.gx, vol.gy, vol.gz);
    if (r.hit) {
      float D = fabsf(__fadd_rn(r.tfar, -r.tnear));
      // tex_pos = eye + dir*tnear + G/2
      float tx = __fadd_rn(__fadd_rn(r.ox, __fmul_rn(r.dx, r.tnear)), __fmul_rn(vol.gx, 0.5f));
      float ty = __fadd_rn(__fadd_rn(r.oy, __fmul_rn(r.dy, r.tnear)), __fmul_rn(vol.gy, 0.5f));
      float tz = __fadd_rn(__fadd_rn(r.oz, __fmul_rn(r.dz, r.tnear)), __fmul_rn(vol.gz, 0.5f));
      float kx = (float)vol.w / vol.gx, ky = (float)vol.h / vol.gy, kz = (float)vol.d / vol.gz;
      float dr = 0.f, dg = 0.f, db = 0.f, da = 0.f;
      float idx = 0.f, idy = 0.f, idz = 0.f, ikx = 0.f, iky = 0.f, ikz = 0.f;
      int bad_cell = -1;
      if (SKIP) {
        idx = 1.0f / r.dx; idy = 1.0f / r.dy; idz = 1.0f / r.dz;
        ikx = 1.0f / kx; iky = 1.0f / ky; ikz = 1.0f / kz;
      }
      for (float s = 0.0f; s < D;) {
        float h = fminf(step, __fadd_rn(D, -s));
        float t = __fadd_rn(s, __fmul_rn(h, 0.5f));
        const float qx = fmaf(r.dx, t, tx), qy = fmaf(r.dy, t, ty), qz = fmaf(r.dz, t, tz);
        int ix = 0, iy = 0, iz = 0; float fx = 0.f, fy = 0.f, fz = 0.f;
        if (SKIP || !HW) vrb_volume_coords(vol, kx, ky, kz, qx, qy, qz, ix, iy, iz, fx, fy, fz);
        if (SKIP) {
          const int cx = ix >> 3, cy = iy >> 3, cz = iz >> 3;
          const int cell = (cz * cells.ch + cy) * cells.cw + cx;
          if (!__ldg(cells.flags + cell)) {
            if (COUNT) ++ns;
            s = __fadd_rn(s, h);
            if (cell != bad_cell) {
              // ray parameter at which the padded index leaves [8c, 8c+8) on each axis (up = p*k + 0.5)
              float bx = (float)((r.dx > 0.0f) ? (cx * 8 + 8) : (cx * 8)), by = (float)((r.dy > 0.0f) ? (cy * 8 + 8) : (cy * 8));
              float bz = (float)((r.dz > 0.0f) ? (cz * 8 + 8) : (cz * 8));
              float ex = (r.dx != 0.0f) ? ((bx - 0.5f) * ikx - tx) * idx : 3.0e38f;
              float ey = (r.dy != 0.0f) ? ((by - 0.5f) * iky - ty) * idy : 3.0e38f;
              float ez = (r.dz != 0.0f) ? ((bz - 0.5f) * ikz - tz) * idz : 3.0e38f;
              const float t_safe = fminf(fminf(ex, ey), ez) - 0.02f * step;
              const float s0 = s; const unsigned int ns0 = ns;
              float t_last = -1.0f;
              while (s < D) {
                float h2 = fminf(step, __fadd_rn(D, -s));
                float t2 = __fadd_rn(s, __fmul_rn(h2, 0.5f));
                if (!(t2 < t_safe)) break;
                t_last = t2; s = __fadd_rn(s, h2);
                if (COUNT) ++ns;
              }
              if (t_last >= 0.0f) {
                int jx, jy, jz; float gx_, gy_, gz_;
                vrb_volume_coords(vol, kx, ky, kz, fmaf(r.dx, t_last, tx), fmaf(r.dy, t_last, ty), fmaf(r.dz, t_last, tz), jx, jy, jz, gx_, gy_, gz_);
                if ((jx >> 3) != cx || (jy >> 3) != cy || (jz >> 3) != cz) { s = s0; ns = ns0; bad_cell = cell; }
              }
            }
            continue;
          }
        }
        // HW: the texture unit's trilinear filter (texel centres at i + 0.5, clamp addressing), as the reference's sampler
        float density = HW ? tex3D<float>(vol.tex3d, qx * kx, qy * ky, qz * kz)
                           : (QUAD ? vrb_fetch_volume_quad(vol, volq, ix, iy, iz, fx, fy, fz) : vrb_fetch_volume(vol, ix, iy, iz, fx, fy, fz));
        float4 src = vrb_sample_tf(tf, tf_n, density);
        if (COUNT) ++ns;
        if (src.w > 0.0f) {
          float a = 1.0f - __expf(-src.w * h);
          float om = (1.0f - da) * a;       // (1 - dst.a) * src.a ; src.rgb is premultiplied by src.a first
          dr = fmaf(om, src.x, dr); dg = fmaf(om, src.y, dg); db = fmaf(om, src.z, db);
          da = da + om;
          if (da > 0.99f) break;
        }
        s = __fadd_rn(s, h);
      }
      vrb_store_pixel(fr, px, py, dr, dg, db, da);
    } else if (fr.zero_miss) vrb_store_pixel(fr, px, py, 0.f, 0.f, 0.f, 0.f);
  }
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) ns += __shfl_xor_sync(0xffffffffu, ns, o);
    if (((threadIdx.y * 8 + threadIdx.x) & 31) == 0 && ns) atomicAdd(counter, (unsigned long long)ns);
  }
}

// The same loop with ShadeBlinnPhong (ray_marching_1p.comp:48-81) on every non-transparent sample.  Its own kernel so that
// the default path keeps its register budget; compositing follows the shader literally (src.rgb = src.rgb * src.a, then
// dst = dst + (1 - dst.a) * src), like the lit marchers.
template <bool COUNT, bool HW>
__global__ void __launch_bounds__(64)
k_rc1pass_lit(VolView vol, const float4* __restrict__ tf_g, int tf_n, FrameView fr, CamView cam, PartView part,
              float step, unsigned long long* counter, const __grid_constant__ PhongView ph) {
  extern __shared__ float4 s_tf[];
  const float4* tf = tf_g;
  if (tf_n <= TF_SMEM_MAX) {
    for (int i = threadIdx.y * 8 + threadIdx.x; i < tf_n + 2; i += 64) s_tf[i] = tf_g[i];
    tf = s_tf;
  }
  __syncthreads();
  int px, py;
  vrb_cta_origin(part, fr.w, 8, 8, px, py);
  px += threadIdx.x; py += threadIdx.y;
  unsigned int ns = 0;
  if (px < fr.w && py < fr.h && vrb_owns_pixel(part, px, py, fr.w)) {
    Ray r = vrb_make_ray(cam, px, py, fr.w, fr.h, vol.gx, vol.gy, vol.gz);
    if (r.hit) {
      float D = fabsf(r.tfar - r.tnear);
      const float tx = (r.ox + r.dx * r.tnear) + (vol.gx * 0.5f), ty = (r.oy + r.dy * r.tnear) + (vol.gy * 0.5f), tz = (r.oz + r.dz * r.tnear) + (vol.gz * 0.5f);
      const float kx = (float)vol.w / vol.gx, ky = (float)vol.h / vol.gy, kz = (float)vol.d / vol.gz;
      float dr = 0.f, dg = 0.f, db = 0.f, da = 0.f;
      for (float s = 0.0f; s < D;) {
        float h = fminf(step, D - s);
        const float t = s + h * 0.5f;
        const float qx = tx + r.dx * t, qy = ty + r.dy * t, qz = tz + r.dz * t;
        float density = HW ? tex3D<float>(vol.tex3d, qx * kx, qy * ky, qz * kz) : vrb_sample_volume(vol, kx, ky, kz, qx, qy, qz);
        float4 src = vrb_sample_tf(tf, tf_n, density);
        if (COUNT) ++ns;
        if (src.w > 0.0f) {
          float dot_diff, spec;
          if (vrb_phong_terms(vol, ph, kx, ky, kz, qx, qy, qz, cam.ex, cam.ey, cam.ez, dot_diff, spec)) {
            const float kad = ph.ka + ph.kd * dot_diff;
            src.x = src.x * kad + ph.isx * ph.ks * spec;
            src.y = src.y * kad + ph.isy * ph.ks * spec;
            src.z = src.z * kad + ph.isz * ph.ks * spec;
          }
          float a = 1.0f - expf(-src.w * h);
          float om = 1.0f - da;
          dr = dr + om * (src.x * a); dg = dg + om * (src.y * a); db = db + om * (src.z * a); da = da + om * a;
          if (da > 0.99f) break;
        }
        s = s + h;
      }
      vrb_store_pixel(fr, px, py, dr, dg, db, da);
    } else if (fr.zero_miss) vrb_store_pixel(fr, px, py, 0.f, 0.f, 0.f, 0.f);
  }
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) ns += __shfl_xor_sync(0xffffffffu, ns, o);
    if (((threadIdx.y * 8 + threadIdx.x) & 31) == 0 && ns) atomicAdd(counter, (unsigned long long)ns);
  }
}

extern "C" int vrb_rc1pass_render_lit(vrb_ctx* c, const vrb_camera* cam, const vrb_rc1pass_params* p, const vrb_lighting* light) {
  VRB_REQUIRE(c && cam && p && light, VRB_ERR_INVALID, "vrb_rc1pass_render_lit: NULL argument");
  if (light->apply_phong != 1) return vrb_rc1pass_render(c, cam, p);
  VRB_REQUIRE(c->d_vol, VRB_ERR_STATE, "vrb_rc1pass_render_lit: no volume uploaded");
  VRB_REQUIRE(c->d_tf_rgbt, VRB_ERR_STATE, "vrb_rc1pass_render_lit: no transfer function uploaded");
  VRB_REQUIRE(c->d_frame, VRB_ERR_STATE, "vrb_rc1pass_render_lit: no frame (vrb_frame_resize)");
  VRB_REQUIRE(p->step_size > 0.0f, VRB_ERR_INVALID, "vrb_rc1pass_render_lit: step_size %g", p->step_size);
  PhongView ph;
  { int rc = vrb_make_phong_view(c, light, &ph, "vrb_rc1pass_render_lit"); if (rc != VRB_OK) return rc; }
  VRB_CUDA(cudaSetDevice(c->device));
  if (!c->d_frame_target) VRB_CUDA(cudaMemsetAsync(c->d_frame, 0, (size_t)c->fw * c->fh * 4 * sizeof(__half), c->stream));
  if (p->count_samples) { int rc = vrb_counters_reset(c); if (rc != VRB_OK) return rc; }
  PartView part;
  dim3 block(8, 8), grid = vrb_make_grid(c, 8, 8, &part);
  { int rc = vrb_vol_tex3d_prepare(c); if (rc != VRB_OK) return rc; }
  VolView vol = c->vol_view();
  vol.atlas = 0;                                  // the gradient taps use the linear layout; keep the density taps on it too
  const size_t smem_bytes = (c->tf_n <= TF_SMEM_MAX) ? (size_t)(c->tf_n + 2) * sizeof(float4) : 0;
#define VRB_RC1_LIT(N, H) k_rc1pass_lit<N, H><<<grid, block, smem_bytes, c->stream>>>(vol, c->d_tf_rgbt, c->tf_n, c->frame_view(), make_cam_view(cam), part, p->step_size, c->d_counter, ph)
  if (vol.tex3d) { if (p->count_samples) VRB_RC1_LIT(true, true); else VRB_RC1_LIT(false, true); }
  else           { if (p->count_samples) VRB_RC1_LIT(true, false); else VRB_RC1_LIT(false, false); }
#undef VRB_RC1_LIT
  VRB_CUDA(cudaGetLastError());
  c->launches++;
  if (p->count_samples) return vrb_counters_fetch(c);
  return VRB_OK;
}

extern "C" int vrb_rc1pass_render(vrb_ctx* c, const vrb_camera* cam, const vrb_rc1pass_params* p) {
  VRB_REQUIRE(c && cam && p, VRB_ERR_INVALID, "vrb_rc1pass_render: NULL argument");
  VRB_REQUIRE(c->d_vol, VRB_ERR_STATE, "vrb_rc1pass_render: no volume uploaded");
  VRB_REQUIRE(c->d_tf_rgbt, VRB_ERR_STATE, "vrb_rc1pass_render: no transfer function uploaded");
  VRB_REQUIRE(c->d_frame, VRB_ERR_STATE, "vrb_rc1pass_render: no frame (vrb_frame_resize)");
  VRB_REQUIRE(p->step_size > 0.0f, VRB_ERR_INVALID, "vrb_rc1pass_render: step_size %g", p->step_size);
  VRB_CUDA(cudaSetDevice(c->device));
  // RayCasting1Pass::Redraw: ClearTexture, then dispatch (misses keep the cleared 0)
  if (!c->d_frame_target) VRB_CUDA(cudaMemsetAsync(c->d_frame, 0, (size_t)c->fw * c->fh * 4 * sizeof(__half), c->stream));
  if (p->count_samples) { int rc = vrb_counters_reset(c); if (rc != VRB_OK) return rc; }
  // VRB_RC1_KERNEL=list: the persistent march kernel of the lit renderers with the compositing done in place (k_list_march
  // MODE 2, march_list.cu: lanes take the next ray of their tile when theirs ends).  Bit-identical frames and loop counts,
  // but measured SLOWER than one fixed ray per lane where neighbouring rays have similar lengths (config 1, V-gauss:
  // 0.422 ms against 0.223 ms, 16 against 29 active lanes per instruction), so k_rc1pass below stays the default.
  { const char* kern = getenv("VRB_RC1_KERNEL");
    if (c->filter_mode != VRB_FILTER_HARDWARE && kern && !strcmp(kern, "list")) {
      VrbKernelTimer timer(c, "k_list_march<rc1pass>");
      int rc = vrb_list_rc1pass(c, cam, p->step_size, p->skip_empty, p->count_samples);
      if (rc != VRB_OK) return rc;
      if (p->count_samples) return vrb_counters_fetch(c);
      return VRB_OK;
    } }
  PartView part;
  dim3 block(8, 8), grid = vrb_make_grid(c, 8, 8, &part);
  { int rc = vrb_vol_tex3d_prepare(c); if (rc != VRB_OK) return rc; }
  if (c->filter_mode != VRB_FILTER_HARDWARE) { int rc = vrb_vol_quads_prepare(c); if (rc != VRB_OK) return rc; }
  VolView vol = c->vol_view();
  FrameView fr = c->frame_view();
  CamView cv = make_cam_view(cam);
  bool smem = c->tf_n <= TF_SMEM_MAX;
  size_t smem_bytes = smem ? (size_t)(c->tf_n + 2) * sizeof(float4) : 0;
  CellView cells{nullptr, 0, 0, 0};
  if (p->skip_empty) {
    int rc = vrb_cells_prepare(c);
    if (rc != VRB_OK) return rc;
    cells.flags = c->d_cell_flags; cells.cw = c->cell_dims[0]; cells.ch = c->cell_dims[1]; cells.cd = c->cell_dims[2];
  }
#define VRB_RC1_LAUNCH(S, N, K, H, Q) k_rc1pass<S, N, K, H, Q><<<grid, block, smem_bytes, c->stream>>>(vol, c->d_vol_quad, c->d_tf_rgbt, c->tf_n, fr, cv, part, p->step_size, c->d_counter, cells)
#define VRB_RC1_LAUNCH2(S, N, K) do { if (vol.tex3d) VRB_RC1_LAUNCH(S, N, K, true, false); else if (c->d_vol_quad) VRB_RC1_LAUNCH(S, N, K, false, true); else VRB_RC1_LAUNCH(S, N, K, false, false); } while (0)
  const int variant = (smem ? 4 : 0) | (p->count_samples ? 2 : 0) | (p->skip_empty ? 1 : 0);
  VrbKernelTimer timer(c, "k_rc1pass");
  switch (variant) {
    case 0: VRB_RC1_LAUNCH2(false, false, false); break;
    case 1: VRB_RC1_LAUNCH2(false, false, true); break;
    case 2: VRB_RC1_LAUNCH2(false, true, false); break;
    case 3: VRB_RC1_LAUNCH2(false, true, true); break;
    case 4: VRB_RC1_LAUNCH2(true, false, false); break;
    case 5: VRB_RC1_LAUNCH2(true, false, true); break;
    case 6: VRB_RC1_LAUNCH2(true, true, false); break;
    default: VRB_RC1_LAUNCH2(true, true, true); break;
  }
#undef VRB_RC1_LAUNCH2
#undef VRB_RC1_LAUNCH
  VRB_CUDA(cudaGetLastError());
  c->launches++;
  if (p->count_samples) return vrb_counters_fetch(c);
  return VRB_OK;
}
