// march_obj.cu -- object-space light cache (PreIlluminationStructuredVolume, cppvolrend/utils/preillumination.cpp:7-80)
// and the march that reads it: _common_shaders/obj_ray_marching.comp, active #else branch (:210-333), which the DOS /
// EBS / VCT renderers dispatch instead of their own marcher while the cache is active (dosrcrenderer.cpp:134-147).
// The cache is RG16F (Iocc, Ishadow), GL_LINEAR, clamp-to-edge: stored here as half2 padded by one replicated texel.
// Compiled with -fmad=false (oracle operation order), blends are explicit fmaf().
#include "vrb_internal.cuh"

void vrb_free_light_cache(vrb_ctx* c) {
  if (c->d_light_cache) cudaFree(c->d_light_cache);
  c->d_light_cache = nullptr;
  c->lc_dims[0] = c->lc_dims[1] = c->lc_dims[2] = 0;
}

int vrb_light_cache_alloc(vrb_ctx* c, int rw, int rh, int rd) {
  if (c->d_light_cache && c->lc_dims[0] == rw && c->lc_dims[1] == rh && c->lc_dims[2] == rd) return VRB_OK;
  vrb_free_light_cache(c);
  VRB_CUDA(cudaMalloc(&c->d_light_cache, (size_t)(rw + 2) * (rh + 2) * (rd + 2) * sizeof(__half2)));
  c->lc_dims[0] = rw; c->lc_dims[1] = rh; c->lc_dims[2] = rd;
  return VRB_OK;
}

__global__ void k_light_cache_pad(__half2* __restrict__ lc, int w, int h, int d) {
  const int pw = w + 2, ph = h + 2, pd = d + 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pw * ph * pd) return;
  const int x = i % pw, y = (i / pw) % ph, z = i / (pw * ph);
  if (x >= 1 && x <= w && y >= 1 && y <= h && z >= 1 && z <= d) return;
  const int sx = min(max(x, 1), w), sy = min(max(y, 1), h), sz = min(max(z, 1), d);
  lc[i] = lc[(size_t)sx + (size_t)pw * ((size_t)sy + (size_t)ph * (size_t)sz)];
}

void vrb_light_cache_finish(vrb_ctx* c) {
  const int n = (c->lc_dims[0] + 2) * (c->lc_dims[1] + 2) * (c->lc_dims[2] + 2);
  k_light_cache_pad<<<(n + 255) / 256, 256, 0, c->stream>>>(c->d_light_cache, c->lc_dims[0], c->lc_dims[1], c->lc_dims[2]);
}

__global__ void k_light_cache_unpad(const __half2* __restrict__ lc, float2* __restrict__ out, int w, int h, int d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * h * d) return;
  const int x = i % w, y = (i / w) % h, z = i / (w * h);
  out[i] = __half22float2(lc[(size_t)(x + 1) + (size_t)(w + 2) * ((size_t)(y + 1) + (size_t)(h + 2) * (size_t)(z + 1))]);
}

extern "C" int vrb_light_cache_read(vrb_ctx* c, float* host_out_rg, int dims_out[3]) {
  VRB_REQUIRE(c, VRB_ERR_INVALID, "vrb_light_cache_read: NULL context");
  VRB_REQUIRE(c->d_light_cache, VRB_ERR_STATE, "vrb_light_cache_read: no light cache");
  if (dims_out) { dims_out[0] = c->lc_dims[0]; dims_out[1] = c->lc_dims[1]; dims_out[2] = c->lc_dims[2]; }
  if (!host_out_rg) return VRB_OK;
  VRB_CUDA(cudaSetDevice(c->device));
  const int n = c->lc_dims[0] * c->lc_dims[1] * c->lc_dims[2];
  float2* tmp = nullptr;
  VRB_CUDA(cudaMalloc(&tmp, (size_t)n * sizeof(float2)));
  k_light_cache_unpad<<<(n + 255) / 256, 256, 0, c->stream>>>(c->d_light_cache, tmp, c->lc_dims[0], c->lc_dims[1], c->lc_dims[2]);
  cudaError_t e = cudaMemcpyAsync(host_out_rg, tmp, (size_t)n * sizeof(float2), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(tmp);
  VRB_CUDA(e);
  c->launches++;
  return VRB_OK;
}

struct LcView { const __half2* tex; int w, h, d; };

// texture(TexVolumeLightCache, tx_pos / VolumeScaledSizes).rg: u = s*N - 0.5, floor, clamp-to-edge (replicated border)
__device__ __forceinline__ float2 lc_fetch(const LcView& L, float sx, float sy, float sz) {
  float ux = sx * (float)L.w - 0.5f, uy = sy * (float)L.h - 0.5f, uz = sz * (float)L.d - 0.5f;
  float flx = floorf(ux), fly = floorf(uy), flz = floorf(uz);
  float fx = ux - flx, fy = uy - fly, fz = uz - flz;
  int ix = min(max((int)flx, -1), L.w - 1), iy = min(max((int)fly, -1), L.h - 1), iz = min(max((int)flz, -1), L.d - 1);
  const int pw = L.w + 2, slice = pw * (L.h + 2);
  const __half2* p = L.tex + (iz + 1) * slice + (iy + 1) * pw + (ix + 1);
  float2 a0 = __half22float2(__ldg(p)), a1 = __half22float2(__ldg(p + 1));
  float2 b0 = __half22float2(__ldg(p + pw)), b1 = __half22float2(__ldg(p + pw + 1));
  float2 c0 = __half22float2(__ldg(p + slice)), c1 = __half22float2(__ldg(p + slice + 1));
  float2 d0 = __half22float2(__ldg(p + slice + pw)), d1 = __half22float2(__ldg(p + slice + pw + 1));
  float2 o;
  o.x = vrb_lerp(vrb_lerp(vrb_lerp(a0.x, a1.x, fx), vrb_lerp(b0.x, b1.x, fx), fy), vrb_lerp(vrb_lerp(c0.x, c1.x, fx), vrb_lerp(d0.x, d1.x, fx), fy), fz);
  o.y = vrb_lerp(vrb_lerp(vrb_lerp(a0.y, a1.y, fx), vrb_lerp(b0.y, b1.y, fx), fy), vrb_lerp(vrb_lerp(c0.y, c1.y, fx), vrb_lerp(d0.y, d1.y, fx), fy), fz);
  return o;
}

// PHONG: ApplyPhongShading == 1, the gradient Blinn-Phong branch of ShadeSample (obj_ray_marching.comp:236-257).
template <bool COUNT, bool PHONG>
__global__ void __launch_bounds__(64)
k_obj_march(VolView vol, const float4* __restrict__ tf_g, int tf_n, FrameView fr, CamView cam, PartView part, LcView lc,
            vrb_obj_params P, float Kambient, float Kdiffuse, const __grid_constant__ PhongView ph, unsigned long long* counter) {
  extern __shared__ float4 s_tf[];
  const float4* tf = tf_g;
  if (tf_n + 2 <= 1026) {
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
      const float D = fabsf(r.tfar - r.tnear);
      float wx = r.ox + r.dx * r.tnear, wy = r.oy + r.dy * r.tnear, wz = r.oz + r.dz * r.tnear;
      wx = wx + vol.gx * 0.5f; wy = wy + vol.gy * 0.5f; wz = wz + vol.gz * 0.5f;
      const float kx = (float)vol.w / vol.gx, ky = (float)vol.h / vol.gy, kz = (float)vol.d / vol.gz;
      const bool Shade = P.apply_occlusion == 1 || P.apply_shadow == 1;
      float cr = 0.f, cg = 0.f, cb = 0.f, ca = 0.f;
      for (float s = 0.0f; s < D;) {
        float h = fminf(P.step_size, D - s);
        float t = s + h * 0.5f;
        float tx = wx + r.dx * t, ty = wy + r.dy * t, tz = wz + r.dz * t;
        float density = vrb_sample_volume(vol, kx, ky, kz, tx, ty, tz);
        float4 src = vrb_sample_tf(tf, tf_n, density);
        if (COUNT) ++ns;
        if (src.w > 0.0f && Shade) {
          float2 IaIs = lc_fetch(lc, tx / vol.gx, ty / vol.gy, tz / vol.gz);
          float ka = 0.0f, kd = 0.0f, ks = 0.0f, IOcc = 0.0f, ISdw = 0.0f;
          if (P.apply_occlusion == 1) { ka = Kambient; IOcc = IaIs.x; }
          if (P.apply_shadow == 1) { kd = Kdiffuse; ks = ph.ks; ISdw = IaIs.y; }
          float rr, gg, bb;
          if (PHONG) {                                     // a zero gradient leaves L = clr (:241)
            rr = src.x; gg = src.y; bb = src.z;
            float dot_diff, spec;
            if (vrb_phong_terms(vol, ph, kx, ky, kz, tx, ty, tz, cam.ex, cam.ey, cam.ez, dot_diff, spec)) {
              float kk = (1.0f / (ka + kd));
              rr = kk * (src.x * IOcc * ka + ISdw * (src.x * kd * dot_diff)) + ISdw * (ks * ph.isx * spec);
              gg = kk * (src.y * IOcc * ka + ISdw * (src.y * kd * dot_diff)) + ISdw * (ks * ph.isy * spec);
              bb = kk * (src.z * IOcc * ka + ISdw * (src.z * kd * dot_diff)) + ISdw * (ks * ph.isz * spec);
            }
          } else {
            float kk = (1.0f / (ka + kd));
            rr = kk * (src.x * IOcc * ka + src.x * ISdw * kd);
            gg = kk * (src.y * IOcc * ka + src.y * ISdw * kd);
            bb = kk * (src.z * IOcc * ka + src.z * ISdw * kd);
          }
          float a = 1.0f - expf(-src.w * h);
          float om = 1.0f - ca;
          cr = cr + om * (rr * a); cg = cg + om * (gg * a); cb = cb + om * (bb * a); ca = ca + om * a;
          if (ca > 0.99f) break;
        }
        s = s + h;
      }
      vrb_store_pixel(fr, px, py, cr, cg, cb, ca);
    } else if (fr.zero_miss) vrb_store_pixel(fr, px, py, 0.f, 0.f, 0.f, 0.f);
  }
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) ns += __shfl_xor_sync(0xffffffffu, ns, o);
    if (((threadIdx.y * 8 + threadIdx.x) & 31) == 0 && ns) atomicAdd(counter, (unsigned long long)ns);
  }
}

extern "C" int vrb_obj_march_render(vrb_ctx* c, const vrb_camera* cam, const vrb_lighting* light, const vrb_obj_params* p) {
  VRB_REQUIRE(c && cam && light && p, VRB_ERR_INVALID, "vrb_obj_march_render: NULL argument");
  VRB_REQUIRE(c->d_vol, VRB_ERR_STATE, "vrb_obj_march_render: no volume uploaded");
  VRB_REQUIRE(c->d_tf_rgbt, VRB_ERR_STATE, "vrb_obj_march_render: no transfer function uploaded");
  VRB_REQUIRE(c->d_light_cache, VRB_ERR_STATE, "vrb_obj_march_render: no light cache (vrb_*_light_cache_build)");
  VRB_REQUIRE(c->d_frame, VRB_ERR_STATE, "vrb_obj_march_render: no frame (vrb_frame_resize)");
  VRB_REQUIRE(p->step_size > 0.0f, VRB_ERR_INVALID, "vrb_obj_march_render: step_size %g", p->step_size);
  VRB_CUDA(cudaSetDevice(c->device));
  PhongView ph;                             // light->apply_phong == 1: TexVolumeGradient must exist (vrb_gradient_build)
  { int rc = vrb_make_phong_view(c, light, &ph, "vrb_obj_march_render"); if (rc != VRB_OK) return rc; }
  if (!c->d_frame_target) VRB_CUDA(cudaMemsetAsync(c->d_frame, 0, (size_t)c->fw * c->fh * 4 * sizeof(__half), c->stream));
  if (p->count_samples) { int rc = vrb_counters_reset(c); if (rc != VRB_OK) return rc; }
  PartView part;
  dim3 block(8, 8), grid = vrb_make_grid(c, 8, 8, &part);
  size_t smem = (c->tf_n + 2 <= 1026) ? (size_t)(c->tf_n + 2) * sizeof(float4) : 0;
  LcView lc; lc.tex = c->d_light_cache; lc.w = c->lc_dims[0]; lc.h = c->lc_dims[1]; lc.d = c->lc_dims[2];
  VolView vol = c->vol_view();
  vol.tex3d = 0;                         // this marcher has no hardware-filter variant
  const FrameView fr = c->frame_view();
  const CamView cv = make_cam_view(cam);
  if (ph.grad) {
    if (p->count_samples) k_obj_march<true, true><<<grid, block, smem, c->stream>>>(vol, c->d_tf_rgbt, c->tf_n, fr, cv, part, lc, *p, light->ka, light->kd, ph, c->d_counter);
    else                  k_obj_march<false, true><<<grid, block, smem, c->stream>>>(vol, c->d_tf_rgbt, c->tf_n, fr, cv, part, lc, *p, light->ka, light->kd, ph, c->d_counter);
  } else {
    if (p->count_samples) k_obj_march<true, false><<<grid, block, smem, c->stream>>>(vol, c->d_tf_rgbt, c->tf_n, fr, cv, part, lc, *p, light->ka, light->kd, ph, c->d_counter);
    else                  k_obj_march<false, false><<<grid, block, smem, c->stream>>>(vol, c->d_tf_rgbt, c->tf_n, fr, cv, part, lc, *p, light->ka, light->kd, ph, c->d_counter);
  }
  VRB_CUDA(cudaGetLastError());
  c->launches++;
  if (p->count_samples) return vrb_counters_fetch(c);
  return VRB_OK;
}
