// march_vct_body.cuh -- device code of the rc1pvctsg marcher, included once per filter mode (VCT_HW = 0: software fp32
// blends, compiled with -fmad=false, bit-reproducible against the oracle; VCT_HW = 1: texture units).

// trilinear RG fetch at normalised coordinates from one padded level (oracle arithmetic: u = s*N - 0.5, floor, clamp)
// BRICK: the level is a window of the whole volume's level; coordinates, floor and clamp-to-edge are the WHOLE volume's
// (identical weights and texels), the window offset is subtracted last.  The window is sized by the host so that every
// tap of an owned sample falls inside it; the final clamp to the padded array only guards memory.
template <bool BRICK>
__device__ __forceinline__ float2 sv_fetch(const SvLevel& L, float sx, float sy, float sz) {
  const int gw = BRICK ? L.gw : L.w, gh = BRICK ? L.gh : L.h, gd = BRICK ? L.gd : L.d;
  float ux = sx * (float)gw - 0.5f, uy = sy * (float)gh - 0.5f, uz = sz * (float)gd - 0.5f;
  float flx = floorf(ux), fly = floorf(uy), flz = floorf(uz);
  float fx = ux - flx, fy = uy - fly, fz = uz - flz;
  int ix = (int)flx, iy = (int)fly, iz = (int)flz;
  int x0 = min(max(ix, 0), gw - 1), x1 = min(max(ix + 1, 0), gw - 1);
  int y0 = min(max(iy, 0), gh - 1), y1 = min(max(iy + 1, 0), gh - 1);
  int z0 = min(max(iz, 0), gd - 1), z1 = min(max(iz + 1, 0), gd - 1);
  if (BRICK) {
    x0 = min(max(x0 - L.ox, -1), L.w); x1 = min(max(x1 - L.ox, -1), L.w);
    y0 = min(max(y0 - L.oy, -1), L.h); y1 = min(max(y1 - L.oy, -1), L.h);
    z0 = min(max(z0 - L.oz, -1), L.d); z1 = min(max(z1 - L.oz, -1), L.d);
  }
  const int pw = L.w + 2;
  const long long slice = (long long)pw * (L.h + 2);
  const __half2* b = L.tex + slice + pw + 1;
  const __half2* r00 = b + (long long)z0 * slice + (long long)y0 * pw;
  const __half2* r10 = b + (long long)z0 * slice + (long long)y1 * pw;
  const __half2* r01 = b + (long long)z1 * slice + (long long)y0 * pw;
  const __half2* r11 = b + (long long)z1 * slice + (long long)y1 * pw;
  float2 a0 = __half22float2(__ldg(r00 + x0)), a1 = __half22float2(__ldg(r00 + x1));
  float2 b0 = __half22float2(__ldg(r10 + x0)), b1 = __half22float2(__ldg(r10 + x1));
  float2 c0 = __half22float2(__ldg(r01 + x0)), c1 = __half22float2(__ldg(r01 + x1));
  float2 d0 = __half22float2(__ldg(r11 + x0)), d1 = __half22float2(__ldg(r11 + x1));
  float2 o;
  o.x = vrb_lerp(vrb_lerp(vrb_lerp(a0.x, a1.x, fx), vrb_lerp(b0.x, b1.x, fx), fy), vrb_lerp(vrb_lerp(c0.x, c1.x, fx), vrb_lerp(d0.x, d1.x, fx), fy), fz);
  o.y = vrb_lerp(vrb_lerp(vrb_lerp(a0.y, a1.y, fx), vrb_lerp(b0.y, b1.y, fx), fy), vrb_lerp(vrb_lerp(c0.y, c1.y, fx), vrb_lerp(d0.y, d1.y, fx), fy), fz);
  return o;
}
template <bool BRICK>
__device__ __forceinline__ float2 sv_texture_lod(const VctConst& C, v3f s, float lod) {
#if VCT_HW
  return tex3DLod<float2>(C.sv_tex, s.x, s.y, s.z, lod);
#endif
  const int maxl = C.n_levels - 1;
  if (!(lod > 0.0f)) return sv_fetch<BRICK>(C.lev[0], s.x, s.y, s.z);
  if (lod >= (float)maxl) return sv_fetch<BRICK>(C.lev[maxl], s.x, s.y, s.z);
  int l0 = (int)floorf(lod);
  float f = lod - (float)l0;
  float2 a = sv_fetch<BRICK>(C.lev[l0], s.x, s.y, s.z);
  if (f == 0.0f) return a;
  float2 b = sv_fetch<BRICK>(C.lev[l0 + 1], s.x, s.y, s.z);
  return make_float2(vrb_lerp(a.x, b.x, f), vrb_lerp(a.y, b.y, f));
}
__device__ __forceinline__ float lut_fetch(const VctConst& C, float sx, float sy) {
#if VCT_HW
  return tex2D<float>(C.lut_tex, sx, sy);
#endif
  float ux = sx * (float)C.lut_w - 0.5f, uy = sy * (float)C.lut_h - 0.5f;
  float flx = floorf(ux), fly = floorf(uy);
  float fx = ux - flx, fy = uy - fly;
  int ix = (int)flx, iy = (int)fly;
  int x0 = min(max(ix, 0), C.lut_w - 1), x1 = min(max(ix + 1, 0), C.lut_w - 1);
  int y0 = min(max(iy, 0), C.lut_h - 1), y1 = min(max(iy + 1, 0), C.lut_h - 1);
  const int pw = C.lut_w + 2;
  const __half* b = C.lut + pw + 1;
  float a = vrb_lerp(__half2float(__ldg(b + x0 + pw * y0)), __half2float(__ldg(b + x1 + pw * y0)), fx);
  float c = vrb_lerp(__half2float(__ldg(b + x0 + pw * y1)), __half2float(__ldg(b + x1 + pw * y1)), fx);
  return vrb_lerp(a, c, fy);
}

// EvaluationVoxelConeTracing (vct_ray_bbox_marching.comp:97-144)
template <bool BRICK>
__device__ float vct_cone(const VctConst& C, v3f tex_pos, unsigned int& ntaps) {
  float Tvd = 1.0f;
  v3f realpos = tex_pos - (C.VSS * 0.5f);
  v3f cone_vec = vnrm(C.light_pos - realpos);
  float apex_distance = C.P.cone_initial_step;
  float step_size = C.P.cone_step_size;
  for (int is = 0; is < C.P.cone_number_of_samples; ++is) {
    float xl_x = (apex_distance + step_size * 0.5f);
    float mm_level = log2f((2.0f * xl_x * C.P.tan_cone_apex_angle) / 1.0f);
    v3f wpos = (tex_pos + cone_vec * xl_x);
    if (wpos.x < 0 || wpos.x > C.VSS.x || wpos.y < 0 || wpos.y > C.VSS.y || wpos.z < 0 || wpos.z > C.VSS.z) break;
#if VCT_HW
    float2 g = sv_texture_lod<BRICK>(C, vm(fmaf(wpos.x, C.sv_scale.x, C.sv_bias.x), fmaf(wpos.y, C.sv_scale.y, C.sv_bias.y), fmaf(wpos.z, C.sv_scale.z, C.sv_bias.z)), mm_level);
#else
    float2 g = sv_texture_lod<BRICK>(C, wpos / C.VSS, mm_level);
#endif
    float opacity = lut_fetch(C, (g.x + 0.5f) / C.P.volume_max_density, (g.y + 0.5f) / C.P.volume_max_stddev);
    opacity = 1.0f - powf(1.0f - opacity, step_size * C.corr_fact);
    Tvd *= (1.0f - opacity);
    apex_distance = apex_distance + step_size;
    step_size = step_size * C.P.cone_step_increase_rate;
    ++ntaps;
  }
  return Tvd;
}

template <bool COUNT, bool PHONG>
__global__ void __launch_bounds__(64)
k_vct(VolView vol, const float4* __restrict__ tf_g, int tf_n, FrameView fr, CamView cam, PartView part, const __grid_constant__ VctConst C,
      unsigned long long* counter) {
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
  unsigned int ns = 0, ntaps = 0;
  if (px < fr.w && py < fr.h && vrb_owns_pixel(part, px, py, fr.w)) {
    Ray r = vrb_make_ray(cam, px, py, fr.w, fr.h, C.VSS.x, C.VSS.y, C.VSS.z);
    if (r.hit) {
      float D = fabsf(r.tfar - r.tnear);
      v3f dir = vm(r.dx, r.dy, r.dz);
      v3f wd = vm(r.ox, r.oy, r.oz) + dir * r.tnear;
      wd = wd + (C.VSS * 0.5f);
      float kx = (float)vol.w / vol.gx, ky = (float)vol.h / vol.gy, kz = (float)vol.d / vol.gz;
      float dr = 0.f, dg = 0.f, db = 0.f, da = 0.f;
      const float step = C.P.step_size;
      for (float s = 0.0f; s < D;) {
        float h = fminf(step, D - s);
        v3f tx = wd + dir * (s + h * 0.5f);
#if VCT_HW
        float density = tex3D<float>(vol.tex3d, tx.x * kx, tx.y * ky, tx.z * kz);
#else
        float density = vrb_sample_volume(vol, kx, ky, kz, tx.x, tx.y, tx.z);
#endif
        float4 src = vrb_sample_tf(tf, tf_n, density);
        if (COUNT) ++ns;
        if (src.w > 0.0f) {
          float ka = 0.0f, kd = 0.0f, ks = 0.0f, Ivd = 0.0f;
          if (C.P.apply_occlusion == 1) ka = C.ka;
          if (C.P.apply_shadow == 1) { kd = C.kd; ks = C.ph.ks; Ivd = vct_cone<false>(C, tx, ntaps); }
          float cr, cg, cb;
          if (PHONG) {                          // ApplyPhongShading == 1 (:163-182); a zero gradient leaves L = clr
            cr = src.x; cg = src.y; cb = src.z;
            float dot_diff, spec;
            if (vrb_phong_terms(vol, C.ph, kx, ky, kz, tx.x, tx.y, tx.z, cam.ex, cam.ey, cam.ez, dot_diff, spec)) {
              float kk = (1.0f / (ka + kd));
              cr = kk * (src.x * ka + Ivd * (src.x * kd * dot_diff)) + Ivd * (ks * C.ph.isx * spec);
              cg = kk * (src.y * ka + Ivd * (src.y * kd * dot_diff)) + Ivd * (ks * C.ph.isy * spec);
              cb = kk * (src.z * ka + Ivd * (src.z * kd * dot_diff)) + Ivd * (ks * C.ph.isz * spec);
            }
          } else {
            float kk = (1.0f / (ka + kd));
            cr = kk * (src.x * ka + src.x * Ivd * kd);
            cg = kk * (src.y * ka + src.y * Ivd * kd);
            cb = kk * (src.z * ka + src.z * Ivd * kd);
          }
          float a = 1.0f - expf(-src.w * h);
          float om = 1.0f - da;
          dr = dr + om * (cr * a); dg = dg + om * (cg * a); db = db + om * (cb * a); da = da + om * a;
          if (da > 0.99f) break;
        }
        s = s + h;
      }
      vrb_store_pixel(fr, px, py, dr, dg, db, da);
    } else if (fr.zero_miss) vrb_store_pixel(fr, px, py, 0.f, 0.f, 0.f, 0.f);
  }
  if (COUNT) {
    unsigned long long nt64 = ntaps;
    for (int o = 16; o > 0; o >>= 1) { ns += __shfl_xor_sync(0xffffffffu, ns, o); nt64 += __shfl_xor_sync(0xffffffffu, nt64, o); }
    if (((threadIdx.y * 8 + threadIdx.x) & 31) == 0 && ns) { atomicAdd(counter, (unsigned long long)ns); atomicAdd(counter + 1, nt64); }
  }
}


#if !VCT_HW
// Deferred frame (shade_list.cuh, march_list.cu), exact filter mode: one list entry per lane.  ShadeSample (:146-189) for
// the entry's position; the colour the compositing multiplies by alpha replaces the TF colour in the entry.  (Splitting the 50
// cone steps of an entry over 8 lanes, product taken in step order, was measured too: 2.09 ms against 1.70 ms at config
// 5-1gpu -- the kernel is bound by its taps, not by the length of the chain -- and removed.)
template <bool PHONG>
__global__ void __launch_bounds__(128)
k_vct_shade(VolView vol, CamView cam, const __grid_constant__ VctConst C, ShadeListView L, unsigned n_entries, int count, unsigned long long* counter) {
  const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
  const float4 A = (e < n_entries) ? L.a[e] : make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
  unsigned int ntaps = 0;
  if (__float_as_int(A.w) >= 0) {                  // slots reserved but never written keep pixel == -1
    const float4 src = L.b[e];
    const v3f tx = vm(A.x, A.y, A.z);
    float ka = 0.0f, kd = 0.0f, ks = 0.0f, Ivd = 0.0f;
    if (C.P.apply_occlusion == 1) ka = C.ka;
    if (C.P.apply_shadow == 1) { kd = C.kd; ks = C.ph.ks; Ivd = vct_cone<false>(C, tx, ntaps); }
    float cr, cg, cb;
    if (PHONG) {                          // ApplyPhongShading == 1 (:163-182); a zero gradient leaves L = clr
      cr = src.x; cg = src.y; cb = src.z;
      const float kx = (float)vol.w / vol.gx, ky = (float)vol.h / vol.gy, kz = (float)vol.d / vol.gz;
      float dot_diff, spec;
      if (vrb_phong_terms(vol, C.ph, kx, ky, kz, tx.x, tx.y, tx.z, cam.ex, cam.ey, cam.ez, dot_diff, spec)) {
        float kk = (1.0f / (ka + kd));
        cr = kk * (src.x * ka + Ivd * (src.x * kd * dot_diff)) + Ivd * (ks * C.ph.isx * spec);
        cg = kk * (src.y * ka + Ivd * (src.y * kd * dot_diff)) + Ivd * (ks * C.ph.isy * spec);
        cb = kk * (src.z * ka + Ivd * (src.z * kd * dot_diff)) + Ivd * (ks * C.ph.isz * spec);
      }
    } else {
      float kk = (1.0f / (ka + kd));
      cr = kk * (src.x * ka + src.x * Ivd * kd);
      cg = kk * (src.y * ka + src.y * Ivd * kd);
      cb = kk * (src.z * ka + src.z * Ivd * kd);
    }
    L.b[e] = make_float4(cr, cg, cb, src.w);
  }
  if (count) {
    unsigned long long nt64 = ntaps;
    for (int o = 16; o > 0; o >>= 1) nt64 += __shfl_xor_sync(0xffffffffu, nt64, o);
    if ((threadIdx.x & 31) == 0 && nt64) atomicAdd(counter + 1, nt64);
  }
}

static int vct_deferred_launch(vrb_ctx* c, const vrb_camera* cam, const VctConst& C, int count_samples) {
  ListFrame f;
  int rc = vrb_list_march(c, cam, C.P.step_size, 0, count_samples, &f);
  if (rc != VRB_OK) return rc;
  if (f.n_entries) {
    const unsigned blocks = (f.n_entries + 127u) / 128u;
    VrbKernelTimer timer(c, "k_vct_shade");
    if (C.ph.grad) k_vct_shade<true><<<blocks, 128, 0, c->stream>>>(c->vol_view(), make_cam_view(cam), C, f.L, f.n_entries, count_samples, c->d_counter);
    else           k_vct_shade<false><<<blocks, 128, 0, c->stream>>>(c->vol_view(), make_cam_view(cam), C, f.L, f.n_entries, count_samples, c->d_counter);
    VRB_CUDA(cudaGetLastError());
    c->launches++;
  }
  return vrb_list_composite(c, cam, 0, f);
}
#endif

// K13 rc1pvctsg/lightcachecomputation.comp (:45-118): Iao = 1, Ivd = the cone WITHOUT the leave-the-volume cut
// (CUT_WHEN_AWAY_FROM_VOLUME is not defined there, :3), evaluated at the cache voxel centres.
__global__ void __launch_bounds__(64)
k_vct_light_cache(const __grid_constant__ VctConst C, v3f cell, int rw, int rh, int rd, __half2* __restrict__ cache) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rw * rh * rd) return;
  const int x = i % rw, y = (i / rw) % rh, z = i / (rw * rh);
  v3f tex_pos = vm(((float)x + 0.5f) * cell.x, ((float)y + 0.5f) * cell.y, ((float)z + 0.5f) * cell.z);
  float Ivd = 1.0f;
  if (C.P.apply_shadow == 1) {
    float Tvd = 1.0f;
    v3f realpos = tex_pos - (C.VSS * 0.5f);
    v3f cone_vec = vnrm(C.light_pos - realpos);
    float apex_distance = C.P.cone_initial_step;
    float step_size = C.P.cone_step_size;
    for (int is = 0; is < C.P.cone_number_of_samples; ++is) {
      float xl_x = (apex_distance + step_size * 0.5f);
      float mm_level = log2f((2.0f * xl_x * C.P.tan_cone_apex_angle) / 1.0f);
      v3f wpos = (tex_pos + cone_vec * xl_x);
#if VCT_HW
      float2 g = sv_texture_lod<false>(C, vm(wpos.x * C.inv_VSS.x, wpos.y * C.inv_VSS.y, wpos.z * C.inv_VSS.z), mm_level);
#else
      float2 g = sv_texture_lod<false>(C, wpos / C.VSS, mm_level);
#endif
      float opacity = lut_fetch(C, (g.x + 0.5f) / C.P.volume_max_density, (g.y + 0.5f) / C.P.volume_max_stddev);
      opacity = 1.0f - powf(1.0f - opacity, step_size * C.corr_fact);
      Tvd *= (1.0f - opacity);
      apex_distance = apex_distance + step_size;
      step_size = step_size * C.P.cone_step_increase_rate;
    }
    Ivd = Tvd;
  }
  cache[(size_t)(x + 1) + (size_t)(rw + 2) * ((size_t)(y + 1) + (size_t)(rh + 2) * (size_t)(z + 1))] = __floats2half2_rn(1.0f, Ivd);
}

static int vct_light_cache_launch(vrb_ctx* c, const VctConst& C, int rw, int rh, int rd) {
  v3f cell; cell.x = c->scale[0] * ((float)c->vw / (float)rw); cell.y = c->scale[1] * ((float)c->vh / (float)rh); cell.z = c->scale[2] * ((float)c->vd / (float)rd);
  const int n = rw * rh * rd;
  k_vct_light_cache<<<(n + 63) / 64, 64, 0, c->stream>>>(C, cell, rw, rh, rd, c->d_light_cache);
  VRB_CUDA(cudaGetLastError());
  return VRB_OK;
}

static int vct_launch(vrb_ctx* c, const vrb_camera* cam, const VctConst& C, int count_samples) {
  PartView part;
  dim3 block(8, 8), grid = vrb_make_grid(c, 8, 8, &part);
  size_t smem = (c->tf_n + 2 <= 1026) ? (size_t)(c->tf_n + 2) * sizeof(float4) : 0;
  VrbKernelTimer timer(c, "k_vct");
  if (C.ph.grad) {
    if (count_samples) k_vct<true, true><<<grid, block, smem, c->stream>>>(c->vol_view(), c->d_tf_rgbt, c->tf_n, c->frame_view(), make_cam_view(cam), part, C, c->d_counter);
    else               k_vct<false, true><<<grid, block, smem, c->stream>>>(c->vol_view(), c->d_tf_rgbt, c->tf_n, c->frame_view(), make_cam_view(cam), part, C, c->d_counter);
  } else {
    if (count_samples) k_vct<true, false><<<grid, block, smem, c->stream>>>(c->vol_view(), c->d_tf_rgbt, c->tf_n, c->frame_view(), make_cam_view(cam), part, C, c->d_counter);
    else               k_vct<false, false><<<grid, block, smem, c->stream>>>(c->vol_view(), c->d_tf_rgbt, c->tf_n, c->frame_view(), make_cam_view(cam), part, C, c->d_counter);
  }
  VRB_CUDA(cudaGetLastError());
  return VRB_OK;
}


// ---- sort-last brick of the VCT renderer -----------------------------------------------
// The ray, its sample positions and the compositing arithmetic are k_vct's; a sample is composited by the brick that owns
// its voxel cell (as k_rc1pass_brick in sort_last.cu).  The volume texels and the pyramid levels are windows of the whole
// volume's arrays (owned cells + a halo sized for the cone reach), addressed with the whole volume's coordinates.
// MODE 0: independent segment; MODE 1: opacity of the segment only (no cones); MODE 2: starts from the opacity of the
// bricks in front (their MODE-1 buffers, local or peer memory), so the 0.99 cut falls where it does on one GPU.
__device__ __forceinline__ float vct_brick_sample(const VolView& v, const VctBrick& b, float px, float py, float pz) {
#if VCT_HW
  return tex3D<float>(v.tex3d, px * b.kx - b.offx, py * b.ky - b.offy, pz * b.kz - b.offz);
#else
  // whole-volume padded index (u + 1) with its clamp, then the window offset (an integer: the subtraction is exact)
  float ux = fmaf(px, b.kx, 0.5f), uy = fmaf(py, b.ky, 0.5f), uz = fmaf(pz, b.kz, 0.5f);
  ux = fminf(fmaxf(ux, 0.0f), (float)b.nx + 0.999f) - b.offx;
  uy = fminf(fmaxf(uy, 0.0f), (float)b.ny + 0.999f) - b.offy;
  uz = fminf(fmaxf(uz, 0.0f), (float)b.nz + 0.999f) - b.offz;
  ux = fminf(fmaxf(ux, 0.0f), (float)v.w + 0.999f);
  uy = fminf(fmaxf(uy, 0.0f), (float)v.h + 0.999f);
  uz = fminf(fmaxf(uz, 0.0f), (float)v.d + 0.999f);
  float flx, fly, flz;
  int ix = vrb_floor_pos(ux, &flx), iy = vrb_floor_pos(uy, &fly), iz = vrb_floor_pos(uz, &flz);
  return vrb_fetch_volume(v, ix, iy, iz, ux - flx, uy - fly, uz - flz);
#endif
}

template <bool COUNT, int MODE>
__global__ void __launch_bounds__(64)
k_vct_brick(VolView vol, const __grid_constant__ VctBrick B, const float4* __restrict__ tf_g, int tf_n, float4* __restrict__ partial,
            float* __restrict__ alpha_out, const __grid_constant__ VctFront front, int W, int H, CamView cam, const __grid_constant__ VctConst C,
            int jump, unsigned long long* counter) {
  extern __shared__ float4 s_tf[];
  const float4* tf = tf_g;
  if (tf_n + 2 <= 1026) {
    for (int i = threadIdx.y * 8 + threadIdx.x; i < tf_n + 2; i += 64) s_tf[i] = tf_g[i];
    tf = s_tf;
  }
  __syncthreads();
  const int px = blockIdx.x * 8 + threadIdx.x, py = vrb_center_out_row(blockIdx.y, gridDim.y) * 8 + threadIdx.y;
  unsigned int ns = 0, ntaps = 0;
  if (px < W && py < H) {
    float dr = 0.f, dg = 0.f, db = 0.f, da = 0.f;
    bool touched = false;
    if (MODE == 2) {
      for (int k = 0; k < front.n; ++k) { float a = front.p[k][(size_t)py * W + px]; da = fmaf(1.0f - da, a, da); }
    }
    Ray r = vrb_make_ray(cam, px, py, W, H, C.VSS.x, C.VSS.y, C.VSS.z);
    if (r.hit && !(MODE == 2 && da > 0.99f)) {
      float D = fabsf(r.tfar - r.tnear);
      v3f dir = vm(r.dx, r.dy, r.dz);
      v3f wd = vm(r.ox, r.oy, r.oz) + dir * r.tnear;
      wd = wd + (C.VSS * 0.5f);
      const float step = C.P.step_size;
      // jump: start at the first sample that can be owned (s = k0 * step, bit-identical to walking there when the host found
      // every multiple of the step exact: vrb_step_multiples_exact), stop after the last one
      float s = 0.0f, s_end = 3.0e38f;
      bool any = true;
      if (jump) { int k0 = 0; any = vrb_brick_span(wd.x, wd.y, wd.z, dir.x, dir.y, dir.z, D, step, B.lo, B.hi, B.kx, B.ky, B.kz, k0, s_end); s = (float)k0 * step; }
      for (; any && s < D && s <= s_end;) {
        float h = fminf(step, D - s);
        v3f tx = wd + dir * (s + h * 0.5f);
        int cx = min(max((int)floorf(tx.x * B.kx), 0), B.nx - 1);
        int cy = min(max((int)floorf(tx.y * B.ky), 0), B.ny - 1);
        int cz = min(max((int)floorf(tx.z * B.kz), 0), B.nz - 1);
        if (cx >= B.lo[0] && cx < B.hi[0] && cy >= B.lo[1] && cy < B.hi[1] && cz >= B.lo[2] && cz < B.hi[2]) {
          float density = vct_brick_sample(vol, B, tx.x, tx.y, tx.z);
          float4 src = vrb_sample_tf(tf, tf_n, density);
          if (COUNT) ++ns;
          touched = true;
          if (src.w > 0.0f) {
            float a = 1.0f - expf(-src.w * h);
            float om = 1.0f - da;
            if (MODE != 1) {
              float ka = 0.0f, kd = 0.0f, Ivd = 0.0f;
              if (C.P.apply_occlusion == 1) ka = C.ka;
              if (C.P.apply_shadow == 1) { kd = C.kd; Ivd = vct_cone<true>(C, tx, ntaps); }
              float kk = (1.0f / (ka + kd));
              float cr = kk * (src.x * ka + src.x * Ivd * kd);
              float cg = kk * (src.y * ka + src.y * Ivd * kd);
              float cb = kk * (src.z * ka + src.z * Ivd * kd);
              dr = dr + om * (cr * a); dg = dg + om * (cg * a); db = db + om * (cb * a);
            }
            da = da + om * a;
            if (da > 0.99f) break;
          }
        }
        s = s + h;
      }
    }
    if (MODE == 1) alpha_out[(size_t)py * W + px] = da;
    else partial[(size_t)py * W + px] = make_float4(dr, dg, db, (MODE == 2 && !touched) ? 0.0f : da);
  }
  if (COUNT) {
    unsigned long long nt64 = ntaps;
    for (int o = 16; o > 0; o >>= 1) { ns += __shfl_xor_sync(0xffffffffu, ns, o); nt64 += __shfl_xor_sync(0xffffffffu, nt64, o); }
    if (((threadIdx.y * 8 + threadIdx.x) & 31) == 0 && ns) { atomicAdd(counter, (unsigned long long)ns); atomicAdd(counter + 1, nt64); }
  }
}

static int vct_brick_launch(vrb_ctx* c, const vrb_camera* cam, const VctConst& C, const VctBrick& B, const VctFront& front, int mode, int count_samples) {
  dim3 block(8, 8), grid((c->fw + 7) / 8, (c->fh + 7) / 8);
  size_t smem = (c->tf_n + 2 <= 1026) ? (size_t)(c->tf_n + 2) * sizeof(float4) : 0;
  const float diag = sqrtf(C.VSS.x * C.VSS.x + C.VSS.y * C.VSS.y + C.VSS.z * C.VSS.z);
  const char* je = getenv("VRB_BRICK_JUMP");
  const int jump = (!(je && je[0] == '0') && vrb_step_multiples_exact(C.P.step_size, diag)) ? 1 : 0;
#define VRB_VCT_BRICK(CNT, MD) k_vct_brick<CNT, MD><<<grid, block, smem, c->stream>>>(c->vol_view(), B, c->d_tf_rgbt, c->tf_n, (float4*)c->d_partial, \
      (float*)c->d_brick_alpha, front, c->fw, c->fh, make_cam_view(cam), C, jump, c->d_counter)
  if (mode == 0) { if (count_samples) VRB_VCT_BRICK(true, 0); else VRB_VCT_BRICK(false, 0); }
  else if (mode == 1) { if (count_samples) VRB_VCT_BRICK(true, 1); else VRB_VCT_BRICK(false, 1); }
  else { if (count_samples) VRB_VCT_BRICK(true, 2); else VRB_VCT_BRICK(false, 2); }
#undef VRB_VCT_BRICK
  VRB_CUDA(cudaGetLastError());
  return VRB_OK;
}
