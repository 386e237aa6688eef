// march_dos_body.cuh -- device code of the rc1pdosct marcher, included once per filter mode (DOS_HW = 0: software fp32
// blends, file compiled with -fmad=false, bit-reproducible against the oracle; DOS_HW = 1: texture units).

// trilinear fetch from one padded fp16 level at normalised coordinates (same arithmetic as the oracle's tex3d:
// u = s*N - 0.5, floor, clamp-to-edge through the replicated border)
__device__ __forceinline__ float dos_level_fetch(const LevelView& L, float sx, float sy, float sz) {
  float ux = sx * (float)L.w - 0.5f, uy = sy * (float)L.h - 0.5f, uz = sz * (float)L.d - 0.5f;
  float flx = floorf(ux), fly = floorf(uy), flz = floorf(uz);
  float fx = ux - flx, fy = uy - fly, fz = uz - flz;
  int ix = (int)flx, iy = (int)fly, iz = (int)flz;
  int x0 = min(max(ix, 0), L.w - 1), x1 = min(max(ix + 1, 0), L.w - 1);
  int y0 = min(max(iy, 0), L.h - 1), y1 = min(max(iy + 1, 0), L.h - 1);
  int z0 = min(max(iz, 0), L.d - 1), z1 = min(max(iz + 1, 0), L.d - 1);
  const int pw = L.w + 2;
  const long long slice = (long long)pw * (L.h + 2);
  const __half* b = L.tex + slice + pw + 1;      // texel (0,0,0)
  const __half* r00 = b + (long long)z0 * slice + (long long)y0 * pw;
  const __half* r10 = b + (long long)z0 * slice + (long long)y1 * pw;
  const __half* r01 = b + (long long)z1 * slice + (long long)y0 * pw;
  const __half* r11 = b + (long long)z1 * slice + (long long)y1 * pw;
  float c00 = vrb_lerp(__half2float(__ldg(r00 + x0)), __half2float(__ldg(r00 + x1)), fx);
  float c10 = vrb_lerp(__half2float(__ldg(r10 + x0)), __half2float(__ldg(r10 + x1)), fx);
  float c01 = vrb_lerp(__half2float(__ldg(r01 + x0)), __half2float(__ldg(r01 + x1)), fx);
  float c11 = vrb_lerp(__half2float(__ldg(r11 + x0)), __half2float(__ldg(r11 + x1)), fx);
  return vrb_lerp(vrb_lerp(c00, c10, fy), vrb_lerp(c01, c11, fy), fz);
}

// textureLod on the GL_LINEAR_MIPMAP_LINEAR pyramid
__device__ __forceinline__ float dos_texture_lod(const DosConst& C, d3 s, float lod) {
#if DOS_HW
  return tex3DLod<float>(C.pyr_tex, s.x, s.y, s.z, lod);
#endif
  const int maxl = C.n_levels - 1;
  if (!(lod > 0.0f)) return dos_level_fetch(C.lev[0], s.x, s.y, s.z);
  if (lod >= (float)maxl) return dos_level_fetch(C.lev[maxl], s.x, s.y, s.z);
  int l0 = (int)floorf(lod);
  float f = lod - (float)l0;
  float a = dos_level_fetch(C.lev[l0], s.x, s.y, s.z);
  if (f == 0.0f) return a;
  float b = dos_level_fetch(C.lev[l0 + 1], s.x, s.y, s.z);
  return vrb_lerp(a, b, f);
}

// GetGaussianExtinction (ray_bbox_marching.comp:92-112)
__device__ __forceinline__ float dos_gaussian_extinction(const DosConst& C, d3 tp, float mip) {
#if DOS_HW
  float rg = dos_texture_lod(C, m3(tp.x * C.inv_VSS.x, tp.y * C.inv_VSS.y, tp.z * C.inv_VSS.z), mip);
#else
  float rg = dos_texture_lod(C, tp / C.VSS, mip);
#endif
  if (tp.x < 0.0f || tp.x > C.VSS.x || tp.y < 0.0f || tp.y > C.VSS.y || tp.z < 0.0f || tp.z > C.VSS.z) {
    float sg = exp2f(mip);                                  // pow(2.0, mipmaplevel): mip is a small integer, exact
    d3 c = m3(fminf(fmaxf(tp.x, 0.0f), C.VSS.x) - tp.x, fminf(fmaxf(tp.y, 0.0f), C.VSS.y) - tp.y, fminf(fmaxf(tp.z, 0.0f), C.VSS.z) - tp.z);
    float dist = c.x * c.x + c.y * c.y + c.z * c.z;
    rg = rg * expf(-(dist) / (2.0f * sg * sg));
  }
  return rg;
}

// Cone1Ray* -> Cone3Ray* -> Cone7Ray* (ray_bbox_marching.comp:124-322 occlusion, :345-531 shadow): same structure for
// both samplers.  swap_uv reproduces Cone1RayShadow's (k, v, u) parameter list (:481 vs its call at :561).
__device__ float dos_cone(const DosConst& C, const ConeView& K, d3 pos0, d3 k, d3 u, d3 v, bool swap_uv, unsigned int& ntaps) {
  if (swap_uv) { d3 t = u; u = v; v = t; }
  float rays[7], last[7];
  float track = K.initial_step;
  rays[0] = 0.0f; last[0] = 0.0f;
  int id = 0;
  for (int i = 0; i < K.counts[0]; ++i, ++id) {
    float4 si = __ldg(K.sections + id);
    float amptau = dos_gaussian_extinction(C, pos0 + k * track, si.y) * si.w;
    rays[0] += (last[0] + amptau) * si.z * K.ui_weight;
    last[0] = amptau;
    track += si.x;
  }
  ntaps += K.counts[0] + 3 * K.counts[1] + 7 * K.counts[2];
  if (!(K.counts[1] + K.counts[2] > 0)) return expf(-rays[0]);
  rays[2] = rays[0]; rays[1] = rays[0];
  last[2] = last[0]; last[1] = last[0];
  {
    d3 vk[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) vk[i] = k * K.axes[i][2] + u * K.axes[i][1] + v * K.axes[i][0];
    for (int s = 0; s < K.counts[1]; ++s, ++id) {
      float4 si = __ldg(K.sections + id);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float amptau = dos_gaussian_extinction(C, pos0 + vk[i] * track, si.y) * si.w;
        rays[i] += (last[i] + amptau) * si.z * K.ui_weight;
        last[i] = amptau;
      }
      track += si.x;
    }
  }
  if (!(K.counts[2] > 0)) return (expf(-rays[0]) + expf(-rays[1]) + expf(-rays[2])) / 3.0f;
  rays[6] = rays[5] = rays[2];
  rays[4] = rays[3] = rays[1];
  float avg = (rays[2] + rays[1] + rays[0]) / 3.0f;
  rays[2] = rays[1] = rays[0];
  rays[0] = avg;
  last[6] = last[5] = last[2];
  last[4] = last[3] = last[1];
  float avgt = (last[2] + last[1] + last[0]) / 3.0f;
  last[2] = last[1] = last[0];
  last[0] = avgt;
  {
    d3 vk[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) vk[i] = k * K.axes[3 + i][2] + u * K.axes[3 + i][1] + v * K.axes[3 + i][0];
    for (int s = 0; s < K.counts[2]; ++s, ++id) {
      float4 si = __ldg(K.sections + id);
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        float amptau = si.w * dos_gaussian_extinction(C, pos0 + vk[i] * track, si.y);
        rays[i] += (last[i] + amptau) * si.z * K.ui_weight;
        last[i] = amptau;
      }
      track += si.x;
    }
  }
  return (expf(-rays[0]) + (expf(-rays[1]) + expf(-rays[2]) + expf(-rays[3]) + expf(-rays[4]) + expf(-rays[5]) + expf(-rays[6])) *
                               K.ray7_adj_weight) / (1.0f + K.ray7_adj_weight * 6.0f);
}

// ShadowEvaluationKernel (:533-562)
__device__ float dos_shadow(const DosConst& C, d3 pos0, unsigned int& ntaps) {
  d3 k = m3(0.f, 0.f, 0.f), u = k, v = k;
  if (C.P.type_of_shadow == 0 || C.P.type_of_shadow == 1) {
    d3 half = m3(C.VSS.x / 2.0f, C.VSS.y / 2.0f, C.VSS.z / 2.0f);
    d3 cone_vec = nrm3(C.light_pos - (pos0 - half));
    k = cone_vec;
    u = nrm3(cross3(k, C.light_right));
    v = nrm3(cross3(k, u));
    if (C.P.type_of_shadow == 1 && dot3(cone_vec, C.light_fwd) < C.P.spot_cos) return 0.0f;
  } else if (C.P.type_of_shadow == 2) {
    k = C.light_fwd; v = C.light_up; u = C.light_right;
  }
  return dos_cone(C, C.sdw, pos0, k, u, v, true, ntaps);
}

template <bool COUNT, bool PHONG>
__global__ void __launch_bounds__(64)
k_dos(VolView vol, const float4* __restrict__ tf_g, int tf_n, FrameView fr, CamView cam, PartView part,
      const __grid_constant__ DosConst C, unsigned long long* counter) {
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
    // camera_dir is normalised once in main (:673-674); RayAABBIntersection normalises it again into r.Dir
    float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
    float vx = (fx / (float)fr.w) * 2.0f - 1.0f, vy = (fy / (float)fr.h) * 2.0f - 1.0f;
    float cx = vx * cam.tan_fovy * cam.aspect, cy = vy * cam.tan_fovy, cz = -1.0f;
    d3 cdir = nrm3(m3(cx * cam.m[0] + cy * cam.m[1] + cz * cam.m[2], cx * cam.m[3] + cy * cam.m[4] + cz * cam.m[5],
                      cx * cam.m[6] + cy * cam.m[7] + cz * cam.m[8]));
    Ray r = vrb_make_ray(cam, px, py, fr.w, fr.h, C.VSS.x, C.VSS.y, C.VSS.z);
    if (r.hit) {
      d3 v_right = nrm3(cross3(cdir, m3(0.f, 1.f, 0.f)));
      d3 v_up = nrm3(cross3(-cdir, v_right));
      float D = fabsf(r.tfar - r.tnear);
      d3 dir = m3(r.dx, r.dy, r.dz);
      d3 half = C.VSS * 0.5f;
      d3 wd = m3(r.ox, r.oy, r.oz) + dir * r.tnear;
      wd = wd + half;
      float kx = (float)vol.w / vol.gx, ky = (float)vol.h / vol.gy, kz = (float)vol.d / vol.gz;
      float dr = 0.f, dg = 0.f, db = 0.f, da = 0.f;
      const float step = C.P.step_size;
      for (float s = 0.0f; s < D;) {
        float h = fminf(step, D - s);
        d3 tx = wd + dir * (s + h * 0.5f);
#if DOS_HW
        float density = tex3D<float>(vol.tex3d, tx.x * kx, tx.y * ky, tx.z * kz);
#else
        float density = vrb_sample_volume(vol, kx, ky, kz, tx.x, tx.y, tx.z);
#endif
        float4 src = vrb_sample_tf(tf, tf_n, density);
        if (COUNT) ++ns;
        if (src.w > 0.0f) {
          float ka = 0.0f, kd = 0.0f, ks = 0.0f, IOcc = 0.0f, ISdw = 0.0f;
          if (C.P.apply_occlusion == 1) {
            ka = C.ka;
            d3 kvec = nrm3(C.eye - (tx - half));            // OcclusionEvaluationKernel (:324-333)
            IOcc = dos_cone(C, C.occ, tx, kvec, v_up, v_right, false, ntaps);
          }
          if (C.P.apply_shadow == 1) { kd = C.kd; ks = C.ph.ks; ISdw = dos_shadow(C, tx, ntaps); }
          float cr, cg, cb;
          if (PHONG) {                                   // ApplyPhongShading == 1 (:629-648); a zero gradient leaves L = clr
            cr = src.x; cg = src.y; cb = src.z;
            float dot_diff, spec;
            if (vrb_phong_terms(vol, C.ph, kx, ky, kz, tx.x, tx.y, tx.z, C.eye.x, C.eye.y, C.eye.z, dot_diff, spec)) {
              const float f = ((1.0f / (ka + kd)) * (IOcc * ka + ISdw * kd * dot_diff));
              const float sp = (ISdw * ks * spec);
              cr = src.x * f + C.ph.isx * sp; cg = src.y * f + C.ph.isy * sp; cb = src.z * f + C.ph.isz * sp;
            }
          } else {
            float kk = (1.0f / (ka + kd));
            cr = kk * (src.x * IOcc * ka + src.x * ISdw * kd);
            cg = kk * (src.y * IOcc * ka + src.y * ISdw * kd);
            cb = kk * (src.z * IOcc * ka + src.z * ISdw * kd);
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


// K6 lightcachecomputation.comp main (:523-546): one (Iocc, Ishadow) pair per light-cache voxel.  Same cone functions as
// the marcher; the occlusion frame is built from EyeCamUp (:280-293).
__global__ void __launch_bounds__(64)
k_dos_light_cache(const __grid_constant__ DosConst C, d3 eye_up, d3 cell, int rw, int rh, int rd, __half2* __restrict__ cache) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rw * rh * rd) return;
  const int x = i % rw, y = (i / rw) % rh, z = i / (rw * rh);
  float Idao = 1.0f, Idcs = 1.0f;
  unsigned int ntaps = 0;
  d3 tex_pos = m3(((float)x + 0.5f) * cell.x, ((float)y + 0.5f) * cell.y, ((float)z + 0.5f) * cell.z);
  d3 realpos = tex_pos - (C.VSS * 0.5f);
  if (C.P.apply_occlusion == 1) {
    d3 cone_vec = nrm3(C.eye - realpos);
    d3 v_right = nrm3(cross3(-cone_vec, eye_up));
    d3 v_up = nrm3(cross3(cone_vec, v_right));
    Idao = dos_cone(C, C.occ, tex_pos, cone_vec, v_up, v_right, false, ntaps);
  }
  if (C.P.apply_shadow == 1) Idcs = dos_shadow(C, tex_pos, ntaps);
  cache[(size_t)(x + 1) + (size_t)(rw + 2) * ((size_t)(y + 1) + (size_t)(rh + 2) * (size_t)(z + 1))] = __floats2half2_rn(Idao, Idcs);
}

static int dos_light_cache_launch(vrb_ctx* c, const DosConst& C, const float eye_up[3], int rw, int rh, int rd) {
  d3 up; up.x = eye_up[0]; up.y = eye_up[1]; up.z = eye_up[2];
  // VolumeScales * (VolumeDimensions / LightCacheDimensions)
  d3 cell; cell.x = c->scale[0] * ((float)c->vw / (float)rw); cell.y = c->scale[1] * ((float)c->vh / (float)rh); cell.z = c->scale[2] * ((float)c->vd / (float)rd);
  const int n = rw * rh * rd;
  k_dos_light_cache<<<(n + 63) / 64, 64, 0, c->stream>>>(C, up, cell, rw, rh, rd, c->d_light_cache);
  VRB_CUDA(cudaGetLastError());
  return VRB_OK;
}

static int dos_launch(vrb_ctx* c, const vrb_camera* cam, const DosConst& C, int count_samples) {
  PartView part;
  dim3 block(8, 8), grid = vrb_make_grid(c, 8, 8, &part);
  size_t smem = (c->tf_n + 2 <= 1026) ? (size_t)(c->tf_n + 2) * sizeof(float4) : 0;
  VrbKernelTimer timer(c, "k_dos");
  if (C.ph.grad) {
    if (count_samples) k_dos<true, true><<<grid, block, smem, c->stream>>>(c->vol_view(), c->d_tf_rgbt, c->tf_n, c->frame_view(), make_cam_view(cam), part, C, c->d_counter);
    else               k_dos<false, true><<<grid, block, smem, c->stream>>>(c->vol_view(), c->d_tf_rgbt, c->tf_n, c->frame_view(), make_cam_view(cam), part, C, c->d_counter);
  } else {
    if (count_samples) k_dos<true, false><<<grid, block, smem, c->stream>>>(c->vol_view(), c->d_tf_rgbt, c->tf_n, c->frame_view(), make_cam_view(cam), part, C, c->d_counter);
    else               k_dos<false, false><<<grid, block, smem, c->stream>>>(c->vol_view(), c->d_tf_rgbt, c->tf_n, c->frame_view(), make_cam_view(cam), part, C, c->d_counter);
  }
  VRB_CUDA(cudaGetLastError());
  return VRB_OK;
}
