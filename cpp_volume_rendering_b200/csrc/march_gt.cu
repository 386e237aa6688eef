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

struct g3 { float x, y, z; };
__device__ __forceinline__ g3 gm(float x, float y, float z) { g3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ g3 operator+(g3 a, g3 b) { return gm(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ g3 operator-(g3 a, g3 b) { return gm(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ g3 operator-(g3 a) { return gm(-a.x, -a.y, -a.z); }
__device__ __forceinline__ g3 operator*(g3 a, float s) { return gm(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float gdot(g3 a, g3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ g3 gcross(g3 a, g3 b) { return gm(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
__device__ __forceinline__ g3 gnrm(g3 a) { float r = 1.0f / sqrtf(gdot(a, a)); return a * r; }

struct GtConst {
  g3 G;
  vrb_gt_params P;
  float ka, kd;
  g3 light_pos, light_fwd, light_up, light_right;
  const float* occ_rays; const float* sdw_rays;    // n x 3, fp16-rounded
};

// extinction (TF .a) at texture-space position p: trilinear volume tap + linear TF lookup of the .w channel only
__device__ __forceinline__ float gt_extinction(const VolView& vol, float kx, float ky, float kz, const float* __restrict__ tfw, int tf_n, g3 p) {
  float d = vrb_sample_volume(vol, kx, ky, kz, p.x, p.y, p.z);
  float up = fmaf(d, (float)tf_n, 0.5f);
  up = fminf(fmaxf(up, 0.0f), (float)tf_n + 0.5f);
  float fl; int i = vrb_floor_pos(up, &fl);
  return vrb_lerp(tfw[i], tfw[i + 1], up - fl);
}

// ConeOcclusionEvaluationRayCasting / ConeShadowsEvaluationRayCasting share this loop (gt_ray_marching.comp:118-166,203-249)
__device__ float gt_cone(const GtConst& C, const VolView& vol, float kx, float ky, float kz, const float* __restrict__ tfw, int tf_n,
                         const float* __restrict__ table, int nrays, float dist_eval, g3 tx, g3 v_right, g3 v_up, g3 v_dir,
                         unsigned long long& nsteps) {
  float S = 0.0f, Sw = 0.0f;
  const float gap = C.P.light_ray_initial_gap, lstep = C.P.light_ray_step_size;
  for (int rayid = 0; rayid < nrays; ++rayid) {
    g3 c = gm(__ldg(table + 3 * rayid), __ldg(table + 3 * rayid + 1), __ldg(table + 3 * rayid + 2));
    g3 w = gnrm(v_right * c.x + v_up * c.y + v_dir * c.z);
    float Vt = 1.0f;
    float s = gap;
    float st0 = gt_extinction(vol, kx, ky, kz, tfw, tf_n, tx + w * s);
    while (s < dist_eval) {
      float h = fminf(lstep, dist_eval - s);
      g3 at = tx + w * (s + h);
      if (at.x < 0.0f || at.x > C.G.x || at.y < 0.0f || at.y > C.G.y || at.z < 0.0f || at.z > C.G.z) break;
      float st1 = gt_extinction(vol, kx, ky, kz, tfw, tf_n, at);
      Vt *= expf(-((st0 + st1) * 0.5f) * h);
      ++nsteps;
      if ((1 - Vt) > 0.99f) break;
      st0 = st1;
      s = s + h;
    }
    float rw = gdot(v_dir, w);
    S += Vt * rw;
    Sw += rw;
  }
  return (S / Sw);
}

template <bool COUNT>
__global__ void __launch_bounds__(64)
k_gt(VolView vol, const float4* __restrict__ tf_g, int tf_n, FrameView fr, CamView cam, PartView part, const __grid_constant__ GtConst C,
     unsigned long long* counter) {
  extern __shared__ float4 s_tf[];                 // tf_n + 2 RGBA texels, then tf_n + 2 extinction floats
  float* s_tfw = reinterpret_cast<float*>(s_tf + (tf_n + 2));
  for (int i = threadIdx.y * 8 + threadIdx.x; i < tf_n + 2; i += 64) { float4 t = tf_g[i]; s_tf[i] = t; s_tfw[i] = t.w; }
  __syncthreads();
  int px = blockIdx.x * 8 + threadIdx.x, py = vrb_center_out_row(blockIdx.y, gridDim.y) * 8 + threadIdx.y;
  unsigned int ns = 0;
  unsigned long long nsec = 0;
  if (px < fr.w && py < fr.h && vrb_owns_pixel(part, px, py, fr.w)) {
    float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
    float vx = (fx / (float)fr.w) * 2.0f - 1.0f, vy = (fy / (float)fr.h) * 2.0f - 1.0f;
    float cx = vx * cam.tan_fovy * cam.aspect, cy = vy * cam.tan_fovy, cz = -1.0f;
    g3 cdir = gnrm(gm(cx * cam.m[0] + cy * cam.m[1] + cz * cam.m[2], cx * cam.m[3] + cy * cam.m[4] + cz * cam.m[5],
                      cx * cam.m[6] + cy * cam.m[7] + cz * cam.m[8]));
    Ray r = vrb_make_ray(cam, px, py, fr.w, fr.h, C.G.x, C.G.y, C.G.z);
    if (r.hit) {
      g3 v_right = gnrm(gcross(cdir, gm(0.f, 1.f, 0.f)));
      g3 v_up = gnrm(gcross(-cdir, v_right));
      float D = r.tfar - r.tnear;
      g3 dir = gm(r.dx, r.dy, r.dz);
      g3 half = C.G * 0.5f;
      g3 tex_pos = (gm(r.ox, r.oy, r.oz) + dir * r.tnear) + half;
      float kx = (float)vol.w / vol.gx, ky = (float)vol.h / vol.gy, kz = (float)vol.d / vol.gz;
      float cr = 0.f, cg = 0.f, cb = 0.f, ca = 0.f;
      const float step = C.P.step_size;
      float s = 0.0f;
      while (s < D) {
        float h = fminf(step, D - s);
        g3 sp = tex_pos + dir * (s + h * 0.5f);
        float density = vrb_sample_volume(vol, kx, ky, kz, sp.x, sp.y, sp.z);
        float4 src = vrb_sample_tf(s_tf, tf_n, density);
        if (COUNT) ++ns;
        bool done = false;
        if (src.w > 0.0f) {
          float ka = 0.0f, kd = 0.0f, IOcc = 0.0f, ISdw = 0.0f;
          if (C.P.apply_occlusion == 1) {
            ka = C.ka;
            g3 v_dir = gnrm(-cdir);
            IOcc = gt_cone(C, vol, kx, ky, kz, s_tfw, tf_n, C.occ_rays, C.P.occ_num_rays, C.P.occ_cone_distance, sp, v_right, v_up, v_dir, nsec);
          }
          if (C.P.apply_shadow == 1) {
            kd = C.kd;
            g3 l_dir = gm(0.f, 0.f, 0.f), l_up = l_dir, l_right = l_dir;
            bool dark = false;
            if (C.P.shadow_type == 0 || C.P.shadow_type == 1) {
              l_dir = gnrm(C.light_pos - (sp - half));
              l_up = gnrm(gcross(l_dir, C.light_right));
              l_right = gnrm(gcross(l_dir, l_up));
              // reference quirk: the spot test compares a cosine with 30.0 (gt_ray_marching.comp:192): always dark
              if (C.P.shadow_type == 1 && gdot(l_dir, C.light_fwd) < 30.0f) dark = true;
            } else if (C.P.shadow_type == 2) {
              l_dir = C.light_fwd; l_up = C.light_up; l_right = C.light_right;
            }
            ISdw = dark ? 0.0f : gt_cone(C, vol, kx, ky, kz, s_tfw, tf_n, C.sdw_rays, C.P.sdw_num_rays, C.P.sdw_cone_distance, sp, l_right, l_up, l_dir, nsec);
          }
          float kk = (1.0f / (ka + kd));
          float rr = kk * (src.x * IOcc * ka + src.x * ISdw * kd);
          float gg = kk * (src.y * IOcc * ka + src.y * ISdw * kd);
          float bb = kk * (src.z * IOcc * ka + src.z * ISdw * kd);
          float a = 1.0f - expf(-src.w * h);
          float om = 1.0f - ca;
          cr = cr + om * (rr * a); cg = cg + om * (gg * a); cb = cb + om * (bb * a); ca = ca + om * a;
          if (ca > 0.99f) done = true;
        }
        // imageStore(OutputFrag) at the end of every dispatch (rgba16f), imageLoad at the start of the next
        cr = __half2float(__float2half_rn(cr)); cg = __half2float(__float2half_rn(cg));
        cb = __half2float(__float2half_rn(cb)); ca = __half2float(__float2half_rn(ca));
        if (done) break;
        s = s + h;
        if (!(s < D)) break;
        s = __half2float(__float2half_rn(s));       // StateFrag is rg16f
      }
      vrb_store_pixel(fr, px, py, cr, cg, cb, ca);
    }
  }
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) { ns += __shfl_xor_sync(0xffffffffu, ns, o); nsec += __shfl_xor_sync(0xffffffffu, nsec, o); }
    if (((threadIdx.y * 8 + threadIdx.x) & 31) == 0 && ns) { atomicAdd(counter, (unsigned long long)ns); atomicAdd(counter + 1, nsec); }
  }
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
  VRB_CUDA(cudaMemsetAsync(c->d_frame, 0, (size_t)c->fw * c->fh * 4 * sizeof(__half), c->stream));
  if (p->count_samples) { int rc = vrb_counters_reset(c); if (rc != VRB_OK) return rc; }
  dim3 block(8, 8), grid((c->fw + 7) / 8, (c->fh + 7) / 8);
  size_t smem = (size_t)(c->tf_n + 2) * (sizeof(float4) + sizeof(float));
  if (p->count_samples) k_gt<true><<<grid, block, smem, c->stream>>>(c->vol_view(), c->d_tf_rgbt, c->tf_n, c->frame_view(), make_cam_view(cam), c->part, C, c->d_counter);
  else                  k_gt<false><<<grid, block, smem, c->stream>>>(c->vol_view(), c->d_tf_rgbt, c->tf_n, c->frame_view(), make_cam_view(cam), c->part, C, c->d_counter);
  VRB_CUDA(cudaGetLastError());
  c->launches++;
  if (p->count_samples) return vrb_counters_fetch(c);
  return VRB_OK;
}
