// march_gt_body.cuh -- device code of the rc1pcrtgt marcher, included once per filter mode (GT_HW = 0: software fp32
// blends, compiled with -fmad=false, bit-reproducible against the oracle; GT_HW = 1: texture-unit trilinear).
#ifndef GT_ILP
#define GT_ILP 4
#endif
#if GT_HW
#define GT_SAMPLE(vol, kx, ky, kz, x, y, z) tex3D<float>((vol).tex3d, (x) * (kx), (y) * (ky), (z) * (kz))
#else
#define GT_SAMPLE(vol, kx, ky, kz, x, y, z) vrb_sample_volume(vol, kx, ky, kz, x, y, z)
#endif

// extinction (TF .a) at texture-space position p: trilinear volume tap + linear TF lookup of the .w channel only
__device__ __forceinline__ float gt_extinction(const VolView& vol, float kx, float ky, float kz, const float* __restrict__ tfw, int tf_n, g3 p) {
  float d = GT_SAMPLE(vol, kx, ky, kz, p.x, p.y, p.z);
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
  // GT_ILP secondary rays are marched together by one thread: their fetch chains are independent, which hides the
  // tap latency; every ray's own arithmetic and the order of the S / Sw accumulation are those of the one-ray loop.
  for (int ray0 = 0; ray0 < nrays; ray0 += GT_ILP) {
    g3 w[GT_ILP];
    float Vt[GT_ILP], sp[GT_ILP], st0[GT_ILP];
    bool act[GT_ILP];
    bool any = false;
#pragma unroll
    for (int j = 0; j < GT_ILP; ++j) {
      const int rayid = min(ray0 + j, nrays - 1);
      g3 c = gm(__ldg(table + 3 * rayid), __ldg(table + 3 * rayid + 1), __ldg(table + 3 * rayid + 2));
      w[j] = gnrm(v_right * c.x + v_up * c.y + v_dir * c.z);
      Vt[j] = 1.0f;
      sp[j] = gap;
      act[j] = (ray0 + j < nrays);
      st0[j] = 0.0f;
      if (act[j]) st0[j] = gt_extinction(vol, kx, ky, kz, tfw, tf_n, tx + w[j] * gap);
      act[j] = act[j] && (gap < dist_eval);
      any = any || act[j];
    }
    while (any) {
      any = false;
      // phase 1: the GT_ILP taps are issued back to back (finished rays re-read the cone origin, a cached tap, so that
      // the fetches stay unconditional and independent); phase 2: each ray's own update, in the one-ray loop's order
      float hh[GT_ILP], st1[GT_ILP];
      bool inside[GT_ILP];
#pragma unroll
      for (int j = 0; j < GT_ILP; ++j) {
        hh[j] = fminf(lstep, dist_eval - sp[j]);
        g3 at = tx + w[j] * (sp[j] + hh[j]);
        inside[j] = !(at.x < 0.0f || at.x > C.G.x || at.y < 0.0f || at.y > C.G.y || at.z < 0.0f || at.z > C.G.z);
        const bool use = act[j] && inside[j];
        g3 q = gm(use ? at.x : tx.x, use ? at.y : tx.y, use ? at.z : tx.z);
        st1[j] = gt_extinction(vol, kx, ky, kz, tfw, tf_n, q);
      }
#pragma unroll
      for (int j = 0; j < GT_ILP; ++j) {
        if (act[j]) {
          if (!inside[j]) { act[j] = false; }
          else {
            Vt[j] *= expf(-((st0[j] + st1[j]) * 0.5f) * hh[j]);
            ++nsteps;
            if ((1 - Vt[j]) > 0.99f) act[j] = false;
            else { st0[j] = st1[j]; sp[j] = sp[j] + hh[j]; act[j] = sp[j] < dist_eval; }
          }
          any = any || act[j];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < GT_ILP; ++j) {
      if (ray0 + j < nrays) {
        float rw = gdot(v_dir, w[j]);
        S += Vt[j] * rw;
        Sw += rw;
      }
    }
  }
  return (S / Sw);
}

template <bool COUNT, bool PHONG>
__global__ void __launch_bounds__(64)
k_gt(VolView vol, const float4* __restrict__ tf_g, int tf_n, FrameView fr, CamView cam, PartView part, const __grid_constant__ GtConst C,
     unsigned long long* counter) {
  extern __shared__ float4 s_tf[];                 // tf_n + 2 RGBA texels, then tf_n + 2 extinction floats
  float* s_tfw = reinterpret_cast<float*>(s_tf + (tf_n + 2));
  for (int i = threadIdx.y * 8 + threadIdx.x; i < tf_n + 2; i += 64) { float4 t = tf_g[i]; s_tf[i] = t; s_tfw[i] = t.w; }
  __syncthreads();
  int px, py;
  vrb_cta_origin(part, fr.w, 8, 8, px, py);
  px += threadIdx.x; py += threadIdx.y;
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
        float density = GT_SAMPLE(vol, kx, ky, kz, sp.x, sp.y, sp.z);
        float4 src = vrb_sample_tf(s_tf, tf_n, density);
        if (COUNT) ++ns;
        bool done = false;
        if (src.w > 0.0f) {
          float ka = 0.0f, kd = 0.0f, ks = 0.0f, IOcc = 0.0f, ISdw = 0.0f;
          if (C.P.apply_occlusion == 1) {
            ka = C.ka;
            g3 v_dir = gnrm(-cdir);
            IOcc = gt_cone(C, vol, kx, ky, kz, s_tfw, tf_n, C.occ_rays, C.P.occ_num_rays, C.P.occ_cone_distance, sp, v_right, v_up, v_dir, nsec);
          }
          if (C.P.apply_shadow == 1) {
            kd = C.kd; ks = C.ph.ks;
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
          float rr, gg, bb;
          if (PHONG) {                                   // ApplyGradientPhongShading == 1 (:277-295), specular colour vec3(1)
            rr = src.x; gg = src.y; bb = src.z;
            float dot_diff, spec;
            if (vrb_phong_terms(vol, C.ph, kx, ky, kz, sp.x, sp.y, sp.z, cam.ex, cam.ey, cam.ez, dot_diff, spec)) {
              const float f = ((1.0f / (ka + kd)) * (IOcc * ka + ISdw * kd * dot_diff));
              const float sc = (ISdw * ks * spec);
              rr = src.x * f + 1.0f * sc; gg = src.y * f + 1.0f * sc; bb = src.z * f + 1.0f * sc;
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
    } else if (fr.zero_miss) vrb_store_pixel(fr, px, py, 0.f, 0.f, 0.f, 0.f);
  }
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) { ns += __shfl_xor_sync(0xffffffffu, ns, o); nsec += __shfl_xor_sync(0xffffffffu, nsec, o); }
    if (((threadIdx.y * 8 + threadIdx.x) & 31) == 0 && ns) { atomicAdd(counter, (unsigned long long)ns); atomicAdd(counter + 1, nsec); }
  }
}


static int gt_launch(vrb_ctx* c, const vrb_camera* cam, const GtConst& C, int count_samples) {
  PartView part;
  dim3 block(8, 8), grid = vrb_make_grid(c, 8, 8, &part);
  size_t smem = (size_t)(c->tf_n + 2) * (sizeof(float4) + sizeof(float));
  VrbKernelTimer timer(c, "k_gt");
  if (C.ph.grad) {
    if (count_samples) k_gt<true, true><<<grid, block, smem, c->stream>>>(c->vol_view(), c->d_tf_rgbt, c->tf_n, c->frame_view(), make_cam_view(cam), part, C, c->d_counter);
    else               k_gt<false, true><<<grid, block, smem, c->stream>>>(c->vol_view(), c->d_tf_rgbt, c->tf_n, c->frame_view(), make_cam_view(cam), part, C, c->d_counter);
  } else {
    if (count_samples) k_gt<true, false><<<grid, block, smem, c->stream>>>(c->vol_view(), c->d_tf_rgbt, c->tf_n, c->frame_view(), make_cam_view(cam), part, C, c->d_counter);
    else               k_gt<false, false><<<grid, block, smem, c->stream>>>(c->vol_view(), c->d_tf_rgbt, c->tf_n, c->frame_view(), make_cam_view(cam), part, C, c->d_counter);
  }
  VRB_CUDA(cudaGetLastError());
  return VRB_OK;
}
#undef GT_SAMPLE
