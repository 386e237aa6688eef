// march_gt_deferred.cuh -- the shading kernel of the exact-filter rc1pcrtgt frame (shade_list.cuh, march_list.cu).
// Included by march_gt.cu (-fmad=false) after march_gt_common.cuh.
//
// A shaded sample of the ground-truth renderer costs N_occ + N_sdw secondary rays of up to dist / step trapezoid steps
// (ConeOcclusionEvaluationRayCasting / ConeShadowsEvaluationRayCasting, gt_ray_marching.comp:109-252), and the rays of one
// sample end after very different numbers of steps (they leave the volume, or saturate at 1 - Vt > 0.99).  With one thread per
// primary ray (round 1's k_gt) a warp ran at 18 of 32 active lanes.  Here the unit of work is one SECONDARY RAY:
//   * a warp takes a group of consecutive list entries (neighbouring samples) = G x (N_occ + N_sdw) ray tasks;
//   * every lane marches one task; when at least a quarter of the lanes have finished, the idle lanes take the next tasks
//     of the group together (one ballot, no atomics), so the step loop runs with 24-32 active lanes until the group drains;
//   * each ray's transmittance Vt and cosine weight go to shared memory; when the group is done one lane per entry and cone
//     adds them up IN RAY ORDER (S += Vt * w, Sw += w: the shader's own association) and writes ShadeSample's colour
//     (:254-302) into the entry.  Groups are handed out through an atomic cursor, so no warp is left with a long tail.
// Every ray's own arithmetic is the shader's, operation by operation; volume taps read the 2x2 quad copy of the volume
// (two LDG.64 per trilinear footprint) when it exists.
namespace gt_deferred {

#define GT_TASK_CAP 256          // ray tasks a warp holds results for (shared memory is taken from the L1 the ray taps live in: keep it small)
#define GT_MAX_GROUP 8           // entries per group (one finishing lane each)
#define GT_WARPS 4

struct GroupSmem {
  float vt[GT_TASK_CAP];         // transmittance of the task's ray
  float rw[GT_TASK_CAP];         // dot(cone axis, ray direction)
  float frame[GT_MAX_GROUP][21]; // per entry: tx (3), occlusion frame right/up/dir (9), shadow frame right/up/dir (9)
  unsigned char dark[GT_MAX_GROUP];
};

template <bool QUAD>
__device__ __forceinline__ float gt_ext(const VolView& vol, const uint2* __restrict__ vq, float kx, float ky, float kz,
                                        const float* __restrict__ tfw, int tf_n, g3 p) {
  int ix, iy, iz; float fx, fy, fz;
  vrb_volume_coords(vol, kx, ky, kz, p.x, p.y, p.z, ix, iy, iz, fx, fy, fz);
  const float d = QUAD ? vrb_fetch_volume_quad(vol, vq, ix, iy, iz, fx, fy, fz) : vrb_fetch_volume(vol, ix, iy, iz, fx, fy, fz);
  float up = fmaf(d, (float)tf_n, 0.5f);
  up = fminf(fmaxf(up, 0.0f), (float)tf_n + 0.5f);
  float fl; const int i = vrb_floor_pos(up, &fl);
  return vrb_lerp(tfw[i], tfw[i + 1], up - fl);
}

template <bool PHONG, bool QUAD>
__global__ void __launch_bounds__(GT_WARPS * 32)
k_gt_shade(VolView vol, const uint2* __restrict__ volq, const float4* __restrict__ tf_g, int tf_n, FrameView fr, CamView cam,
           const __grid_constant__ GtConst C, ShadeListView L, unsigned n_entries, int group, int count, unsigned long long* counter) {
  extern __shared__ float s_raw[];
  float* s_tfw = s_raw;                                                    // tf_n + 2 extinction floats
  GroupSmem* groups = reinterpret_cast<GroupSmem*>(s_raw + ((tf_n + 2 + 3) & ~3));
  for (int i = threadIdx.x; i < tf_n + 2; i += blockDim.x) s_tfw[i] = tf_g[i].w;
  __syncthreads();
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  GroupSmem& S = groups[warp];
  const float kx = (float)vol.w / vol.gx, ky = (float)vol.h / vol.gy, kz = (float)vol.d / vol.gz;
  const g3 half = C.G * 0.5f;
  const int n_occ = (C.P.apply_occlusion == 1) ? C.P.occ_num_rays : 0;
  const int n_sdw = (C.P.apply_shadow == 1) ? C.P.sdw_num_rays : 0;
  const int NT = n_occ + n_sdw;
  const float gap = C.P.light_ray_initial_gap, lstep = C.P.light_ray_step_size;
  const unsigned n_groups = (n_entries + (unsigned)group - 1u) / (unsigned)group;
  unsigned long long nsteps = 0;
  for (;;) {
    unsigned g = 0;
    if (lane == 0) g = atomicAdd(&L.counters[2], 1u);
    g = __shfl_sync(0xffffffffu, g, 0);
    if (g >= n_groups) break;
    const unsigned e0 = g * (unsigned)group;
    const int nE = (int)min((unsigned)group, n_entries - e0);
    // ---- per entry: position and the two cone frames (ShadeSample, :254-276; main, :385-396)
    if ((int)lane < nE) {
      const float4 A = L.a[e0 + lane];
      const bool valid = __float_as_int(A.w) >= 0;      // slots reserved but never written keep pixel == -1 (shade_list.cuh)
      const g3 sp = gm(A.x, A.y, A.z);
      float cdx, cdy, cdz;
      vrb_list_camera_dir(cam, fr.w, fr.h, valid ? __float_as_int(A.w) : 0, cdx, cdy, cdz);
      const g3 cdir = gm(cdx, cdy, cdz);
      const g3 v_right = gnrm(gcross(cdir, gm(0.f, 1.f, 0.f)));
      const g3 v_up = gnrm(gcross(-cdir, v_right));
      const g3 v_dir = gnrm(-cdir);
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
      float* f = S.frame[lane];
      f[0] = sp.x; f[1] = sp.y; f[2] = sp.z;
      f[3] = v_right.x; f[4] = v_right.y; f[5] = v_right.z; f[6] = v_up.x; f[7] = v_up.y; f[8] = v_up.z; f[9] = v_dir.x; f[10] = v_dir.y; f[11] = v_dir.z;
      f[12] = l_right.x; f[13] = l_right.y; f[14] = l_right.z; f[15] = l_up.x; f[16] = l_up.y; f[17] = l_up.z; f[18] = l_dir.x; f[19] = l_dir.y; f[20] = l_dir.z;
      S.dark[lane] = (dark ? 1 : 0) | (valid ? 0 : 2);
    }
    __syncwarp();
    // ---- the group's ray tasks
    const int ntasks = nE * NT;
    int next = 0, task = -1;
    bool have = false;
    g3 tx = gm(0.f, 0.f, 0.f), w = tx;
    float Vt = 1.0f, sp_ = 0.0f, st0 = 0.0f, dist_eval = 0.0f;
    for (;;) {
      const unsigned idle = __ballot_sync(0xffffffffu, !have);
      // refill when a quarter of the lanes idle (or nothing runs): the set-up code is paid once for many lanes
      if (next < ntasks && (__popc(idle) >= 8 || idle == 0xffffffffu)) {
        const int mine = next + __popc(idle & ((1u << lane) - 1u));
        next += __popc(idle);
        if (!have && mine < ntasks) {
          task = mine;
          const int en = task / NT, r = task - en * NT;
          const bool occ = r < n_occ;
          const int rayid = occ ? r : r - n_occ;
          const float* f = S.frame[en];
          tx = gm(f[0], f[1], f[2]);
          const float* tb = (occ ? C.occ_rays : C.sdw_rays) + 3 * rayid;
          const g3 c = gm(__ldg(tb), __ldg(tb + 1), __ldg(tb + 2));
          const float* q = f + (occ ? 3 : 12);
          const g3 a_r = gm(q[0], q[1], q[2]), a_u = gm(q[3], q[4], q[5]), a_d = gm(q[6], q[7], q[8]);
          w = gnrm(a_r * c.x + a_u * c.y + a_d * c.z);
          S.rw[task] = gdot(a_d, w);
          dist_eval = occ ? C.P.occ_cone_distance : C.P.sdw_cone_distance;
          Vt = 1.0f; sp_ = gap;
          if ((!occ && (S.dark[en] & 1)) || (S.dark[en] & 2)) { S.vt[task] = 0.0f; }       // never read: unused slot, or the shadow term of a dark sample (0)
          else {
            st0 = gt_ext<QUAD>(vol, volq, kx, ky, kz, s_tfw, tf_n, tx + w * gap);
            if (gap < dist_eval) have = true; else S.vt[task] = Vt;
          }
        }
      }
      if (!__any_sync(0xffffffffu, have)) { if (next >= ntasks) break; else continue; }
      if (have) {
        // one trapezoid step of the ray (:133-166, :217-249)
        const float hh = fminf(lstep, dist_eval - sp_);
        const g3 at = tx + w * (sp_ + hh);
        if (at.x < 0.0f || at.x > C.G.x || at.y < 0.0f || at.y > C.G.y || at.z < 0.0f || at.z > C.G.z) have = false;
        else {
          const float st1 = gt_ext<QUAD>(vol, volq, kx, ky, kz, s_tfw, tf_n, at);
          Vt *= expf(-((st0 + st1) * 0.5f) * hh);
          ++nsteps;
          if ((1 - Vt) > 0.99f) have = false;
          else { st0 = st1; sp_ = sp_ + hh; have = sp_ < dist_eval; }
        }
        if (!have) S.vt[task] = Vt;
      }
    }
    __syncwarp();
    // ---- per entry: the cosine-weighted means in ray order, then the colour (:254-302)
    if ((int)lane < nE && !(S.dark[lane] & 2)) {
      const unsigned e = e0 + lane;
      const float4 B = L.b[e];
      float ka = 0.0f, kd = 0.0f, ks = 0.0f, IOcc = 0.0f, ISdw = 0.0f;
      const float* vt = S.vt + (int)lane * NT;
      const float* rw = S.rw + (int)lane * NT;
      if (C.P.apply_occlusion == 1) {
        ka = C.ka;
        float Sv = 0.0f, Sw = 0.0f;
        for (int j = 0; j < n_occ; ++j) { Sv += vt[j] * rw[j]; Sw += rw[j]; }
        IOcc = (Sv / Sw);
      }
      if (C.P.apply_shadow == 1) {
        kd = C.kd; ks = C.ph.ks;
        if (!(S.dark[lane] & 1)) {
          float Sv = 0.0f, Sw = 0.0f;
          for (int j = 0; j < n_sdw; ++j) { Sv += vt[n_occ + j] * rw[n_occ + j]; Sw += rw[n_occ + j]; }
          ISdw = (Sv / Sw);
        }
      }
      float rr, gg, bb;
      if (PHONG) {                                   // ApplyGradientPhongShading == 1 (:277-295), specular colour vec3(1)
        rr = B.x; gg = B.y; bb = B.z;
        const float* f = S.frame[lane];
        float dot_diff, spec;
        if (vrb_phong_terms(vol, C.ph, kx, ky, kz, f[0], f[1], f[2], cam.ex, cam.ey, cam.ez, dot_diff, spec)) {
          const float ff = ((1.0f / (ka + kd)) * (IOcc * ka + ISdw * kd * dot_diff));
          const float sc = (ISdw * ks * spec);
          rr = B.x * ff + 1.0f * sc; gg = B.y * ff + 1.0f * sc; bb = B.z * ff + 1.0f * sc;
        }
      } else {
        const float kk = (1.0f / (ka + kd));
        rr = kk * (B.x * IOcc * ka + B.x * ISdw * kd);
        gg = kk * (B.y * IOcc * ka + B.y * ISdw * kd);
        bb = kk * (B.z * IOcc * ka + B.z * ISdw * kd);
      }
      L.b[e] = make_float4(rr, gg, bb, B.w);
    }
    __syncwarp();
  }
  if (count) {
    for (int o = 16; o > 0; o >>= 1) nsteps += __shfl_xor_sync(0xffffffffu, nsteps, o);
    if (lane == 0 && nsteps) atomicAdd(counter + 1, nsteps);
  }
}


// ---- variant "entry": one list entry per lane, the 32 lanes of a warp trace the SAME secondary-ray index from 32
// neighbouring samples at the same time (parallel rays one or two voxels apart: their taps share cache lines), ILP rays in
// flight per lane.  Lanes only idle while the longest of the 32 parallel rays finishes.  Same per-ray arithmetic.
template <bool QUAD, int ILP>
__device__ float gt_cone_e(const GtConst& C, const VolView& vol, const uint2* __restrict__ vq, float kx, float ky, float kz,
                           const float* __restrict__ tfw, int tf_n, const float* __restrict__ table, int nrays, float dist_eval,
                           g3 tx, g3 v_right, g3 v_up, g3 v_dir, unsigned long long& nsteps) {
  float S = 0.0f, Sw = 0.0f;
  const float gap = C.P.light_ray_initial_gap, lstep = C.P.light_ray_step_size;
  for (int ray0 = 0; ray0 < nrays; ray0 += ILP) {
    g3 w[ILP];
    float Vt[ILP], sp[ILP], st0[ILP];
    bool act[ILP];
    bool any = false;
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
      const int rayid = min(ray0 + j, nrays - 1);
      const g3 c = gm(__ldg(table + 3 * rayid), __ldg(table + 3 * rayid + 1), __ldg(table + 3 * rayid + 2));
      w[j] = gnrm(v_right * c.x + v_up * c.y + v_dir * c.z);
      Vt[j] = 1.0f; sp[j] = gap; st0[j] = 0.0f;
      act[j] = (ray0 + j < nrays);
      if (act[j]) st0[j] = gt_ext<QUAD>(vol, vq, kx, ky, kz, tfw, tf_n, tx + w[j] * gap);
      act[j] = act[j] && (gap < dist_eval);
      any = any || act[j];
    }
    while (any) {
      any = false;
      float hh[ILP], st1[ILP];
      bool inside[ILP];
#pragma unroll
      for (int j = 0; j < ILP; ++j) {
        hh[j] = fminf(lstep, dist_eval - sp[j]);
        const g3 at = tx + w[j] * (sp[j] + hh[j]);
        inside[j] = !(at.x < 0.0f || at.x > C.G.x || at.y < 0.0f || at.y > C.G.y || at.z < 0.0f || at.z > C.G.z);
        const bool use = act[j] && inside[j];
        st1[j] = 0.0f;
        if (ILP == 1) { if (use) st1[j] = gt_ext<QUAD>(vol, vq, kx, ky, kz, tfw, tf_n, at); }
        else st1[j] = gt_ext<QUAD>(vol, vq, kx, ky, kz, tfw, tf_n, gm(use ? at.x : tx.x, use ? at.y : tx.y, use ? at.z : tx.z));
      }
#pragma unroll
      for (int j = 0; j < ILP; ++j) {
        if (act[j]) {
          if (!inside[j]) act[j] = false;
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
    for (int j = 0; j < ILP; ++j) {
      if (ray0 + j < nrays) {
        const float rw = gdot(v_dir, w[j]);
        S += Vt[j] * rw;
        Sw += rw;
      }
    }
  }
  return (S / Sw);
}

template <bool PHONG, bool QUAD, int ILP>
__global__ void __launch_bounds__(128)
k_gt_shade_e(VolView vol, const uint2* __restrict__ volq, const float4* __restrict__ tf_g, int tf_n, FrameView fr, CamView cam,
             const __grid_constant__ GtConst C, ShadeListView L, unsigned n_entries, int count, unsigned long long* counter) {
  extern __shared__ float s_raw[];
  float* s_tfw = s_raw;
  for (int i = threadIdx.x; i < tf_n + 2; i += blockDim.x) s_tfw[i] = tf_g[i].w;
  __syncthreads();
  const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
  const float4 A = (e < n_entries) ? L.a[e] : make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
  unsigned long long nsteps = 0;
  if (__float_as_int(A.w) >= 0) {
    const float kx = (float)vol.w / vol.gx, ky = (float)vol.h / vol.gy, kz = (float)vol.d / vol.gz;
    const g3 half = C.G * 0.5f;
    const float4 B = L.b[e];
    const g3 sp = gm(A.x, A.y, A.z);
    float cdx, cdy, cdz;
    vrb_list_camera_dir(cam, fr.w, fr.h, __float_as_int(A.w), cdx, cdy, cdz);
    const g3 cdir = gm(cdx, cdy, cdz);
    const g3 v_right = gnrm(gcross(cdir, gm(0.f, 1.f, 0.f)));
    const g3 v_up = gnrm(gcross(-cdir, v_right));
    float ka = 0.0f, kd = 0.0f, ks = 0.0f, IOcc = 0.0f, ISdw = 0.0f;
    if (C.P.apply_occlusion == 1) {
      ka = C.ka;
      const g3 v_dir = gnrm(-cdir);
      IOcc = gt_cone_e<QUAD, ILP>(C, vol, volq, kx, ky, kz, s_tfw, tf_n, C.occ_rays, C.P.occ_num_rays, C.P.occ_cone_distance, sp, v_right, v_up, v_dir, nsteps);
    }
    if (C.P.apply_shadow == 1) {
      kd = C.kd; ks = C.ph.ks;
      g3 l_dir = gm(0.f, 0.f, 0.f), l_up = l_dir, l_right = l_dir;
      bool dark = false;
      if (C.P.shadow_type == 0 || C.P.shadow_type == 1) {
        l_dir = gnrm(C.light_pos - (sp - half));
        l_up = gnrm(gcross(l_dir, C.light_right));
        l_right = gnrm(gcross(l_dir, l_up));
        if (C.P.shadow_type == 1 && gdot(l_dir, C.light_fwd) < 30.0f) dark = true;      // reference quirk (gt_ray_marching.comp:192)
      } else if (C.P.shadow_type == 2) {
        l_dir = C.light_fwd; l_up = C.light_up; l_right = C.light_right;
      }
      ISdw = dark ? 0.0f : gt_cone_e<QUAD, ILP>(C, vol, volq, kx, ky, kz, s_tfw, tf_n, C.sdw_rays, C.P.sdw_num_rays, C.P.sdw_cone_distance, sp, l_right, l_up, l_dir, nsteps);
    }
    float rr, gg, bb;
    if (PHONG) {
      rr = B.x; gg = B.y; bb = B.z;
      float dot_diff, spec;
      if (vrb_phong_terms(vol, C.ph, kx, ky, kz, sp.x, sp.y, sp.z, cam.ex, cam.ey, cam.ez, dot_diff, spec)) {
        const float ff = ((1.0f / (ka + kd)) * (IOcc * ka + ISdw * kd * dot_diff));
        const float sc = (ISdw * ks * spec);
        rr = B.x * ff + 1.0f * sc; gg = B.y * ff + 1.0f * sc; bb = B.z * ff + 1.0f * sc;
      }
    } else {
      const float kk = (1.0f / (ka + kd));
      rr = kk * (B.x * IOcc * ka + B.x * ISdw * kd);
      gg = kk * (B.y * IOcc * ka + B.y * ISdw * kd);
      bb = kk * (B.z * IOcc * ka + B.z * ISdw * kd);
    }
    L.b[e] = make_float4(rr, gg, bb, B.w);
  }
  if (count) {
    for (int o = 16; o > 0; o >>= 1) nsteps += __shfl_xor_sync(0xffffffffu, nsteps, o);
    if ((threadIdx.x & 31) == 0 && nsteps) atomicAdd(counter + 1, nsteps);
  }
}

}  // namespace gt_deferred
