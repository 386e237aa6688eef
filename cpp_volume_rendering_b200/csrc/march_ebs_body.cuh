// march_ebs_body.cuh -- device functions + kernel of the EBS marcher, included once per SAT layout (EBS_PACK = 1, 2, 4)
// from march_ebs.cu inside a namespace.  See march_ebs.cu for the description.
// One axis of a GL_LINEAR lookup: texel indices i0, i1 (clamped to the edge) and the blend weight f.
// For the packed SAT layouts the neighbour index is implicit (texel x carries x+1 with it); when the lookup falls off
// the low edge (i < 0, both indices clamp to 0) the weight is forced to 0 so that a + 0*(b-a) == a exactly, which is
// what lerp(S[0], S[0], f) gives in the linear layout.
struct Axis { int i0, i1; float f; int ic; };
__device__ __forceinline__ Axis sat_axis(float x, float inv, int n) {
  Axis a;
  float s = x * inv;
  float u = s * (float)n - 0.5f;
  float fl = floorf(u);
  a.f = u - fl;
  int i = (int)fl;
  a.i0 = min(max(i, 0), n - 1);
  a.i1 = min(max(i + 1, 0), n - 1);
  a.ic = min(max(i, -1), n - 1);     // unclamped low texel, limited to the one-texel gutter of the gather atlas
  return a;
}
// PACK 1: linear fp32 SAT, 8 scalar loads.  PACK 2: texel x holds {S[x], S[x+1]} (float2), 4 LDG.64.
// PACK 4: texel (x,y) holds {S[x,y], S[x+1,y], S[x,y+1], S[x+1,y+1]} (float4), 2 LDG.128.
// The blend arithmetic (and therefore every result bit) is identical in the three variants.

__device__ __forceinline__ float sat_tri(const EbsConst& E, const Axis& X, const Axis& Y, const Axis& Z) {
  if (EBS_PACK == 1) {
    const float* z0 = E.sat + (long long)Z.i0 * E.sslice;
    const float* z1 = E.sat + (long long)Z.i1 * E.sslice;
    const int r0 = Y.i0 * E.sw, r1 = Y.i1 * E.sw;
    float c00 = vrb_lerp(__ldg(z0 + r0 + X.i0), __ldg(z0 + r0 + X.i1), X.f);
    float c10 = vrb_lerp(__ldg(z0 + r1 + X.i0), __ldg(z0 + r1 + X.i1), X.f);
    float c01 = vrb_lerp(__ldg(z1 + r0 + X.i0), __ldg(z1 + r0 + X.i1), X.f);
    float c11 = vrb_lerp(__ldg(z1 + r1 + X.i0), __ldg(z1 + r1 + X.i1), X.f);
    return vrb_lerp(vrb_lerp(c00, c10, Y.f), vrb_lerp(c01, c11, Y.f), Z.f);
  } else if (EBS_PACK == 2) {
    const float2* base = reinterpret_cast<const float2*>(E.sat_packed);
    const float2* z0 = base + (long long)Z.i0 * E.sslice;
    const float2* z1 = base + (long long)Z.i1 * E.sslice;
    const int r0 = Y.i0 * E.sw, r1 = Y.i1 * E.sw;
    const float fx = (X.i1 == X.i0 && X.i0 == 0) ? 0.0f : X.f;
    float2 a = __ldg(z0 + r0 + X.i0), b = __ldg(z0 + r1 + X.i0), c = __ldg(z1 + r0 + X.i0), d = __ldg(z1 + r1 + X.i0);
    float c00 = vrb_lerp(a.x, a.y, fx), c10 = vrb_lerp(b.x, b.y, fx), c01 = vrb_lerp(c.x, c.y, fx), c11 = vrb_lerp(d.x, d.y, fx);
    return vrb_lerp(vrb_lerp(c00, c10, Y.f), vrb_lerp(c01, c11, Y.f), Z.f);
  } else if (EBS_PACK == 8) {
    // texture-unit path: the SAT lives in a 2-D cudaArray atlas (one (w+2)x(h+2) tile per z slice, replicated one-texel
    // gutter = clamp-to-edge); ONE tex2Dgather per z slice returns the four texels of the xy footprint unfiltered
    // (exact fp32), the blend stays in fp32 ALU in the oracle's order.  Gather order: x=(i0,j1) y=(i1,j1) z=(i1,j0) w=(i0,j0).
    const float gx = (float)(X.ic + 2), gy = (float)(Y.ic + 2);      // padded index (ic+1) + 1.0: midway between the 2x2 centres
    const int t0 = Z.i0 % E.atlas_tiles_x, u0 = Z.i0 / E.atlas_tiles_x;
    const int t1 = Z.i1 % E.atlas_tiles_x, u1 = Z.i1 / E.atlas_tiles_x;
    float4 a = tex2Dgather<float4>(E.sat_tex, gx + (float)(t0 * E.atlas_tile_w), gy + (float)(u0 * E.atlas_tile_h), 0);
    float4 c = tex2Dgather<float4>(E.sat_tex, gx + (float)(t1 * E.atlas_tile_w), gy + (float)(u1 * E.atlas_tile_h), 0);
    float c00 = vrb_lerp(a.w, a.z, X.f), c10 = vrb_lerp(a.x, a.y, X.f), c01 = vrb_lerp(c.w, c.z, X.f), c11 = vrb_lerp(c.x, c.y, X.f);
    return vrb_lerp(vrb_lerp(c00, c10, Y.f), vrb_lerp(c01, c11, Y.f), Z.f);
  } else {
    const float4* base = reinterpret_cast<const float4*>(E.sat_packed);
    const float fx = (X.i1 == X.i0 && X.i0 == 0) ? 0.0f : X.f;
    const float fy = (Y.i1 == Y.i0 && Y.i0 == 0) ? 0.0f : Y.f;
    const long long o = (long long)Y.i0 * E.sw + X.i0;
    float4 a = __ldg(base + (long long)Z.i0 * E.sslice + o), c = __ldg(base + (long long)Z.i1 * E.sslice + o);
    float c00 = vrb_lerp(a.x, a.y, fx), c10 = vrb_lerp(a.z, a.w, fx), c01 = vrb_lerp(c.x, c.y, fx), c11 = vrb_lerp(c.z, c.w, fx);
    return vrb_lerp(vrb_lerp(c00, c10, fy), vrb_lerp(c01, c11, fy), Z.f);
  }
}
// EvaluateSAT3D (ebs_ray_bbox_marching.comp:87-100)
__device__ __forceinline__ float eval_sat3d(const EbsConst& E, f3 p1, f3 p2) {
  Axis x1 = sat_axis(p1.x, E.inv_vol_scaled.x, E.sw), x2 = sat_axis(p2.x, E.inv_vol_scaled.x, E.sw);
  Axis y1 = sat_axis(p1.y, E.inv_vol_scaled.y, E.sh), y2 = sat_axis(p2.y, E.inv_vol_scaled.y, E.sh);
  Axis z1 = sat_axis(p1.z, E.inv_vol_scaled.z, E.sd), z2 = sat_axis(p2.z, E.inv_vol_scaled.z, E.sd);
  float V1 = sat_tri(E, x2, y2, z2);
  float V2 = sat_tri(E, x1, y2, z2);
  float V3 = sat_tri(E, x2, y2, z1);
  float V4 = sat_tri(E, x1, y2, z1);
  float V5 = sat_tri(E, x2, y1, z2);
  float V6 = sat_tri(E, x1, y1, z2);
  float V7 = sat_tri(E, x2, y1, z1);
  float V8 = sat_tri(E, x1, y1, z1);
  return (V1 - V2 - V3 + V4 - V5 + V6 + V7 - V8);
}
__device__ __forceinline__ float eval_ao_sat3d(const EbsConst& E, f3 p1, f3 p2) {
  p1 = clamp3(p1 + E.VS, E.MinSAT, E.MaxSAT);
  p2 = clamp3(p2 + E.VS, E.MinSAT, E.MaxSAT);
  return eval_sat3d(E, p1, p2);
}
// ExtinctionAmbientOcclusion (:112-146)
__device__ float ebs_ambient_occlusion(const EbsConst& E, f3 tx, unsigned int& nq) {
  nq += (unsigned int)E.P.amb_occ_shells;
  const float R = E.P.amb_occ_radius;
  float SAT_Sh0 = eval_ao_sat3d(E, tx - R * E.VS, tx + R * E.VS);
  float tshi = SAT_Sh0 * (1.0f / (R * R));
  float SAT_Shi = SAT_Sh0;
  for (int ith = 1; ith < E.P.amb_occ_shells; ++ith) {
    float r1 = R * (float)(ith + 1);
    float S1 = eval_ao_sat3d(E, tx - r1 * E.VS, tx + r1 * E.VS);
    tshi = tshi + (S1 - SAT_Shi) * (1.0f / (r1 * r1));
    SAT_Shi = S1;
  }
  float rshi = R * (float)E.P.amb_occ_shells;
  float W_A = 1.0f / (rshi * rshi);
  float Stau = W_A * tshi;
  return expf(-(Stau));
}
// EvaluateShadowSAT3D (:148-188, texture() branch)
__device__ __forceinline__ float eval_shadow_sat3d(const EbsConst& E, f3 p1, f3 p2) {
  float volquery = ((fabsf(p1.x - p2.x) / E.VS.x)) * ((fabsf(p1.y - p2.y) / E.VS.y)) * ((fabsf(p1.z - p2.z) / E.VS.z));
  p1 = clamp3(p1 + E.VS, E.MinSAT, E.MaxSAT);
  p2 = clamp3(p2 + E.VS, E.MinSAT, E.MaxSAT);
  return ((eval_sat3d(E, p1, p2) / volquery)) * E.P.sdw_ui_weight;
}

// ConeZAxis / ConeYAxis / ConeXAxis (:190-430) as ONE body over the marching axis AX.  The three shader functions differ
// only in which components play "marching axis" and "lateral": in each, the cone axis is projected onto the two planes that
// contain the marching axis, each projection is rotated by -/+ the cone angle (lateral' = lat cs - ax sn, axis' = lat sn +
// ax cs) and normalised; per box the lateral extents are those directions scaled to the box's mean depth, snapped outward
// to whole voxels.  A zero component adds exactly 0 to a norm, so working on the (lateral, axis) pairs gives the shader's
// bits for all three (tests/test_ebs_gpu.py::test_ebs_all_three_dominant_light_axes, full-size config 2).
struct L2 { float lat, ax; };
__device__ __forceinline__ L2 n2(float lat, float ax) {
  const float r = 1.0f / sqrtf(lat * lat + ax * ax);
  L2 o; o.lat = lat * r; o.ax = ax * r; return o;
}
template <int I> __device__ __forceinline__ float comp(const f3& a) { return I == 0 ? a.x : (I == 1 ? a.y : a.z); }
template <int A, int U, int V> __device__ __forceinline__ f3 from_axes(float a, float u, float v) {
  f3 o;
  o.x = (A == 0) ? a : (U == 0 ? u : v);
  o.y = (A == 1) ? a : (U == 1 ? u : v);
  o.z = (A == 2) ? a : (U == 2 ? u : v);
  return o;
}
template <int AX>
__device__ float ebs_cone_axis(const EbsConst& E, f3 pos, f3 cv, unsigned int& nq) {
  constexpr int U = (AX + 1) % 3, V = (AX + 2) % 3;
  const float ca = comp<AX>(cv);
  float Stau = 0.0f;
  float signal = 1.0f; if (ca < 0) signal = -1.0f;
  const L2 pu = n2(comp<U>(cv), ca), pv = n2(comp<V>(cv), ca);
  const L2 u1 = n2(pu.lat * E.n_cs - pu.ax * E.n_sn, pu.lat * E.n_sn + pu.ax * E.n_cs);
  const L2 u2 = n2(pu.lat * E.p_cs - pu.ax * E.p_sn, pu.lat * E.p_sn + pu.ax * E.p_cs);
  const L2 v1 = n2(pv.lat * E.n_cs - pv.ax * E.n_sn, pv.lat * E.n_sn + pv.ax * E.n_cs);
  const L2 v2 = n2(pv.lat * E.p_cs - pv.ax * E.p_sn, pv.lat * E.p_sn + pv.ax * E.p_cs);
  const float vsa = comp<AX>(E.VS), vsu = comp<U>(E.VS), vsv = comp<V>(E.VS);
  const float pa = comp<AX>(pos), lo = comp<AX>(E.MinVol), hi = comp<AX>(E.MaxVol);
  const float si = E.P.sdw_sample_interval * signal * vsa;
  float a_pos = E.P.sdw_initial_step * signal * vsa;
  while ((a_pos / ca) < E.P.sdw_cone_max_distance && (pa + (a_pos + si) > lo && pa + (a_pos + si) < hi)) {
    const float a_mean = fabsf(a_pos + si * 0.5f);
    const float p_u1 = u1.lat * (a_mean / fabsf(u1.ax)), p_u2 = u2.lat * (a_mean / fabsf(u2.ax));
    const float p_v1 = v1.lat * (a_mean / fabsf(v1.ax)), p_v2 = v2.lat * (a_mean / fabsf(v2.ax));
    float ulo = fminf(p_u1, p_u2), uhi = fmaxf(p_u1, p_u2);
    float vlo = fminf(p_v1, p_v2), vhi = fmaxf(p_v1, p_v2);
    const float udiff = fabsf(uhi - ulo), vdiff = fabsf(vhi - vlo);
    const float us = (ceilf(udiff / vsu) - (udiff / vsu)) * 0.5f;
    const float vs = (ceilf(vdiff / vsv) - (vdiff / vsv)) * 0.5f;
    ulo = ulo - us * vsu; uhi = uhi + us * vsu;
    vlo = vlo - vs * vsv; vhi = vhi + vs * vsv;
    const float alo = fminf(a_pos, a_pos + si), ahi = fmaxf(a_pos, a_pos + si);
    ++nq;
    Stau += eval_shadow_sat3d(E, pos + from_axes<AX, U, V>(alo, ulo, vlo), pos + from_axes<AX, U, V>(ahi, uhi, vhi));
    a_pos = a_pos + si;
  }
  return Stau;
}
// ExtinctionDirectionalShadows (:432-456)
__device__ float ebs_directional_shadows(const EbsConst& E, f3 tx, unsigned int& nq) {
  f3 realpos = tx - (E.VSS * 0.5f);
  f3 cone_vec = mk3(0.f, 0.f, 0.f);
  if (E.P.type_of_shadow == 0) cone_vec = norm3(E.light_pos - realpos);
  else if (E.P.type_of_shadow == 1) cone_vec = norm3(E.light_fwd);
  float ax = fabsf(cone_vec.x), ay = fabsf(cone_vec.y), az = fabsf(cone_vec.z);
  float Stau;
  if (az > ax && az > ay) Stau = ebs_cone_axis<2>(E, tx, cone_vec, nq);
  else if (ay > ax) Stau = ebs_cone_axis<1>(E, tx, cone_vec, nq);
  else Stau = ebs_cone_axis<0>(E, tx, cone_vec, nq);
  return expf(-Stau);
}

template <bool COUNT>
__global__ void __launch_bounds__(64, EBS_MIN_BLOCKS)
k_ebs(VolView vol, const float4* __restrict__ tf_g, int tf_n, FrameView fr, CamView cam, PartView part, EbsConst E,
      unsigned long long* counter) {
  extern __shared__ float4 s_tf[];
  const float4* tf = tf_g;
  if (tf_n + 2 <= 1026) {
    for (int i = threadIdx.y * 8 + threadIdx.x; i < tf_n + 2; i += 64) s_tf[i] = tf_g[i];
    __syncthreads();
    tf = s_tf;
  }
  int px, py;
  vrb_cta_origin(part, fr.w, 8, 8, px, py);
  px += threadIdx.x; py += threadIdx.y;
  unsigned int ns = 0, nq = 0;
  if (px < fr.w && py < fr.h && vrb_owns_pixel(part, px, py, fr.w)) {
    Ray r = vrb_make_ray(cam, px, py, fr.w, fr.h, E.VSS.x, E.VSS.y, E.VSS.z);
    if (r.hit) {
      float D = fabsf(r.tfar - r.tnear);
      f3 dir = mk3(r.dx, r.dy, r.dz);
      f3 wd = mk3(r.ox, r.oy, r.oz) + dir * r.tnear;
      wd = wd + (E.VSS * 0.5f);
      float kx = (float)vol.w / vol.gx, ky = (float)vol.h / vol.gy, kz = (float)vol.d / vol.gz;
      float dr = 0.f, dg = 0.f, db = 0.f, da = 0.f;
      const float step = E.P.step_size;
      for (float s = 0.0f; s < D;) {
        float h = fminf(step, D - s);
        f3 tx = wd + dir * (s + h * 0.5f);
        float density = vrb_sample_volume(vol, kx, ky, kz, tx.x, tx.y, tx.z);
        float4 src = vrb_sample_tf(tf, tf_n, density);
        if (COUNT) ++ns;
        if (src.w > 0.0f) {
          // ShadeSample (:500-551)
          float ka = 0.0f, kd = 0.0f, ks = 0.0f, IOcc = 0.0f, ISdw = 0.0f;
          if (E.P.apply_occlusion == 1) { ka = E.ka; IOcc = ebs_ambient_occlusion(E, tx, nq); }
          if (E.P.apply_shadow == 1) { kd = E.kd; ks = E.ph.ks; ISdw = ebs_directional_shadows(E, tx, nq); }
          float cr, cg, cb;
          if (E.ph.grad) {                                   // ApplyPhongShading == 1 (:524-544); a zero gradient leaves L = clr
            cr = src.x; cg = src.y; cb = src.z;
            float dot_diff, spec;
            if (vrb_phong_terms(vol, E.ph, kx, ky, kz, tx.x, tx.y, tx.z, cam.ex, cam.ey, cam.ez, dot_diff, spec)) {
              float k = (1.0f / (ka + kd));
              cr = k * (src.x * IOcc * ka + ISdw * (src.x * kd * dot_diff)) + ISdw * (ks * E.ph.isx * spec);
              cg = k * (src.y * IOcc * ka + ISdw * (src.y * kd * dot_diff)) + ISdw * (ks * E.ph.isy * spec);
              cb = k * (src.z * IOcc * ka + ISdw * (src.z * kd * dot_diff)) + ISdw * (ks * E.ph.isz * spec);
            }
          } else {
            float k = (1.0f / (ka + kd));
            cr = k * (src.x * IOcc * ka + src.x * ISdw * kd);
            cg = k * (src.y * IOcc * ka + src.y * ISdw * kd);
            cb = k * (src.z * IOcc * ka + src.z * ISdw * kd);
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
    unsigned long long nq64 = nq;
    for (int o = 16; o > 0; o >>= 1) { ns += __shfl_xor_sync(0xffffffffu, ns, o); nq64 += __shfl_xor_sync(0xffffffffu, nq64, o); }
    if (((threadIdx.y * 8 + threadIdx.x) & 31) == 0 && ns) { atomicAdd(counter, (unsigned long long)ns); atomicAdd(counter + 1, nq64); }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// M lanes per ray.  The lighting of a sample (AO shells + shadow boxes: hundreds of dependent-latency SAT queries) does
// not depend on the compositing state, so the M lanes of a ray shade M CONSECUTIVE samples at the same time and the
// group then composites them in ray order (every lane redundantly, values exchanged with shuffles).  Same arithmetic
// per sample, same compositing order and 0.99 cut => identical pixels; samples shaded past the cut are discarded (at
// most M-1 per ray).  Why: with one thread per ray a CTA lives for milliseconds (ms-long dependent chains), which caps
// the speed-up when the frame is split over GPUs; M lanes cut that chain by M, and the M samples of a ray are 0.5 voxel
// apart, so one gather instruction touches fewer distinct texel quads.
// The round loop of the cooperative marcher: M lanes (gbase .. gbase+M-1 of the warp) shade the next M samples of one ray
// and composite them in order.  State (s, colour, done) is identical on all lanes of the group.
template <bool COUNT, int M>
__device__ __forceinline__ void ebs_coop_rounds(const VolView& vol, const float4* __restrict__ tf, int tf_n, const EbsConst& E,
                                                f3 wd, f3 dir, float D, float kx, float ky, float kz, int sub, int gbase,
                                                bool& done, float& s, float& dr, float& dg, float& db, float& da,
                                                unsigned int& ns, unsigned long long& nq_used) {
  const float step = E.P.step_size;
  while (!__all_sync(0xffffffffu, done)) {
    // the next M ray parameters, by the same sequential additions as the one-sample-at-a-time loop
    float my_s = 0.f, my_h = 0.f; bool my_valid = false;
    float sj = s;
#pragma unroll
    for (int j = 0; j < M; ++j) {
      bool v = sj < D;
      float h = fminf(step, D - sj);
      if (j == sub) { my_valid = v; my_s = sj; my_h = h; }
      if (v) sj = sj + h;
    }
    int flag = 0;                                    // 0: past the end of the ray, 1: transparent sample, 2: shaded sample
    float tr = 0.f, tg = 0.f, tb = 0.f, ta = 0.f;
    unsigned int nq = 0;
    if (!done && my_valid) {
      f3 tx = wd + dir * (my_s + my_h * 0.5f);
      float density = vrb_sample_volume(vol, kx, ky, kz, tx.x, tx.y, tx.z);
      float4 src = vrb_sample_tf(tf, tf_n, density);
      flag = 1;
      if (src.w > 0.0f) {
        float ka = 0.0f, kd = 0.0f, IOcc = 0.0f, ISdw = 0.0f;
        if (E.P.apply_occlusion == 1) { ka = E.ka; IOcc = ebs_ambient_occlusion(E, tx, nq); }
        if (E.P.apply_shadow == 1) { kd = E.kd; ISdw = ebs_directional_shadows(E, tx, nq); }
        float k = (1.0f / (ka + kd));
        float cr = k * (src.x * IOcc * ka + src.x * ISdw * kd);
        float cg = k * (src.y * IOcc * ka + src.y * ISdw * kd);
        float cb = k * (src.z * IOcc * ka + src.z * ISdw * kd);
        float a = 1.0f - expf(-src.w * my_h);
        tr = cr * a; tg = cg * a; tb = cb * a; ta = a;
        flag = 2;
      }
    }
    // ordered compositing of the M samples (all lanes of the group keep identical state)
#pragma unroll
    for (int j = 0; j < M; ++j) {
      const int src_lane = gbase + j;
      int f = __shfl_sync(0xffffffffu, flag, src_lane);
      float r_ = __shfl_sync(0xffffffffu, tr, src_lane), g_ = __shfl_sync(0xffffffffu, tg, src_lane);
      float b_ = __shfl_sync(0xffffffffu, tb, src_lane), a_ = __shfl_sync(0xffffffffu, ta, src_lane);
      if (!done) {
        if (f == 0) { done = true; }
        else {
          if (COUNT) { ++ns; if (j == sub) nq_used += nq; }
          if (f == 2) {
            float om = 1.0f - da;
            dr = dr + om * r_; dg = dg + om * g_; db = db + om * b_; da = da + om * a_;
            if (da > 0.99f) done = true;
          }
        }
      }
    }
    s = sj;
  }
}

template <bool COUNT, int M>
__global__ void __launch_bounds__(64, EBS_MIN_BLOCKS)
k_ebs_coop(VolView vol, const float4* __restrict__ tf_g, int tf_n, FrameView fr, CamView cam, PartView part, EbsConst E,
           unsigned long long* counter, const unsigned int* __restrict__ cta_order, unsigned int* __restrict__ cta_cost) {
  extern __shared__ float4 s_tf[];
  const float4* tf = tf_g;
  const int tid = threadIdx.x;
  const long long t_start = cta_cost ? clock64() : 0;
  // launch slot -> logical CTA: heaviest CTAs of the previous frame first (cta_order.cu)
  const unsigned int cta = cta_order ? __ldg(cta_order + blockIdx.x) : blockIdx.x;
  if (tf_n + 2 <= 1026) {
    for (int i = tid; i < tf_n + 2; i += 64) s_tf[i] = tf_g[i];
    __syncthreads();
    tf = s_tf;
  }
  constexpr int RPB = 64 / M;                       // rays per CTA
  constexpr int TW = (M >= 16) ? 2 : (M >= 4) ? 4 : 8, TH = RPB / TW;   // pixel tile of the CTA; a warp covers TW x (TH/2)
  const int ray = tid / M, sub = tid % M;
  const int lane = tid & 31, gbase = lane - sub;    // first lane of this ray's group
  int px, py;
  vrb_cta_origin_linear(part, fr.w, fr.h, TW, TH, cta, gridDim.x, cta_order != nullptr, px, py);
  px += ray % TW; py += ray / TW;
  unsigned int ns = 0;
  unsigned long long nq_used = 0;
  bool done = true;
  float D = 0.f, kx = 0.f, ky = 0.f, kz = 0.f;
  f3 dir = mk3(0.f, 0.f, 0.f), wd = dir;
  bool hit = false;
  if (px < fr.w && py < fr.h && vrb_owns_pixel(part, px, py, fr.w)) {
    Ray r = vrb_make_ray(cam, px, py, fr.w, fr.h, E.VSS.x, E.VSS.y, E.VSS.z);
    if (r.hit) {
      hit = true; done = false;
      D = fabsf(r.tfar - r.tnear);
      dir = mk3(r.dx, r.dy, r.dz);
      wd = mk3(r.ox, r.oy, r.oz) + dir * r.tnear;
      wd = wd + (E.VSS * 0.5f);
      kx = (float)vol.w / vol.gx; ky = (float)vol.h / vol.gy; kz = (float)vol.d / vol.gz;
    }
  }
  float dr = 0.f, dg = 0.f, db = 0.f, da = 0.f;
  float s = 0.0f;
  ebs_coop_rounds<COUNT, M>(vol, tf, tf_n, E, wd, dir, D, kx, ky, kz, sub, gbase, done, s, dr, dg, db, da, ns, nq_used);
  if (hit && sub == 0) vrb_store_pixel(fr, px, py, dr, dg, db, da);
  else if (!hit && sub == 0 && fr.zero_miss && px < fr.w && py < fr.h && vrb_owns_pixel(part, px, py, fr.w)) vrb_store_pixel(fr, px, py, 0.f, 0.f, 0.f, 0.f);
  if (cta_cost && lane == 0) atomicMax(cta_cost + cta, (unsigned int)min((clock64() - t_start) >> 6, 0xffffffffLL));
  if (COUNT) {
    if (sub != 0) ns = 0;
    for (int o = 16; o > 0; o >>= 1) { ns += __shfl_xor_sync(0xffffffffu, ns, o); nq_used += __shfl_xor_sync(0xffffffffu, nq_used, o); }
    if (lane == 0 && ns) { atomicAdd(counter, (unsigned long long)ns); atomicAdd(counter + 1, nq_used); }
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Deferred frame (shade_list.cuh, march_list.cu): one list entry per lane.  ShadeSample (:500-551) for the entry's
// position; the colour that the compositing multiplies by alpha replaces the TF colour in the entry.  Every entry costs
// 15 AO shells + its shadow boxes; the list holds neighbouring rays' samples next to each other, so the SAT gathers of
// a warp stay as coherent as with the M-lanes-per-ray kernel, without a longest ray or a longest CTA.
template <bool COUNT>
__global__ void __launch_bounds__(128, 4)
k_ebs_shade(VolView vol, CamView cam, EbsConst E, ShadeListView L, unsigned n_entries, unsigned long long* counter) {
  const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
  const float4 A = (e < n_entries) ? L.a[e] : make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
  unsigned int nq = 0;
  if (__float_as_int(A.w) >= 0) {                  // slots reserved but never written keep pixel == -1
    const float4 src = L.b[e];
    const f3 tx = mk3(A.x, A.y, A.z);
    float ka = 0.0f, kd = 0.0f, ks = 0.0f, IOcc = 0.0f, ISdw = 0.0f;
    if (E.P.apply_occlusion == 1) { ka = E.ka; IOcc = ebs_ambient_occlusion(E, tx, nq); }
    if (E.P.apply_shadow == 1) { kd = E.kd; ks = E.ph.ks; ISdw = ebs_directional_shadows(E, tx, nq); }
    float cr, cg, cb;
    if (E.ph.grad) {                                   // ApplyPhongShading == 1 (:524-544); a zero gradient leaves L = clr
      cr = src.x; cg = src.y; cb = src.z;
      const float kx = (float)vol.w / vol.gx, ky = (float)vol.h / vol.gy, kz = (float)vol.d / vol.gz;
      float dot_diff, spec;
      if (vrb_phong_terms(vol, E.ph, kx, ky, kz, tx.x, tx.y, tx.z, cam.ex, cam.ey, cam.ez, dot_diff, spec)) {
        float k = (1.0f / (ka + kd));
        cr = k * (src.x * IOcc * ka + ISdw * (src.x * kd * dot_diff)) + ISdw * (ks * E.ph.isx * spec);
        cg = k * (src.y * IOcc * ka + ISdw * (src.y * kd * dot_diff)) + ISdw * (ks * E.ph.isy * spec);
        cb = k * (src.z * IOcc * ka + ISdw * (src.z * kd * dot_diff)) + ISdw * (ks * E.ph.isz * spec);
      }
    } else {
      float k = (1.0f / (ka + kd));
      cr = k * (src.x * IOcc * ka + src.x * ISdw * kd);
      cg = k * (src.y * IOcc * ka + src.y * ISdw * kd);
      cb = k * (src.z * IOcc * ka + src.z * ISdw * kd);
    }
    L.b[e] = make_float4(cr, cg, cb, src.w);
  }
  if (COUNT) {
    unsigned long long nq64 = nq;
    for (int o = 16; o > 0; o >>= 1) nq64 += __shfl_xor_sync(0xffffffffu, nq64, o);
    if ((threadIdx.x & 31) == 0 && nq64) atomicAdd(counter + 1, nq64);
  }
}

// K9 rc1pextbsd/lightcachecomputation.comp main (:443-469): one (Iao, Ids) pair per light-cache voxel, the marcher's own
// occlusion / shadow functions evaluated at the cache voxel centres.
__global__ void __launch_bounds__(64, EBS_MIN_BLOCKS)
k_ebs_light_cache(EbsConst E, f3 cell, int rw, int rh, int rd, __half2* __restrict__ cache) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rw * rh * rd) return;
  const int x = i % rw, y = (i / rw) % rh, z = i / (rw * rh);
  f3 tex_pos = mk3(((float)x + 0.5f) * cell.x, ((float)y + 0.5f) * cell.y, ((float)z + 0.5f) * cell.z);
  unsigned int nq = 0;
  float Iao = 1.0f, Ids = 1.0f;
  if (E.P.apply_occlusion == 1) Iao = ebs_ambient_occlusion(E, tex_pos, nq);
  if (E.P.apply_shadow == 1) Ids = ebs_directional_shadows(E, tex_pos, nq);
  cache[(size_t)(x + 1) + (size_t)(rw + 2) * ((size_t)(y + 1) + (size_t)(rh + 2) * (size_t)(z + 1))] = __floats2half2_rn(Iao, Ids);
}
