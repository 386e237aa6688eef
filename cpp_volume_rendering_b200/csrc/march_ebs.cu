// march_ebs.cu -- extinction-based shading marcher (Schlegel et al. 2011) for sm_100a.
// Replaces the dispatch of rc1pextbsd/ebs_ray_bbox_marching.comp made by RC1PExtinctionBasedShading::Redraw
// (ebsrenderer.cpp:249-260); uniforms as uploaded by ebsrenderer.cpp:125-247.
//
// Numerics: the SAT queries are fp32 differences of prefix sums as large as 1e7-1e8, so the reference RESULT contains
// the cancellation noise of its exact sequence of fp32 operations.  This file is therefore compiled with -fmad=false
// (see Makefile) and written in the shader's operation order; the only fused operations are the explicit fmaf() of
// the GL_LINEAR blend (vrb_lerp), which the oracle spells the same way.  cos/sin of the cone angle depend on uniforms
// only and are evaluated once on the host.
#include "vrb_internal.cuh"
#include <cmath>

struct f3 { float x, y, z; };
__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ f3 operator*(float s, f3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ f3 clamp3(f3 a, f3 lo, f3 hi) {
  return mk3(fminf(fmaxf(a.x, lo.x), hi.x), fminf(fmaxf(a.y, lo.y), hi.y), fminf(fmaxf(a.z, lo.z), hi.z));
}
__device__ __forceinline__ f3 norm3(f3 a) {
  float d = a.x * a.x + a.y * a.y + a.z * a.z;
  float r = 1.0f / sqrtf(d);
  return mk3(a.x * r, a.y * r, a.z * r);
}

struct EbsConst {
  f3 VS, VSS, MinSAT, MaxSAT, MinVol, MaxVol, inv_vol_scaled, InvG;
  float p_cs, p_sn, n_cs, n_sn;
  const float* sat; int sw, sh, sd; long long sslice;
  vrb_ebs_params P;
  float ka, kd;
  f3 light_pos, light_fwd;
};

struct Axis { int i0, i1; float f; };
__device__ __forceinline__ Axis sat_axis(float x, float inv, int n) {
  Axis a;
  float s = x * inv;
  float u = s * (float)n - 0.5f;
  float fl = floorf(u);
  a.f = u - fl;
  int i = (int)fl;
  a.i0 = min(max(i, 0), n - 1);
  a.i1 = min(max(i + 1, 0), n - 1);
  return a;
}
__device__ __forceinline__ float sat_tri(const EbsConst& E, const Axis& X, const Axis& Y, const Axis& Z) {
  const float* z0 = E.sat + (long long)Z.i0 * E.sslice;
  const float* z1 = E.sat + (long long)Z.i1 * E.sslice;
  const int r0 = Y.i0 * E.sw, r1 = Y.i1 * E.sw;
  float c00 = vrb_lerp(__ldg(z0 + r0 + X.i0), __ldg(z0 + r0 + X.i1), X.f);
  float c10 = vrb_lerp(__ldg(z0 + r1 + X.i0), __ldg(z0 + r1 + X.i1), X.f);
  float c01 = vrb_lerp(__ldg(z1 + r0 + X.i0), __ldg(z1 + r0 + X.i1), X.f);
  float c11 = vrb_lerp(__ldg(z1 + r1 + X.i0), __ldg(z1 + r1 + X.i1), X.f);
  return vrb_lerp(vrb_lerp(c00, c10, Y.f), vrb_lerp(c01, c11, Y.f), Z.f);
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

// ConeZAxis / ConeYAxis / ConeXAxis (:190-430).  The lateral extents use slightly different rotation formulas per
// axis in the shader; they are kept as written.
__device__ float ebs_cone_z(const EbsConst& E, f3 pos, f3 cv, unsigned int& nq) {
  float Stau = 0.0f;
  float signal = 1.0f; if (cv.z < 0) signal = -1.0f;
  f3 proj_y = norm3(mk3(0.0f, cv.y, cv.z));
  f3 proj_x = norm3(mk3(cv.x, 0.0f, cv.z));
  f3 pj_x1 = norm3(mk3(proj_x.x * E.n_cs - proj_x.z * E.n_sn, 0.0f, proj_x.x * E.n_sn + proj_x.z * E.n_cs));
  f3 pj_x2 = norm3(mk3(proj_x.x * E.p_cs - proj_x.z * E.p_sn, 0.0f, proj_x.x * E.p_sn + proj_x.z * E.p_cs));
  f3 pj_y1 = norm3(mk3(0.0f, proj_y.y * E.n_cs - proj_y.z * E.n_sn, proj_y.y * E.n_sn + proj_y.z * E.n_cs));
  f3 pj_y2 = norm3(mk3(0.0f, proj_y.y * E.p_cs - proj_y.z * E.p_sn, proj_y.y * E.p_sn + proj_y.z * E.p_cs));
  float si = E.P.sdw_sample_interval * signal * E.VS.z;
  float z_pos = E.P.sdw_initial_step * signal * E.VS.z;
  while ((z_pos / cv.z) < E.P.sdw_cone_max_distance &&
         (pos.z + (z_pos + si) > E.MinVol.z && pos.z + (z_pos + si) < E.MaxVol.z)) {
    float z_mean = fabsf(z_pos + si * 0.5f);
    float p_x1 = pj_x1.x * (z_mean / fabsf(pj_x1.z));
    float p_x2 = pj_x2.x * (z_mean / fabsf(pj_x2.z));
    float p_y1 = pj_y1.y * (z_mean / fabsf(pj_y1.z));
    float p_y2 = pj_y2.y * (z_mean / fabsf(pj_y2.z));
    float x1 = fminf(p_x1, p_x2), x2 = fmaxf(p_x1, p_x2);
    float y1 = fminf(p_y1, p_y2), y2 = fmaxf(p_y1, p_y2);
    float xdiff = fabsf(x2 - x1), ydiff = fabsf(y2 - y1);
    float xs = (ceilf(xdiff / E.VS.x) - (xdiff / E.VS.x)) * 0.5f;
    float ys = (ceilf(ydiff / E.VS.y) - (ydiff / E.VS.y)) * 0.5f;
    x1 = x1 - xs * E.VS.x; x2 = x2 + xs * E.VS.x;
    y1 = y1 - ys * E.VS.y; y2 = y2 + ys * E.VS.y;
    float z1 = fminf(z_pos, z_pos + si), z2 = fmaxf(z_pos, z_pos + si);
    ++nq;
    Stau += eval_shadow_sat3d(E, pos + mk3(x1, y1, z1), pos + mk3(x2, y2, z2));
    z_pos = z_pos + si;
  }
  return Stau;
}
__device__ float ebs_cone_y(const EbsConst& E, f3 pos, f3 cv, unsigned int& nq) {
  float Stau = 0.0f;
  float signal = 1.0f; if (cv.y < 0) signal = -1.0f;
  f3 proj_x = norm3(mk3(cv.x, cv.y, 0.0f));
  f3 proj_z = norm3(mk3(0.0f, cv.y, cv.z));
  f3 pj_x1 = norm3(mk3(proj_x.x * E.n_cs - proj_x.y * E.n_sn, proj_x.x * E.n_sn + proj_x.y * E.n_cs, 0.0f));
  f3 pj_x2 = norm3(mk3(proj_x.x * E.p_cs - proj_x.y * E.p_sn, proj_x.x * E.p_sn + proj_x.y * E.p_cs, 0.0f));
  f3 pj_z1 = norm3(mk3(0.0f, proj_z.z * E.n_sn + proj_z.y * E.n_cs, proj_z.z * E.n_cs - proj_z.y * E.n_sn));
  f3 pj_z2 = norm3(mk3(0.0f, proj_z.z * E.p_sn + proj_z.y * E.p_cs, proj_z.z * E.p_cs - proj_z.y * E.p_sn));
  float si = E.P.sdw_sample_interval * signal * E.VS.y;
  float y_pos = E.P.sdw_initial_step * signal * E.VS.y;
  while ((y_pos / cv.y) < E.P.sdw_cone_max_distance &&
         (pos.y + (y_pos + si) > E.MinVol.y && pos.y + (y_pos + si) < E.MaxVol.y)) {
    float y_mean = fabsf(y_pos + si * 0.5f);
    float p_x1 = pj_x1.x * (y_mean / fabsf(pj_x1.y));
    float p_x2 = pj_x2.x * (y_mean / fabsf(pj_x2.y));
    float p_z1 = pj_z1.z * (y_mean / fabsf(pj_z1.y));
    float p_z2 = pj_z2.z * (y_mean / fabsf(pj_z2.y));
    float x1 = fminf(p_x1, p_x2), x2 = fmaxf(p_x1, p_x2);
    float z1 = fminf(p_z1, p_z2), z2 = fmaxf(p_z1, p_z2);
    float xdiff = fabsf(x2 - x1), zdiff = fabsf(z2 - z1);
    float xs = (ceilf(xdiff / E.VS.x) - (xdiff / E.VS.x)) * 0.5f;
    float zs = (ceilf(zdiff / E.VS.z) - (zdiff / E.VS.z)) * 0.5f;
    x1 = x1 - xs * E.VS.x; x2 = x2 + xs * E.VS.x;
    z1 = z1 - zs * E.VS.z; z2 = z2 + zs * E.VS.z;
    float y1 = fminf(y_pos, y_pos + si), y2 = fmaxf(y_pos, y_pos + si);
    ++nq;
    Stau += eval_shadow_sat3d(E, pos + mk3(x1, y1, z1), pos + mk3(x2, y2, z2));
    y_pos = y_pos + si;
  }
  return Stau;
}
__device__ float ebs_cone_x(const EbsConst& E, f3 pos, f3 cv, unsigned int& nq) {
  float Stau = 0.0f;
  float signal = 1.0f; if (cv.x < 0) signal = -1.0f;
  f3 proj_y = norm3(mk3(cv.x, cv.y, 0.0f));
  f3 proj_z = norm3(mk3(cv.x, 0.0f, cv.z));
  f3 pj_y1 = norm3(mk3(proj_y.y * E.n_sn + proj_y.x * E.n_cs, proj_y.y * E.n_cs - proj_y.x * E.n_sn, 0.0f));
  f3 pj_y2 = norm3(mk3(proj_y.y * E.p_sn + proj_y.x * E.p_cs, proj_y.y * E.p_cs - proj_y.x * E.p_sn, 0.0f));
  f3 pj_z1 = norm3(mk3(proj_z.z * E.n_sn + proj_z.x * E.n_cs, 0.0f, proj_z.z * E.n_cs - proj_z.x * E.n_sn));
  f3 pj_z2 = norm3(mk3(proj_z.z * E.p_sn + proj_z.x * E.p_cs, 0.0f, proj_z.z * E.p_cs - proj_z.x * E.p_sn));
  float si = E.P.sdw_sample_interval * signal * E.VS.x;
  float x_pos = E.P.sdw_initial_step * signal * E.VS.x;
  while ((x_pos / cv.x) < E.P.sdw_cone_max_distance &&
         (pos.x + (x_pos + si) > E.MinVol.x && pos.x + (x_pos + si) < E.MaxVol.x)) {
    float x_mean = fabsf(x_pos + si * 0.5f);
    float p_y1 = pj_y1.y * (x_mean / fabsf(pj_y1.x));
    float p_y2 = pj_y2.y * (x_mean / fabsf(pj_y2.x));
    float p_z1 = pj_z1.z * (x_mean / fabsf(pj_z1.x));
    float p_z2 = pj_z2.z * (x_mean / fabsf(pj_z2.x));
    float y1 = fminf(p_y1, p_y2), y2 = fmaxf(p_y1, p_y2);
    float z1 = fminf(p_z1, p_z2), z2 = fmaxf(p_z1, p_z2);
    float ydiff = fabsf(y2 - y1), zdiff = fabsf(z2 - z1);
    float ys = (ceilf(ydiff / E.VS.y) - (ydiff / E.VS.y)) * 0.5f;
    float zs = (ceilf(zdiff / E.VS.z) - (zdiff / E.VS.z)) * 0.5f;
    y1 = y1 - ys * E.VS.y; y2 = y2 + ys * E.VS.y;
    z1 = z1 - zs * E.VS.z; z2 = z2 + zs * E.VS.z;
    float x1 = fminf(x_pos, x_pos + si), x2 = fmaxf(x_pos, x_pos + si);
    ++nq;
    Stau += eval_shadow_sat3d(E, pos + mk3(x1, y1, z1), pos + mk3(x2, y2, z2));
    x_pos = x_pos + si;
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
  if (az > ax && az > ay) Stau = ebs_cone_z(E, tx, cone_vec, nq);
  else if (ay > ax) Stau = ebs_cone_y(E, tx, cone_vec, nq);
  else Stau = ebs_cone_x(E, tx, cone_vec, nq);
  return expf(-Stau);
}

template <bool COUNT>
__global__ void __launch_bounds__(64)
k_ebs(VolView vol, const float4* __restrict__ tf_g, int tf_n, FrameView fr, CamView cam, PartView part, EbsConst E,
      unsigned long long* counter) {
  extern __shared__ float4 s_tf[];
  const float4* tf = tf_g;
  if (tf_n + 2 <= 1026) {
    for (int i = threadIdx.y * 8 + threadIdx.x; i < tf_n + 2; i += 64) s_tf[i] = tf_g[i];
    __syncthreads();
    tf = s_tf;
  }
  int px = blockIdx.x * 8 + threadIdx.x, py = blockIdx.y * 8 + threadIdx.y;
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
          // ShadeSample (:500-551), ApplyPhongShading == 0
          float ka = 0.0f, kd = 0.0f, IOcc = 0.0f, ISdw = 0.0f;
          if (E.P.apply_occlusion == 1) { ka = E.ka; IOcc = ebs_ambient_occlusion(E, tx, nq); }
          if (E.P.apply_shadow == 1) { kd = E.kd; ISdw = ebs_directional_shadows(E, tx, nq); }
          float k = (1.0f / (ka + kd));
          float cr = k * (src.x * IOcc * ka + src.x * ISdw * kd);
          float cg = k * (src.y * IOcc * ka + src.y * ISdw * kd);
          float cb = k * (src.z * IOcc * ka + src.z * ISdw * kd);
          float a = 1.0f - expf(-src.w * h);
          float om = 1.0f - da;
          dr = dr + om * (cr * a); dg = dg + om * (cg * a); db = db + om * (cb * a); da = da + om * a;
          if (da > 0.99f) break;
        }
        s = s + h;
      }
      vrb_store_pixel(fr, px, py, dr, dg, db, da);
    }
  }
  if (COUNT) {
    unsigned long long nq64 = nq;
    for (int o = 16; o > 0; o >>= 1) { ns += __shfl_xor_sync(0xffffffffu, ns, o); nq64 += __shfl_xor_sync(0xffffffffu, nq64, o); }
    if (((threadIdx.y * 8 + threadIdx.x) & 31) == 0 && ns) { atomicAdd(counter, (unsigned long long)ns); atomicAdd(counter + 1, nq64); }
  }
}

static f3 h3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }

extern "C" int vrb_ebs_render(vrb_ctx* c, const vrb_camera* cam, const vrb_lighting* light, const vrb_ebs_params* p) {
  VRB_REQUIRE(c && cam && light && p, VRB_ERR_INVALID, "vrb_ebs_render: NULL argument");
  VRB_REQUIRE(c->d_vol, VRB_ERR_STATE, "vrb_ebs_render: no volume uploaded");
  VRB_REQUIRE(c->d_tf_rgbt, VRB_ERR_STATE, "vrb_ebs_render: no transfer function uploaded");
  VRB_REQUIRE(c->d_sat, VRB_ERR_STATE, "vrb_ebs_render: no SAT (vrb_sat_build)");
  VRB_REQUIRE(c->d_frame, VRB_ERR_STATE, "vrb_ebs_render: no frame (vrb_frame_resize)");
  VRB_REQUIRE(p->step_size > 0.0f, VRB_ERR_INVALID, "vrb_ebs_render: step_size %g", p->step_size);
  VRB_REQUIRE(p->amb_occ_shells >= 1 && p->amb_occ_shells <= 4096, VRB_ERR_INVALID, "vrb_ebs_render: amb_occ_shells %d", p->amb_occ_shells);
  VRB_REQUIRE(p->sdw_sample_interval > 0.0f, VRB_ERR_INVALID, "vrb_ebs_render: sdw_sample_interval %g", p->sdw_sample_interval);
  VRB_CUDA(cudaSetDevice(c->device));
  EbsConst E;
  E.VS = h3(c->scale[0], c->scale[1], c->scale[2]);
  E.VSS = h3((float)c->vw * E.VS.x, (float)c->vh * E.VS.y, (float)c->vd * E.VS.z);
  E.MinSAT = h3(E.VS.x * 0.5f, E.VS.y * 0.5f, E.VS.z * 0.5f);
  E.MaxSAT = h3(E.VSS.x + E.VS.x * 1.5f, E.VSS.y + E.VS.y * 1.5f, E.VSS.z + E.VS.z * 1.5f);
  E.MinVol = E.MinSAT;
  E.MaxVol = h3(E.VSS.x - E.VS.x * 0.5f, E.VSS.y - E.VS.y * 0.5f, E.VSS.z - E.VS.z * 0.5f);
  E.inv_vol_scaled = h3(1.0f / (E.VSS.x + E.VS.x * 2.0f), 1.0f / (E.VSS.y + E.VS.y * 2.0f), 1.0f / (E.VSS.z + E.VS.z * 2.0f));
  E.InvG = h3(1.0f / E.VSS.x, 1.0f / E.VSS.y, 1.0f / E.VSS.z);
  E.p_cs = cosf(p->sdw_cone_angle_rad); E.p_sn = sinf(p->sdw_cone_angle_rad);
  E.n_cs = cosf(-p->sdw_cone_angle_rad); E.n_sn = sinf(-p->sdw_cone_angle_rad);
  E.sat = c->d_sat; E.sw = c->sat_w; E.sh = c->sat_h; E.sd = c->sat_d; E.sslice = (long long)c->sat_w * c->sat_h;
  E.P = *p;
  E.ka = light->ka; E.kd = light->kd;
  E.light_pos = h3(light->light_pos[0], light->light_pos[1], light->light_pos[2]);
  E.light_fwd = h3(light->light_forward[0], light->light_forward[1], light->light_forward[2]);

  VRB_CUDA(cudaMemsetAsync(c->d_frame, 0, (size_t)c->fw * c->fh * 4 * sizeof(__half), c->stream));
  if (p->count_samples) { int rc = vrb_counters_reset(c); if (rc != VRB_OK) return rc; }
  dim3 block(8, 8), grid((c->fw + 7) / 8, (c->fh + 7) / 8);
  size_t smem = (c->tf_n + 2 <= 1026) ? (size_t)(c->tf_n + 2) * sizeof(float4) : 0;
  if (p->count_samples) k_ebs<true><<<grid, block, smem, c->stream>>>(c->vol_view(), c->d_tf_rgbt, c->tf_n, c->frame_view(), make_cam_view(cam), c->part, E, c->d_counter);
  else                  k_ebs<false><<<grid, block, smem, c->stream>>>(c->vol_view(), c->d_tf_rgbt, c->tf_n, c->frame_view(), make_cam_view(cam), c->part, E, c->d_counter);
  VRB_CUDA(cudaGetLastError());
  c->launches++;
  if (p->count_samples) return vrb_counters_fetch(c);
  return VRB_OK;
}
