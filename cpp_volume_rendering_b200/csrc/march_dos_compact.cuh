// march_dos_compact.cuh -- k_dos_compact: the exact-filter rc1pdosct marcher of round 2 (included by march_dos.cu, which is
// compiled with -fmad=false).  Same arithmetic, operation by operation, as dos_exact::k_dos (march_dos_body.cuh), which
// follows rc1pdosct/ray_bbox_marching.comp:124-322 (occlusion cones), :345-562 (shadow cones), :607-734 (ShadeSample, main);
// what changes is how the work is laid out on the machine:
//
//  * ray compaction per warp.  A lane (= one ray of the 8x4 pixel patch) marches until its NEXT sample with alpha > 0 and
//    parks there; when every live lane of the warp is parked, all of them evaluate their cones together.  k_dos shaded a
//    sample as soon as one lane met it, the other lanes idling (20.7 of 32 lanes active on config 3).
//  * the cone taps of neighbouring rays stay neighbours (same section, same pyramid level, positions a fraction of a texel
//    apart), so a warp-wide fetch touches one or two cache lines; lanes are NOT spread over the taps of one cone (32 taps
//    along a cone are 32 different lines: the L1 wavefront rate would bound that).
//  * 2x2 texel quads (d_pyr_quad): a trilinear footprint is two 8-byte loads instead of eight 2-byte loads, and the
//    replicated border replaces the six index clamps by one clamp per axis.
//  * the sections come from shared memory with the pyramid level resolved on the host (all levels of a
//    ConeGaussianSampler are integers, conegaussiansampler.cpp:384; a table with a fractional level takes the old kernel),
//    runs of sections on the same level keep the level's constants in registers and are unrolled four taps deep with all
//    eight loads in flight before the first blend.
//  * pos / VolumeGridSize is a multiplication when the three sizes are powers of two (bit-identical); floor() is a
//    round-down add against 1.5 * 2^23 instead of the conversion pipe.

struct DosLevelQ {
  const uint2* q;        // quad copy of the level, padded grid
  int pw, pslice;        // padded row / slice pitch in quads
  float fw, fh, fd;      // POW2 = false: level resolution as floats; POW2 = true: resolution / VolumeGridSize (an exact power of two)
  int w, h, d;
  int pad_;
};

struct DosFast {
  DosLevelQ lev[VRB_MAX_LEVELS];
  const float4* packed;  // occlusion sections then shadow sections: {interval, d_integral, amplitude, bits}
  int n_occ, n_sdw;      // bits = level | (mip + 64) << 8 | (end of the run of sections on this level) << 16
  d3 inv_vss;
  int count;
};

namespace dos_compact {

struct LevRegs { const uint2* q; int pw, pslice; float fw, fh, fd; int w, h, d; };

__device__ __forceinline__ LevRegs load_level(const DosFast& F, int l) {
  LevRegs L;
  L.q = F.lev[l].q; L.pw = F.lev[l].pw; L.pslice = F.lev[l].pslice;
  L.fw = F.lev[l].fw; L.fh = F.lev[l].fh; L.fd = F.lev[l].fd;
  L.w = F.lev[l].w; L.h = F.lev[l].h; L.d = F.lev[l].d;
  return L;
}

// POW2: VolumeGridSize and every level's resolution are powers of two, so (pos / size) * N == pos * (N / size) bit for bit
// (both factors only change the exponent) and the level carries N / size; otherwise the division is done as written.
template <bool POW2>
__device__ __forceinline__ d3 norm_pos(const DosConst& C, const DosFast& F, d3 tp) {
  if (POW2) return tp;
  return tp / C.VSS;
}

__device__ __forceinline__ bool outside(const DosConst& C, d3 tp) {
  return tp.x < 0.0f || tp.x > C.VSS.x || tp.y < 0.0f || tp.y > C.VSS.y || tp.z < 0.0f || tp.z > C.VSS.z;
}

// one axis of the GL_LINEAR footprint: u = s*N - 0.5, floor, fraction; returns the padded index of the lower texel
// clamped to [0, N] (both texels then fall on the replicated border when the footprint is outside)
__device__ __forceinline__ int axis(float s, float fn, int n, float& f) {
  const float u = s * fn - 0.5f;
  const float t = __fadd_rd(u, 12582912.0f);          // 1.5 * 2^23 + floor(u) for |u| < 2^22
  f = u - (t - 12582912.0f);
  return min(max(__float_as_int(t) - 0x4B400000 + 1, 0), n);
}

struct Tap { uint2 q0, q1; float fx, fy, fz; };

__device__ __forceinline__ void tap_issue(const LevRegs& L, d3 s, Tap& t) {
  const int jx = axis(s.x, L.fw, L.w, t.fx), jy = axis(s.y, L.fh, L.h, t.fy), jz = axis(s.z, L.fd, L.d, t.fz);
  const int o0 = jz * L.pslice + jy * L.pw + jx, o1 = o0 + L.pslice;     // levels hold fewer than 2^31 quads (checked on the host)
  t.q0 = __ldg(L.q + o0);
  t.q1 = __ldg(L.q + o1);
}

__device__ __forceinline__ float2 h2f(unsigned v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }

__device__ __forceinline__ float tap_blend(const Tap& t) {
  const float2 a = h2f(t.q0.x), b = h2f(t.q0.y), c = h2f(t.q1.x), d = h2f(t.q1.y);
  const float c00 = vrb_lerp(a.x, a.y, t.fx), c10 = vrb_lerp(b.x, b.y, t.fx);
  const float c01 = vrb_lerp(c.x, c.y, t.fx), c11 = vrb_lerp(d.x, d.y, t.fx);
  return vrb_lerp(vrb_lerp(c00, c10, t.fy), vrb_lerp(c01, c11, t.fy), t.fz);
}

// CONSIDER_BORDERS falloff of GetGaussianExtinction (:99-110) for a tap outside the volume; 2 * sg * sg with
// sg = pow(2, mip) is a power of two, so the division is an exact scaling
__device__ __noinline__ float tap_border(const DosConst& C, d3 tp, float rg, int mip) {
  const d3 c = m3(fminf(fmaxf(tp.x, 0.0f), C.VSS.x) - tp.x, fminf(fmaxf(tp.y, 0.0f), C.VSS.y) - tp.y, fminf(fmaxf(tp.z, 0.0f), C.VSS.z) - tp.z);
  const float dist = c.x * c.x + c.y * c.y + c.z * c.z;
  const float inv = __int_as_float((127 - (2 * mip + 1)) << 23);
  return rg * expf(-(dist) * inv);
}

// Cone1Ray* -> Cone3Ray* -> Cone7Ray*; u and v already swapped for the shadow cone (see dos_cone in march_dos_body.cuh)
template <bool HAS7, bool POW2>
__device__ __forceinline__ float cone_eval(const DosConst& C, const DosFast& F, const ConeView& K, const float4* __restrict__ sec,
                                           d3 pos0, d3 k, d3 u, d3 v) {
  const float uiw = K.ui_weight;
  float track = K.initial_step;
  float r0 = 0.0f, l0 = 0.0f;
  int id = 0;
  const int n1 = K.counts[0];
  while (id < n1) {
    const int pk = __float_as_int(sec[id].w);
    const LevRegs L = load_level(F, pk & 0xff);
    const int mip = ((pk >> 8) & 0xff) - 64;
    const int end = min(pk >> 16, n1);
    for (; id + 4 <= end; id += 4) {
      float4 si[4];
      float t[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) si[j] = sec[id + j];
      t[0] = track; t[1] = t[0] + si[0].x; t[2] = t[1] + si[1].x; t[3] = t[2] + si[2].x; track = t[3] + si[3].x;
      Tap tp[4];
      unsigned out = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const d3 p = pos0 + k * t[j];
        out |= outside(C, p) ? (1u << j) : 0u;
        tap_issue(L, norm_pos<POW2>(C, F, p), tp[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float rg = tap_blend(tp[j]);
        if (out & (1u << j)) rg = tap_border(C, pos0 + k * t[j], rg, mip);
        const float amptau = rg * si[j].z;
        r0 += (l0 + amptau) * si[j].y * uiw;
        l0 = amptau;
      }
    }
    for (; id < end; ++id) {
      const float4 si = sec[id];
      const d3 p = pos0 + k * track;
      Tap tp;
      tap_issue(L, norm_pos<POW2>(C, F, p), tp);
      float rg = tap_blend(tp);
      if (outside(C, p)) rg = tap_border(C, p, rg, mip);
      const float amptau = rg * si.z;
      r0 += (l0 + amptau) * si.y * uiw;
      l0 = amptau;
      track += si.x;
    }
  }
  if (!(K.counts[1] + K.counts[2] > 0)) return expf(-r0);
  float r[3] = {r0, r0, r0}, l[3] = {l0, l0, l0};
  {
    d3 vk[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) vk[i] = k * K.axes[i][2] + u * K.axes[i][1] + v * K.axes[i][0];
    const int n13 = n1 + K.counts[1];
    while (id < n13) {
      const int pk = __float_as_int(sec[id].w);
      const LevRegs L = load_level(F, pk & 0xff);
      const int mip = ((pk >> 8) & 0xff) - 64;
      const int end = min(pk >> 16, n13);
      for (; id < end; ++id) {
        const float4 si = sec[id];
        Tap tp[3];
        unsigned out = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const d3 p = pos0 + vk[i] * track;
          out |= outside(C, p) ? (1u << i) : 0u;
          tap_issue(L, norm_pos<POW2>(C, F, p), tp[i]);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float rg = tap_blend(tp[i]);
          if (out & (1u << i)) rg = tap_border(C, pos0 + vk[i] * track, rg, mip);
          const float amptau = rg * si.z;
          r[i] += (l[i] + amptau) * si.y * uiw;
          l[i] = amptau;
        }
        track += si.x;
      }
    }
  }
  if (!HAS7 || !(K.counts[2] > 0)) return (expf(-r[0]) + expf(-r[1]) + expf(-r[2])) / 3.0f;
  float rays[7], last[7];
  rays[6] = rays[5] = r[2];
  rays[4] = rays[3] = r[1];
  const float avg = (r[2] + r[1] + r[0]) / 3.0f;
  rays[2] = rays[1] = r[0];
  rays[0] = avg;
  last[6] = last[5] = l[2];
  last[4] = last[3] = l[1];
  const float avgt = (l[2] + l[1] + l[0]) / 3.0f;
  last[2] = last[1] = l[0];
  last[0] = avgt;
  {
    const int n137 = n1 + K.counts[1] + K.counts[2];
    while (id < n137) {
      const int pk = __float_as_int(sec[id].w);
      const LevRegs L = load_level(F, pk & 0xff);
      const int mip = ((pk >> 8) & 0xff) - 64;
      const int end = min(pk >> 16, n137);
      for (; id < end; ++id) {
        const float4 si = sec[id];
        Tap tp[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) {
          const d3 vk = k * K.axes[3 + i][2] + u * K.axes[3 + i][1] + v * K.axes[3 + i][0];
          tap_issue(L, norm_pos<POW2>(C, F, pos0 + vk * track), tp[i]);
        }
#pragma unroll
        for (int i = 0; i < 7; ++i) {
          const d3 vk = k * K.axes[3 + i][2] + u * K.axes[3 + i][1] + v * K.axes[3 + i][0];
          const d3 p = pos0 + vk * track;
          float rg = tap_blend(tp[i]);
          if (outside(C, p)) rg = tap_border(C, p, rg, mip);
          const float amptau = si.z * rg;
          rays[i] += (last[i] + amptau) * si.y * uiw;
          last[i] = amptau;
        }
        track += si.x;
      }
    }
  }
  return (expf(-rays[0]) + (expf(-rays[1]) + expf(-rays[2]) + expf(-rays[3]) + expf(-rays[4]) + expf(-rays[5]) + expf(-rays[6])) *
                               K.ray7_adj_weight) / (1.0f + K.ray7_adj_weight * 6.0f);
}

// ShadowEvaluationKernel (:533-562); Cone1RayShadow's parameter list is (k, v, u) (:481 vs its call at :561)
template <bool HAS7, bool POW2>
__device__ __forceinline__ float shadow_eval(const DosConst& C, const DosFast& F, const float4* __restrict__ sec, d3 pos0, bool& ran) {
  d3 k = m3(0.f, 0.f, 0.f), u = k, v = k;
  ran = true;
  if (C.P.type_of_shadow == 0 || C.P.type_of_shadow == 1) {
    const d3 half = m3(C.VSS.x / 2.0f, C.VSS.y / 2.0f, C.VSS.z / 2.0f);
    const d3 cone_vec = nrm3(C.light_pos - (pos0 - half));
    k = cone_vec;
    u = nrm3(cross3(k, C.light_right));
    v = nrm3(cross3(k, u));
    if (C.P.type_of_shadow == 1 && dot3(cone_vec, C.light_fwd) < C.P.spot_cos) { ran = false; return 0.0f; }
  } else if (C.P.type_of_shadow == 2) {
    k = C.light_fwd; v = C.light_up; u = C.light_right;
  }
  return cone_eval<HAS7, POW2>(C, F, C.sdw, sec, pos0, k, v, u);
}

template <bool PHONG, bool HAS7, bool POW2>
__global__ void __launch_bounds__(64)
k_dos_compact(VolView vol, const float4* __restrict__ tf_g, int tf_n, FrameView fr, CamView cam, PartView part,
              const __grid_constant__ DosConst C, const __grid_constant__ DosFast F, unsigned long long* counter) {
  extern __shared__ float4 s_mem[];
  const int tid = threadIdx.y * 8 + threadIdx.x;
  const int tf_sh = (tf_n + 2 <= 1026) ? tf_n + 2 : 0;
  const float4* tf = tf_g;
  if (tf_sh) {
    for (int i = tid; i < tf_sh; i += 64) s_mem[i] = tf_g[i];
    tf = s_mem;
  }
  float4* s_sec = s_mem + tf_sh;
  for (int i = tid; i < F.n_occ + F.n_sdw; i += 64) s_sec[i] = F.packed[i];
  __syncthreads();
  const float4* sec_occ = s_sec;
  const float4* sec_sdw = s_sec + F.n_occ;

  int px, py;
  vrb_cta_origin(part, fr.w, 8, 8, px, py);
  px += threadIdx.x; py += threadIdx.y;
  unsigned int ns = 0, ntaps = 0;
  const bool mine = px < fr.w && py < fr.h && vrb_owns_pixel(part, px, py, fr.w);
  bool alive = false;
  d3 cdir = m3(0.f, 0.f, 0.f), dir = cdir, wd = cdir;
  float D = 0.0f, s = 0.0f;
  if (mine) {
    // camera_dir is normalised once in main (:673-674); RayAABBIntersection normalises it again into r.Dir
    const float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
    const float vx = (fx / (float)fr.w) * 2.0f - 1.0f, vy = (fy / (float)fr.h) * 2.0f - 1.0f;
    const float cx = vx * cam.tan_fovy * cam.aspect, cy = vy * cam.tan_fovy, cz = -1.0f;
    cdir = nrm3(m3(cx * cam.m[0] + cy * cam.m[1] + cz * cam.m[2], cx * cam.m[3] + cy * cam.m[4] + cz * cam.m[5],
                   cx * cam.m[6] + cy * cam.m[7] + cz * cam.m[8]));
    const Ray r = vrb_make_ray(cam, px, py, fr.w, fr.h, C.VSS.x, C.VSS.y, C.VSS.z);
    alive = r.hit;
    if (!alive && fr.zero_miss) vrb_store_pixel(fr, px, py, 0.f, 0.f, 0.f, 0.f);
    D = fabsf(r.tfar - r.tnear);
    dir = m3(r.dx, r.dy, r.dz);
    wd = m3(r.ox, r.oy, r.oz) + dir * r.tnear;
    wd = wd + C.VSS * 0.5f;
  }
  const bool hit = alive;
  const float kx = (float)vol.w / vol.gx, ky = (float)vol.h / vol.gy, kz = (float)vol.d / vol.gz;
  const float step = C.P.step_size;
  float dr = 0.f, dg = 0.f, db = 0.f, da = 0.f;
  for (;;) {
    // phase 1: every live lane walks to its next sample with alpha > 0 (or to the end of its ray)
    bool pending = false;
    float4 src = make_float4(0.f, 0.f, 0.f, 0.f);
    float h = 0.0f;
    d3 tx = wd;
    if (alive) {
      while (s < D) {
        h = fminf(step, D - s);
        tx = wd + dir * (s + h * 0.5f);
        const float density = vrb_sample_volume(vol, kx, ky, kz, tx.x, tx.y, tx.z);
        src = vrb_sample_tf(tf, tf_n, density);
        ++ns;
        if (src.w > 0.0f) { pending = true; break; }
        s = s + h;
      }
      alive = pending;
    }
    if (!__any_sync(0xffffffffu, pending)) break;
    // phase 2: the parked lanes shade together
    if (pending) {
      const d3 half = C.VSS * 0.5f;
      float ka = 0.0f, kd = 0.0f, ks = 0.0f, IOcc = 0.0f, ISdw = 0.0f;
      if (C.P.apply_occlusion == 1) {
        ka = C.ka;
        const d3 v_right = nrm3(cross3(cdir, m3(0.f, 1.f, 0.f)));
        const d3 v_up = nrm3(cross3(-cdir, v_right));
        const d3 kvec = nrm3(C.eye - (tx - half));            // OcclusionEvaluationKernel (:324-333)
        IOcc = cone_eval<HAS7, POW2>(C, F, C.occ, sec_occ, tx, kvec, v_up, v_right);
        ntaps += C.occ.counts[0] + 3 * C.occ.counts[1] + 7 * C.occ.counts[2];
      }
      if (C.P.apply_shadow == 1) {
        kd = C.kd; ks = C.ph.ks;
        bool ran;
        ISdw = shadow_eval<HAS7, POW2>(C, F, sec_sdw, tx, ran);
        if (ran) ntaps += C.sdw.counts[0] + 3 * C.sdw.counts[1] + 7 * C.sdw.counts[2];
      }
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
        const float kk = (1.0f / (ka + kd));
        cr = kk * (src.x * IOcc * ka + src.x * ISdw * kd);
        cg = kk * (src.y * IOcc * ka + src.y * ISdw * kd);
        cb = kk * (src.z * IOcc * ka + src.z * ISdw * kd);
      }
      const float a = 1.0f - expf(-src.w * h);
      const float om = 1.0f - da;
      dr = dr + om * (cr * a); dg = dg + om * (cg * a); db = db + om * (cb * a); da = da + om * a;
      if (da > 0.99f) alive = false;
      else s = s + h;
    }
  }
  if (hit) vrb_store_pixel(fr, px, py, dr, dg, db, da);
  if (F.count) {
    unsigned long long nt64 = ntaps;
    for (int o = 16; o > 0; o >>= 1) { ns += __shfl_xor_sync(0xffffffffu, ns, o); nt64 += __shfl_xor_sync(0xffffffffu, nt64, o); }
    if ((tid & 31) == 0 && ns) { atomicAdd(counter, (unsigned long long)ns); atomicAdd(counter + 1, nt64); }
  }
}

}  // namespace dos_compact
