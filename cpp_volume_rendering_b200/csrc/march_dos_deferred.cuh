// march_dos_deferred.cuh -- the exact-filter rc1pdosct frame as three kernels (shade_list.cuh): k_dos_march appends every
// sample with alpha > 0, k_dos_shade evaluates the occlusion and shadow cones of one list entry per lane
// (dos_compact::cone_eval / shadow_eval, march_dos_compact.cuh), k_dos_composite replays ShadeSample's compositing
// (rc1pdosct/ray_bbox_marching.comp:607-656, :700-716) per ray in march order.  Included by march_dos.cu (-fmad=false).
namespace dos_deferred {
using namespace dos_compact;

struct RaySetup { d3 cdir, dir, wd; float D; bool hit; };

// pixel -> camera_dir, ray, entry point in texture space (main, :658-699)
__device__ __forceinline__ RaySetup ray_setup(const CamView& cam, const FrameView& fr, const DosConst& C, int px, int py) {
  RaySetup R;
  const float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
  const float vx = (fx / (float)fr.w) * 2.0f - 1.0f, vy = (fy / (float)fr.h) * 2.0f - 1.0f;
  const float cx = vx * cam.tan_fovy * cam.aspect, cy = vy * cam.tan_fovy, cz = -1.0f;
  R.cdir = nrm3(m3(cx * cam.m[0] + cy * cam.m[1] + cz * cam.m[2], cx * cam.m[3] + cy * cam.m[4] + cz * cam.m[5],
                   cx * cam.m[6] + cy * cam.m[7] + cz * cam.m[8]));
  const Ray r = vrb_make_ray(cam, px, py, fr.w, fr.h, C.VSS.x, C.VSS.y, C.VSS.z);
  R.hit = r.hit;
  R.D = fabsf(r.tfar - r.tnear);
  R.dir = m3(r.dx, r.dy, r.dz);
  R.wd = m3(r.ox, r.oy, r.oz) + R.dir * r.tnear;
  R.wd = R.wd + C.VSS * 0.5f;
  return R;
}

__global__ void __launch_bounds__(64)
k_dos_march(VolView vol, const float4* __restrict__ tf_g, int tf_n, FrameView fr, CamView cam, PartView part,
            const __grid_constant__ DosConst C, ShadeListView L, int count, unsigned long long* counter) {
  extern __shared__ float4 s_tf[];
  const int tid = threadIdx.y * 8 + threadIdx.x;
  const float4* tf = tf_g;
  if (tf_n + 2 <= 1026) {
    for (int i = tid; i < tf_n + 2; i += 64) s_tf[i] = tf_g[i];
    tf = s_tf;
  }
  __syncthreads();
  const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x;
  const unsigned warp_id = cta * 2u + (unsigned)(tid >> 5), lane = (unsigned)(tid & 31);
  if (lane == 0) L.head[warp_id] = VRB_SL_NONE;
  int px, py;
  vrb_cta_origin(part, fr.w, 8, 8, px, py);
  px += threadIdx.x; py += threadIdx.y;
  const bool mine = px < fr.w && py < fr.h && vrb_owns_pixel(part, px, py, fr.w);
  RaySetup R;
  R.hit = false; R.D = 0.0f; R.cdir = R.dir = R.wd = m3(0.f, 0.f, 0.f);
  if (mine) R = ray_setup(cam, fr, C, px, py);
  bool alive = R.hit;
  const float kx = (float)vol.w / vol.gx, ky = (float)vol.h / vol.gy, kz = (float)vol.d / vol.gz;
  const float step = C.P.step_size;
  float s = 0.0f, da = 0.0f;
  unsigned ns = 0, last_chunk = VRB_SL_NONE;
  for (;;) {
    bool pending = false;
    float4 src = make_float4(0.f, 0.f, 0.f, 0.f);
    float h = 0.0f;
    d3 tx = R.wd;
    if (alive) {
      while (s < R.D) {
        h = fminf(step, R.D - s);
        tx = R.wd + R.dir * (s + h * 0.5f);
        const float density = vrb_sample_volume(vol, kx, ky, kz, tx.x, tx.y, tx.z);
        src = vrb_sample_tf(tf, tf_n, density);
        ++ns;
        if (src.w > 0.0f) { pending = true; break; }
        s = s + h;
      }
      alive = pending;
    }
    const unsigned e = sl_append(L, pending, last_chunk, warp_id, lane);
    if (!__any_sync(0xffffffffu, pending)) break;
    if (pending) {
      const float a = 1.0f - expf(-src.w * h);
      if (e != VRB_SL_NONE) {
        L.a[e] = make_float4(tx.x, tx.y, tx.z, __int_as_float(py * fr.w + px));
        L.b[e] = make_float4(src.x, src.y, src.z, a);
      }
      const float om = 1.0f - da;
      da = da + om * a;
      if (da > 0.99f) alive = false;
      else s = s + h;
    }
  }
  if (count) {
    for (int o = 16; o > 0; o >>= 1) ns += __shfl_xor_sync(0xffffffffu, ns, o);
    if (lane == 0 && ns) atomicAdd(counter, (unsigned long long)ns);
  }
}

// one list entry per lane: ShadeSample's lighting terms (:607-650) -> L.b[e].rgb = the colour the compositing multiplies by alpha
template <bool PHONG, bool HAS7, bool POW2>
__global__ void __launch_bounds__(128)
k_dos_shade(VolView vol, FrameView fr, CamView cam, const __grid_constant__ DosConst C, const __grid_constant__ DosFast F,
            ShadeListView L, unsigned n_entries, unsigned long long* counter) {
  extern __shared__ float4 s_sec[];
  for (int i = threadIdx.x; i < F.n_occ + F.n_sdw; i += blockDim.x) s_sec[i] = F.packed[i];
  __syncthreads();
  const float4* sec_occ = s_sec;
  const float4* sec_sdw = s_sec + F.n_occ;
  const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned ntaps = 0;
  if (e < n_entries) {
    const float4 A = L.a[e], B = L.b[e];
    const d3 tx = m3(A.x, A.y, A.z);
    const d3 half = C.VSS * 0.5f;
    float ka = 0.0f, kd = 0.0f, ks = 0.0f, IOcc = 0.0f, ISdw = 0.0f;
    if (C.P.apply_occlusion == 1) {
      ka = C.ka;
      const int pix = __float_as_int(A.w);
      const int px = pix % fr.w, py = pix / fr.w;
      const float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
      const float vx = (fx / (float)fr.w) * 2.0f - 1.0f, vy = (fy / (float)fr.h) * 2.0f - 1.0f;
      const float cx = vx * cam.tan_fovy * cam.aspect, cy = vy * cam.tan_fovy, cz = -1.0f;
      const d3 cdir = nrm3(m3(cx * cam.m[0] + cy * cam.m[1] + cz * cam.m[2], cx * cam.m[3] + cy * cam.m[4] + cz * cam.m[5],
                              cx * cam.m[6] + cy * cam.m[7] + cz * cam.m[8]));
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
      cr = B.x; cg = B.y; cb = B.z;
      const float kx = (float)vol.w / vol.gx, ky = (float)vol.h / vol.gy, kz = (float)vol.d / vol.gz;
      float dot_diff, spec;
      if (vrb_phong_terms(vol, C.ph, kx, ky, kz, tx.x, tx.y, tx.z, C.eye.x, C.eye.y, C.eye.z, dot_diff, spec)) {
        const float f = ((1.0f / (ka + kd)) * (IOcc * ka + ISdw * kd * dot_diff));
        const float sp = (ISdw * ks * spec);
        cr = B.x * f + C.ph.isx * sp; cg = B.y * f + C.ph.isy * sp; cb = B.z * f + C.ph.isz * sp;
      }
    } else {
      const float kk = (1.0f / (ka + kd));
      cr = kk * (B.x * IOcc * ka + B.x * ISdw * kd);
      cg = kk * (B.y * IOcc * ka + B.y * ISdw * kd);
      cb = kk * (B.z * IOcc * ka + B.z * ISdw * kd);
    }
    L.b[e] = make_float4(cr, cg, cb, B.w);
  }
  if (F.count) {
    unsigned long long nt64 = ntaps;
    for (int o = 16; o > 0; o >>= 1) nt64 += __shfl_xor_sync(0xffffffffu, nt64, o);
    if ((threadIdx.x & 31) == 0 && nt64) atomicAdd(counter + 1, nt64);
  }
}

// one lane per ray, same grid as k_dos_march: front-to-back over the ray's entries in march order (:700-716)
__global__ void __launch_bounds__(64)
k_dos_composite(FrameView fr, CamView cam, PartView part, float gx, float gy, float gz, ShadeListView L) {
  const int tid = threadIdx.y * 8 + threadIdx.x;
  const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x;
  const unsigned warp_id = cta * 2u + (unsigned)(tid >> 5), lane = (unsigned)(tid & 31);
  int px, py;
  vrb_cta_origin(part, fr.w, 8, 8, px, py);
  px += threadIdx.x; py += threadIdx.y;
  const bool mine = px < fr.w && py < fr.h && vrb_owns_pixel(part, px, py, fr.w);
  bool hit = false;
  if (mine) hit = vrb_make_ray(cam, px, py, fr.w, fr.h, gx, gy, gz).hit;
  float dr = 0.f, dg = 0.f, db = 0.f, da = 0.f;
  unsigned h = L.head[warp_id];
  while (h != VRB_SL_NONE) {
    const uint4 H = L.hdr[h];
    if (H.x & (1u << lane)) {
      const float4 o = L.b[H.y + __popc(H.x & ((1u << lane) - 1u))];
      const float om = 1.0f - da;
      dr = dr + om * (o.x * o.w); dg = dg + om * (o.y * o.w); db = db + om * (o.z * o.w); da = da + om * o.w;
    }
    h = H.z;
  }
  if (hit) vrb_store_pixel(fr, px, py, dr, dg, db, da);
  else if (mine && fr.zero_miss) vrb_store_pixel(fr, px, py, 0.f, 0.f, 0.f, 0.f);
}

}  // namespace dos_deferred
