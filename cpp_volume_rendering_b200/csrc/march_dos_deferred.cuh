// march_dos_deferred.cuh -- the shading kernel of the exact-filter rc1pdosct frame (shade_list.cuh, march_list.cu): k_dos_shade
// evaluates the occlusion and shadow cones of one list entry per lane (dos_compact::cone_eval / shadow_eval,
// march_dos_compact.cuh) and leaves ShadeSample's colour (rc1pdosct/ray_bbox_marching.comp:607-656) in the entry; the march
// that builds the list and the compositing that consumes it (:700-716) are the renderer-independent kernels of
// march_list.cu.  Included by march_dos.cu (-fmad=false).
namespace dos_deferred {
using namespace dos_compact;

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
  const float4 A = (e < n_entries) ? L.a[e] : make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
  if (__float_as_int(A.w) >= 0) {                  // slots reserved but never written keep pixel == -1 (shade_list.cuh)
    const float4 B = L.b[e];
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

}  // namespace dos_deferred
