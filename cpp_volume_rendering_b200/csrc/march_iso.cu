// march_iso.cu -- adaptive-step isosurface ray caster, sm_100a.
// Replaces the dispatch of rc1pisoadapt/ray_marching_1p_iso_adapt.comp (main :91-173, ShadeBlinnPhong :48-88) made by
// RayCasting1PassIsoAdapt::Redraw (rc1pisoadaptrenderer.cpp:168-179); uniforms as uploaded by Update (:113-165).
// The one marcher of the reference whose step depends on the data: a small step while the previous sample is within
// StepSizeRange of the isovalue, a large one otherwise; a sign change between two samples is a hit, refined by linear
// interpolation, shaded (optionally Blinn-Phong on the gradient texture) and composited front to back.  No transfer
// function is involved.  One thread per pixel, 8x8 tiles; fp32 in the shader's operation order (-fmad=false).
#include "vrb_internal.cuh"

template <bool COUNT, bool HW>
__global__ void __launch_bounds__(64)
k_iso(VolView vol, FrameView fr, CamView cam, PartView part, const __grid_constant__ vrb_iso_params P, const __grid_constant__ PhongView ph,
      unsigned long long* counter) {
  int px, py;
  vrb_cta_origin(part, fr.w, 8, 8, px, py);
  px += threadIdx.x; py += threadIdx.y;
  unsigned int ns = 0;
  if (px < fr.w && py < fr.h && vrb_owns_pixel(part, px, py, fr.w)) {
    Ray r = vrb_make_ray(cam, px, py, fr.w, fr.h, vol.gx, vol.gy, vol.gz);
    if (r.hit) {
      const float D = fabsf(r.tfar - r.tnear);
      const float tx = (r.ox + r.dx * r.tnear) + (vol.gx * 0.5f), ty = (r.oy + r.dy * r.tnear) + (vol.gy * 0.5f), tz = (r.oz + r.dz * r.tnear) + (vol.gz * 0.5f);
      const float kx = (float)vol.w / vol.gx, ky = (float)vol.h / vol.gy, kz = (float)vol.d / vol.gz;
      float dr = 0.f, dg = 0.f, db = 0.f, da = 0.f;
      float prev = HW ? tex3D<float>(vol.tex3d, tx * kx, ty * ky, tz * kz) : vrb_sample_volume(vol, kx, ky, kz, tx, ty, tz);
      for (float s = 0.0f; s < D;) {
        const float cur_step = (fabsf(prev - P.isovalue) < P.step_size_range) ? P.step_size_small : P.step_size_large;
        const float h = fminf(cur_step, D - s);
        const float t = s + h;                                  // sampled at the END of the interval (:137)
        const float qx = tx + r.dx * t, qy = ty + r.dy * t, qz = tz + r.dz * t;
        const float density = HW ? tex3D<float>(vol.tex3d, qx * kx, qy * ky, qz * kz) : vrb_sample_volume(vol, kx, ky, kz, qx, qy, qz);
        if (COUNT) ++ns;
        if ((prev <= P.isovalue && P.isovalue < density) || (prev >= P.isovalue && P.isovalue > density)) {
          const float u = (P.isovalue - prev) / (density - prev);
          const float tt = s + u * h;
          const float hx = tx + r.dx * tt, hy = ty + r.dy * tt, hz = tz + r.dz * tt;      // refined hit position
          float cr = P.color[0], cg = P.color[1], cb = P.color[2];
          const float ca = P.color[3];
          if (ph.grad) {
            float dot_diff, spec;
            if (vrb_phong_terms(vol, ph, kx, ky, kz, hx, hy, hz, cam.ex, cam.ey, cam.ez, dot_diff, spec)) {
              const float kad = ph.ka + ph.kd * dot_diff;
              cr = cr * kad + ph.isx * ph.ks * spec;
              cg = cg * kad + ph.isy * ph.ks * spec;
              cb = cb * kad + ph.isz * ph.ks * spec;
            }
          }
          const float om = 1.0f - da;
          dr = dr + om * (cr * ca); dg = dg + om * (cg * ca); db = db + om * (cb * ca); da = da + om * ca;
          if (da > 0.99f) break;
        }
        prev = density;
        s = s + h;
      }
      vrb_store_pixel(fr, px, py, dr, dg, db, da);
    } else if (fr.zero_miss) vrb_store_pixel(fr, px, py, 0.f, 0.f, 0.f, 0.f);
  }
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) ns += __shfl_xor_sync(0xffffffffu, ns, o);
    if (((threadIdx.y * 8 + threadIdx.x) & 31) == 0 && ns) atomicAdd(counter, (unsigned long long)ns);
  }
}

extern "C" int vrb_iso_render(vrb_ctx* c, const vrb_camera* cam, const vrb_lighting* light, const vrb_iso_params* p) {
  VRB_REQUIRE(c && cam && light && p, VRB_ERR_INVALID, "vrb_iso_render: NULL argument");
  VRB_REQUIRE(c->d_vol, VRB_ERR_STATE, "vrb_iso_render: no volume uploaded");
  VRB_REQUIRE(c->d_frame, VRB_ERR_STATE, "vrb_iso_render: no frame (vrb_frame_resize)");
  VRB_REQUIRE(p->step_size_small > 0.0f && p->step_size_large > 0.0f, VRB_ERR_INVALID, "vrb_iso_render: step sizes %g, %g must be positive",
              p->step_size_small, p->step_size_large);
  PhongView ph;
  { int rc = vrb_make_phong_view(c, light, &ph, "vrb_iso_render"); if (rc != VRB_OK) return rc; }
  VRB_CUDA(cudaSetDevice(c->device));
  if (!c->d_frame_target) VRB_CUDA(cudaMemsetAsync(c->d_frame, 0, (size_t)c->fw * c->fh * 4 * sizeof(__half), c->stream));
  if (p->count_samples) { int rc = vrb_counters_reset(c); if (rc != VRB_OK) return rc; }
  PartView part;
  dim3 block(8, 8), grid = vrb_make_grid(c, 8, 8, &part);
  { int rc = vrb_vol_tex3d_prepare(c); if (rc != VRB_OK) return rc; }
  VolView vol = c->vol_view();
#define VRB_ISO(N, H) k_iso<N, H><<<grid, block, 0, c->stream>>>(vol, c->frame_view(), make_cam_view(cam), part, *p, ph, c->d_counter)
  if (vol.tex3d) { if (p->count_samples) VRB_ISO(true, true); else VRB_ISO(false, true); }
  else           { if (p->count_samples) VRB_ISO(true, false); else VRB_ISO(false, false); }
#undef VRB_ISO
  VRB_CUDA(cudaGetLastError());
  c->launches++;
  if (p->count_samples) return vrb_counters_fetch(c);
  return VRB_OK;
}
