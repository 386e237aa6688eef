// march_rc1pass.cu -- single-pass ray casting for sm_100a.
// Replaces the dispatch of rc1pass/ray_marching_1p.comp (main at :85-179) made by RayCasting1Pass::Redraw
// (rc1prenderer.cpp:140-151).  One thread per pixel, 8x4 pixel tile per warp (two warps per 8x8 CTA, the reference's
// work-group shape, ray_marching_1p.comp:38) so that the 32 rays of a warp walk the same few cache lines.
#include "vrb_internal.cuh"

#define TF_SMEM_MAX 1024   // transfer functions up to this many texels are staged in shared memory

template <bool TF_SMEM, bool COUNT>
__global__ void __launch_bounds__(64)
k_rc1pass(VolView vol, const float4* __restrict__ tf_g, int tf_n, FrameView fr, CamView cam, PartView part,
          float step, unsigned long long* counter) {
  extern __shared__ float4 s_tf[];
  const float4* tf = tf_g;
  if (TF_SMEM) {
    for (int i = threadIdx.y * 8 + threadIdx.x; i < tf_n + 2; i += 64) s_tf[i] = tf_g[i];
    __syncthreads();
    tf = s_tf;
  }
  int px = blockIdx.x * 8 + threadIdx.x, py = vrb_center_out_row(blockIdx.y, gridDim.y) * 8 + threadIdx.y;
  unsigned int ns = 0;
  if (px < fr.w && py < fr.h && vrb_owns_pixel(part, px, py, fr.w)) {
    Ray r = vrb_make_ray(cam, px, py, fr.w, fr.h, vol.gx, vol.gy, vol.gz);
    if (r.hit) {
      float D = fabsf(__fadd_rn(r.tfar, -r.tnear));
      // tex_pos = eye + dir*tnear + G/2
      float tx = __fadd_rn(__fadd_rn(r.ox, __fmul_rn(r.dx, r.tnear)), __fmul_rn(vol.gx, 0.5f));
      float ty = __fadd_rn(__fadd_rn(r.oy, __fmul_rn(r.dy, r.tnear)), __fmul_rn(vol.gy, 0.5f));
      float tz = __fadd_rn(__fadd_rn(r.oz, __fmul_rn(r.dz, r.tnear)), __fmul_rn(vol.gz, 0.5f));
      float kx = (float)vol.w / vol.gx, ky = (float)vol.h / vol.gy, kz = (float)vol.d / vol.gz;
      float dr = 0.f, dg = 0.f, db = 0.f, da = 0.f;
      for (float s = 0.0f; s < D;) {
        float h = fminf(step, __fadd_rn(D, -s));
        float t = __fadd_rn(s, __fmul_rn(h, 0.5f));
        float density = vrb_sample_volume(vol, kx, ky, kz, fmaf(r.dx, t, tx), fmaf(r.dy, t, ty), fmaf(r.dz, t, tz));
        float4 src = vrb_sample_tf(tf, tf_n, density);
        if (COUNT) ++ns;
        if (src.w > 0.0f) {
          float a = 1.0f - __expf(-src.w * h);
          float om = (1.0f - da) * a;       // (1 - dst.a) * src.a ; src.rgb is premultiplied by src.a first
          dr = fmaf(om, src.x, dr); dg = fmaf(om, src.y, dg); db = fmaf(om, src.z, db);
          da = da + om;
          if (da > 0.99f) break;
        }
        s = __fadd_rn(s, h);
      }
      vrb_store_pixel(fr, px, py, dr, dg, db, da);
    }
  }
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) ns += __shfl_xor_sync(0xffffffffu, ns, o);
    if (((threadIdx.y * 8 + threadIdx.x) & 31) == 0 && ns) atomicAdd(counter, (unsigned long long)ns);
  }
}

extern "C" int vrb_rc1pass_render(vrb_ctx* c, const vrb_camera* cam, const vrb_rc1pass_params* p) {
  VRB_REQUIRE(c && cam && p, VRB_ERR_INVALID, "vrb_rc1pass_render: NULL argument");
  VRB_REQUIRE(c->d_vol, VRB_ERR_STATE, "vrb_rc1pass_render: no volume uploaded");
  VRB_REQUIRE(c->d_tf_rgbt, VRB_ERR_STATE, "vrb_rc1pass_render: no transfer function uploaded");
  VRB_REQUIRE(c->d_frame, VRB_ERR_STATE, "vrb_rc1pass_render: no frame (vrb_frame_resize)");
  VRB_REQUIRE(p->step_size > 0.0f, VRB_ERR_INVALID, "vrb_rc1pass_render: step_size %g", p->step_size);
  VRB_CUDA(cudaSetDevice(c->device));
  // RayCasting1Pass::Redraw: ClearTexture, then dispatch (misses keep the cleared 0)
  VRB_CUDA(cudaMemsetAsync(c->d_frame, 0, (size_t)c->fw * c->fh * 4 * sizeof(__half), c->stream));
  if (p->count_samples) { int rc = vrb_counters_reset(c); if (rc != VRB_OK) return rc; }
  dim3 block(8, 8), grid((c->fw + 7) / 8, (c->fh + 7) / 8);
  VolView vol = c->vol_view();
  FrameView fr = c->frame_view();
  CamView cv = make_cam_view(cam);
  bool smem = c->tf_n <= TF_SMEM_MAX;
  size_t smem_bytes = smem ? (size_t)(c->tf_n + 2) * sizeof(float4) : 0;
  if (smem) {
    if (p->count_samples) k_rc1pass<true, true><<<grid, block, smem_bytes, c->stream>>>(vol, c->d_tf_rgbt, c->tf_n, fr, cv, c->part, p->step_size, c->d_counter);
    else                  k_rc1pass<true, false><<<grid, block, smem_bytes, c->stream>>>(vol, c->d_tf_rgbt, c->tf_n, fr, cv, c->part, p->step_size, c->d_counter);
  } else {
    if (p->count_samples) k_rc1pass<false, true><<<grid, block, smem_bytes, c->stream>>>(vol, c->d_tf_rgbt, c->tf_n, fr, cv, c->part, p->step_size, c->d_counter);
    else                  k_rc1pass<false, false><<<grid, block, smem_bytes, c->stream>>>(vol, c->d_tf_rgbt, c->tf_n, fr, cv, c->part, p->step_size, c->d_counter);
  }
  VRB_CUDA(cudaGetLastError());
  c->launches++;
  if (p->count_samples) return vrb_counters_fetch(c);
  return VRB_OK;
}
