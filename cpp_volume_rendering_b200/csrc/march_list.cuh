// march_list.cuh -- the two renderer-independent kernels of a deferred lit frame (shade_list.cuh): vrb_list_march appends
// every sample with alpha > 0 of every ray to the shading list, vrb_list_composite replays the shaders' front-to-back
// arithmetic over the shaded entries.  Between the two a renderer runs its own shade kernel over list.b[0 .. n_entries).
// Implemented in march_list.cu.
#pragma once
#include "vrb_internal.cuh"

struct ListFrame {
  ShadeListView L;
  unsigned n_entries;
};

// rc1pcrtgt keeps its running colour and ray parameter in rgba16f / rg16f images between dispatches
// (gt_ray_marching.comp:375,406-471; crtgtrenderer.cpp:272-325): every step rounds them to fp16, and D = tfar - tnear.
#define VRB_LIST_GT 1

int vrb_list_march(vrb_ctx* c, const vrb_camera* cam, float step, int flags, int count_samples, ListFrame* out);
int vrb_list_composite(vrb_ctx* c, const vrb_camera* cam, int flags, const ListFrame& f);
// rc1pass (no list: composited in place) through the same persistent march kernel; exact filter mode
int vrb_list_rc1pass(vrb_ctx* c, const vrb_camera* cam, float step, int skip_requested, int count_samples);

#ifdef __CUDACC__
// camera_dir of a pixel: normalize(vec3(x tan aspect, y tan, -1) * mat3(ViewMatrix)) as every lit shader computes it before
// the ray set-up (e.g. rc1pdosct/ray_bbox_marching.comp:666-672).  Explicit roundings: independent of -fmad.
__device__ __forceinline__ void vrb_list_camera_dir(const CamView& cam, int fw, int fh, int pix, float& cx_, float& cy_, float& cz_) {
  const int px = pix % fw, py = pix / fw;
  const float fx = __fadd_rn((float)px, 0.5f), fy = __fadd_rn((float)py, 0.5f);
  const float vx = __fadd_rn(__fmul_rn(__fdiv_rn(fx, (float)fw), 2.0f), -1.0f), vy = __fadd_rn(__fmul_rn(__fdiv_rn(fy, (float)fh), 2.0f), -1.0f);
  const float cx = __fmul_rn(__fmul_rn(vx, cam.tan_fovy), cam.aspect), cy = __fmul_rn(vy, cam.tan_fovy), cz = -1.0f;
  float dx = __fadd_rn(__fadd_rn(__fmul_rn(cx, cam.m[0]), __fmul_rn(cy, cam.m[1])), __fmul_rn(cz, cam.m[2]));
  float dy = __fadd_rn(__fadd_rn(__fmul_rn(cx, cam.m[3]), __fmul_rn(cy, cam.m[4])), __fmul_rn(cz, cam.m[5]));
  float dz = __fadd_rn(__fadd_rn(__fmul_rn(cx, cam.m[6]), __fmul_rn(cy, cam.m[7])), __fmul_rn(cz, cam.m[8]));
  const float r = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))));
  cx_ = __fmul_rn(dx, r); cy_ = __fmul_rn(dy, r); cz_ = __fmul_rn(dz, r);
}
#endif
