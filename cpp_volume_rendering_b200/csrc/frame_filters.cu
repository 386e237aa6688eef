// frame_filters.cu -- pixel multi-scaling of vis::RenderFrameToScreen (libs/vis_utils/renderoutputframe.cpp:89-145,
// 265-540).  The marchers render into a frame of (W*m, H*m) (m > 0: MULTIPLE_RAYS_PER_PIXEL,
// DOWN_SCALING_RENDER) or (W/|m|, H/|m|) (m < 0: UP_SCALING_RENDER); one of three image-space passes then produces the
// W x H "filtered screen output":
//   multisample_filter.comp   one GL_LINEAR fetch of the rendered frame at the output pixel's centre
//   downscaling_filter.comp   separable reconstruction kernel evaluated at the output centre over the rendered texels,
//                             scaled by the two resolution ratios
//   upscaling_filter.comp     the same kernels as an interpolator of the low-resolution frame
// with the six kernels of renderoutputframe/*_filter.comp (box, hat, Catmull-Rom, Mitchell-Netravali, cardinal
// B-spline 3, cardinal O-MOMS 3).  The two cardinal kernels come with an in-place recursive digital filter
// (cbs_ / comoms_digital_filter.comp) run over rows, then columns: after the down-scale on the filtered frame, before the
// up-scale on the rendered frame; every step of the recursion goes through an rgba16f imageStore, i.e. is rounded to
// fp16, and that is kept.  texelFetch outside the texture is undefined in GL 4.3 without robust access; here (and in
// the oracle) it returns zero.  fp32 in the shaders' operation order (compiled with -fmad=false).
#include "vrb_internal.cuh"
#include <algorithm>

__device__ __forceinline__ float4 ff_load(const __half* __restrict__ img, int w, int x, int y) {
  uint2 pk = reinterpret_cast<const uint2*>(img)[(size_t)y * w + x];
  float2 a = __half22float2(*reinterpret_cast<__half2*>(&pk.x)), b = __half22float2(*reinterpret_cast<__half2*>(&pk.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 ff_fetch0(const __half* __restrict__ img, int w, int h, int x, int y) {
  if (x < 0 || y < 0 || x >= w || y >= h) return make_float4(0.f, 0.f, 0.f, 0.f);
  return ff_load(img, w, x, y);
}
__device__ __forceinline__ void ff_store(__half* __restrict__ img, int w, int x, int y, float4 v) {
  __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
  uint2 pk; pk.x = *reinterpret_cast<unsigned int*>(&lo); pk.y = *reinterpret_cast<unsigned int*>(&hi);
  reinterpret_cast<uint2*>(img)[(size_t)y * w + x] = pk;
}

// renderoutputframe/{box,hat,catmullrom,mitchellnetravali,cardinalbspline,cardinalomoms}_filter.comp
template <int K> __device__ __forceinline__ float ff_support() { return K == VRB_KERNEL_BOX ? 1.0f : (K == VRB_KERNEL_HAT ? 2.0f : 4.0f); }
template <int K> __device__ __forceinline__ float ff_k0(float u) {
  if (K == VRB_KERNEL_CATMULL_ROM) return ((.5f * u - .5f) * u) * u;
  if (K == VRB_KERNEL_MITCHELL_NETRAVALI) return (((7 / 18.0f) * u - 1 / 3.0f) * u) * u;
  if (K == VRB_KERNEL_CARDINAL_BSPLINE_3) return ((u)*u) * u;
  return ((.875f * u) * u + .125f) * u;
}
template <int K> __device__ __forceinline__ float ff_k1(float u) {
  if (K == VRB_KERNEL_CATMULL_ROM) return ((-1.5f * u + 2.0f) * u + .5f) * u;
  if (K == VRB_KERNEL_MITCHELL_NETRAVALI) return (((-7 / 6.0f) * u + 1.5f) * u + 0.5f) * u + 1 / 18.0f;
  if (K == VRB_KERNEL_CARDINAL_BSPLINE_3) return ((-3.0f * u + 3.0f) * u + 3.0f) * u + 1.0f;
  return ((-2.625f * u + 2.625f) * u + 2.25f) * u + 1.0f;
}
template <int K> __device__ __forceinline__ float ff_weight(float x) {
  if (K == VRB_KERNEL_BOX) return (x <= -0.5f || x > 0.5f) ? 0.0f : 1.0f;
  x = fabsf(x);
  if (K == VRB_KERNEL_HAT) return x > 1.0f ? 0.0f : 1.0f - x;
  return x > 2.0f ? 0.0f : (x > 1.0f ? ff_k0<K>(2.0f - x) : ff_k1<K>(1.0f - x));
}

// multisample_filter.comp: texture(TexGeneratedFrame, (storePos + 0.5) / size), GL_LINEAR, clamp to edge
__global__ void __launch_bounds__(64)
k_ff_multisample(const __half* __restrict__ src, int sw, int sh, __half* __restrict__ dst, int dw, int dh) {
  const int x = blockIdx.x * 8 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if (x >= dw || y >= dh) return;
  const float fx = ((float)x + 0.5f) / (float)dw, fy = ((float)y + 0.5f) / (float)dh;
  const float ux = fx * (float)sw - 0.5f, uy = fy * (float)sh - 0.5f;
  const float flx = floorf(ux), fly = floorf(uy);
  const float tx = ux - flx, ty = uy - fly;
  const int ix = (int)flx, iy = (int)fly;
  const int x0 = min(max(ix, 0), sw - 1), x1 = min(max(ix + 1, 0), sw - 1), y0 = min(max(iy, 0), sh - 1), y1 = min(max(iy + 1, 0), sh - 1);
  const float4 a = ff_load(src, sw, x0, y0), b = ff_load(src, sw, x1, y0), c = ff_load(src, sw, x0, y1), d = ff_load(src, sw, x1, y1);
  float4 o;
  o.x = vrb_lerp(vrb_lerp(a.x, b.x, tx), vrb_lerp(c.x, d.x, tx), ty);
  o.y = vrb_lerp(vrb_lerp(a.y, b.y, tx), vrb_lerp(c.y, d.y, tx), ty);
  o.z = vrb_lerp(vrb_lerp(a.z, b.z, tx), vrb_lerp(c.z, d.z, tx), ty);
  o.w = vrb_lerp(vrb_lerp(a.w, b.w, tx), vrb_lerp(c.w, d.w, tx), ty);
  ff_store(dst, dw, x, y, o);
}

// downscaling_filter.comp (TexGenerated = src, Target = dst)
template <int K>
__global__ void __launch_bounds__(64)
k_ff_downscale(const __half* __restrict__ src, int sw, int sh, __half* __restrict__ dst, int dw, int dh) {
  const int j_c = blockIdx.x * 8 + threadIdx.x, j_r = blockIdx.y * 8 + threadIdx.y;
  if (j_c >= dw || j_r >= dh) return;
  const float s_r = (float)dh / (float)sh, s_c = (float)dw / (float)sw;
  const int n_r = sh, n_c = sw;
  const float kr = 0.5f * ff_support<K>();
  const float x_r = ((float)j_r + 0.5f) / (float)dh;
  const int il_r = (int)ceilf((x_r - kr / (float)dh) * (float)n_r - 0.5f), ir_r = (int)floorf((x_r + kr / (float)dh) * (float)n_r - 0.5f);
  const float x_c = ((float)j_c + 0.5f) / (float)dw;
  const int il_c = (int)ceilf((x_c - kr / (float)dw) * (float)n_c - 0.5f), ir_c = (int)floorf((x_c + kr / (float)dw) * (float)n_c - 0.5f);
  float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i_r = il_r; i_r <= ir_r; ++i_r) {
    const float wr = ff_weight<K>((x_r - ((float)i_r + 0.5f) / (float)n_r) * (float)dh);
    for (int i_c = il_c; i_c <= ir_c; ++i_c) {
      const float wgt = wr * ff_weight<K>((x_c - ((float)i_c + 0.5f) / (float)n_c) * (float)dw);
      const float4 t = ff_fetch0(src, sw, sh, i_c, i_r);
      f.x += wgt * t.x; f.y += wgt * t.y; f.z += wgt * t.z; f.w += wgt * t.w;
    }
  }
  const float sc = (s_r * s_c);
  f.x *= sc; f.y *= sc; f.z *= sc; f.w *= sc;
  ff_store(dst, dw, j_c, j_r, f);
}

// upscaling_filter.comp
template <int K>
__global__ void __launch_bounds__(64)
k_ff_upscale(const __half* __restrict__ src, int sw, int sh, __half* __restrict__ dst, int dw, int dh) {
  const int j_c = blockIdx.x * 8 + threadIdx.x, j_r = blockIdx.y * 8 + threadIdx.y;
  if (j_c >= dw || j_r >= dh) return;
  const float kr = 0.5f * ff_support<K>();
  const float x_r = ((float)j_r + 0.5f) / (float)dh;
  const float xi_r = x_r * (float)sh - 0.5f;
  const int il_r = (int)ceilf(xi_r - kr), ir_r = (int)floorf(xi_r + kr);
  const float x_c = ((float)j_c + 0.5f) / (float)dw;
  const float xi_c = x_c * (float)sw - 0.5f;
  const int il_c = (int)ceilf(xi_c - kr), ir_c = (int)floorf(xi_c + kr);
  float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i_r = il_r; i_r <= ir_r; ++i_r) {
    const float wr = ff_weight<K>(xi_r - (float)i_r);
    for (int i_c = il_c; i_c <= ir_c; ++i_c) {
      const float wgt = wr * ff_weight<K>(xi_c - (float)i_c);
      const float4 t = ff_fetch0(src, sw, sh, i_c, i_r);
      f.x += wgt * t.x; f.y += wgt * t.y; f.z += wgt * t.z; f.w += wgt * t.w;
    }
  }
  ff_store(dst, dw, j_c, j_r, f);
}

// cbs_digital_filter.comp / comoms_digital_filter.comp: one thread per row (direction 0) or column (direction 1), the
// LU recursion in place with an fp16 round trip at every step (imageStore into rgba16f, imageLoad back)
struct DigitalL { float L[9]; int m; };
__device__ __forceinline__ float4 ff_r16(float4 v) {
  return make_float4(__half2float(__float2half_rn(v.x)), __half2float(__float2half_rn(v.y)), __half2float(__float2half_rn(v.z)), __half2float(__float2half_rn(v.w)));
}
__global__ void __launch_bounds__(64)
k_ff_digital(__half* __restrict__ img, int w, int h, int direction, const __grid_constant__ DigitalL D) {
  const int line = blockIdx.x * blockDim.x + threadIdx.x;
  const int nlines = direction == 0 ? h : w, nn = direction == 0 ? w : h;
  if (line >= nlines) return;
  const int m = D.m;
  const float p_inv = 1.0f;
  const float L_inf = D.L[m - 1], v_inv = L_inf / (1.f + L_inf);
#define FF_AT(i) (direction == 0 ? (i) : line), (direction == 0 ? line : (i))
  // texels outside the image: imageLoad returns 0, imageStore is dropped (images narrower than m)
  auto ld = [&](int i) -> float4 { return (i >= 0 && i < nn) ? ff_load(img, w, FF_AT(i)) : make_float4(0.f, 0.f, 0.f, 0.f); };
  auto st = [&](int i, float4 v) { if (i >= 0 && i < nn) ff_store(img, w, FF_AT(i), v); };
  for (int x = 1; x < m; ++x) {
    const float4 a = ld(x), b = ld(x - 1); const float l = D.L[x - 1];
    st(x, make_float4(a.x - (l * b.x), a.y - (l * b.y), a.z - (l * b.z), a.w - (l * b.w)));
  }
  float4 prev = ld(m - 1);
  for (int x = m; x < nn; ++x) {
    const float4 a = ff_load(img, w, FF_AT(x));
    const float4 o = ff_r16(make_float4(a.x - (L_inf * prev.x), a.y - (L_inf * prev.y), a.z - (L_inf * prev.z), a.w - (L_inf * prev.w)));
    ff_store(img, w, FF_AT(x), o);
    prev = o;
  }
  {
    const float4 a = ld(nn - 1);
    st(nn - 1, make_float4(a.x * p_inv * v_inv, a.y * p_inv * v_inv, a.z * p_inv * v_inv, a.w * p_inv * v_inv));
  }
  float4 next = ld(nn - 1);
  for (int x = nn - 2; x >= m - 1; --x) {
    const float4 a = ld(x);
    const float4 o = ff_r16(make_float4(L_inf * (p_inv * a.x - next.x), L_inf * (p_inv * a.y - next.y), L_inf * (p_inv * a.z - next.z), L_inf * (p_inv * a.w - next.w)));
    st(x, o);
    next = o;
  }
  for (int x = m - 2; x >= 0; --x) {
    const float4 a = ld(x), b = ld(x + 1); const float l = D.L[x];
    st(x, make_float4(l * (p_inv * a.x - b.x), l * (p_inv * a.y - b.y), l * (p_inv * a.z - b.z), l * (p_inv * a.w - b.w)));
  }
#undef FF_AT
}

static int ff_digital(vrb_ctx* c, __half* img, int w, int h, int kernel) {
  DigitalL D;
  if (kernel == VRB_KERNEL_CARDINAL_BSPLINE_3) {
    const float L[8] = {.2f, .26315789f, .26760563f, .26792453f, .26794742f, .26794907f, .26794918f, .26794919f};
    for (int i = 0; i < 8; ++i) D.L[i] = L[i];
    D.L[8] = 0.f; D.m = 8;
  } else {
    const float L[9] = {.23529412f, .33170732f, .34266611f, .34395774f, .34411062f, .34412872f, .34413087f, .34413112f, .34413115f};
    for (int i = 0; i < 9; ++i) D.L[i] = L[i];
    D.m = 9;
  }
  k_ff_digital<<<(h + 63) / 64, 64, 0, c->stream>>>(img, w, h, 0, D);       // FilterDirection 0: along x, one thread per row
  k_ff_digital<<<(w + 63) / 64, 64, 0, c->stream>>>(img, w, h, 1, D);       // FilterDirection 1: along y, one thread per column
  VRB_CUDA(cudaGetLastError());
  c->launches += 2;
  return VRB_OK;
}

// RenderFrameToScreen::UpdateScreenResolutionMultiScaling (renderoutputframe.cpp:89-145)
extern "C" int vrb_frame_resize_multiscaling(vrb_ctx* c, int s_w, int s_h, int mw, int mh) {
  VRB_REQUIRE(c, VRB_ERR_INVALID, "vrb_frame_resize_multiscaling: ctx is NULL");
  VRB_REQUIRE(s_w > 0 && s_h > 0 && s_w <= 16384 && s_h <= 16384, VRB_ERR_INVALID, "vrb_frame_resize_multiscaling: bad size %dx%d", s_w, s_h);
  VRB_REQUIRE(mw != 0 && mh != 0 && abs(mw) <= 16 && abs(mh) <= 16, VRB_ERR_INVALID, "vrb_frame_resize_multiscaling: bad multipliers %d, %d", mw, mh);
  const int i_w = mw < 0 ? s_w / abs(mw) : s_w * mw, i_h = mh < 0 ? s_h / abs(mh) : s_h * mh;
  VRB_REQUIRE(i_w > 0 && i_h > 0 && i_w <= 16384 && i_h <= 16384, VRB_ERR_INVALID, "vrb_frame_resize_multiscaling: rendered frame would be %dx%d", i_w, i_h);
  int rc = vrb_frame_resize(c, i_w, i_h);
  if (rc != VRB_OK) return rc;
  if (!c->d_filtered || c->flw != s_w || c->flh != s_h) {
    if (c->d_filtered) { VRB_CUDA(cudaStreamSynchronize(c->stream)); VRB_CUDA(cudaFree(c->d_filtered)); c->d_filtered = nullptr; }
    VRB_CUDA(cudaMalloc(&c->d_filtered, (size_t)s_w * s_h * 4 * sizeof(__half)));
    VRB_CUDA(cudaMemsetAsync(c->d_filtered, 0, (size_t)s_w * s_h * 4 * sizeof(__half), c->stream));
    c->flw = s_w; c->flh = s_h;
  }
  return VRB_OK;
}

void vrb_free_filtered(vrb_ctx* c) {
  if (c->d_filtered) cudaFree(c->d_filtered);
  c->d_filtered = nullptr; c->flw = c->flh = 0;
}

extern "C" int vrb_frame_filter(vrb_ctx* c, int pass, int kernel) {
  VRB_REQUIRE(c, VRB_ERR_INVALID, "vrb_frame_filter: ctx is NULL");
  VRB_REQUIRE(c->d_frame && c->d_filtered, VRB_ERR_STATE, "vrb_frame_filter: no multi-scaling frames (vrb_frame_resize_multiscaling)");
  VRB_REQUIRE(pass >= VRB_FILTER_PASS_MULTISAMPLE && pass <= VRB_FILTER_PASS_UPSCALE, VRB_ERR_INVALID, "vrb_frame_filter: pass %d", pass);
  VRB_REQUIRE(kernel >= VRB_KERNEL_BOX && kernel <= VRB_KERNEL_CARDINAL_OMOMS3, VRB_ERR_INVALID, "vrb_frame_filter: kernel %d", kernel);
  VRB_CUDA(cudaSetDevice(c->device));
  __half* src = c->d_frame; const int sw = c->fw, sh = c->fh;
  __half* dst = c->d_filtered; const int dw = c->flw, dh = c->flh;
  const dim3 block(8, 8), grid((dw + 7) / 8, (dh + 7) / 8);
  const bool cardinal = kernel == VRB_KERNEL_CARDINAL_BSPLINE_3 || kernel == VRB_KERNEL_CARDINAL_OMOMS3;
#define FF_BY_KERNEL(KERN)                                                                                            \
  switch (kernel) {                                                                                                  \
    case VRB_KERNEL_BOX: KERN<VRB_KERNEL_BOX><<<grid, block, 0, c->stream>>>(src, sw, sh, dst, dw, dh); break;        \
    case VRB_KERNEL_HAT: KERN<VRB_KERNEL_HAT><<<grid, block, 0, c->stream>>>(src, sw, sh, dst, dw, dh); break;        \
    case VRB_KERNEL_CATMULL_ROM: KERN<VRB_KERNEL_CATMULL_ROM><<<grid, block, 0, c->stream>>>(src, sw, sh, dst, dw, dh); break; \
    case VRB_KERNEL_MITCHELL_NETRAVALI: KERN<VRB_KERNEL_MITCHELL_NETRAVALI><<<grid, block, 0, c->stream>>>(src, sw, sh, dst, dw, dh); break; \
    case VRB_KERNEL_CARDINAL_BSPLINE_3: KERN<VRB_KERNEL_CARDINAL_BSPLINE_3><<<grid, block, 0, c->stream>>>(src, sw, sh, dst, dw, dh); break; \
    default: KERN<VRB_KERNEL_CARDINAL_OMOMS3><<<grid, block, 0, c->stream>>>(src, sw, sh, dst, dw, dh); break;       \
  }
  if (pass == VRB_FILTER_PASS_MULTISAMPLE) {
    k_ff_multisample<<<grid, block, 0, c->stream>>>(src, sw, sh, dst, dw, dh);
  } else if (pass == VRB_FILTER_PASS_DOWNSCALE) {
    FF_BY_KERNEL(k_ff_downscale)
    if (cardinal) { int rc = ff_digital(c, dst, dw, dh, kernel); if (rc != VRB_OK) return rc; }
  } else {
    if (cardinal) { int rc = ff_digital(c, src, sw, sh, kernel); if (rc != VRB_OK) return rc; }   // in place on the rendered frame
    FF_BY_KERNEL(k_ff_upscale)
  }
#undef FF_BY_KERNEL
  VRB_CUDA(cudaGetLastError());
  c->launches++;
  return VRB_OK;
}

extern "C" int vrb_filtered_frame_info(vrb_ctx* c, void** dev, int* w, int* h) {
  VRB_REQUIRE(c, VRB_ERR_INVALID, "vrb_filtered_frame_info: ctx is NULL");
  VRB_REQUIRE(c->d_filtered, VRB_ERR_STATE, "vrb_filtered_frame_info: no filtered frame (vrb_frame_resize_multiscaling)");
  if (dev) *dev = c->d_filtered;
  if (w) *w = c->flw;
  if (h) *h = c->flh;
  return VRB_OK;
}

__global__ void k_ff_to_f32(const __half* __restrict__ in, float* __restrict__ out, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    uint2 pk = reinterpret_cast<const uint2*>(in)[i];
    float2 a = __half22float2(*reinterpret_cast<__half2*>(&pk.x)), b = __half22float2(*reinterpret_cast<__half2*>(&pk.y));
    reinterpret_cast<float4*>(out)[i] = make_float4(a.x, a.y, b.x, b.y);
  }
}

extern "C" int vrb_filtered_frame_read_rgba32f(vrb_ctx* c, float* host_out) {
  VRB_REQUIRE(c && host_out, VRB_ERR_INVALID, "vrb_filtered_frame_read_rgba32f: NULL argument");
  VRB_REQUIRE(c->d_filtered, VRB_ERR_STATE, "vrb_filtered_frame_read_rgba32f: no filtered frame (vrb_frame_resize_multiscaling)");
  VRB_CUDA(cudaSetDevice(c->device));
  const size_t n4 = (size_t)c->flw * c->flh;
  float* tmp = nullptr;
  VRB_CUDA(cudaMalloc(&tmp, n4 * 4 * sizeof(float)));
  k_ff_to_f32<<<(int)std::min<size_t>((n4 + 255) / 256, 148 * 8), 256, 0, c->stream>>>(c->d_filtered, tmp, n4);
  c->launches++;
  cudaError_t e = cudaMemcpyAsync(host_out, tmp, n4 * 4 * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
  cudaError_t e2 = cudaStreamSynchronize(c->stream);
  cudaFree(tmp);
  VRB_REQUIRE(e == cudaSuccess && e2 == cudaSuccess, VRB_ERR_CUDA, "vrb_filtered_frame_read_rgba32f: copy failed");
  return VRB_OK;
}
