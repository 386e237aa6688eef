// vrb_api.cu -- context, inputs and output frame of libvrb200.so (C ABI in include/vrb200.h).
#include "vrb_internal.cuh"
#include <cstdarg>
#include <cstring>
#include <vector>

static thread_local std::string g_last_error;

void vrb_set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap; va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

extern "C" const char* vrb_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* vrb_version(void) { return "vrb200 0.1 (sm_100a)"; }

extern "C" int vrb_ctx_create(int device, vrb_ctx** out) {
  VRB_REQUIRE(out != nullptr, VRB_ERR_INVALID, "vrb_ctx_create: out is NULL");
  int ndev = 0;
  VRB_CUDA(cudaGetDeviceCount(&ndev));
  VRB_REQUIRE(device >= 0 && device < ndev, VRB_ERR_INVALID, "vrb_ctx_create: device %d of %d", device, ndev);
  VRB_CUDA(cudaSetDevice(device));
  vrb_ctx* c = new vrb_ctx();
  c->device = device;
  cudaError_t e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete c; vrb_set_error("cudaStreamCreate: %s", cudaGetErrorString(e)); return VRB_ERR_CUDA; }
  c->stream = c->own_stream;
  e = cudaMalloc(&c->d_counter, 2 * sizeof(unsigned long long));
  if (e != cudaSuccess) { cudaStreamDestroy(c->own_stream); delete c; vrb_set_error("cudaMalloc: %s", cudaGetErrorString(e)); return VRB_ERR_CUDA; }
  if (const char* o = getenv("VRB_SAT_ORDER")) c->sat_order = (o[0] == 's' || o[0] == 'S' || o[0] == '1') ? VRB_SAT_ORDER_SCAN : VRB_SAT_ORDER_REFERENCE;
  if (const char* f = getenv("VRB_FILTER")) c->filter_mode = (f[0] == 'h' || f[0] == 'H' || f[0] == '1') ? VRB_FILTER_HARDWARE : VRB_FILTER_EXACT;
  *out = c;
  return VRB_OK;
}

void vrb_free_vol_atlas(vrb_ctx* c) {
  if (c->vol_tex) cudaDestroyTextureObject(c->vol_tex);
  if (c->vol_array) cudaFreeArray(c->vol_array);
  c->vol_tex = 0; c->vol_array = nullptr; c->vol_atlas_tiles_x = 0;
  if (c->vol_tex3d) cudaDestroyTextureObject(c->vol_tex3d);
  if (c->vol_array3d) cudaFreeArray(c->vol_array3d);
  c->vol_tex3d = 0; c->vol_array3d = nullptr;
}

// Hardware-filtered volume texture: what the reference binds as GL_R16F / GL_LINEAR / GL_CLAMP_TO_EDGE
// (libs/volvis_utils/utils.cpp:20-56).  Built lazily, only in VRB_FILTER_HARDWARE mode.
int vrb_vol_tex3d_prepare(vrb_ctx* c) {
  if (c->filter_mode != 1 || c->vol_tex3d) return VRB_OK;
  VRB_REQUIRE(c->d_vol, VRB_ERR_STATE, "no volume uploaded");
  cudaChannelFormatDesc fd = cudaCreateChannelDescHalf();
  cudaExtent ext = make_cudaExtent((size_t)c->vw, (size_t)c->vh, (size_t)c->vd);
  VRB_CUDA(cudaMalloc3DArray(&c->vol_array3d, &fd, ext, cudaArrayDefault));
  const size_t pw = (size_t)c->vw + 2, ph = (size_t)c->vh + 2;
  cudaMemcpy3DParms cp; memset(&cp, 0, sizeof(cp));
  cp.srcPtr = make_cudaPitchedPtr((void*)(c->d_vol + pw * ph + pw + 1), pw * sizeof(__half), (size_t)c->vw, ph);
  cp.dstArray = c->vol_array3d;
  cp.extent = ext;
  cp.kind = cudaMemcpyDeviceToDevice;
  VRB_CUDA(cudaMemcpy3DAsync(&cp, c->stream));
  cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
  rd.resType = cudaResourceTypeArray; rd.res.array.array = c->vol_array3d;
  cudaTextureDesc td; memset(&td, 0, sizeof(td));
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
  VRB_CUDA(cudaCreateTextureObject(&c->vol_tex3d, &rd, &td, nullptr));
  return VRB_OK;
}

extern "C" int vrb_ctx_set_filter(vrb_ctx* c, int mode) {
  VRB_REQUIRE(c, VRB_ERR_INVALID, "vrb_ctx_set_filter: NULL context");
  VRB_REQUIRE(mode == VRB_FILTER_EXACT || mode == VRB_FILTER_HARDWARE, VRB_ERR_INVALID, "vrb_ctx_set_filter: mode %d", mode);
  c->filter_mode = mode;
  return VRB_OK;
}
extern "C" int vrb_ctx_get_filter(const vrb_ctx* c) { return c ? c->filter_mode : -1; }

// fp16 gather atlas of the padded volume: padded slice z -> tile (z % T, z / T).  Returns VRB_ERR_UNSUPPORTED when the
// atlas would exceed the 32768 x 32768 texture-gather limit (e.g. 1024^3 bricks): the marchers then use plain loads.
static int build_vol_atlas(vrb_ctx* c) {
  // Off by default: measured on B200 the two-gather path is NOT faster than eight 16-bit loads for the fp16 volume
  // (cfg1 0.311 vs 0.298 ms, cfg4 1751 vs 1346 ms; L1 hit rate of the linear layout is already ~95 %), unlike the
  // fp32 SAT where it wins 1.6x.  VRB_VOL_GATHER=1 enables it for experiments.
  const char* env = getenv("VRB_VOL_GATHER");
  if (!env || atoi(env) == 0) return VRB_ERR_UNSUPPORTED;
  const int pw = c->vw + 2, ph = c->vh + 2, pd = c->vd + 2;
  int T = 1;
  while ((long long)T * T < pd) ++T;
  T = std::min(T, 32768 / pw);
  if (T < 1) return VRB_ERR_UNSUPPORTED;
  const int rows = (pd + T - 1) / T;
  if ((long long)rows * ph > 32768) return VRB_ERR_UNSUPPORTED;
  cudaChannelFormatDesc fd = cudaCreateChannelDescHalf();
  VRB_CUDA(cudaMallocArray(&c->vol_array, &fd, (size_t)T * pw, (size_t)rows * ph, cudaArrayTextureGather));
  for (int z = 0; z < pd; ++z) {
    const __half* src = c->d_vol + (size_t)z * pw * ph;
    VRB_CUDA(cudaMemcpy2DToArrayAsync(c->vol_array, (size_t)(z % T) * pw * sizeof(__half), (size_t)(z / T) * ph, src, (size_t)pw * sizeof(__half),
                                      (size_t)pw * sizeof(__half), ph, cudaMemcpyDeviceToDevice, c->stream));
  }
  cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
  rd.resType = cudaResourceTypeArray; rd.res.array.array = c->vol_array;
  cudaTextureDesc td; memset(&td, 0, sizeof(td));
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
  VRB_CUDA(cudaCreateTextureObject(&c->vol_tex, &rd, &td, nullptr));
  c->vol_atlas_tiles_x = T;
  return VRB_OK;
}

static void free_volume(vrb_ctx* c) {
  vrb_free_vol_atlas(c);
  if (c->d_raw) cudaFree(c->d_raw);
  if (c->d_vol) cudaFree(c->d_vol);
  if (c->d_sat) cudaFree(c->d_sat);
  if (c->d_sat_slab64) cudaFree(c->d_sat_slab64);
  c->d_sat_slab64 = nullptr;
  if (c->d_sat_packed) cudaFree(c->d_sat_packed);
  c->d_sat_packed = nullptr;
  vrb_free_sat_atlas(c);
  c->d_raw = nullptr; c->d_vol = nullptr; c->d_sat = nullptr;
  c->sat_w = c->sat_h = c->sat_d = 0;
  vrb_free_pyramid(c);      // every pre-pass product derives from the volume
  vrb_free_vct(c);
  vrb_free_gradient(c);
  vrb_free_cells(c);
  vrb_free_vol_quads(c);
  vrb_free_light_cache(c);
  vrb_free_cta_order(c);
}

extern "C" int vrb_ctx_destroy(vrb_ctx* c) {
  if (!c) return VRB_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  free_volume(c);
  if (c->d_tf_rgbt) cudaFree(c->d_tf_rgbt);
  if (c->d_tf_rgba) cudaFree(c->d_tf_rgba);
  if (c->d_frame) cudaFree(c->d_frame);
  for (int i = 0; i < 2; ++i) if (c->d_frame_extra[i]) cudaFree(c->d_frame_extra[i]);
  if (c->d_partial) cudaFree(c->d_partial);
  if (c->d_brick_alpha) cudaFree(c->d_brick_alpha);
  vrb_free_filtered(c);
  if (c->d_counter) cudaFree(c->d_counter);
  for (int i = 0; i < 2; ++i) if (c->d_cone_sections[i]) cudaFree(c->d_cone_sections[i]);
  for (int i = 0; i < 2; ++i) if (c->ev_kern[i]) cudaEventDestroy(c->ev_kern[i]);
  if (c->d_dos_packed) cudaFree(c->d_dos_packed);
  vrb_free_shade_list(c);
  for (int i = 0; i < 2; ++i) if (c->d_gt_rays[i]) cudaFree(c->d_gt_rays[i]);
  if (c->copy_stream) {
    cudaStreamSynchronize(c->copy_stream);
    for (int i = 0; i < 2; ++i) { if (c->d_stage[i]) cudaFree(c->d_stage[i]); cudaEventDestroy(c->ev_ready[i]); cudaEventDestroy(c->ev_copied[i]); }
    cudaStreamDestroy(c->copy_stream);
  }
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
  return VRB_OK;
}

extern "C" int vrb_ctx_set_stream(vrb_ctx* c, void* s) {
  VRB_REQUIRE(c, VRB_ERR_INVALID, "ctx is NULL");
  c->stream = s ? (cudaStream_t)s : c->own_stream;
  return VRB_OK;
}

extern "C" int vrb_ctx_synchronize(vrb_ctx* c) {
  VRB_REQUIRE(c, VRB_ERR_INVALID, "ctx is NULL");
  VRB_CUDA(cudaSetDevice(c->device));
  VRB_CUDA(cudaStreamSynchronize(c->stream));
  return VRB_OK;
}

extern "C" int vrb_ctx_set_partition(vrb_ctx* c, const vrb_partition* p) {
  VRB_REQUIRE(c && p, VRB_ERR_INVALID, "NULL argument");
  VRB_REQUIRE(p->nranks >= 1 && p->rank >= 0 && p->rank < p->nranks && p->tile_w > 0 && p->tile_h > 0,
              VRB_ERR_INVALID, "bad partition rank %d/%d tile %dx%d", p->rank, p->nranks, p->tile_w, p->tile_h);
  c->part = PartView{p->rank, p->nranks, p->tile_w, p->tile_h, 0};
  return VRB_OK;
}

extern "C" uint64_t vrb_launch_count(const vrb_ctx* c) { return c ? c->launches : 0; }
extern "C" uint64_t vrb_last_sample_count(const vrb_ctx* c) { return c ? c->last_samples : 0; }
extern "C" uint64_t vrb_last_aux_count(const vrb_ctx* c) { return c ? c->last_aux : 0; }
extern "C" int vrb_step_multiples_exact_f(float step_size, float longest_ray) {
  return vrb_step_multiples_exact(step_size, longest_ray) ? 1 : 0;
}
extern "C" int vrb_ctx_set_kernel_timing(vrb_ctx* c, int on) {
  VRB_REQUIRE(c, VRB_ERR_INVALID, "vrb_ctx_set_kernel_timing: ctx is NULL");
  VRB_CUDA(cudaSetDevice(c->device));
  if (on && !c->ev_kern[0]) { VRB_CUDA(cudaEventCreate(&c->ev_kern[0])); VRB_CUDA(cudaEventCreate(&c->ev_kern[1])); }
  c->time_kernels = on != 0;
  c->kern_timed = false;
  return VRB_OK;
}
extern "C" int vrb_last_kernel_ms(vrb_ctx* c, float* ms, const char** name) {
  VRB_REQUIRE(c && ms, VRB_ERR_INVALID, "vrb_last_kernel_ms: NULL argument");
  VRB_REQUIRE(c->kern_timed, VRB_ERR_STATE, "vrb_last_kernel_ms: no render call was timed (vrb_ctx_set_kernel_timing)");
  VRB_CUDA(cudaEventSynchronize(c->ev_kern[1]));
  VRB_CUDA(cudaEventElapsedTime(ms, c->ev_kern[0], c->ev_kern[1]));
  if (name) *name = c->kern_name;
  return VRB_OK;
}
extern "C" float vrb_last_prepass_ms(const vrb_ctx* c) { return c ? c->last_prepass_ms : 0.f; }
extern "C" int vrb_sat_layout(const vrb_ctx* c) { return c ? ((c->sat_pack == 8 && !c->sat_tex) ? 1 : c->sat_pack) : 0; }

// ---------------------------------------------------------------------------------------------------------
// volume: raw voxels -> padded fp16 texels, value = half(float(double(v)/max)) exactly as the reference's
// GetNormalizedSample -> (GLfloat) -> GL_R16F chain (structuredgridvolume.cpp:121-151, utils.cpp:20-56).
// One thread per padded texel; clamp-to-edge is materialised as the replicated border.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_volume_to_padded_f16(const T* __restrict__ raw, __half* __restrict__ out, int w, int h, int d, double maxv) {
  int pw = w + 2, ph = h + 2, pd = d + 2;
  long long n = (long long)pw * ph * pd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int x = (int)(i % pw); long long r = i / pw;
    int y = (int)(r % ph); int z = (int)(r / ph);
    int sx = min(max(x - 1, 0), w - 1), sy = min(max(y - 1, 0), h - 1), sz = min(max(z - 1, 0), d - 1);
    T v = raw[(size_t)sx + (size_t)w * ((size_t)sy + (size_t)h * (size_t)sz)];
    float f = __double2float_rn(__ddiv_rn((double)v, maxv));
    out[i] = __float2half_rn(f);
  }
}

static int volume_finish(vrb_ctx* c, int w, int h, int d, int bpv, const float scale[3]) {
  size_t np = (size_t)(w + 2) * (h + 2) * (d + 2);
  VRB_CUDA(cudaMalloc(&c->d_vol, np * sizeof(__half)));
  int threads = 256;
  int blocks = (int)std::min<size_t>((np + threads - 1) / threads, 148 * 32);
  if (bpv == 1)
    k_volume_to_padded_f16<uint8_t><<<blocks, threads, 0, c->stream>>>((const uint8_t*)c->d_raw, c->d_vol, w, h, d, 255.0);
  else
    k_volume_to_padded_f16<uint16_t><<<blocks, threads, 0, c->stream>>>((const uint16_t*)c->d_raw, c->d_vol, w, h, d, 65535.0);
  VRB_CUDA(cudaGetLastError());
  c->launches++;
  c->vw = w; c->vh = h; c->vd = d; c->bpv = bpv;
  c->scale[0] = scale ? scale[0] : 1.0f; c->scale[1] = scale ? scale[1] : 1.0f; c->scale[2] = scale ? scale[2] : 1.0f;
  int rc = build_vol_atlas(c);
  if (rc == VRB_ERR_UNSUPPORTED) { vrb_free_vol_atlas(c); rc = VRB_OK; }
  return rc;
}

static int volume_upload_common(vrb_ctx* c, const void* vox, bool on_device, int w, int h, int d, int bpv, const float scale[3]) {
  VRB_REQUIRE(c && vox, VRB_ERR_INVALID, "vrb_volume_upload: NULL argument");
  VRB_REQUIRE(w > 0 && h > 0 && d > 0 && w <= 4096 && h <= 4096 && d <= 4096, VRB_ERR_INVALID,
              "vrb_volume_upload: bad resolution %dx%dx%d", w, h, d);
  VRB_REQUIRE(bpv == 1 || bpv == 2, VRB_ERR_UNSUPPORTED, "vrb_volume_upload: bytes_per_voxel %d (1 or 2)", bpv);
  VRB_CUDA(cudaSetDevice(c->device));
  free_volume(c);
  size_t bytes = (size_t)w * h * d * bpv;
  VRB_CUDA(cudaMalloc(&c->d_raw, bytes));
  VRB_CUDA(cudaMemcpyAsync(c->d_raw, vox, bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
  int rc = volume_finish(c, w, h, d, bpv, scale);
  if (rc != VRB_OK) return rc;
  if (!on_device) VRB_CUDA(cudaStreamSynchronize(c->stream));   // host array is borrowed for the call only
  return VRB_OK;
}

extern "C" int vrb_volume_upload(vrb_ctx* c, const void* vox, int w, int h, int d, int bpv, const float scale[3]) {
  return volume_upload_common(c, vox, false, w, h, d, bpv, scale);
}
extern "C" int vrb_volume_upload_device(vrb_ctx* c, const void* vox, int w, int h, int d, int bpv, const float scale[3]) {
  return volume_upload_common(c, vox, true, w, h, d, bpv, scale);
}

// ---------------------------------------------------------------------------------------------------------
// transfer function textures: GL_FLOAT client array -> RGBA16F texels (kept as fp16-rounded float4), padded by one
// replicated texel at both ends (clamp-to-edge).
// ---------------------------------------------------------------------------------------------------------
static int upload_tf_table(vrb_ctx* c, const float* src, int n, float4** dst) {
  std::vector<float4> h((size_t)n + 2);
  for (int i = 0; i < n + 2; ++i) {
    int s = std::min(std::max(i - 1, 0), n - 1);
    float4 t;
    t.x = __half2float(__float2half_rn(src[4 * s + 0]));
    t.y = __half2float(__float2half_rn(src[4 * s + 1]));
    t.z = __half2float(__float2half_rn(src[4 * s + 2]));
    t.w = __half2float(__float2half_rn(src[4 * s + 3]));
    h[i] = t;
  }
  if (*dst) { VRB_CUDA(cudaFree(*dst)); *dst = nullptr; }
  VRB_CUDA(cudaMalloc(dst, h.size() * sizeof(float4)));
  VRB_CUDA(cudaMemcpyAsync(*dst, h.data(), h.size() * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
  VRB_CUDA(cudaStreamSynchronize(c->stream));
  return VRB_OK;
}

extern "C" int vrb_tf_upload(vrb_ctx* c, const float* rgbt, const float* rgba, int n) {
  VRB_REQUIRE(c && rgbt, VRB_ERR_INVALID, "vrb_tf_upload: NULL argument");
  VRB_REQUIRE(n >= 1 && n <= (1 << 20), VRB_ERR_INVALID, "vrb_tf_upload: bad size %d", n);
  VRB_CUDA(cudaSetDevice(c->device));
  int rc = upload_tf_table(c, rgbt, n, &c->d_tf_rgbt);
  if (rc != VRB_OK) return rc;
  if (rgba) {
    rc = upload_tf_table(c, rgba, n, &c->d_tf_rgba);
    if (rc != VRB_OK) return rc;
  } else if (c->d_tf_rgba) {
    // no opacity table for this transfer function: drop the previous one (it may have another size), so that
    // vrb_extcoef_build reports VRB_ERR_STATE instead of reading a stale table
    VRB_CUDA(cudaFree(c->d_tf_rgba));
    c->d_tf_rgba = nullptr;
  }
  c->tf_n = n;
  c->cell_flags_valid = false;
  return VRB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// output frame
// ---------------------------------------------------------------------------------------------------------
extern "C" int vrb_frame_resize(vrb_ctx* c, int w, int h) {
  VRB_REQUIRE(c, VRB_ERR_INVALID, "ctx is NULL");
  VRB_REQUIRE(w > 0 && h > 0 && w <= 16384 && h <= 16384, VRB_ERR_INVALID, "vrb_frame_resize: bad size %dx%d", w, h);
  VRB_CUDA(cudaSetDevice(c->device));
  if (w == c->fw && h == c->fh && c->d_frame) return VRB_OK;
  if (c->d_frame) { VRB_CUDA(cudaStreamSynchronize(c->stream)); VRB_CUDA(cudaFree(c->d_frame)); c->d_frame = nullptr; }
  for (int i = 0; i < 2; ++i) if (c->d_frame_extra[i]) { VRB_CUDA(cudaFree(c->d_frame_extra[i])); c->d_frame_extra[i] = nullptr; }
  vrb_free_filtered(c);          // UpdateScreenResolution deletes m_filtered_screen_output with the old frame (renderoutputframe.cpp:80-81)
  c->d_frame_target = nullptr;
  VRB_CUDA(cudaMalloc(&c->d_frame, (size_t)w * h * 4 * sizeof(__half)));
  c->fw = w; c->fh = h;
  VRB_CUDA(cudaMemsetAsync(c->d_frame, 0, (size_t)w * h * 4 * sizeof(__half), c->stream));
  return VRB_OK;
}

extern "C" int vrb_frame_clear(vrb_ctx* c) {
  VRB_REQUIRE(c && c->d_frame, VRB_ERR_STATE, "vrb_frame_clear: no frame");
  VRB_CUDA(cudaSetDevice(c->device));
  VRB_CUDA(cudaMemsetAsync(c->d_frame, 0, (size_t)c->fw * c->fh * 4 * sizeof(__half), c->stream));
  return VRB_OK;
}

__global__ void k_frame_to_f32(const __half* __restrict__ in, float* __restrict__ out, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    uint2 pk = reinterpret_cast<const uint2*>(in)[i];
    __half2 lo = *reinterpret_cast<__half2*>(&pk.x), hi = *reinterpret_cast<__half2*>(&pk.y);
    float2 a = __half22float2(lo), b = __half22float2(hi);
    reinterpret_cast<float4*>(out)[i] = make_float4(a.x, a.y, b.x, b.y);
  }
}

extern "C" int vrb_frame_read_rgba32f(vrb_ctx* c, float* host_out) {
  VRB_REQUIRE(c && host_out, VRB_ERR_INVALID, "NULL argument");
  VRB_REQUIRE(c->d_frame, VRB_ERR_STATE, "vrb_frame_read_rgba32f: no frame");
  VRB_CUDA(cudaSetDevice(c->device));
  size_t n4 = (size_t)c->fw * c->fh;
  float* tmp = nullptr;
  VRB_CUDA(cudaMallocAsync(&tmp, n4 * 4 * sizeof(float), c->stream));
  int blocks = (int)std::min<size_t>((n4 + 255) / 256, 148 * 8);
  k_frame_to_f32<<<blocks, 256, 0, c->stream>>>(c->frame_ptr(), tmp, n4);
  c->launches++;
  VRB_CUDA(cudaGetLastError());
  VRB_CUDA(cudaMemcpyAsync(host_out, tmp, n4 * 4 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  VRB_CUDA(cudaFreeAsync(tmp, c->stream));
  VRB_CUDA(cudaStreamSynchronize(c->stream));
  return VRB_OK;
}

// Pipelined read-back: conversion behind the render on the context's stream, device->host copy on a copy stream, so
// that the copy of frame i overlaps the render of frame i+1 (double-buffered staging).
extern "C" int vrb_frame_read_rgba32f_async(vrb_ctx* c, float* host_out) {
  VRB_REQUIRE(c && host_out, VRB_ERR_INVALID, "NULL argument");
  VRB_REQUIRE(c->d_frame, VRB_ERR_STATE, "vrb_frame_read_rgba32f_async: no frame");
  VRB_CUDA(cudaSetDevice(c->device));
  const size_t n4 = (size_t)c->fw * c->fh;
  if (!c->copy_stream) {
    VRB_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      VRB_CUDA(cudaEventCreateWithFlags(&c->ev_ready[i], cudaEventDisableTiming));
      VRB_CUDA(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
    }
  }
  if (c->stage_px != n4) {
    VRB_CUDA(cudaStreamSynchronize(c->copy_stream));
    for (int i = 0; i < 2; ++i) {
      if (c->d_stage[i]) { VRB_CUDA(cudaFree(c->d_stage[i])); c->d_stage[i] = nullptr; }
      VRB_CUDA(cudaMalloc(&c->d_stage[i], n4 * 4 * sizeof(float)));
      c->stage_busy[i] = false;
    }
    c->stage_px = n4;
  }
  const int i = (int)(c->stage_next++ & 1u);
  if (c->stage_busy[i]) VRB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copied[i], 0));   // staging buffer still being copied out
  int blocks = (int)std::min<size_t>((n4 + 255) / 256, 148 * 8);
  k_frame_to_f32<<<blocks, 256, 0, c->stream>>>(c->frame_ptr(), c->d_stage[i], n4);
  VRB_CUDA(cudaGetLastError());
  c->launches++;
  VRB_CUDA(cudaEventRecord(c->ev_ready[i], c->stream));
  VRB_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_ready[i], 0));
  VRB_CUDA(cudaMemcpyAsync(host_out, c->d_stage[i], n4 * 4 * sizeof(float), cudaMemcpyDeviceToHost, c->copy_stream));
  VRB_CUDA(cudaEventRecord(c->ev_copied[i], c->copy_stream));
  c->stage_busy[i] = true;
  return VRB_OK;
}

extern "C" int vrb_frame_read_wait(vrb_ctx* c, int max_in_flight) {
  VRB_REQUIRE(c, VRB_ERR_INVALID, "NULL argument");
  VRB_REQUIRE(max_in_flight == 0 || max_in_flight == 1, VRB_ERR_INVALID, "vrb_frame_read_wait: max_in_flight %d (0 or 1)", max_in_flight);
  if (!c->copy_stream) return VRB_OK;
  VRB_CUDA(cudaSetDevice(c->device));
  // reads complete in issue order; the most recent one used slot (stage_next - 1) & 1
  const int newest = (int)((c->stage_next - 1u) & 1u), oldest = newest ^ 1;
  if (c->stage_busy[oldest]) { VRB_CUDA(cudaEventSynchronize(c->ev_copied[oldest])); c->stage_busy[oldest] = false; }
  if (max_in_flight == 0 && c->stage_busy[newest]) { VRB_CUDA(cudaEventSynchronize(c->ev_copied[newest])); c->stage_busy[newest] = false; }
  return VRB_OK;
}

// Sort-first without a frame reduce (see the header): redirect the marchers' pixel stores.
extern "C" int vrb_frame_set_target(vrb_ctx* c, void* rgba16f) {
  VRB_REQUIRE(c, VRB_ERR_INVALID, "vrb_frame_set_target: NULL context");
  VRB_REQUIRE(c->d_frame, VRB_ERR_STATE, "vrb_frame_set_target: no frame (vrb_frame_resize)");
  c->d_frame_target = (__half*)rgba16f;
  return VRB_OK;
}

extern "C" int vrb_frame_extra(vrb_ctx* c, int index, void** dev) {
  VRB_REQUIRE(c && dev, VRB_ERR_INVALID, "vrb_frame_extra: NULL argument");
  VRB_REQUIRE(index == 0 || index == 1, VRB_ERR_INVALID, "vrb_frame_extra: index %d (0 or 1)", index);
  VRB_REQUIRE(c->d_frame, VRB_ERR_STATE, "vrb_frame_extra: no frame (vrb_frame_resize)");
  VRB_CUDA(cudaSetDevice(c->device));
  if (!c->d_frame_extra[index]) {
    const size_t bytes = (size_t)c->fw * c->fh * 4 * sizeof(__half);
    VRB_CUDA(cudaMalloc(&c->d_frame_extra[index], bytes));
    VRB_CUDA(cudaMemsetAsync(c->d_frame_extra[index], 0, bytes, c->stream));
    VRB_CUDA(cudaStreamSynchronize(c->stream));
  }
  *dev = c->d_frame_extra[index];
  return VRB_OK;
}

extern "C" int vrb_frame_device_ptr(vrb_ctx* c, void** dev, int* w, int* h) {
  VRB_REQUIRE(c && dev, VRB_ERR_INVALID, "NULL argument");
  VRB_REQUIRE(c->d_frame, VRB_ERR_STATE, "vrb_frame_device_ptr: no frame");
  *dev = c->d_frame;
  if (w) *w = c->fw;
  if (h) *h = c->fh;
  return VRB_OK;
}
