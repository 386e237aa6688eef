// peaks.cu -- measured rooflines the march kernels are reported against (MEASURED_PEAKS.json has HBM only).
//   vrb_measure_l1_bandwidth : L1-resident 128-bit loads, one 32 KB window per CTA (fits L1), all SMs busy.
//   vrb_measure_hbm_bandwidth: plain device-to-device copy of a buffer much larger than L2 (read + write bytes).
//   vrb_measure_gather_rate  : tex2Dgather on an R32F 2-D array (the SAT atlas format of march_ebs), tex-cache
//                              resident, 32 coherent lanes (8x4 pixel footprints), all SMs busy: the texture-pipe
//                              ceiling the box queries of k_ebs_coop are reported against.
#include "vrb_internal.cuh"
#include <vector>
#include <cstring>

__global__ void __launch_bounds__(1024) k_l1_peak(const float4* __restrict__ buf, float* __restrict__ sink, int iters, int window4) {
  // every CTA owns a private 32 KB window; after the first pass all loads hit L1
  const float4* p = buf + (size_t)blockIdx.x * window4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int i = threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float4 v = __ldca(p + ((i + k * 1024) & (window4 - 1)));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    i = (i + 37) & (window4 - 1);
  }
  if (acc.x + acc.y + acc.z + acc.w == 12345.678f) sink[0] = acc.x;
}

extern "C" int vrb_measure_l1_bandwidth(vrb_ctx* c, double* gb_per_s) {
  VRB_REQUIRE(c && gb_per_s, VRB_ERR_INVALID, "vrb_measure_l1_bandwidth: NULL argument");
  VRB_CUDA(cudaSetDevice(c->device));
  cudaDeviceProp prop;
  VRB_CUDA(cudaGetDeviceProperties(&prop, c->device));
  const int ctas = prop.multiProcessorCount * 2, window4 = 2048 /* float4 = 32 KB */, iters = 2000;
  float4* buf = nullptr; float* sink = nullptr;
  VRB_CUDA(cudaMalloc(&buf, (size_t)ctas * window4 * sizeof(float4)));
  VRB_CUDA(cudaMalloc(&sink, sizeof(float)));
  VRB_CUDA(cudaMemsetAsync(buf, 0, (size_t)ctas * window4 * sizeof(float4), c->stream));
  cudaEvent_t e0, e1;
  VRB_CUDA(cudaEventCreate(&e0)); VRB_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    VRB_CUDA(cudaEventRecord(e0, c->stream));
    k_l1_peak<<<ctas, 1024, 0, c->stream>>>(buf, sink, iters, window4);
    VRB_CUDA(cudaEventRecord(e1, c->stream));
    VRB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f; VRB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
    c->launches++;
  }
  VRB_CUDA(cudaGetLastError());
  double bytes = (double)ctas * 1024.0 * iters * 8.0 * 16.0;
  *gb_per_s = bytes / (best * 1e-3) / 1e9;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf); cudaFree(sink);
  return VRB_OK;
}

extern "C" int vrb_measure_hbm_bandwidth(vrb_ctx* c, double* gb_per_s) {
  VRB_REQUIRE(c && gb_per_s, VRB_ERR_INVALID, "vrb_measure_hbm_bandwidth: NULL argument");
  VRB_CUDA(cudaSetDevice(c->device));
  const size_t bytes = (size_t)1 << 30;
  void *a = nullptr, *b = nullptr;
  VRB_CUDA(cudaMalloc(&a, bytes)); VRB_CUDA(cudaMalloc(&b, bytes));
  VRB_CUDA(cudaMemsetAsync(a, 1, bytes, c->stream));
  cudaEvent_t e0, e1;
  VRB_CUDA(cudaEventCreate(&e0)); VRB_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    VRB_CUDA(cudaEventRecord(e0, c->stream));
    VRB_CUDA(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice, c->stream));
    VRB_CUDA(cudaEventRecord(e1, c->stream));
    VRB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f; VRB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  *gb_per_s = 2.0 * (double)bytes / (best * 1e-3) / 1e9;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(a); cudaFree(b);
  return VRB_OK;
}

__global__ void __launch_bounds__(256) k_gather_peak(cudaTextureObject_t tex, float* __restrict__ sink, int iters, int side) {
  // a warp covers an 8x4 texel patch (as the rays of a warp do); every CTA walks its own corner of a small array
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float x = (float)((blockIdx.x * 7 + warp * 9 + (lane & 7)) % (side - 24) + 1);
  float y = (float)((blockIdx.x * 3 + warp * 5 + (lane >> 3)) % (side - 24) + 1);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int it = 0; it < iters; ++it) {
    // the footprint moves every iteration (a loop-invariant fetch would be hoisted out of the loop)
    const float xo = x + (float)((it & 3) * 4), yo = y + (float)(((it >> 2) & 1) * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float4 v = tex2Dgather<float4>(tex, xo + (float)(k & 3), yo + (float)(k >> 2) * 4.0f, 0);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  if (acc.x + acc.y + acc.z + acc.w == 12345.678f) sink[0] = acc.x;
}

extern "C" int vrb_measure_gather_rate(vrb_ctx* c, double* ggathers_per_s) {
  VRB_REQUIRE(c && ggathers_per_s, VRB_ERR_INVALID, "vrb_measure_gather_rate: NULL argument");
  VRB_CUDA(cudaSetDevice(c->device));
  cudaDeviceProp prop;
  VRB_CUDA(cudaGetDeviceProperties(&prop, c->device));
  const int side = 64;                                  // 64x64 R32F = 16 KB: stays in the texture cache
  cudaArray_t arr = nullptr; cudaTextureObject_t tex = 0; float* sink = nullptr;
  cudaChannelFormatDesc fd = cudaCreateChannelDesc<float>();
  VRB_CUDA(cudaMallocArray(&arr, &fd, side, side, cudaArrayTextureGather));
  std::vector<float> host((size_t)side * side, 1.0f);
  VRB_CUDA(cudaMemcpy2DToArray(arr, 0, 0, host.data(), side * sizeof(float), side * sizeof(float), side, cudaMemcpyHostToDevice));
  cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
  rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
  cudaTextureDesc td; memset(&td, 0, sizeof(td));
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
  VRB_CUDA(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  VRB_CUDA(cudaMalloc(&sink, sizeof(float)));
  const int ctas = prop.multiProcessorCount * 8, iters = 2000;
  cudaEvent_t e0, e1;
  VRB_CUDA(cudaEventCreate(&e0)); VRB_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    VRB_CUDA(cudaEventRecord(e0, c->stream));
    k_gather_peak<<<ctas, 256, 0, c->stream>>>(tex, sink, iters, side);
    VRB_CUDA(cudaEventRecord(e1, c->stream));
    VRB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f; VRB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
    c->launches++;
  }
  VRB_CUDA(cudaGetLastError());
  *ggathers_per_s = (double)ctas * 256.0 * iters * 8.0 / (best * 1e-3) / 1e9;     // lane-level gathers (16 B each)
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaDestroyTextureObject(tex); cudaFreeArray(arr); cudaFree(sink);
  return VRB_OK;
}

// ---- trilinear tex3D ceiling (hardware filter mode) -------------------------------------------------------------------
// What one tex3D<float>() of the hardware-filter marchers costs at best: a cache-resident R16F 3-D array, GL_LINEAR,
// the 32 lanes of a warp on an 8x4 patch of neighbouring texels at fractional positions (as the rays of a warp are).
__global__ void __launch_bounds__(256) k_tex3d_peak(cudaTextureObject_t tex, float* __restrict__ sink, int iters, int side) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float x = (float)((blockIdx.x * 7 + warp * 9 + (lane & 7)) % (side - 12) + 1) + 0.37f;
  const float y = (float)((blockIdx.x * 3 + warp * 5 + (lane >> 3)) % (side - 12) + 1) + 0.61f;
  float z = (float)((blockIdx.x + warp) % (side - 12) + 1) + 0.13f;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    const float zo = z + (float)(it & 7) * 0.5f;        // the footprint moves along the "ray" every iteration
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += tex3D<float>(tex, x + (float)(k & 3) * 0.5f, y + (float)(k >> 2) * 0.5f, zo + (float)k * 0.25f);
  }
  if (acc == 12345.678f) sink[0] = acc;
}

extern "C" int vrb_measure_tex3d_rate(vrb_ctx* c, double* gfetches_per_s) {
  VRB_REQUIRE(c && gfetches_per_s, VRB_ERR_INVALID, "vrb_measure_tex3d_rate: NULL argument");
  VRB_CUDA(cudaSetDevice(c->device));
  cudaDeviceProp prop;
  VRB_CUDA(cudaGetDeviceProperties(&prop, c->device));
  const int side = 24;                                  // 24^3 R16F = 27 KB: stays in the texture cache
  cudaArray_t arr = nullptr; cudaTextureObject_t tex = 0; float* sink = nullptr;
  cudaChannelFormatDesc fd = cudaCreateChannelDescHalf();
  VRB_CUDA(cudaMalloc3DArray(&arr, &fd, make_cudaExtent(side, side, side), cudaArrayDefault));
  std::vector<__half> host((size_t)side * side * side);
  for (size_t i = 0; i < host.size(); ++i) host[i] = __float2half((float)(i % 97) / 97.0f);
  cudaMemcpy3DParms cp; memset(&cp, 0, sizeof(cp));
  cp.srcPtr = make_cudaPitchedPtr((void*)host.data(), side * sizeof(__half), side, side);
  cp.dstArray = arr; cp.extent = make_cudaExtent(side, side, side); cp.kind = cudaMemcpyHostToDevice;
  cudaError_t e = cudaMemcpy3D(&cp);
  if (e != cudaSuccess) { cudaFreeArray(arr); vrb_set_error("vrb_measure_tex3d_rate: cudaMemcpy3D: %s", cudaGetErrorString(e)); return VRB_ERR_CUDA; }
  cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
  rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
  cudaTextureDesc td; memset(&td, 0, sizeof(td));
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
  e = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
  if (e != cudaSuccess) { cudaFreeArray(arr); vrb_set_error("vrb_measure_tex3d_rate: cudaCreateTextureObject: %s", cudaGetErrorString(e)); return VRB_ERR_CUDA; }
  e = cudaMalloc(&sink, sizeof(float));
  if (e != cudaSuccess) { cudaDestroyTextureObject(tex); cudaFreeArray(arr); vrb_set_error("vrb_measure_tex3d_rate: cudaMalloc: %s", cudaGetErrorString(e)); return VRB_ERR_CUDA; }
  const int ctas = prop.multiProcessorCount * 8, iters = 2000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, c->stream);
    k_tex3d_peak<<<ctas, 256, 0, c->stream>>>(tex, sink, iters, side);
    cudaEventRecord(e1, c->stream);
    cudaEventSynchronize(e1);
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
    c->launches++;
  }
  e = cudaGetLastError();
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaDestroyTextureObject(tex); cudaFreeArray(arr); cudaFree(sink);
  VRB_CUDA(e);
  *gfetches_per_s = (double)ctas * 256.0 * iters * 8.0 / (best * 1e-3) / 1e9;     // lane-level trilinear fetches
  return VRB_OK;
}

// ---- scattered 16-bit load ceiling (exact filter mode) ------------------------------------------------------------------
// What the eight fp16 taps of a software trilinear fetch cost at best: an L1-resident padded fp16 brick per CTA, the 32
// lanes of a warp on neighbouring texels (an 8x4 patch walking along z), eight LDG.U16 per "sample" at the corner offsets
// of the footprint.  Reported as lane-level 16-bit loads per second.
__global__ void __launch_bounds__(256) k_ldg16_peak(const __half* __restrict__ buf, float* __restrict__ sink, int iters, int side) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const __half* brick = buf + (size_t)blockIdx.x * side * side * side;       // side^3 halves per CTA: 16^3 * 2 B = 8 KB
  const int x = (warp * 3 + (lane & 7)) % (side - 1), y = (warp * 5 + (lane >> 3)) % (side - 1);
  const int slice = side * side;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    const int z = (it + warp) % (side - 1);
    const __half* p = brick + z * slice + y * side + x;
    acc += __half2float(__ldg(p)) + __half2float(__ldg(p + 1)) + __half2float(__ldg(p + side)) + __half2float(__ldg(p + side + 1)) +
           __half2float(__ldg(p + slice)) + __half2float(__ldg(p + slice + 1)) + __half2float(__ldg(p + slice + side)) +
           __half2float(__ldg(p + slice + side + 1));
  }
  if (acc == 12345.678f) sink[0] = acc;
}

extern "C" int vrb_measure_ldg16_rate(vrb_ctx* c, double* gloads_per_s) {
  VRB_REQUIRE(c && gloads_per_s, VRB_ERR_INVALID, "vrb_measure_ldg16_rate: NULL argument");
  VRB_CUDA(cudaSetDevice(c->device));
  cudaDeviceProp prop;
  VRB_CUDA(cudaGetDeviceProperties(&prop, c->device));
  const int side = 16, ctas = prop.multiProcessorCount * 8, iters = 4000;
  __half* buf = nullptr; float* sink = nullptr;
  const size_t bytes = (size_t)ctas * side * side * side * sizeof(__half);
  VRB_CUDA(cudaMalloc(&buf, bytes));
  cudaError_t e = cudaMalloc(&sink, sizeof(float));
  if (e != cudaSuccess) { cudaFree(buf); vrb_set_error("vrb_measure_ldg16_rate: cudaMalloc: %s", cudaGetErrorString(e)); return VRB_ERR_CUDA; }
  cudaMemsetAsync(buf, 0, bytes, c->stream);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, c->stream);
    k_ldg16_peak<<<ctas, 256, 0, c->stream>>>(buf, sink, iters, side);
    cudaEventRecord(e1, c->stream);
    cudaEventSynchronize(e1);
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
    c->launches++;
  }
  e = cudaGetLastError();
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf); cudaFree(sink);
  VRB_CUDA(e);
  *gloads_per_s = (double)ctas * 256.0 * iters * 8.0 / (best * 1e-3) / 1e9;
  return VRB_OK;
}
