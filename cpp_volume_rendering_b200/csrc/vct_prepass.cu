// vct_prepass.cu -- pre-passes of the voxel-cone-tracing renderer, which the reference runs on the CPU in double:
//   VCTPreProcessing::PreProcessSuperVoxels        (rc1pvctsg/preprocessingstages.cpp:35-145)
//   VCTPreProcessing::PreProcessPreIntegrationTable (:147-202, O(maxDensity^2 * maxStdDev) -- hours for 16-bit data)
//
// Super voxels: level 0 holds mean = normalised * 255 (for every storage type) and stddev 0; level l is
// the plain mean of the 8 child means and the deviation of those 8 means, all in fp64 like the reference.  Means are
// kept in an fp64 scratch level for the next reduction; what the marcher samples is the RG16F copy, padded by one
// replicated texel.  HBM-bound streaming kernels: level 1 reads b_v bytes per voxel and writes (8 + 4)/8 bytes.
// Pre-integration LUT: one thread per (density, stddev) entry, the Gaussian-weighted sum runs over all densities in
// the reference's order (fp64), so the result only differs through exp() rounding.
#include "vrb_internal.cuh"
#include <cstring>
#include <algorithm>
#include <cmath>
#include <vector>

__device__ __forceinline__ void atomic_max_nonneg_double(double* addr, double v) {
  // non-negative doubles order like their bit patterns
  atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

template <typename T>
__global__ void __launch_bounds__(256)
k_sv_level0(const T* __restrict__ raw, __half2* __restrict__ out, int w, int h, int d, double maxv) {
  const int pw = w + 2, ph = h + 2, pd = d + 2;
  const long long n = (long long)pw * ph * pd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int x = (int)(i % pw), y = (int)((i / pw) % ph), z = (int)(i / ((long long)pw * ph));
    int sx = min(max(x - 1, 0), w - 1), sy = min(max(y - 1, 0), h - 1), sz = min(max(z - 1, 0), d - 1);
    double v = (double)raw[(size_t)sx + (size_t)w * ((size_t)sy + (size_t)h * sz)];
    double mean = __dmul_rn(__ddiv_rn(v, maxv), 255.0);
    out[i] = __floats2half2_rn(__double2float_rn(mean), 0.0f);
  }
}

// level l (>= 1) from level l-1.  FROM_RAW: the children are raw voxels (level 0 is never materialised in fp64).
template <typename T, bool FROM_RAW>
__global__ void __launch_bounds__(256)
k_sv_reduce(const T* __restrict__ raw, const double* __restrict__ prev, double* __restrict__ mean_out, __half2* __restrict__ out,
            int pw_, int ph_, int cw, int ch, int cd, double maxv, double* max_stddev) {
  const long long n = (long long)cw * ch * cd;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double sd = 0.0;
  if (i < n) {
    const int ix = (int)(i % cw), iy = (int)((i / cw) % ch), iz = (int)(i / ((long long)cw * ch));
    double vm[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      // reference order vm0..vm7: (x, y, z+1 fastest) -- lw + (k>>2), lh + ((k>>1)&1), ld + (k&1)
      const size_t id = (size_t)(2 * ix + (k >> 2)) + (size_t)(2 * iy + ((k >> 1) & 1)) * pw_ + (size_t)(2 * iz + (k & 1)) * pw_ * ph_;
      vm[k] = FROM_RAW ? __dmul_rn(__ddiv_rn((double)raw[id], maxv), 255.0) : prev[id];
    }
    double vmn = (vm[0] + vm[1] + vm[2] + vm[3] + vm[4] + vm[5] + vm[6] + vm[7]) / 8.0;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { double t = vm[k] - vmn; acc = acc + t * t; }
    sd = sqrt(acc / 8.0);
    mean_out[i] = vmn;
    out[(size_t)(ix + 1) + (size_t)(cw + 2) * ((iy + 1) + (size_t)(ch + 2) * (iz + 1))] = __floats2half2_rn(__double2float_rn(vmn), __double2float_rn(sd));
  }
  // block max -> one atomic
  for (int o = 16; o > 0; o >>= 1) sd = fmax(sd, __shfl_xor_sync(0xffffffffu, sd, o));
  if ((threadIdx.x & 31) == 0 && sd > 0.0) atomic_max_nonneg_double(max_stddev, sd);
}

__global__ void __launch_bounds__(256) k_sv_pad(__half2* __restrict__ lev, int w, int h, int d) {
  const int pw = w + 2, ph = h + 2, pd = d + 2;
  const long long n = (long long)pw * ph * pd;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = (int)(i % pw), y = (int)((i / pw) % ph), z = (int)(i / ((long long)pw * ph));
  if (!(x == 0 || y == 0 || z == 0 || x == pw - 1 || y == ph - 1 || z == pd - 1)) return;
  int sx = min(max(x, 1), w), sy = min(max(y, 1), h), sz = min(max(z, 1), d);
  lev[i] = lev[(long long)sx + (long long)pw * (sy + (long long)ph * sz)];
}

// LUT[iw][ih] (x = density fastest), padded by one replicated texel; R16F
__global__ void __launch_bounds__(128)
k_preint(const float* __restrict__ opc, int dens_val, int w, int h, __half* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)w * h) return;
  const int iw = (int)(i % w), ih = (int)(i / w);
  const double mean = (double)iw, stddev = (double)ih;
  double SumG;
  if (fabs(stddev) > 0.0001) {
    const double nf = 1.0 / (stddev * sqrt(2.0 * 3.14159265358979323846264338327950288));
    const double den = 2.0 * stddev * stddev;
    double sg = 0.0, sw = 0.0;
    for (int k = 0; k < dens_val; ++k) {
      double dx = (double)k - mean;
      double W = nf * exp(-(dx * dx) / den);
      sg += W * (double)__ldg(opc + k);
      sw += W;
    }
    SumG = sg / sw;
  } else {
    SumG = (double)__ldg(opc + iw);
  }
  out[(size_t)(iw + 1) + (size_t)(w + 2) * (ih + 1)] = __float2half_rn(__double2float_rn(SumG));
}

__global__ void k_preint_pad(__half* __restrict__ t, int w, int h) {
  const int pw = w + 2, ph = h + 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pw * ph) return;
  const int x = i % pw, y = i / pw;
  if (!(x == 0 || y == 0 || x == pw - 1 || y == ph - 1)) return;
  t[i] = t[min(max(x, 1), w) + pw * min(max(y, 1), h)];
}

static void free_vct(vrb_ctx* c) {
  if (c->sv_tex) cudaDestroyTextureObject(c->sv_tex);
  if (c->sv_mip) cudaFreeMipmappedArray(c->sv_mip);
  if (c->preint_tex) cudaDestroyTextureObject(c->preint_tex);
  if (c->preint_array) cudaFreeArray(c->preint_array);
  c->sv_tex = 0; c->sv_mip = nullptr; c->preint_tex = 0; c->preint_array = nullptr;
  for (int l = 0; l < VRB_MAX_LEVELS; ++l) { if (c->d_sv[l]) cudaFree(c->d_sv[l]); c->d_sv[l] = nullptr; }
  c->sv_levels = 0;
  if (c->d_sv_top_means) cudaFree(c->d_sv_top_means);
  c->d_sv_top_means = nullptr;
  if (c->d_preint) cudaFree(c->d_preint);
  c->d_preint = nullptr; c->preint_w = c->preint_h = 0; c->sv_max_stddev = 0.0f;
}
void vrb_free_vct(vrb_ctx* c) { free_vct(c); }

// VRB_FILTER_HARDWARE: the super-voxel pyramid as the reference binds it (GL_RG16F, GL_LINEAR_MIPMAP_LINEAR,
// clamp-to-edge; preprocessingstages.cpp:35-137) and the pre-integration LUT (GL_R16F 2-D, GL_LINEAR; :139-202).
int vrb_sv_tex_prepare(vrb_ctx* c) {
  if (c->sv_tex && c->preint_tex) return VRB_OK;
  VRB_REQUIRE(c->sv_levels > 0 && c->d_preint, VRB_ERR_STATE, "no super-voxel pyramid / LUT");
  cudaChannelFormatDesc fd2 = cudaCreateChannelDescHalf2();
  cudaExtent ext0 = make_cudaExtent((size_t)c->sv_dims[0][0], (size_t)c->sv_dims[0][1], (size_t)c->sv_dims[0][2]);
  VRB_CUDA(cudaMallocMipmappedArray(&c->sv_mip, &fd2, ext0, (unsigned)c->sv_levels, cudaArrayDefault));
  for (int l = 0; l < c->sv_levels; ++l) {
    cudaArray_t lev;
    VRB_CUDA(cudaGetMipmappedArrayLevel(&lev, c->sv_mip, (unsigned)l));
    const size_t w = (size_t)c->sv_dims[l][0], h = (size_t)c->sv_dims[l][1], d = (size_t)c->sv_dims[l][2];
    const size_t pw = w + 2, ph = h + 2;
    cudaMemcpy3DParms cp; memset(&cp, 0, sizeof(cp));
    cp.srcPtr = make_cudaPitchedPtr((void*)(c->d_sv[l] + pw * ph + pw + 1), pw * sizeof(__half2), w, ph);
    cp.dstArray = lev;
    cp.extent = make_cudaExtent(w, h, d);
    cp.kind = cudaMemcpyDeviceToDevice;
    VRB_CUDA(cudaMemcpy3DAsync(&cp, c->stream));
  }
  cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
  rd.resType = cudaResourceTypeMipmappedArray; rd.res.mipmap.mipmap = c->sv_mip;
  cudaTextureDesc td; memset(&td, 0, sizeof(td));
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModeLinear; td.mipmapFilterMode = cudaFilterModeLinear;
  td.readMode = cudaReadModeElementType; td.normalizedCoords = 1;
  td.minMipmapLevelClamp = 0.0f; td.maxMipmapLevelClamp = (float)(c->sv_levels - 1);
  VRB_CUDA(cudaCreateTextureObject(&c->sv_tex, &rd, &td, nullptr));

  cudaChannelFormatDesc fd1 = cudaCreateChannelDescHalf();
  VRB_CUDA(cudaMallocArray(&c->preint_array, &fd1, (size_t)c->preint_w, (size_t)c->preint_h, cudaArrayDefault));
  const size_t lpw = (size_t)c->preint_w + 2;
  VRB_CUDA(cudaMemcpy2DToArrayAsync(c->preint_array, 0, 0, c->d_preint + lpw + 1, lpw * sizeof(__half),
                                    (size_t)c->preint_w * sizeof(__half), (size_t)c->preint_h, cudaMemcpyDeviceToDevice, c->stream));
  memset(&rd, 0, sizeof(rd));
  rd.resType = cudaResourceTypeArray; rd.res.array.array = c->preint_array;
  memset(&td, 0, sizeof(td));
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeElementType; td.normalizedCoords = 1;
  VRB_CUDA(cudaCreateTextureObject(&c->preint_tex, &rd, &td, nullptr));
  return VRB_OK;
}

// Levels 0..n-1 of the mean/stddev pyramid of the uploaded array (max_levels <= 0: down to 1 voxel).  With a brick the
// uploaded array is a window of a larger volume: level l of the window equals the WHOLE volume's level l on the window's
// texels as long as the window starts (and, away from the volume's end, stops) on multiples of 2^l, which is checked.
// The fp64 means of the last level stay in c->d_sv_top_means; *max_sd = largest deviation over the levels built.
static int sv_build_levels(vrb_ctx* c, int max_levels, const vrb_brick* brick, double* max_sd_out, const char* who) {
  VRB_REQUIRE(c->d_raw, VRB_ERR_STATE, "%s: no volume uploaded", who);
  VRB_CUDA(cudaSetDevice(c->device));
  free_vct(c);
  const int dens_val = c->bpv == 1 ? 255 : 65535;
  const double maxv = (double)dens_val;
  int w = c->vw, h = c->vh, d = c->vd;
  int nlev = 1;
  { int a = w / 2, b = h / 2, e = d / 2; while ((long long)a * b * e >= 1) { ++nlev; a /= 2; b /= 2; e /= 2; } }
  if (max_levels > 0) { VRB_REQUIRE(max_levels <= nlev, VRB_ERR_INVALID, "%s: %d levels asked, the array has %d", who, max_levels, nlev); nlev = max_levels; }
  VRB_REQUIRE(nlev <= VRB_MAX_LEVELS, VRB_ERR_INVALID, "%s: too many levels", who);
  int g[3] = {w, h, d}, off[3] = {0, 0, 0};
  if (brick) {
    for (int a = 0; a < 3; ++a) {
      g[a] = brick->global_dims[a]; off[a] = brick->origin[a] - brick->ghost_lo[a];
      const int dim = a == 0 ? w : (a == 1 ? h : d), end = off[a] + dim, al = 1 << (nlev - 1);
      VRB_REQUIRE(off[a] % al == 0 && (end == g[a] || end % al == 0), VRB_ERR_INVALID,
                  "%s: window [%d, %d) of axis %d is not aligned to 2^%d (levels of the window would not be levels of the volume)", who, off[a], end, a, nlev - 1);
    }
  }
  double* d_max = nullptr;
  VRB_CUDA(cudaMalloc(&d_max, sizeof(double)));
  VRB_CUDA(cudaMemsetAsync(d_max, 0, sizeof(double), c->stream));
  double *d_mean_prev = nullptr, *d_mean_cur = nullptr;
  int rc = VRB_OK;
  for (int l = 0; l < nlev && rc == VRB_OK; ++l) {
    const size_t np = (size_t)(w + 2) * (h + 2) * (d + 2);
    if (cudaMalloc(&c->d_sv[l], np * sizeof(__half2)) != cudaSuccess) { vrb_set_error("%s: cudaMalloc level %d failed", who, l); rc = VRB_ERR_CUDA; break; }
    c->sv_levels = l + 1;
    c->sv_dims[l][0] = w; c->sv_dims[l][1] = h; c->sv_dims[l][2] = d;
    for (int a = 0; a < 3; ++a) { c->sv_gdims[l][a] = g[a]; c->sv_off[l][a] = off[a]; }
    if (l == 0) {
      int blocks = (int)std::min<size_t>((np + 255) / 256, 148 * 32);
      if (c->bpv == 1) k_sv_level0<uint8_t><<<blocks, 256, 0, c->stream>>>((const uint8_t*)c->d_raw, c->d_sv[0], w, h, d, maxv);
      else             k_sv_level0<uint16_t><<<blocks, 256, 0, c->stream>>>((const uint16_t*)c->d_raw, c->d_sv[0], w, h, d, maxv);
      c->launches++;
    } else {
      const int pw = c->sv_dims[l - 1][0], ph = c->sv_dims[l - 1][1];
      const size_t n = (size_t)w * h * d;
      if (cudaMalloc(&d_mean_cur, n * sizeof(double)) != cudaSuccess) { vrb_set_error("%s: cudaMalloc scratch failed", who); rc = VRB_ERR_CUDA; break; }
      const unsigned blocks = (unsigned)((n + 255) / 256);
      if (l == 1) {
        if (c->bpv == 1) k_sv_reduce<uint8_t, true><<<blocks, 256, 0, c->stream>>>((const uint8_t*)c->d_raw, nullptr, d_mean_cur, c->d_sv[l], pw, ph, w, h, d, maxv, d_max);
        else             k_sv_reduce<uint16_t, true><<<blocks, 256, 0, c->stream>>>((const uint16_t*)c->d_raw, nullptr, d_mean_cur, c->d_sv[l], pw, ph, w, h, d, maxv, d_max);
      } else {
        k_sv_reduce<uint8_t, false><<<blocks, 256, 0, c->stream>>>(nullptr, d_mean_prev, d_mean_cur, c->d_sv[l], pw, ph, w, h, d, maxv, d_max);
      }
      k_sv_pad<<<(unsigned)((np + 255) / 256), 256, 0, c->stream>>>(c->d_sv[l], w, h, d);
      c->launches += 2;
      if (cudaStreamSynchronize(c->stream) != cudaSuccess) { vrb_set_error("%s: level %d failed", who, l); rc = VRB_ERR_CUDA; break; }
      if (d_mean_prev) cudaFree(d_mean_prev);
      d_mean_prev = d_mean_cur; d_mean_cur = nullptr;
    }
    w /= 2; h /= 2; d /= 2;
    for (int a = 0; a < 3; ++a) { g[a] /= 2; off[a] /= 2; }
  }
  if (d_mean_cur) cudaFree(d_mean_cur);
  c->d_sv_top_means = d_mean_prev;                       // nullptr when only level 0 was built
  double max_sd = 0.0;
  if (rc == VRB_OK && (cudaMemcpyAsync(&max_sd, d_max, sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
                       cudaStreamSynchronize(c->stream) != cudaSuccess)) { vrb_set_error("%s: max stddev read-back failed", who); rc = VRB_ERR_CUDA; }
  cudaFree(d_max);
  if (rc != VRB_OK) return rc;
  *max_sd_out = max_sd;
  return VRB_OK;
}

// pre-integration table: w = ceil(maxDensity), h = ceil(maxStdDev) (preprocessingstages.cpp:160-164)
static int preint_build(vrb_ctx* c, const float* opc_by_density, int n_opc, double max_sd, const char* who) {
  const int dens_val = c->bpv == 1 ? 255 : 65535;
  VRB_REQUIRE(n_opc == dens_val + 1, VRB_ERR_INVALID, "%s: need %d opacity entries (GetOpc(i, maxDensity), i = 0..maxDensity), got %d", who, dens_val + 1, n_opc);
  VRB_REQUIRE(c->sv_levels > 0, VRB_ERR_STATE, "%s: no super-voxel pyramid", who);
  VRB_CUDA(cudaSetDevice(c->device));
  if (c->preint_tex) { cudaDestroyTextureObject(c->preint_tex); c->preint_tex = 0; }
  if (c->preint_array) { cudaFreeArray(c->preint_array); c->preint_array = nullptr; }
  if (c->sv_tex) { cudaDestroyTextureObject(c->sv_tex); c->sv_tex = 0; }     // vrb_sv_tex_prepare rebuilds both together
  if (c->sv_mip) { cudaFreeMipmappedArray(c->sv_mip); c->sv_mip = nullptr; }
  if (c->d_preint) { cudaFree(c->d_preint); c->d_preint = nullptr; }
  c->sv_max_stddev = (float)max_sd;
  const int lw = dens_val, lh = (int)std::ceil(max_sd);
  VRB_REQUIRE(lh >= 1, VRB_ERR_UNSUPPORTED, "%s: homogeneous volume (max stddev 0): the reference would create an empty look-up texture", who);
  float* d_opc = nullptr;
  VRB_CUDA(cudaMalloc(&d_opc, (size_t)n_opc * sizeof(float)));
  VRB_CUDA(cudaMemcpyAsync(d_opc, opc_by_density, (size_t)n_opc * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  VRB_CUDA(cudaMalloc(&c->d_preint, (size_t)(lw + 2) * (lh + 2) * sizeof(__half)));
  const long long ne = (long long)lw * lh;
  k_preint<<<(unsigned)((ne + 127) / 128), 128, 0, c->stream>>>(d_opc, dens_val, lw, lh, c->d_preint);
  k_preint_pad<<<((lw + 2) * (lh + 2) + 255) / 256, 256, 0, c->stream>>>(c->d_preint, lw, lh);
  c->launches += 2;
  VRB_CUDA(cudaGetLastError());
  VRB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_opc);
  c->preint_w = lw; c->preint_h = lh;
  return VRB_OK;
}

extern "C" int vrb_vct_build(vrb_ctx* c, const float* opc_by_density, int n_opc) {
  VRB_REQUIRE(c && opc_by_density, VRB_ERR_INVALID, "vrb_vct_build: NULL argument");
  VRB_REQUIRE(c->d_raw, VRB_ERR_STATE, "vrb_vct_build: no volume uploaded");
  VRB_REQUIRE(n_opc == (c->bpv == 1 ? 256 : 65536), VRB_ERR_INVALID, "vrb_vct_build: need %d opacity entries (GetOpc(i, maxDensity), i = 0..maxDensity), got %d",
              c->bpv == 1 ? 256 : 65536, n_opc);
  double max_sd = 0.0;
  int rc = sv_build_levels(c, 0, nullptr, &max_sd, "vrb_vct_build");
  if (rc != VRB_OK) return rc;
  return preint_build(c, opc_by_density, n_opc, max_sd, "vrb_vct_build");
}

// ---- sort-last bricks: the pyramid of a window of the volume --------------------------
extern "C" int vrb_sv_build_brick(vrb_ctx* c, const vrb_brick* brick, int n_levels, double* local_max_stddev) {
  VRB_REQUIRE(c && brick && local_max_stddev, VRB_ERR_INVALID, "vrb_sv_build_brick: NULL argument");
  VRB_REQUIRE(n_levels >= 2, VRB_ERR_INVALID, "vrb_sv_build_brick: n_levels %d (at least 2)", n_levels);
  VRB_REQUIRE(c->d_raw, VRB_ERR_STATE, "vrb_sv_build_brick: no volume uploaded");
  { int rc = vrb_brick_check(c, brick, "vrb_sv_build_brick"); if (rc != VRB_OK) return rc; }
  return sv_build_levels(c, n_levels, brick, local_max_stddev, "vrb_sv_build_brick");
}

// fp64 means of the OWNED texels of the last level built (x fastest); dims_out = their extent
extern "C" int vrb_sv_top_means_read(vrb_ctx* c, const vrb_brick* brick, double* host_out, size_t cap_doubles, int dims_out[3], int origin_out[3]) {
  VRB_REQUIRE(c && brick && dims_out && origin_out, VRB_ERR_INVALID, "vrb_sv_top_means_read: NULL argument");
  VRB_REQUIRE(c->sv_levels >= 2 && c->d_sv_top_means, VRB_ERR_STATE, "vrb_sv_top_means_read: no brick pyramid (vrb_sv_build_brick)");
  const int T = c->sv_levels - 1, al = 1 << T;
  int lo[3], n[3];
  for (int a = 0; a < 3; ++a) {
    const int o = brick->origin[a], e = brick->origin[a] + brick->owned[a];
    VRB_REQUIRE(o % al == 0 && (e == brick->global_dims[a] || e % al == 0), VRB_ERR_INVALID,
                "vrb_sv_top_means_read: owned range [%d, %d) of axis %d is not aligned to 2^%d", o, e, a, T);
    const int glo = o >> T, ghi = (e == brick->global_dims[a]) ? c->sv_gdims[T][a] : (e >> T);
    lo[a] = glo - c->sv_off[T][a]; n[a] = ghi - glo;
    VRB_REQUIRE(lo[a] >= 0 && n[a] >= 0 && lo[a] + n[a] <= c->sv_dims[T][a], VRB_ERR_STATE, "vrb_sv_top_means_read: owned texels outside the window (axis %d)", a);
    dims_out[a] = n[a]; origin_out[a] = glo;
  }
  const size_t total = (size_t)n[0] * n[1] * n[2];
  if (!host_out) return VRB_OK;                           // size query
  VRB_REQUIRE(cap_doubles >= total, VRB_ERR_INVALID, "vrb_sv_top_means_read: buffer holds %zu doubles, need %zu", cap_doubles, total);
  if (total == 0) return VRB_OK;
  VRB_CUDA(cudaSetDevice(c->device));
  cudaMemcpy3DParms cp; memset(&cp, 0, sizeof(cp));
  const size_t w = (size_t)c->sv_dims[T][0], h = (size_t)c->sv_dims[T][1];
  cp.srcPtr = make_cudaPitchedPtr(c->d_sv_top_means, w * sizeof(double), w, h);
  cp.srcPos = make_cudaPos((size_t)lo[0] * sizeof(double), (size_t)lo[1], (size_t)lo[2]);
  cp.dstPtr = make_cudaPitchedPtr(host_out, (size_t)n[0] * sizeof(double), (size_t)n[0], (size_t)n[1]);
  cp.extent = make_cudaExtent((size_t)n[0] * sizeof(double), (size_t)n[1], (size_t)n[2]);
  cp.kind = cudaMemcpyDeviceToHost;
  VRB_CUDA(cudaMemcpy3DAsync(&cp, c->stream));
  VRB_CUDA(cudaStreamSynchronize(c->stream));
  return VRB_OK;
}

// the levels ABOVE the bricks' last level: means of one whole level (gathered from all bricks) -> largest deviation of
// every coarser level, with the kernel the single-GPU build uses (same fp64 operation order)
extern "C" int vrb_sv_reduce_top(vrb_ctx* c, const double* level_means, int w, int h, int d, double* max_stddev) {
  VRB_REQUIRE(c && level_means && max_stddev, VRB_ERR_INVALID, "vrb_sv_reduce_top: NULL argument");
  VRB_REQUIRE(w >= 1 && h >= 1 && d >= 1, VRB_ERR_INVALID, "vrb_sv_reduce_top: bad dims %dx%dx%d", w, h, d);
  VRB_CUDA(cudaSetDevice(c->device));
  double *prev = nullptr, *cur = nullptr, *d_max = nullptr; __half2* scratch = nullptr;
  int rc = VRB_OK;
  const size_t n0 = (size_t)w * h * d;
  if (cudaMalloc(&prev, n0 * sizeof(double)) != cudaSuccess || cudaMalloc(&d_max, sizeof(double)) != cudaSuccess ||
      cudaMalloc(&scratch, (size_t)(w / 2 + 2) * (h / 2 + 2) * (d / 2 + 2) * sizeof(__half2)) != cudaSuccess) { vrb_set_error("vrb_sv_reduce_top: cudaMalloc failed"); rc = VRB_ERR_CUDA; }
  if (rc == VRB_OK) {
    cudaMemcpyAsync(prev, level_means, n0 * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    cudaMemsetAsync(d_max, 0, sizeof(double), c->stream);
    int pw = w, ph = h, cw = w / 2, ch = h / 2, cd = d / 2;
    while ((long long)cw * ch * cd >= 1) {
      const size_t n = (size_t)cw * ch * cd;
      if (cudaMalloc(&cur, n * sizeof(double)) != cudaSuccess) { vrb_set_error("vrb_sv_reduce_top: cudaMalloc failed"); rc = VRB_ERR_CUDA; break; }
      k_sv_reduce<uint8_t, false><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(nullptr, prev, cur, scratch, pw, ph, cw, ch, cd, 255.0, d_max);
      c->launches++;
      if (cudaStreamSynchronize(c->stream) != cudaSuccess) { vrb_set_error("vrb_sv_reduce_top: reduce failed"); rc = VRB_ERR_CUDA; break; }
      cudaFree(prev); prev = cur; cur = nullptr;
      pw = cw; ph = ch; cw /= 2; ch /= 2; cd /= 2;
    }
  }
  double m = 0.0;
  if (rc == VRB_OK && (cudaMemcpyAsync(&m, d_max, sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
                       cudaStreamSynchronize(c->stream) != cudaSuccess)) { vrb_set_error("vrb_sv_reduce_top: read-back failed"); rc = VRB_ERR_CUDA; }
  if (prev) cudaFree(prev); if (cur) cudaFree(cur); if (d_max) cudaFree(d_max); if (scratch) cudaFree(scratch);
  if (rc == VRB_OK) *max_stddev = m;
  return rc;
}

// the LUT for a deviation range agreed between the bricks (max over the local levels of every brick and the top levels)
extern "C" int vrb_preint_build(vrb_ctx* c, const float* opc_by_density, int n_opc, double max_stddev) {
  VRB_REQUIRE(c && opc_by_density, VRB_ERR_INVALID, "vrb_preint_build: NULL argument");
  VRB_REQUIRE(max_stddev >= 0.0 && max_stddev < 1e6, VRB_ERR_INVALID, "vrb_preint_build: max_stddev %g", max_stddev);
  return preint_build(c, opc_by_density, n_opc, max_stddev, "vrb_preint_build");
}

extern "C" int vrb_vct_info(vrb_ctx* c, int* n_levels, int* dims_xyz, int cap_levels, int* lut_w, int* lut_h, float* max_stddev) {
  VRB_REQUIRE(c, VRB_ERR_INVALID, "vrb_vct_info: ctx is NULL");
  if (n_levels) *n_levels = c->sv_levels;
  if (dims_xyz)
    for (int l = 0; l < c->sv_levels && l < cap_levels; ++l)
      for (int k = 0; k < 3; ++k) dims_xyz[3 * l + k] = c->sv_dims[l][k];
  if (lut_w) *lut_w = c->preint_w;
  if (lut_h) *lut_h = c->preint_h;
  if (max_stddev) *max_stddev = c->sv_max_stddev;
  return VRB_OK;
}

__global__ void k_sv_unpad(const __half2* __restrict__ lev, float* __restrict__ out, int w, int h, int d) {
  const long long n = (long long)w * h * d;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = (int)(i % w), y = (int)((i / w) % h), z = (int)(i / ((long long)w * h));
  float2 v = __half22float2(lev[(long long)(x + 1) + (long long)(w + 2) * ((y + 1) + (long long)(h + 2) * (z + 1))]);
  out[2 * i] = v.x; out[2 * i + 1] = v.y;
}
__global__ void k_lut_unpad(const __half* __restrict__ t, float* __restrict__ out, int w, int h) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * h) return;
  out[i] = __half2float(t[(i % w + 1) + (w + 2) * (i / w + 1)]);
}

// level >= 0: super-voxel level as w*h*d x (mean, stddev) floats; level == -1: the LUT as w*h floats
extern "C" int vrb_vct_read(vrb_ctx* c, int level, float* host_out) {
  VRB_REQUIRE(c && host_out, VRB_ERR_INVALID, "vrb_vct_read: NULL argument");
  VRB_REQUIRE(level >= -1 && level < c->sv_levels, VRB_ERR_INVALID, "vrb_vct_read: level %d of %d", level, c->sv_levels);
  VRB_REQUIRE(level >= 0 || c->d_preint, VRB_ERR_STATE, "vrb_vct_read: no LUT");
  VRB_CUDA(cudaSetDevice(c->device));
  size_t n;
  float* tmp = nullptr;
  if (level >= 0) {
    const int w = c->sv_dims[level][0], h = c->sv_dims[level][1], d = c->sv_dims[level][2];
    n = (size_t)w * h * d * 2;
    VRB_CUDA(cudaMalloc(&tmp, n * sizeof(float)));
    k_sv_unpad<<<(unsigned)((n / 2 + 255) / 256), 256, 0, c->stream>>>(c->d_sv[level], tmp, w, h, d);
  } else {
    n = (size_t)c->preint_w * c->preint_h;
    VRB_CUDA(cudaMalloc(&tmp, n * sizeof(float)));
    k_lut_unpad<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_preint, tmp, c->preint_w, c->preint_h);
  }
  c->launches++;
  cudaError_t e = cudaMemcpyAsync(host_out, tmp, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
  cudaError_t e2 = cudaStreamSynchronize(c->stream);
  cudaFree(tmp);
  VRB_REQUIRE(e == cudaSuccess && e2 == cudaSuccess, VRB_ERR_CUDA, "vrb_vct_read: copy failed");
  return VRB_OK;
}
