// extcoef_pyramid.cu -- Gaussian extinction-coefficient mip pyramid of the directional-occlusion renderer, sm_100a.
// Replaces ExtinctionCoefficientVolume::BuildMipMappedTexture (rc1pdosct/extcoefvolumegenerator.cpp:20-40,92-408) and
// its GLSL passes glslextgen/gen_extcoefvol_{samesize,anysize}.comp (level 0: 7^3 Gaussian of TF opacity at trilinear
// volume samples), gen_extcoefvol_*_mmlevel.comp (level i from level i-1, sigma doubled) and backtotau.comp
// (opacity -> extinction on every level, only after all levels exist).
//
// Layout: every level is an fp16 array (the reference's GL_R16F storage, rounding included) padded by one replicated
// texel on each side, so the marcher's trilinear fetch needs no index clamping.  The whole 128^3 default pyramid is
// 4.8 MB: L2-resident.  The 343 Gaussian weights only depend on sigma and are evaluated once on the host with the
// same fp32 expression the shader uses; out-of-volume taps add 0 to the numerator but still count in sum(w).
#include "vrb_internal.cuh"
#include <cstring>
#include <cmath>
#include <vector>

// the 7x7x7 Gaussian weights of one level: a launch parameter (constant bank of THAT launch), so that two contexts building
// pyramids on one device at the same time cannot overwrite each other's weights
struct GaussW { float w[343]; };

// trilinear fetch at normalised coordinates s in [0,1]^3 from a padded fp16 level (GL_LINEAR, CLAMP_TO_EDGE)
__device__ __forceinline__ float level_tex3d(const LevelView& L, float sx, float sy, float sz) {
  float ux = fmaf(sx, (float)L.w, 0.5f), uy = fmaf(sy, (float)L.h, 0.5f), uz = fmaf(sz, (float)L.d, 0.5f);
  ux = fminf(fmaxf(ux, 0.0f), (float)L.w + 0.999f);
  uy = fminf(fmaxf(uy, 0.0f), (float)L.h + 0.999f);
  uz = fminf(fmaxf(uz, 0.0f), (float)L.d + 0.999f);
  float flx, fly, flz;
  int ix = vrb_floor_pos(ux, &flx), iy = vrb_floor_pos(uy, &fly), iz = vrb_floor_pos(uz, &flz);
  float fx = ux - flx, fy = uy - fly, fz = uz - flz;
  const int pw = L.w + 2;
  const long long slice = (long long)pw * (L.h + 2);
  const __half* p = L.tex + ((long long)iz * slice + (long long)iy * pw + ix);
  const __half* q = p + slice;
  float c00 = vrb_lerp(__half2float(__ldg(p)), __half2float(__ldg(p + 1)), fx);
  float c10 = vrb_lerp(__half2float(__ldg(p + pw)), __half2float(__ldg(p + pw + 1)), fx);
  float c01 = vrb_lerp(__half2float(__ldg(q)), __half2float(__ldg(q + 1)), fx);
  float c11 = vrb_lerp(__half2float(__ldg(q + pw)), __half2float(__ldg(q + pw + 1)), fx);
  return vrb_lerp(vrb_lerp(c00, c10, fy), vrb_lerp(c01, c11, fy), fz);
}

// LEVEL0: source = TF opacity of the volume texture; otherwise source = previous pyramid level (opacity).
template <bool LEVEL0>
__global__ void __launch_bounds__(256)
k_extcoef_level(LevelView src, const float4* __restrict__ tf_rgba, int tf_n, __half* __restrict__ dst, int w, int h, int d,
                float gx, float gy, float gz, float sigma, float vsx, float vsy, float vsz, const __grid_constant__ GaussW gw) {
  extern __shared__ float s_opacity[];     // LEVEL0: padded TF opacity table (tf_n + 2 entries)
  if (LEVEL0) {
    for (int i = threadIdx.x; i < tf_n + 2; i += blockDim.x) s_opacity[i] = tf_rgba[i].w;
    __syncthreads();
  }
  const long long n = (long long)w * h * d;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = (int)(i % w), y = (int)((i / w) % h), z = (int)(i / ((long long)w * h));
  // voxel size of this level: grid size / resolution (ExtCoefVoxelSize, extcoefvolumegenerator.cpp:233; LevelVoxelSize in the
  // level shaders), except for the base level of the same-size build, whose shader uses the volume's own VoxelSize uniform
  // (gen_extcoefvol_samesize.comp:45): passed in vs* > 0
  const float vx = vsx > 0.0f ? vsx : gx / (float)w, vy = vsy > 0.0f ? vsy : gy / (float)h, vz = vsz > 0.0f ? vsz : gz / (float)d;
  const float px = ((float)x + 0.5f) * vx, py = ((float)y + 0.5f) * vy, pz = ((float)z + 0.5f) * vz;
  float sum_wkck = 0.0f, sum_wk = 0.0f;
  int t = 0;
  for (int a = -3; a < 4; ++a) {
    const float sx = (px + (float)a * sigma) / gx;
    const bool ox = sx < 0.0f || sx > 1.0f;
    for (int b = -3; b < 4; ++b) {
      const float sy = (py + (float)b * sigma) / gy;
      const bool oy = ox || sy < 0.0f || sy > 1.0f;
#pragma unroll
      for (int c = -3; c < 4; ++c, ++t) {
        const float sz = (pz + (float)c * sigma) / gz;
        const float wk = gw.w[t];
        float ck = 0.0f;
        if (!(oy || sz < 0.0f || sz > 1.0f)) {
          float v = level_tex3d(src, sx, sy, sz);
          if (LEVEL0) {
            float up = fmaf(v, (float)tf_n, 0.5f);
            up = fminf(fmaxf(up, 0.0f), (float)tf_n + 0.5f);
            float fl; int k = vrb_floor_pos(up, &fl);
            ck = vrb_lerp(s_opacity[k], s_opacity[k + 1], up - fl);
          } else {
            ck = v;
          }
        }
        sum_wkck += wk * ck;
        sum_wk += wk;
      }
    }
  }
  const int pw = w + 2, ph = h + 2;
  dst[(long long)(x + 1) + (long long)pw * ((y + 1) + (long long)ph * (z + 1))] = __float2half_rn(sum_wkck / sum_wk);
}

// replicate the outermost texels into the one-texel border (clamp-to-edge), optionally applying backtotau first
__global__ void __launch_bounds__(256)
k_extcoef_finish(__half* __restrict__ lev, int w, int h, int d, int to_tau, int pad_only) {
  const int pw = w + 2, ph = h + 2, pd = d + 2;
  const long long n = (long long)pw * ph * pd;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = (int)(i % pw), y = (int)((i / pw) % ph), z = (int)(i / ((long long)pw * ph));
  const bool border = x == 0 || y == 0 || z == 0 || x == pw - 1 || y == ph - 1 || z == pd - 1;
  if (pad_only) {
    if (!border) return;
    int sx = min(max(x, 1), w), sy = min(max(y, 1), h), sz = min(max(z, 1), d);
    lev[i] = lev[(long long)sx + (long long)pw * (sy + (long long)ph * sz)];
    return;
  }
  if (to_tau && !border) {
    float op = __half2float(lev[i]);
    lev[i] = __float2half_rn(-1.0f * logf(1.0f - op));      // backtotau.comp:22-34
  }
}

static void free_pyramid(vrb_ctx* c) {
  if (c->pyr_tex) cudaDestroyTextureObject(c->pyr_tex);
  if (c->pyr_mip) cudaFreeMipmappedArray(c->pyr_mip);
  c->pyr_tex = 0; c->pyr_mip = nullptr;
  for (int l = 0; l < c->pyr_levels; ++l) if (c->d_pyr[l]) cudaFree(c->d_pyr[l]);
  for (int l = 0; l < VRB_MAX_LEVELS; ++l) {
    if (c->d_pyr_quad[l]) cudaFree(c->d_pyr_quad[l]);
    c->d_pyr[l] = nullptr; c->d_pyr_quad[l] = nullptr;
  }
  c->pyr_quad_valid = false;
  c->pyr_levels = 0;
}
void vrb_free_pyramid(vrb_ctx* c) { free_pyramid(c); }

static void gauss_weights(float sigma, float* w) {
  int t = 0;
  for (int a = -3; a < 4; ++a)
    for (int b = -3; b < 4; ++b)
      for (int cc = -3; cc < 4; ++cc, ++t) {
        float fx = float(a) * sigma, fy = float(b) * sigma, fz = float(cc) * sigma;
        w[t] = (sigma * sigma * sigma) * expf(-(fx * fx + fy * fy + fz * fz) / (2.0f * sigma * sigma));
      }
}

extern "C" int vrb_extcoef_build(vrb_ctx* c, float sigma0, int rw, int rh, int rd) {
  VRB_REQUIRE(c, VRB_ERR_INVALID, "vrb_extcoef_build: ctx is NULL");
  VRB_REQUIRE(c->d_vol, VRB_ERR_STATE, "vrb_extcoef_build: no volume uploaded");
  VRB_REQUIRE(c->d_tf_rgba, VRB_ERR_STATE, "vrb_extcoef_build: no RGBA (opacity) transfer function uploaded");
  VRB_REQUIRE(sigma0 > 0.0f, VRB_ERR_INVALID, "vrb_extcoef_build: sigma0 %g", sigma0);
  // the level-0 kernel stages the opacity table in shared memory (48 KB without opt-in)
  VRB_REQUIRE((size_t)(c->tf_n + 2) * sizeof(float) <= 48 * 1024, VRB_ERR_UNSUPPORTED,
              "vrb_extcoef_build: transfer functions above 12286 texels are not supported (got %d)", c->tf_n);
  const bool same_size = rw <= 0 || rh <= 0 || rd <= 0;                          // GenerateExtinctionCoefficientVolumeSameSize (:92-228)
  if (same_size) { rw = c->vw; rh = c->vh; rd = c->vd; }
  VRB_REQUIRE(rw <= 4096 && rh <= 4096 && rd <= 4096, VRB_ERR_INVALID, "vrb_extcoef_build: bad resolution");
  VRB_CUDA(cudaSetDevice(c->device));
  free_pyramid(c);
  int nlev = 1;
  for (int m = std::max(rw, std::max(rh, rd)); m > 1; m >>= 1) ++nlev;
  VRB_REQUIRE(nlev <= VRB_MAX_LEVELS, VRB_ERR_INVALID, "vrb_extcoef_build: too many levels");
  const float gx = (float)c->vw * c->scale[0], gy = (float)c->vh * c->scale[1], gz = (float)c->vd * c->scale[2];
  LevelView vol; vol.tex = c->d_vol; vol.w = c->vw; vol.h = c->vh; vol.d = c->vd;
  GaussW hw;
  for (int l = 0; l < nlev; ++l) {
    const int w = std::max(1, rw >> l), h = std::max(1, rh >> l), d = std::max(1, rd >> l);
    const size_t np = (size_t)(w + 2) * (h + 2) * (d + 2);
    VRB_CUDA(cudaMalloc(&c->d_pyr[l], np * sizeof(__half)));
    c->pyr_levels = l + 1;
    c->pyr_dims[l][0] = w; c->pyr_dims[l][1] = h; c->pyr_dims[l][2] = d;
    VRB_CUDA(cudaMemsetAsync(c->d_pyr[l], 0, np * sizeof(__half), c->stream));
    const float sigma = l == 0 ? sigma0 : sigma0 * powf(2.0f, (float)l);       // Si (extcoefvolumegenerator.cpp:207)
    gauss_weights(sigma, hw.w);
    const long long n = (long long)w * h * d;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (l == 0) {
      k_extcoef_level<true><<<blocks, 256, (size_t)(c->tf_n + 2) * sizeof(float), c->stream>>>(
          vol, c->d_tf_rgba, c->tf_n, c->d_pyr[0], w, h, d, gx, gy, gz, sigma,
          same_size ? c->scale[0] : 0.0f, same_size ? c->scale[1] : 0.0f, same_size ? c->scale[2] : 0.0f, hw);
    } else {
      LevelView prev; prev.tex = c->d_pyr[l - 1]; prev.w = c->pyr_dims[l - 1][0]; prev.h = c->pyr_dims[l - 1][1]; prev.d = c->pyr_dims[l - 1][2];
      k_extcoef_level<false><<<blocks, 256, 0, c->stream>>>(prev, nullptr, 0, c->d_pyr[l], w, h, d, gx, gy, gz, sigma, 0.0f, 0.0f, 0.0f, hw);
    }
    VRB_CUDA(cudaGetLastError());
    k_extcoef_finish<<<(unsigned)((np + 255) / 256), 256, 0, c->stream>>>(c->d_pyr[l], w, h, d, 0, 1);   // border for the next level's fetches
    VRB_CUDA(cudaGetLastError());
    c->launches += 2;
  }
  // opacity -> extinction on every level, then refresh the replicated border
  for (int l = 0; l < nlev; ++l) {
    const int w = c->pyr_dims[l][0], h = c->pyr_dims[l][1], d = c->pyr_dims[l][2];
    const size_t np = (size_t)(w + 2) * (h + 2) * (d + 2);
    k_extcoef_finish<<<(unsigned)((np + 255) / 256), 256, 0, c->stream>>>(c->d_pyr[l], w, h, d, 1, 0);
    k_extcoef_finish<<<(unsigned)((np + 255) / 256), 256, 0, c->stream>>>(c->d_pyr[l], w, h, d, 0, 1);
    VRB_CUDA(cudaGetLastError());
    c->launches += 2;
  }
  VRB_CUDA(cudaStreamSynchronize(c->stream));
  return VRB_OK;
}

// VRB_FILTER_HARDWARE: the pyramid as the reference binds it, GL_R16F with GL_LINEAR_MIPMAP_LINEAR and clamp-to-edge
// (extcoefvolumegenerator.cpp:101-116); sampled with tex3DLod at normalised coordinates.
int vrb_pyr_tex_prepare(vrb_ctx* c) {
  if (c->pyr_tex) return VRB_OK;
  VRB_REQUIRE(c->pyr_levels > 0, VRB_ERR_STATE, "no extinction pyramid");
  cudaChannelFormatDesc fd = cudaCreateChannelDescHalf();
  cudaExtent ext0 = make_cudaExtent((size_t)c->pyr_dims[0][0], (size_t)c->pyr_dims[0][1], (size_t)c->pyr_dims[0][2]);
  VRB_CUDA(cudaMallocMipmappedArray(&c->pyr_mip, &fd, ext0, (unsigned)c->pyr_levels, cudaArrayDefault));
  for (int l = 0; l < c->pyr_levels; ++l) {
    cudaArray_t lev;
    VRB_CUDA(cudaGetMipmappedArrayLevel(&lev, c->pyr_mip, (unsigned)l));
    const size_t w = (size_t)c->pyr_dims[l][0], h = (size_t)c->pyr_dims[l][1], d = (size_t)c->pyr_dims[l][2];
    const size_t pw = w + 2, ph = h + 2;
    cudaMemcpy3DParms cp; memset(&cp, 0, sizeof(cp));
    cp.srcPtr = make_cudaPitchedPtr((void*)(c->d_pyr[l] + pw * ph + pw + 1), pw * sizeof(__half), w, ph);
    cp.dstArray = lev;
    cp.extent = make_cudaExtent(w, h, d);
    cp.kind = cudaMemcpyDeviceToDevice;
    VRB_CUDA(cudaMemcpy3DAsync(&cp, c->stream));
  }
  cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
  rd.resType = cudaResourceTypeMipmappedArray; rd.res.mipmap.mipmap = c->pyr_mip;
  cudaTextureDesc td; memset(&td, 0, sizeof(td));
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModeLinear; td.mipmapFilterMode = cudaFilterModeLinear;
  td.readMode = cudaReadModeElementType; td.normalizedCoords = 1;
  td.minMipmapLevelClamp = 0.0f; td.maxMipmapLevelClamp = (float)(c->pyr_levels - 1);
  VRB_CUDA(cudaCreateTextureObject(&c->pyr_tex, &rd, &td, nullptr));
  return VRB_OK;
}

// Quad copy of the levels for k_dos_compact: entry (x,y,z) of the padded grid holds the padded texels (x,y), (x+1,y),
// (x,y+1), (x+1,y+1) of slice z as four halves, so a trilinear footprint is two 8-byte loads (slices z and z+1).
__global__ void __launch_bounds__(256)
k_pyr_quads(const __half* __restrict__ lev, uint2* __restrict__ quad, int pw, int ph, int pd) {
  const long long n = (long long)pw * ph * pd;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = (int)(i % pw), y = (int)((i / pw) % ph);
  const int x1 = min(x + 1, pw - 1);
  const long long row1 = (y + 1 < ph) ? (long long)pw : 0;
  const unsigned a = __half_as_ushort(lev[i]), b = __half_as_ushort(lev[i - x + x1]);
  const unsigned cc = __half_as_ushort(lev[i + row1]), d = __half_as_ushort(lev[i - x + x1 + row1]);
  quad[i] = make_uint2(a | (b << 16), cc | (d << 16));
}

void vrb_build_quads(const __half* lev, uint2* quad, int pw, int ph, int pd, cudaStream_t stream) {
  const size_t np = (size_t)pw * ph * pd;
  k_pyr_quads<<<(unsigned)((np + 255) / 256), 256, 0, stream>>>(lev, quad, pw, ph, pd);
}

int vrb_pyr_quads_prepare(vrb_ctx* c) {
  if (c->pyr_quad_valid) return VRB_OK;
  VRB_REQUIRE(c->pyr_levels > 0, VRB_ERR_STATE, "no extinction pyramid");
  for (int l = 0; l < c->pyr_levels; ++l) {
    const int pw = c->pyr_dims[l][0] + 2, ph = c->pyr_dims[l][1] + 2, pd = c->pyr_dims[l][2] + 2;
    const size_t np = (size_t)pw * ph * pd;
    if (!c->d_pyr_quad[l]) VRB_CUDA(cudaMalloc(&c->d_pyr_quad[l], np * sizeof(uint2)));
    k_pyr_quads<<<(unsigned)((np + 255) / 256), 256, 0, c->stream>>>(c->d_pyr[l], c->d_pyr_quad[l], pw, ph, pd);
    VRB_CUDA(cudaGetLastError());
    c->launches++;
  }
  c->pyr_quad_valid = true;
  return VRB_OK;
}

extern "C" int vrb_extcoef_info(vrb_ctx* c, int* n_levels, int* dims, int cap_levels) {
  VRB_REQUIRE(c && n_levels, VRB_ERR_INVALID, "vrb_extcoef_info: NULL argument");
  *n_levels = c->pyr_levels;
  if (dims)
    for (int l = 0; l < c->pyr_levels && l < cap_levels; ++l)
      for (int k = 0; k < 3; ++k) dims[3 * l + k] = c->pyr_dims[l][k];
  return VRB_OK;
}

__global__ void k_level_unpad_f32(const __half* __restrict__ lev, float* __restrict__ out, int w, int h, int d) {
  const long long n = (long long)w * h * d;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = (int)(i % w), y = (int)((i / w) % h), z = (int)(i / ((long long)w * h));
  out[i] = __half2float(lev[(long long)(x + 1) + (long long)(w + 2) * ((y + 1) + (long long)(h + 2) * (z + 1))]);
}

extern "C" int vrb_extcoef_read_level(vrb_ctx* c, int level, float* host_out) {
  VRB_REQUIRE(c && host_out, VRB_ERR_INVALID, "vrb_extcoef_read_level: NULL argument");
  VRB_REQUIRE(level >= 0 && level < c->pyr_levels, VRB_ERR_INVALID, "vrb_extcoef_read_level: level %d of %d", level, c->pyr_levels);
  VRB_CUDA(cudaSetDevice(c->device));
  const int w = c->pyr_dims[level][0], h = c->pyr_dims[level][1], d = c->pyr_dims[level][2];
  const size_t n = (size_t)w * h * d;
  float* tmp = nullptr;
  VRB_CUDA(cudaMalloc(&tmp, n * sizeof(float)));
  k_level_unpad_f32<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_pyr[level], tmp, w, h, d);
  c->launches++;
  cudaError_t e = cudaMemcpyAsync(host_out, tmp, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
  cudaError_t e2 = cudaStreamSynchronize(c->stream);
  cudaFree(tmp);
  VRB_REQUIRE(e == cudaSuccess && e2 == cudaSuccess, VRB_ERR_CUDA, "vrb_extcoef_read_level: copy failed");
  return VRB_OK;
}
