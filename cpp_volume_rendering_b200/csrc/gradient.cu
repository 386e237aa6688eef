// gradient.cu -- the gradient texture of DataManager::GenerateStructuredGradientTexture (libs/volvis_utils/
// datamanager.cpp:332-352) built on the GPU; SURVEY.md section 8f row 1.  Three generators, as in the reference:
//   VRB_GRADIENT_SOBEL_FELDMAN          vis::GenerateSobelFeldmanGradientTexture (libs/volvis_utils/utils.cpp:287-350):
//                                       3x3x3 Sobel-Feldman on GetNormalizedSample (0 outside), double accumulation
//   VRB_GRADIENT_FINITE_DIFFERENCES     vis::GenerateGradientTexture with its default arguments (utils.cpp:146-284):
//                                       central differences at distance 1, normalised in double, NaN -> 0, no filtering
//   VRB_GRADIENT_COMPUTE_SHADER_SOBEL   sobelfeldman_generator.comp (datamanager.cpp:623-717): the same stencil in fp32
//                                       on the R16F volume texels
// What the marchers sample is the GL_RGB16F texture (GL_LINEAR, clamp to edge): here half4 texels (x, y, z, 0) = one
// 64-bit load per tap, padded by one replicated texel like the volume, so the volume's tap indices are reused.
// The fp64 accumulation order is the reference's (compiled with -fmad=false), so modes 1 and 2 match it bit for bit.
#include "vrb_internal.cuh"
#include <algorithm>
#include <cstring>

template <typename T>
__device__ __forceinline__ double grad_norm_sample(const T* __restrict__ raw, int w, int h, int d, double maxv, int x, int y, int z) {
  if (x < 0 || y < 0 || z < 0 || x >= w || y >= h || z >= d) return 0.0;      // StructuredGridVolume::GetNormalizedSample
  return (double)raw[(size_t)x + (size_t)w * ((size_t)y + (size_t)h * (size_t)z)] / maxv;
}
__device__ __forceinline__ float grad_texel(const __half* __restrict__ tex, int w, int h, int d, int x, int y, int z) {
  if (x < 0 || y < 0 || z < 0 || x >= w || y >= h || z >= d) return 0.0f;       // GetScalarValue (sobelfeldman_generator.comp:14-21)
  return __half2float(tex[(size_t)(x + 1) + (size_t)(w + 2) * ((size_t)(y + 1) + (size_t)(h + 2) * (size_t)(z + 1))]);
}

template <typename T, int MODE>
__global__ void __launch_bounds__(256)
k_gradient(const T* __restrict__ raw, const __half* __restrict__ tex, uint2* __restrict__ out, int w, int h, int d, double maxv) {
  const int pw = w + 2, ph = h + 2, pd = d + 2;
  const long long n = (long long)pw * ph * pd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % pw), Y = (int)((i / pw) % ph), Z = (int)(i / ((long long)pw * ph));
    const int x = min(max(X - 1, 0), w - 1), y = min(max(Y - 1, 0), h - 1), z = min(max(Z - 1, 0), d - 1);   // border texels replicate
    float gx, gy, gz;
    if (MODE == VRB_GRADIENT_SOBEL_FELDMAN) {
      double sx = 0.0, sy = 0.0, sz = 0.0;
      for (int v1 = -1; v1 <= 1; ++v1)
        for (int v2 = -1; v2 <= 1; ++v2) {
          const double wp = 4.0 / (double)(1 << (abs(v1) + abs(v2))), wn = -wp;        // 4 / pow(2, |v1| + |v2|)
          sz += grad_norm_sample(raw, w, h, d, maxv, x + v1, y + v2, z - 1) * wp + grad_norm_sample(raw, w, h, d, maxv, x + v1, y + v2, z + 1) * wn;
          sy += grad_norm_sample(raw, w, h, d, maxv, x + v1, y - 1, z + v2) * wp + grad_norm_sample(raw, w, h, d, maxv, x + v1, y + 1, z + v2) * wn;
          sx += grad_norm_sample(raw, w, h, d, maxv, x - 1, y + v2, z + v1) * wp + grad_norm_sample(raw, w, h, d, maxv, x + 1, y + v2, z + v1) * wn;
        }
      gx = (float)sx; gy = (float)sy; gz = (float)sz;
    } else if (MODE == VRB_GRADIENT_FINITE_DIFFERENCES) {
      double dx = grad_norm_sample(raw, w, h, d, maxv, x + 1, y, z) - grad_norm_sample(raw, w, h, d, maxv, x - 1, y, z);
      double dy = grad_norm_sample(raw, w, h, d, maxv, x, y + 1, z) - grad_norm_sample(raw, w, h, d, maxv, x, y - 1, z);
      double dz = grad_norm_sample(raw, w, h, d, maxv, x, y, z + 1) - grad_norm_sample(raw, w, h, d, maxv, x, y, z - 1);
      const double inv = 1.0 / sqrt(dx * dx + dy * dy + dz * dz);      // glm::normalize<double>: v * inversesqrt(dot(v, v))
      dx *= inv; dy *= inv; dz *= inv;
      if (dx != dx) { dx = 0.0; dy = 0.0; dz = 0.0; }
      gx = (float)dx; gy = (float)dy; gz = (float)dz;
    } else {
      float sx = 0.f, sy = 0.f, sz = 0.f;
      for (int v1 = -1; v1 < 2; ++v1)
        for (int v2 = -1; v2 < 2; ++v2) {
          const float wp = 4.0f / (float)(1 << (abs(v1) + abs(v2))), wn = -wp;
          sz = sz + grad_texel(tex, w, h, d, x + v1, y + v2, z - 1) * wp + grad_texel(tex, w, h, d, x + v1, y + v2, z + 1) * wn;
          sy = sy + grad_texel(tex, w, h, d, x + v1, y - 1, z + v2) * wp + grad_texel(tex, w, h, d, x + v1, y + 1, z + v2) * wn;
          sx = sx + grad_texel(tex, w, h, d, x - 1, y + v2, z + v1) * wp + grad_texel(tex, w, h, d, x + 1, y + v2, z + v1) * wn;
        }
      gx = sx; gy = sy; gz = sz;
    }
    __half2 lo = __floats2half2_rn(gx, gy), hi = __floats2half2_rn(gz, 0.0f);
    uint2 pk; pk.x = *reinterpret_cast<unsigned int*>(&lo); pk.y = *reinterpret_cast<unsigned int*>(&hi);
    out[i] = pk;
  }
}

void vrb_free_gradient(vrb_ctx* c) {
  if (c->d_grad) cudaFree(c->d_grad);
  c->d_grad = nullptr; c->grad_mode = VRB_GRADIENT_NONE;
}

extern "C" int vrb_gradient_build(vrb_ctx* c, int mode) {
  VRB_REQUIRE(c, VRB_ERR_INVALID, "vrb_gradient_build: ctx is NULL");
  VRB_REQUIRE(mode >= VRB_GRADIENT_NONE && mode <= VRB_GRADIENT_COMPUTE_SHADER_SOBEL, VRB_ERR_INVALID, "vrb_gradient_build: mode %d", mode);
  VRB_CUDA(cudaSetDevice(c->device));
  if (mode == VRB_GRADIENT_NONE) { VRB_CUDA(cudaStreamSynchronize(c->stream)); vrb_free_gradient(c); return VRB_OK; }
  VRB_REQUIRE(c->d_raw && c->d_vol, VRB_ERR_STATE, "vrb_gradient_build: no volume uploaded");
  VRB_CUDA(cudaStreamSynchronize(c->stream));
  vrb_free_gradient(c);
  const int w = c->vw, h = c->vh, d = c->vd;
  const size_t np = (size_t)(w + 2) * (h + 2) * (d + 2);
  VRB_CUDA(cudaMalloc(&c->d_grad, np * sizeof(uint2)));
  const int blocks = (int)std::min<size_t>((np + 255) / 256, 148 * 64);
  const double maxv = c->bpv == 1 ? 255.0 : 65535.0;
#define VRB_GRAD(T, M) k_gradient<T, M><<<blocks, 256, 0, c->stream>>>((const T*)c->d_raw, c->d_vol, (uint2*)c->d_grad, w, h, d, maxv)
  if (c->bpv == 1) {
    if (mode == VRB_GRADIENT_SOBEL_FELDMAN) VRB_GRAD(uint8_t, VRB_GRADIENT_SOBEL_FELDMAN);
    else if (mode == VRB_GRADIENT_FINITE_DIFFERENCES) VRB_GRAD(uint8_t, VRB_GRADIENT_FINITE_DIFFERENCES);
    else VRB_GRAD(uint8_t, VRB_GRADIENT_COMPUTE_SHADER_SOBEL);
  } else {
    if (mode == VRB_GRADIENT_SOBEL_FELDMAN) VRB_GRAD(uint16_t, VRB_GRADIENT_SOBEL_FELDMAN);
    else if (mode == VRB_GRADIENT_FINITE_DIFFERENCES) VRB_GRAD(uint16_t, VRB_GRADIENT_FINITE_DIFFERENCES);
    else VRB_GRAD(uint16_t, VRB_GRADIENT_COMPUTE_SHADER_SOBEL);
  }
#undef VRB_GRAD
  VRB_CUDA(cudaGetLastError());
  c->launches++;
  c->grad_mode = mode;
  return VRB_OK;
}

extern "C" int vrb_gradient_mode(const vrb_ctx* c) { return c ? c->grad_mode : VRB_GRADIENT_NONE; }

__global__ void k_gradient_unpad(const uint2* __restrict__ g, float* __restrict__ out, int w, int h, int d) {
  const long long n = (long long)w * h * d;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = (int)(i % w), y = (int)((i / w) % h), z = (int)(i / ((long long)w * h));
  uint2 pk = g[(size_t)(x + 1) + (size_t)(w + 2) * ((size_t)(y + 1) + (size_t)(h + 2) * (size_t)(z + 1))];
  float2 a = __half22float2(*reinterpret_cast<__half2*>(&pk.x)), b = __half22float2(*reinterpret_cast<__half2*>(&pk.y));
  out[3 * i] = a.x; out[3 * i + 1] = a.y; out[3 * i + 2] = b.x;
}

// the RGB16F texels as floats, w*h*d*3, x fastest
extern "C" int vrb_gradient_read(vrb_ctx* c, float* host_xyz) {
  VRB_REQUIRE(c && host_xyz, VRB_ERR_INVALID, "vrb_gradient_read: NULL argument");
  VRB_REQUIRE(c->d_grad, VRB_ERR_STATE, "vrb_gradient_read: no gradient texture (vrb_gradient_build)");
  VRB_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)c->vw * c->vh * c->vd;
  float* tmp = nullptr;
  VRB_CUDA(cudaMalloc(&tmp, n * 3 * sizeof(float)));
  k_gradient_unpad<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((const uint2*)c->d_grad, tmp, c->vw, c->vh, c->vd);
  c->launches++;
  cudaError_t e = cudaMemcpyAsync(host_xyz, tmp, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
  cudaError_t e2 = cudaStreamSynchronize(c->stream);
  cudaFree(tmp);
  VRB_REQUIRE(e == cudaSuccess && e2 == cudaSuccess, VRB_ERR_CUDA, "vrb_gradient_read: copy failed");
  return VRB_OK;
}

int vrb_make_phong_view(const vrb_ctx* c, const vrb_lighting* light, PhongView* out, const char* who) {
  memset(out, 0, sizeof(*out));
  if (!light || light->apply_phong != 1) return VRB_OK;
  VRB_REQUIRE(c->d_grad, VRB_ERR_STATE, "%s: apply_phong needs the gradient texture (vrb_gradient_build)", who);
  out->grad = (const uint2*)c->d_grad;
  out->lx = light->light_pos[0]; out->ly = light->light_pos[1]; out->lz = light->light_pos[2];
  out->ka = light->ka; out->kd = light->kd; out->ks = light->ks; out->shininess = light->shininess;
  out->isx = light->ispecular[0]; out->isy = light->ispecular[1]; out->isz = light->ispecular[2];
  return VRB_OK;
}
