// oracle/oracle_gradient.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// CPU restatement of the three gradient generators behind DataManager::GenerateStructuredGradientTexture
// (libs/volvis_utils/datamanager.cpp:332-352):
//   mode 1  vis::GenerateSobelFeldmanGradientTexture   (libs/volvis_utils/utils.cpp:287-350, double accumulation)
//   mode 2  vis::GenerateGradientTexture, default args (utils.cpp:146-284: central differences at distance 1, normalised,
//           no filtering)
//   mode 3  DataManager::GenerateGradientWithComputeShader (datamanager.cpp:623-717) = sobelfeldman_generator.comp,
//           fp32 on the R16F volume texture, R16F image stores, then re-uploaded as RGB16F
// Modes 1 and 2 are pinned bit for bit against the reference's own utils.cpp compiled into oracle/_ref
// (tests/test_gradient.py); mode 3 is pinned against sobelfeldman_generator.comp run on the CPU (tests/test_refglsl.py).
// Output: w*h*d*3 floats rounded to fp16 (the GL_RGB16F texels the shaders sample).
#include "oracle_common.h"

using namespace orc;

namespace {
Tex3D g_grad;
bool g_grad_set = false;
inline double norm_sample(const void* vox, int w, int h, int d, int bpv, int x, int y, int z) {
  // StructuredGridVolume::GetNormalizedSample (structuredgridvolume.cpp:121-151)
  if (x < 0 || y < 0 || z < 0 || x >= w || y >= h || z >= d) return 0.0;
  const size_t i = (size_t)x + (size_t)y * w + (size_t)z * w * h;
  return bpv == 1 ? (double)((const uint8_t*)vox)[i] / (256.0 - 1.0) : (double)((const uint16_t*)vox)[i] / (65536.0 - 1.0);
}
}  // namespace

namespace orc {
const Tex3D* gradient_texture() { return g_grad_set ? &g_grad : nullptr; }
}

extern "C" {

// binds / unbinds (rgb == NULL) TexVolumeGradient for the renderers of this library; the array must outlive the renders
void orc_set_gradient(const float* rgb16f, int w, int h, int d) {
  g_grad_set = rgb16f != nullptr;
  g_grad.w = w; g_grad.h = h; g_grad.d = d; g_grad.c = 3; g_grad.data = rgb16f;
}

int orc_gradient_build(const void* vox, int w, int h, int d, int bpv, int mode, float* out_rgb16f) {
  if (mode == 1) {
#pragma omp parallel for schedule(static)
    for (int z = 0; z < d; ++z)
      for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
          double sx = 0.0, sy = 0.0, sz = 0.0;
          for (int v1 = -1; v1 <= 1; ++v1)
            for (int v2 = -1; v2 <= 1; ++v2) {
              const double wp = 4.0 / std::pow(2.0, std::abs(v1) + std::abs(v2)), wn = -4.0 / std::pow(2.0, std::abs(v1) + std::abs(v2));
              sz += norm_sample(vox, w, h, d, bpv, x + v1, y + v2, z - 1) * wp + norm_sample(vox, w, h, d, bpv, x + v1, y + v2, z + 1) * wn;
              sy += norm_sample(vox, w, h, d, bpv, x + v1, y - 1, z + v2) * wp + norm_sample(vox, w, h, d, bpv, x + v1, y + 1, z + v2) * wn;
              sx += norm_sample(vox, w, h, d, bpv, x - 1, y + v2, z + v1) * wp + norm_sample(vox, w, h, d, bpv, x + 1, y + v2, z + v1) * wn;
            }
          float* o = out_rgb16f + 3 * ((size_t)x + (size_t)y * w + (size_t)z * w * h);
          o[0] = round_f16((float)sx); o[1] = round_f16((float)sy); o[2] = round_f16((float)sz);
        }
    return 0;
  }
  if (mode == 2) {
#pragma omp parallel for schedule(static)
    for (int z = 0; z < d; ++z)
      for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
          double gx = norm_sample(vox, w, h, d, bpv, x + 1, y, z) - norm_sample(vox, w, h, d, bpv, x - 1, y, z);
          double gy = norm_sample(vox, w, h, d, bpv, x, y + 1, z) - norm_sample(vox, w, h, d, bpv, x, y - 1, z);
          double gz = norm_sample(vox, w, h, d, bpv, x, y, z + 1) - norm_sample(vox, w, h, d, bpv, x, y, z - 1);
          // glm::normalize<double>: v * inversesqrt(dot(v, v)) (include/glm/detail/func_geometric.inl:257-266)
          double inv = 1.0 / std::sqrt(gx * gx + gy * gy + gz * gz);
          gx *= inv; gy *= inv; gz *= inv;
          if (gx != gx) { gx = 0.0; gy = 0.0; gz = 0.0; }       // "lm.IsNaN": zero difference -> zero gradient
          float* o = out_rgb16f + 3 * ((size_t)x + (size_t)y * w + (size_t)z * w * h);
          o[0] = round_f16((float)gx); o[1] = round_f16((float)gy); o[2] = round_f16((float)gz);
        }
    return 0;
  }
  if (mode == 3) {
    std::vector<float> tex((size_t)w * h * d);
    volume_to_r16f(vox, tex.size(), bpv, tex.data());
    Tex3D vol; vol.w = w; vol.h = h; vol.d = d; vol.c = 1; vol.data = tex.data();
    auto scalar = [&](int px, int py, int pz) -> float {       // GetScalarValue (sobelfeldman_generator.comp:14-21)
      if (px < 0 || py < 0 || pz < 0 || (float)px > (float)w - 1.0f || (float)py > (float)h - 1.0f || (float)pz > (float)d - 1.0f) return 0.0f;
      return tex3d(vol, (v3((float)px, (float)py, (float)pz) + v3(0.5f, 0.5f, 0.5f)) / v3((float)w, (float)h, (float)d));
    };
#pragma omp parallel for schedule(static)
    for (int z = 0; z < d; ++z)
      for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
          float sx = 0.f, sy = 0.f, sz = 0.f;
          for (int v1 = -1; v1 < 2; ++v1)
            for (int v2 = -1; v2 < 2; ++v2) {
              const float wp = 4.0f / std::pow(2.0f, (float)(std::abs(v1) + std::abs(v2))), wn = -4.0f / std::pow(2.0f, (float)(std::abs(v1) + std::abs(v2)));
              sz = sz + scalar(x + v1, y + v2, z - 1) * wp + scalar(x + v1, y + v2, z + 1) * wn;
              sy = sy + scalar(x + v1, y - 1, z + v2) * wp + scalar(x + v1, y + 1, z + v2) * wn;
              sx = sx + scalar(x - 1, y + v2, z + v1) * wp + scalar(x + 1, y + v2, z + v1) * wn;
            }
          float* o = out_rgb16f + 3 * ((size_t)x + (size_t)y * w + (size_t)z * w * h);
          o[0] = round_f16(sx); o[1] = round_f16(sy); o[2] = round_f16(sz);       // imageStore into r16f, re-uploaded as RGB16F
        }
    return 0;
  }
  return -1;
}

}  // extern "C"
