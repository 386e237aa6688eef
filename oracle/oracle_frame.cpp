// oracle/oracle_frame.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// CPU restatement of the pixel multi-scaling passes of vis::RenderFrameToScreen (libs/vis_utils/renderoutputframe.cpp:
// 265-540) and their shaders under libs/vis_utils/shader/renderoutputframe/: multisample_filter.comp,
// downscaling_filter.comp, upscaling_filter.comp, the six *_filter.comp kernels and the cbs_ / comoms_digital_filter.comp
// recursions (in place, every step rounded to fp16 by the rgba16f imageStore).  Images are H x W x 4 floats holding
// fp16-representable values, row 0 = bottom.  texelFetch outside the texture (undefined in GL 4.3 without robust
// access) returns zero, imageStore outside the image is dropped.  Pinned against the reference's own shaders run on the CPU (tests/test_refglsl.py).
#include "oracle_common.h"

using namespace orc;

namespace {
struct Img { float* p; int w, h; };
inline V4 ld(const Img& im, int x, int y) { const float* q = im.p + 4 * ((size_t)y * im.w + x); return V4{q[0], q[1], q[2], q[3]}; }
inline V4 fetch0(const Img& im, int x, int y) { if (x < 0 || y < 0 || x >= im.w || y >= im.h) return V4{0, 0, 0, 0}; return ld(im, x, y); }
inline void st(const Img& im, int x, int y, V4 v) {
  if (x < 0 || y < 0 || x >= im.w || y >= im.h) return;
  float* q = im.p + 4 * ((size_t)y * im.w + x);
  q[0] = round_f16(v.x); q[1] = round_f16(v.y); q[2] = round_f16(v.z); q[3] = round_f16(v.w);
}
float support(int K) { return K == 0 ? 1.0f : (K == 1 ? 2.0f : 4.0f); }
float k0(int K, float u) {
  if (K == 2) return ((.5f * u - .5f) * u) * u;
  if (K == 3) return (((7 / 18.0f) * u - 1 / 3.0f) * u) * u;
  if (K == 4) return ((u)*u) * u;
  return ((.875f * u) * u + .125f) * u;
}
float k1(int K, float u) {
  if (K == 2) return ((-1.5f * u + 2.0f) * u + .5f) * u;
  if (K == 3) return (((-7 / 6.0f) * u + 1.5f) * u + 0.5f) * u + 1 / 18.0f;
  if (K == 4) return ((-3.0f * u + 3.0f) * u + 3.0f) * u + 1.0f;
  return ((-2.625f * u + 2.625f) * u + 2.25f) * u + 1.0f;
}
float weight(int K, float x) {
  if (K == 0) return (x <= -0.5f || x > 0.5f) ? 0.0f : 1.0f;
  x = std::fabs(x);
  if (K == 1) return x > 1.0f ? 0.0f : 1.0f - x;
  return x > 2.0f ? 0.0f : (x > 1.0f ? k0(K, 2.0f - x) : k1(K, 1.0f - x));
}
void digital(const Img& im, int K) {
  static const float L8[8] = {.2f, .26315789f, .26760563f, .26792453f, .26794742f, .26794907f, .26794918f, .26794919f};
  static const float L9[9] = {.23529412f, .33170732f, .34266611f, .34395774f, .34411062f, .34412872f, .34413087f, .34413112f, .34413115f};
  const float* L = K == 4 ? L8 : L9;
  const int m = K == 4 ? 8 : 9;
  const float p_inv = 1.0f;
  const float L_inf = L[m - 1], v_inv = L_inf / (1.f + L_inf);
  for (int direction = 0; direction < 2; ++direction) {
    const int nlines = direction == 0 ? im.h : im.w, nn = direction == 0 ? im.w : im.h;
#pragma omp parallel for schedule(static)
    for (int line = 0; line < nlines; ++line) {
      auto X = [&](int i) { return direction == 0 ? i : line; };
      auto Y = [&](int i) { return direction == 0 ? line : i; };
      auto sub = [](V4 a, float l, V4 b) { return V4{a.x - (l * b.x), a.y - (l * b.y), a.z - (l * b.z), a.w - (l * b.w)}; };
      auto back = [&](float l, V4 a, V4 b) { return V4{l * (p_inv * a.x - b.x), l * (p_inv * a.y - b.y), l * (p_inv * a.z - b.z), l * (p_inv * a.w - b.w)}; };
      for (int x = 1; x < m; ++x) st(im, X(x), Y(x), sub(fetch0(im, X(x), Y(x)), L[x - 1], fetch0(im, X(x - 1), Y(x - 1))));
      for (int x = m; x < nn; ++x) st(im, X(x), Y(x), sub(fetch0(im, X(x), Y(x)), L_inf, fetch0(im, X(x - 1), Y(x - 1))));
      { V4 a = fetch0(im, X(nn - 1), Y(nn - 1)); st(im, X(nn - 1), Y(nn - 1), V4{a.x * p_inv * v_inv, a.y * p_inv * v_inv, a.z * p_inv * v_inv, a.w * p_inv * v_inv}); }
      for (int x = nn - 2; x >= m - 1; --x) st(im, X(x), Y(x), back(L_inf, fetch0(im, X(x), Y(x)), fetch0(im, X(x + 1), Y(x + 1))));
      for (int x = m - 2; x >= 0; --x) st(im, X(x), Y(x), back(L[x], fetch0(im, X(x), Y(x)), fetch0(im, X(x + 1), Y(x + 1))));
    }
  }
}
}  // namespace

extern "C" {

// pass 1 multisample, 2 down-scale, 3 up-scale.  src is MODIFIED by pass 3 with a cardinal kernel (the digital
// pre-filter runs in place on the rendered frame, renderoutputframe.cpp:472-502).
int orc_frame_filter(float* src, int sw, int sh, float* dst, int dw, int dh, int pass, int K) {
  Img S{src, sw, sh}, D{dst, dw, dh};
  if (pass == 1) {
    Tex3D t; t.w = sw; t.h = sh; t.d = 1; t.c = 4; t.data = src;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dh; ++y)
      for (int x = 0; x < dw; ++x) {
        const float fx = ((float)x + 0.5f) / (float)dw, fy = ((float)y + 0.5f) / (float)dh;
        int x0, x1, y0, y1; float tx, ty;
        lin_coord(fx, sw, &x0, &x1, &tx);
        lin_coord(fy, sh, &y0, &y1, &ty);
        V4 a = ld(S, x0, y0), b = ld(S, x1, y0), c = ld(S, x0, y1), d = ld(S, x1, y1);
        st(D, x, y, V4{lerp(lerp(a.x, b.x, tx), lerp(c.x, d.x, tx), ty), lerp(lerp(a.y, b.y, tx), lerp(c.y, d.y, tx), ty),
                       lerp(lerp(a.z, b.z, tx), lerp(c.z, d.z, tx), ty), lerp(lerp(a.w, b.w, tx), lerp(c.w, d.w, tx), ty)});
      }
    return 0;
  }
  if (K < 0 || K > 5) return -1;
  const bool cardinal = K == 4 || K == 5;
  const float kr = 0.5f * support(K);
  if (pass == 2) {
    const float s_r = (float)dh / (float)sh, s_c = (float)dw / (float)sw;
    const int n_r = sh, n_c = sw;
#pragma omp parallel for schedule(static)
    for (int j_r = 0; j_r < dh; ++j_r)
      for (int j_c = 0; j_c < dw; ++j_c) {
        float x_r = ((float)j_r + 0.5f) / (float)dh;
        int il_r = (int)std::ceil((x_r - kr / (float)dh) * (float)n_r - 0.5f), ir_r = (int)std::floor((x_r + kr / (float)dh) * (float)n_r - 0.5f);
        float x_c = ((float)j_c + 0.5f) / (float)dw;
        int il_c = (int)std::ceil((x_c - kr / (float)dw) * (float)n_c - 0.5f), ir_c = (int)std::floor((x_c + kr / (float)dw) * (float)n_c - 0.5f);
        V4 f{0, 0, 0, 0};
        for (int i_r = il_r; i_r <= ir_r; ++i_r)
          for (int i_c = il_c; i_c <= ir_c; ++i_c) {
            float wgt = weight(K, (x_r - ((float)i_r + 0.5f) / (float)n_r) * (float)dh) * weight(K, (x_c - ((float)i_c + 0.5f) / (float)n_c) * (float)dw);
            V4 t = fetch0(S, i_c, i_r);
            f.x += wgt * t.x; f.y += wgt * t.y; f.z += wgt * t.z; f.w += wgt * t.w;
          }
        const float sc = (s_r * s_c);
        st(D, j_c, j_r, V4{f.x * sc, f.y * sc, f.z * sc, f.w * sc});
      }
    if (cardinal) digital(D, K);
    return 0;
  }
  if (pass == 3) {
    if (cardinal) digital(S, K);
#pragma omp parallel for schedule(static)
    for (int j_r = 0; j_r < dh; ++j_r)
      for (int j_c = 0; j_c < dw; ++j_c) {
        float x_r = ((float)j_r + 0.5f) / (float)dh;
        float xi_r = x_r * (float)sh - 0.5f;
        int il_r = (int)std::ceil(xi_r - kr), ir_r = (int)std::floor(xi_r + kr);
        float x_c = ((float)j_c + 0.5f) / (float)dw;
        float xi_c = x_c * (float)sw - 0.5f;
        int il_c = (int)std::ceil(xi_c - kr), ir_c = (int)std::floor(xi_c + kr);
        V4 f{0, 0, 0, 0};
        for (int i_r = il_r; i_r <= ir_r; ++i_r)
          for (int i_c = il_c; i_c <= ir_c; ++i_c) {
            float wgt = weight(K, xi_r - (float)i_r) * weight(K, xi_c - (float)i_c);
            V4 t = fetch0(S, i_c, i_r);
            f.x += wgt * t.x; f.y += wgt * t.y; f.z += wgt * t.z; f.w += wgt * t.w;
          }
        st(D, j_c, j_r, f);
      }
    return 0;
  }
  return -1;
}

}  // extern "C"
