// oracle/oracle_tf.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// CPU restatement of vis::TransferFunction1D (libs/volvis_utils/transferfunction1d.cpp) and of the
// opacity<->extinction helpers (libs/volvis_utils/transferfunction.h:74-82).
// Pinned against the reference's own TransferFunction1D compiled into oracle/_ref (tests/test_oracle_ref.py).
#include "oracle_common.h"

extern "C" {

// BuildLinear (transferfunction1d.cpp:319-358).  Control-point colours are stored as glm::vec4 (float)
// by TransferControlPoint (transferfunction.cpp:8-24); differences are taken in float, the lerp in double.
// Entries never covered by a segment are uninitialised in the reference (new glm::dvec4[], :124); we zero them
// (SURVEY.md Appendix B).  rgb_pts: n_rgb x (r,g,b,iso); a_pts: n_a x (a,iso).  out: (max_density+1) x 4 doubles.
void orc_tf_build(const double* rgb_pts, int n_rgb, const double* a_pts, int n_a, int max_density, double* out) {
  for (int i = 0; i < (max_density + 1) * 4; ++i) out[i] = 0.0;
  for (int i = 0; i < n_rgb - 1; ++i) {
    int i0 = (int)rgb_pts[4 * i + 3], i1 = (int)rgb_pts[4 * (i + 1) + 3];
    float c0[3], c1[3];
    for (int k = 0; k < 3; ++k) { c0[k] = (float)rgb_pts[4 * i + k]; c1[k] = (float)rgb_pts[4 * (i + 1) + k]; }
    double diff[3];
    for (int k = 0; k < 3; ++k) diff[k] = (double)(float)(c1[k] - c0[k]);
    for (int x = i0; x <= i1; ++x) {
      if (x < 0 || x > max_density) continue;
      double k = (double)(x - i0) / (double)(i1 - i0);
      for (int c = 0; c < 3; ++c) out[4 * x + c] = c0[c] + diff[c] * k;
    }
  }
  for (int i = 0; i < n_a - 1; ++i) {
    int i0 = (int)a_pts[2 * i + 1], i1 = (int)a_pts[2 * (i + 1) + 1];
    float a0 = (float)a_pts[2 * i], a1 = (float)a_pts[2 * (i + 1)];
    double diff = (double)(float)(a1 - a0);
    for (int x = i0; x <= i1; ++x) {
      if (x < 0 || x > max_density) continue;
      double k = (double)(x - i0) / (double)(i1 - i0);
      out[4 * x + 3] = a0 + diff * k;
    }
  }
}

// transferfunction.h:74-82 (float argument, double arithmetic)
static inline double ext_to_opc(float e) { return 1.0 - (double)std::exp(-e); }   // glm::exp(float): fp32 exp, then a double subtraction
static inline double opc_to_ext(float a) { return std::log(1.0 / (1.0 - (double)a)); }

// TransferFunction1D::Get (transferfunction1d.cpp:132-157); returns the 4 channels as float (glm::vec4 cast).
void orc_tf_get(const double* table, int max_density, double value, double max_data_value, float out[4]) {
  if (max_data_value >= 0) value = value * ((double)max_density / max_data_value);
  if (value < 0.0f || value > (float)max_density) { out[0] = out[1] = out[2] = out[3] = 0.0f; return; }
  if (std::fabs(value - (float)max_density) < 0.000001) {
    for (int c = 0; c < 4; ++c) out[c] = (float)table[4 * max_density + c];
  } else {
    int iv = (int)value;
    double t = value - iv;
    for (int c = 0; c < 4; ++c) out[c] = (float)((1.0 - t) * table[4 * iv + c] + t * table[4 * (iv + 1) + c]);
  }
}

// GetExtN / GetOpcN (transferfunction1d.cpp:169-197)
float orc_tf_get_extn(const double* table, int max_density, int ext_type, double normalized) {
  float v[4]; orc_tf_get(table, max_density, normalized, 1.0, v);
  if (!ext_type) return (float)opc_to_ext(v[3]);
  return v[3];
}
float orc_tf_get_opcn(const double* table, int max_density, int ext_type, double normalized) {
  float v[4]; orc_tf_get(table, max_density, normalized, 1.0, v);
  if (ext_type) return (float)ext_to_opc(v[3]);
  return v[3];
}
// GetOpc(value, max_input_value) (transferfunction1d.cpp:159-167)
float orc_tf_get_opc(const double* table, int max_density, int ext_type, double value, double max_input) {
  float v[4]; orc_tf_get(table, max_density, value, max_input, v);
  if (ext_type) return (float)ext_to_opc(v[3]);
  return v[3];
}

// GenerateTexture_1D_RGBt / _RGBA (transferfunction1d.cpp:89-118 / 58-87): float data handed to GL,
// which stores RGBA16F.  We return the fp16-ROUNDED texel values (as float) that the sampler sees.
void orc_tf_texture_rgbt(const double* table, int max_density, int ext_type, float* out) {
  for (int i = 0; i <= max_density; ++i) {
    out[4 * i + 0] = orc::round_f16((float)table[4 * i + 0]);
    out[4 * i + 1] = orc::round_f16((float)table[4 * i + 1]);
    out[4 * i + 2] = orc::round_f16((float)table[4 * i + 2]);
    float v4 = (float)table[4 * i + 3];
    if (!ext_type) v4 = (float)opc_to_ext(v4);
    out[4 * i + 3] = orc::round_f16(v4);
  }
}
void orc_tf_texture_rgba(const double* table, int max_density, int ext_type, float* out) {
  for (int i = 0; i <= max_density; ++i) {
    out[4 * i + 0] = orc::round_f16((float)table[4 * i + 0]);
    out[4 * i + 1] = orc::round_f16((float)table[4 * i + 1]);
    out[4 * i + 2] = orc::round_f16((float)table[4 * i + 2]);
    float v4 = (float)(double)table[4 * i + 3];
    if (ext_type) v4 = (float)ext_to_opc(v4);
    out[4 * i + 3] = orc::round_f16(v4);
  }
}
// The same two textures BEFORE the fp16 rounding: this is what the host hands to the C ABI (the GL_FLOAT
// client array of SetData); the library rounds on upload like the GL driver does.
void orc_tf_floats_rgbt(const double* table, int max_density, int ext_type, float* out) {
  for (int i = 0; i <= max_density; ++i) {
    for (int c = 0; c < 3; ++c) out[4 * i + c] = (float)table[4 * i + c];
    float v4 = (float)table[4 * i + 3];
    if (!ext_type) v4 = (float)opc_to_ext(v4);
    out[4 * i + 3] = v4;
  }
}
void orc_tf_floats_rgba(const double* table, int max_density, int ext_type, float* out) {
  for (int i = 0; i <= max_density; ++i) {
    for (int c = 0; c < 3; ++c) out[4 * i + c] = (float)table[4 * i + c];
    float v4 = (float)(double)table[4 * i + 3];
    if (ext_type) v4 = (float)ext_to_opc(v4);
    out[4 * i + 3] = v4;
  }
}

// fp16 helpers exported for the tests (checked against numpy.float16)
uint16_t orc_f32_to_f16_bits(float f) { return orc::f32_to_f16_bits(f); }
float orc_f16_bits_to_f32(uint16_t h) { return orc::f16_bits_to_f32(h); }
void orc_round_f16_array(const float* in, float* out, size_t n) { for (size_t i = 0; i < n; ++i) out[i] = orc::round_f16(in[i]); }

void orc_look_at(const float eye[3], const float center[3], const float up[3], float out[16]) {
  orc::look_at(eye, center, up, out);
}

void orc_volume_to_r16f(const void* vox, size_t n, int bytes_per_voxel, float* out) {
  orc::volume_to_r16f(vox, n, bytes_per_voxel, out);
}

}  // extern "C"
