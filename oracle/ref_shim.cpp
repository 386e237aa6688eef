// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin extern "C" wrapper around the pieces of the REFERENCE that compile with g++ as they are
// (SURVEY.md F5): vis::SummedAreaTable3D<T> (libs/vis_utils/summedareatable.h:171-303),
// vis::TransferFunction1D (libs/volvis_utils/transferfunction1d.cpp) and vis::StructuredGridVolume
// (libs/volvis_utils/structuredgridvolume.cpp).  The reference sources are compiled from where they lie
// under /root/reference (see oracle/Makefile target `ref`); nothing is copied into this repository.
// gl::Texture1D is replaced by a link-time stub that captures the GL_FLOAT client array handed to
// SetData, so GenerateTexture_1D_RGBt/_RGBA can be pinned too.
//
// Output: oracle/_ref/libref.so (git-ignored; travels to the GPU box).
#include <vis_utils/summedareatable.h>
#include <volvis_utils/transferfunction1d.h>
#include <volvis_utils/structuredgridvolume.h>
#include <gl_utils/texture1d.h>
#include <gl_utils/texture2d.h>
#include <gl_utils/texture3d.h>
#include <vis_utils/colorutils.h>                  // Cie2000Comparison (libs/vis_utils/colorutils.cpp:221-311)
#include <volvis_utils/utils.h>                    // vis::GenerateGradientTexture / GenerateSobelFeldmanGradientTexture / GenerateRTexture
#include <glm/gtc/matrix_transform.hpp>
#include <vector>
#include <cstring>
#include <cstdint>
#include <math_utils/utils.h>                      // oracle/ref_stubs/math_utils/utils.h (see the Makefile's -I order)
#include <conegaussiansampler.h>                   // $(REF)/cppvolrend/structured/rc1pdosct
#include <volvis_utils/camerastatelist.h>           // CameraStateList (libs/volvis_utils/camerastatelist.cpp, compiled in place)
#include <volvis_utils/lightsourcelist.h>           // LightSourceList (libs/volvis_utils/lightsourcelist.cpp, compiled in place)
#include <../cppvolrend/utils/parameterspace.h>     // ParameterSpace + the reference's own ParameterSpaceTest() (cppvolrend/utils/parameterspace.cpp)
#include <file_utils/pvm.h>                        // Pvm, DDSV3 (libs/file_utils/pvm.cpp, compiled in place with ref_stubs/msvc_compat.h)

// libs/math_utils/utils.cpp:149-165 restated (that file does not compile outside MSVC); needed by the reference's
// conegaussiansampler.cpp, which is compiled verbatim.
glm::vec3 RodriguesRotation(glm::vec3 v, float teta, glm::vec3 k) {
  glm::vec3 r = v * glm::cos(teta) + glm::cross(k, v) * glm::sin(teta) + k * glm::dot(k, v) * (1.0f - glm::cos(teta));
  return glm::normalize(r);
}
glm::dvec3 RodriguesRotation(glm::dvec3 v, double teta, glm::dvec3 k) {
  glm::dvec3 r = v * glm::cos(teta) + glm::cross(k, v) * glm::sin(teta) + k * glm::dot(k, v) * (1.0f - glm::cos(teta));
  return glm::normalize(r);
}

// vis::CameraData's constructors live in libs/vis_utils/camera.cpp, which g++ rejects (double * glm::vec3, :186,196); the list
// parser only needs the default constructor and the destructor, linked from here (defaults of camera.cpp:17-29).
namespace vis {
CameraData::CameraData() : cam_setup_name(""), c_type(0), eye(0.0f), center(0.0f), up(0.0f), aspect_ratio(1.0f), field_of_view_y(45.0f), z_near(1.0f), z_far(5000.0f) {}
CameraData::~CameraData() {}
}  // namespace vis

// ---- link-time stub of gl::Texture1D (libs/gl_utils/texture1d.cpp needs a GL context) -----------------
static std::vector<float> g_last_tex1d;
namespace gl {
void ExitOnGLError(const char*) {}   // libs/gl_utils/utils.cpp:11-30 polls glGetError; there is no GL here
Texture1D::Texture1D(unsigned int length) : m_size(length), m_length(length), m_textureID(0) {}
Texture1D::~Texture1D() {}
void Texture1D::GenerateTexture(GLint, GLint, GLint) {}
bool Texture1D::SetData(GLvoid* data, GLint, GLenum, GLenum) {
  g_last_tex1d.assign((float*)data, (float*)data + 4 * (size_t)m_length);
  return true;
}
GLuint Texture1D::GetTextureID() { return 0; }
unsigned int Texture1D::GetLength() { return m_length; }
void Texture1D::DestroyTexture() {}
// gl::Texture3D / gl::Texture2D: same kind of stub; SetData keeps the GL_FLOAT client array (1 or 3 channels) that
// libs/volvis_utils/utils.cpp hands to glTexImage3D.
Texture3D::Texture3D(unsigned int w, unsigned int h, unsigned int d) : m_width(w), m_height(h), m_depth(d), m_textureID(0) {}
Texture3D::~Texture3D() {}
void Texture3D::GenerateTexture(GLint, GLint, GLint, GLint, GLint, bool) {}
unsigned int Texture3D::GetWidth() { return m_width; }
unsigned int Texture3D::GetHeight() { return m_height; }
unsigned int Texture3D::GetDepth() { return m_depth; }
GLuint Texture3D::GetTextureID() { return 0; }
void Texture3D::DestroyTexture() {}
Texture2D::Texture2D(unsigned int w, unsigned int h) : m_width(w), m_height(h), m_textureID(0) {}
Texture2D::~Texture2D() {}
void Texture2D::GenerateTexture(GLint, GLint, GLint, GLint) {}
}  // namespace gl
static std::vector<float> g_last_tex2d;            // GL_RED / GL_FLOAT client array of the last Texture2D::SetData
static int g_last_tex2d_w = 0, g_last_tex2d_h = 0;
namespace gl {
bool Texture2D::SetData(GLvoid* data, GLint, GLenum format, GLenum type) {
  g_last_tex2d.clear(); g_last_tex2d_w = (int)m_width; g_last_tex2d_h = (int)m_height;
  if (format != GL_RED || type != GL_FLOAT || !data) return true;
  g_last_tex2d.assign((float*)data, (float*)data + (size_t)m_width * m_height);
  return true;
}
}  // namespace gl
// read access for ref_shim_vct.cpp
const std::vector<float>& ref_last_tex2d(int* w, int* h) { *w = g_last_tex2d_w; *h = g_last_tex2d_h; return g_last_tex2d; }
static std::vector<float> g_last_tex3d;
static int g_last_tex3d_channels = 0;
namespace gl {
bool Texture3D::SetData(GLvoid* data, GLint, GLenum format, GLenum type) {
  g_last_tex3d_channels = (format == GL_RGB) ? 3 : (format == GL_RED ? 1 : 0);
  if (type != GL_FLOAT || g_last_tex3d_channels == 0) { g_last_tex3d.clear(); return false; }
  const size_t n = (size_t)m_width * m_height * m_depth * g_last_tex3d_channels;
  g_last_tex3d.assign((float*)data, (float*)data + n);
  return true;
}
}  // namespace gl

extern "C" {

// ---- SummedAreaTable3D<double> exactly as RC1PExtinctionBasedShading::GenerateExtinctionSAT3DTex uses it
// (ebsrenderer.cpp:624-723): values laid out x + w*y + w*h*z, BuildSAT, cast to float.
// in: w*h*d doubles (already bordered by the caller); out_f32: float(sat) ; out_f64 optional.
void ref_sat3d_double(const double* in, int w, int h, int d, float* out_f32, double* out_f64) {
  vis::SummedAreaTable3D<double> sat(w, h, d);
  for (int z = 0; z < d; ++z)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) sat.SetValue(in[x + (size_t)w * y + (size_t)w * h * z], x, y, z);
  sat.BuildSAT();
  double* p = sat.GetData();
  size_t n = (size_t)w * h * d;
  if (out_f32) for (size_t i = 0; i < n; ++i) out_f32[i] = (float)p[i];
  if (out_f64) std::memcpy(out_f64, p, n * sizeof(double));
}

// Integer instantiation of the same template: the bit-exact integer oracle BASELINE.json asks for.
void ref_sat3d_u64(const uint64_t* in, int w, int h, int d, uint64_t* out) {
  vis::SummedAreaTable3D<unsigned long long> sat(w, h, d);
  for (int z = 0; z < d; ++z)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) sat.SetValue(in[x + (size_t)w * y + (size_t)w * h * z], x, y, z);
  sat.BuildSAT();
  std::memcpy(out, sat.GetData(), (size_t)w * h * d * sizeof(uint64_t));
}

// Timed variant for the CPU baseline (bench.py --impl reference / cpu_baseline, kind "reference"):
// fill loop equivalent to ebsrenderer.cpp:636-662 from a per-voxel-value extinction table, BuildSAT, float cast.
// Returns nothing; the caller times it.  lut has 256 (bpv=1) or 65536 (bpv=2) float entries.
void ref_sat3d_from_volume(const void* vox, int vw, int vh, int vd, int bpv, const float* lut, float* out_f32) {
  int w = vw + 2, h = vh + 2, d = vd + 2;
  vis::SummedAreaTable3D<double> sat(w, h, d);
  for (int x = 0; x < w; x++)
    for (int y = 0; y < h; y++)
      for (int z = 0; z < d; z++) {
        double val;
        if (x == 0 || y == 0 || z == 0 || x == w - 1 || y == h - 1 || z == d - 1) val = 0.0f;
        else {
          size_t id = (size_t)(x - 1) + (size_t)vw * (y - 1) + (size_t)vw * vh * (z - 1);
          val = bpv == 1 ? lut[((const uint8_t*)vox)[id]] : lut[((const uint16_t*)vox)[id]];
        }
        sat.SetValue(val, x, y, z);
      }
  sat.BuildSAT();
  double* p = sat.GetData();
  size_t n = (size_t)w * h * d;
  for (size_t i = 0; i < n; ++i) out_f32[i] = (float)p[i];
}

// ---- TransferFunction1D --------------------------------------------------------------------------------
void* ref_tf_create(const double* rgb_pts, int n_rgb, const double* a_pts, int n_a, int max_density, int ext_type) {
  vis::TransferFunction1D* tf = new vis::TransferFunction1D(max_density);
  tf->SetExtinctionCoefficientInput(ext_type != 0);
  for (int i = 0; i < n_rgb; ++i)
    tf->AddRGBControlPoint(vis::TransferControlPoint(rgb_pts[4 * i], rgb_pts[4 * i + 1], rgb_pts[4 * i + 2], (int)rgb_pts[4 * i + 3]));
  for (int i = 0; i < n_a; ++i)
    tf->AddAlphaControlPoint(vis::TransferControlPoint(a_pts[2 * i], (int)a_pts[2 * i + 1]));
  tf->Build();
  return tf;
}
void ref_tf_destroy(void* p) { delete (vis::TransferFunction1D*)p; }
void ref_tf_get(void* p, double value, double max_data_value, float out[4]) {
  glm::vec4 v = ((vis::TransferFunction1D*)p)->Get(value, max_data_value);
  out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
float ref_tf_get_extn(void* p, double n) { return ((vis::TransferFunction1D*)p)->GetExtN(n); }
float ref_tf_get_opcn(void* p, double n) { return ((vis::TransferFunction1D*)p)->GetOpcN(n); }
float ref_tf_get_opc(void* p, double v, double mx) { return ((vis::TransferFunction1D*)p)->GetOpc(v, mx); }
// float client arrays handed to glTexImage1D by GenerateTexture_1D_RGBt / _RGBA
int ref_tf_texture_rgbt(void* p, float* out, int cap) {
  gl::Texture1D* t = ((vis::TransferFunction1D*)p)->GenerateTexture_1D_RGBt();
  int n = (int)g_last_tex1d.size(); if (n > cap) n = cap;
  std::memcpy(out, g_last_tex1d.data(), n * sizeof(float)); delete t; return n;
}
int ref_tf_texture_rgba(void* p, float* out, int cap) {
  gl::Texture1D* t = ((vis::TransferFunction1D*)p)->GenerateTexture_1D_RGBA();
  int n = (int)g_last_tex1d.size(); if (n > cap) n = cap;
  std::memcpy(out, g_last_tex1d.data(), n * sizeof(float)); delete t; return n;
}

// ---- ConeGaussianSampler: the reference's own rc1pdosct/conegaussiansampler.cpp, compiled in place -----------------
struct RefConeParams { float cone_half_angle, initial_step; int max_packing; float covered_distance, d_sigma, r_sigma, ui_weight; };
struct RefConeOut { int n_sections; int counts[3]; float ray_axes[10][3]; float ray3_adj_weight, ray7_adj_weight; };
int ref_cone_sampler_compute(const RefConeParams* P, double min_sg_gaussian, float* sections_out, int cap, RefConeOut* out) {
  ConeGaussianSampler s;
  s.SetConeHalfAngle(P->cone_half_angle);
  s.SetInitialStep(P->initial_step);
  s.SetMaxGaussianPacking(P->max_packing);
  s.SetCoveredDistance(P->covered_distance);
  s.SetIntegrationHalfStepMultiplier(P->d_sigma);
  s.SetGaussianSigmaLimitMultiplier(P->r_sigma);
  s.SetUIWeightPercentage(P->ui_weight);
  s.ComputeConeIntegrationSteps(min_sg_gaussian);
  out->n_sections = s.GetNumberOfComputedConeSections();
  out->counts[0] = s.gaussian_samples_1; out->counts[1] = s.gaussian_samples_3; out->counts[2] = s.gaussian_samples_7;
  for (int i = 0; i < 3; ++i) { glm::vec3 a = s.Get3ConeRayID(i); out->ray_axes[i][0] = a.x; out->ray_axes[i][1] = a.y; out->ray_axes[i][2] = a.z; }
  for (int i = 0; i < 7; ++i) { glm::vec3 a = s.Get7ConeRayID(i); out->ray_axes[3 + i][0] = a.x; out->ray_axes[3 + i][1] = a.y; out->ray_axes[3 + i][2] = a.z; }
  out->ray3_adj_weight = (float)s.GetRay3AdjacentWeight();
  out->ray7_adj_weight = (float)s.GetRay7AdjacentWeight();
  gl::Texture1D* t = s.GetConeSectionsInfoTex();          // captured by the gl::Texture1D stub
  int n = (int)g_last_tex1d.size() / 4;
  if (t) delete t;
  if (n > cap) return -1;
  std::memcpy(sections_out, g_last_tex1d.data(), (size_t)n * 4 * sizeof(float));
  return 0;
}

// ---- StructuredGridVolume::GetNormalizedSample (structuredgridvolume.cpp:121-151) ----------------------
double ref_volume_normalized_sample(const void* vox, int w, int h, int d, int bpv, int x, int y, int z) {
  vis::StructuredGridVolume vol("v", w, h, d);
  vol.SetArrayData((void*)vox, bpv == 1 ? vis::DataStorageSize::_8_BITS : vis::DataStorageSize::_16_BITS);
  double r = vol.GetNormalizedSample(x, y, z);
  vol.SetArrayData(nullptr, vis::DataStorageSize::UNKNOWN);  // do not let the dtor free the caller's array
  return r;
}

// ---- vis::GenerateSobelFeldmanGradientTexture (mode 1) / vis::GenerateGradientTexture with its default arguments
// (mode 2), libs/volvis_utils/utils.cpp:146-350, run verbatim; out_rgb = the w*h*d*3 GL_FLOAT values given to SetData
// (before the GL_RGB16F rounding).  Returns the channel count captured (3) or 0.
int ref_gradient_texture(const void* vox, int w, int h, int d, int bpv, int mode, float* out_rgb) {
  vis::StructuredGridVolume vol("v", w, h, d);
  vol.SetArrayData(const_cast<void*>(vox), bpv == 1 ? vis::DataStorageSize::_8_BITS : vis::DataStorageSize::_16_BITS);
  g_last_tex3d.clear(); g_last_tex3d_channels = 0;
  gl::Texture3D* t = (mode == 1) ? vis::GenerateSobelFeldmanGradientTexture(&vol) : vis::GenerateGradientTexture(&vol);
  int ch = g_last_tex3d_channels;
  if (ch == 3 && g_last_tex3d.size() == (size_t)w * h * d * 3) std::memcpy(out_rgb, g_last_tex3d.data(), g_last_tex3d.size() * sizeof(float));
  else ch = 0;
  delete t;
  vol.SetArrayData(nullptr, vis::DataStorageSize::UNKNOWN);     // the volume does not own the caller's voxels
  return ch;
}

// ---- vis::GenerateRTexture (utils.cpp:58-143): the GL_FLOAT array uploaded as the R16F volume texture ----------------
int ref_volume_rtexture(const void* vox, int w, int h, int d, int bpv, float* out_r) {
  vis::StructuredGridVolume vol("v", w, h, d);
  vol.SetArrayData(const_cast<void*>(vox), bpv == 1 ? vis::DataStorageSize::_8_BITS : vis::DataStorageSize::_16_BITS);
  g_last_tex3d.clear(); g_last_tex3d_channels = 0;
  gl::Texture3D* t = vis::GenerateRTexture(&vol, 0, 0, 0, w, h, d);
  int ch = g_last_tex3d_channels;
  if (ch == 1 && g_last_tex3d.size() == (size_t)w * h * d) std::memcpy(out_r, g_last_tex3d.data(), g_last_tex3d.size() * sizeof(float));
  else ch = 0;
  delete t;
  vol.SetArrayData(nullptr, vis::DataStorageSize::UNKNOWN);
  return ch;
}

// ---- Cie2000Comparison of two 8-bit-range sRGB triplets (the "Generate Diff" button, renderingmanager.cpp:696) ------
double ref_cie2000(const double* rgb_a, const double* rgb_b) {
  double a[3] = {rgb_a[0], rgb_a[1], rgb_a[2]}, b[3] = {rgb_b[0], rgb_b[1], rgb_b[2]};
  return Cie2000Comparison(a, b);
}

// ---- DDSV3::readDDSfile (libs/file_utils/pvm.cpp:518-572): the reference's own "DDS v3d" / "DDS v3e" decoder run on a
// file; returns the unpacked size (bytes beyond cap are not copied), -1 when the reference rejects the file.
long long ref_dds_read(const char* path, unsigned char* out, unsigned long long cap) {
  DDSV3 loader;
  unsigned int bytes = 0;
  unsigned char* data = loader.readDDSfile(path, &bytes);
  if (!data) return -1;
  std::memcpy(out, data, bytes < cap ? bytes : cap);
  free(data);
  return (long long)bytes;
}

// ---- Pvm (pvm.cpp:22-109), what VolumeReader::readpvm (reader.cpp:100-160) consumes: dims, components, scale and the
// post-processed voxels (u8, or u16 assembled as byte1 * 256 + byte0).  Only call with files the reference accepts: it
// exit()s on malformed ones.
int ref_pvm_read(const char* path, unsigned int dims[3], double scale[3], void* out, unsigned long long cap_bytes) {
  Pvm f(path);
  f.GetDimensions(&dims[0], &dims[1], &dims[2]);
  f.GetScale(&scale[0], &scale[1], &scale[2]);
  int comp = f.GetComponents();
  unsigned long long n = (unsigned long long)dims[0] * dims[1] * dims[2] * (unsigned long long)comp;
  if ((comp != 1 && comp != 2) || n > cap_bytes || !f.GetData()) return -comp;
  std::memcpy(out, f.GetData(), n);
  return comp;
}

// ---- CameraStateList::ReadCameraStates (camerastatelist.cpp:26-87): 9 floats per state (eye, center, up).
int ref_read_camera_states(const char* path, float* out9, int cap) {
  vis::CameraStateList l;
  if (!l.ReadCameraStates(path)) return -1;
  int n = l.NumberOfCameraStates();
  for (int i = 0; i < n && i < cap; ++i) {
    vis::CameraData* c = l.GetCameraState(i);
    float* o = out9 + 9 * i;
    o[0] = c->eye.x; o[1] = c->eye.y; o[2] = c->eye.z; o[3] = c->center.x; o[4] = c->center.y; o[5] = c->center.z; o[6] = c->up.x; o[7] = c->up.y; o[8] = c->up.z;
  }
  return n;
}

// ---- LightSourceList::ReadLightSourceLists (lightsourcelist.cpp:81-148): 13 floats per light over all lists
// (position, -z_axis = forward, y_axis, x_axis, spot angle), the layout of the host mirror's vrbh_read_light_lists.
int ref_read_light_lists(const char* path, float* out13, int cap, int* n_lists) {
  vis::LightSourceList l;
  if (!l.ReadLightSourceLists(path)) return -1;
  int k = 0;
  *n_lists = l.NumberOfLists();
  for (int i = 0; i < l.NumberOfLists(); ++i)
    for (auto& s : l.GetList(i)->m_lightsources) {
      if (k < cap) {
        float* o = out13 + 13 * k;
        o[0] = s.position.x; o[1] = s.position.y; o[2] = s.position.z;
        o[3] = -s.z_axis.x; o[4] = -s.z_axis.y; o[5] = -s.z_axis.z;
        o[6] = s.y_axis.x; o[7] = s.y_axis.y; o[8] = s.y_axis.z;
        o[9] = s.x_axis.x; o[10] = s.x_axis.y; o[11] = s.x_axis.z; o[12] = s.spot_light_angle;
      }
      ++k;
    }
  return k;
}

// ---- glm::lookAt of the vendored glm 0.9.5.3 (include/glm/gtc/matrix_transform.inl:403-428), what Camera::LookAt returns
// (libs/vis_utils/camera.cpp:281-284); column-major 16 floats.
void ref_glm_look_at(const float* eye, const float* center, const float* up, float* out16) {
  glm::mat4 m = glm::lookAt(glm::vec3(eye[0], eye[1], eye[2]), glm::vec3(center[0], center[1], center[2]), glm::vec3(up[0], up[1], up[2]));
  std::memcpy(out16, &m[0][0], 16 * sizeof(float));
}

// ---- ParameterSpace (cppvolrend/utils/parameterspace.{h,cpp}): the reference's own self-test, and a sweep over numeric ranges
// written as the CSV rows renderingmanager.cpp's evaluation writes (one line per sample point, last dimension fastest).
// kind: 0 float, 1 double, 2 int.  Returns the number of points visited; num_sample_points = GetNumSamplePoints().
int ref_parameterspace_test(void) { return ParameterSpaceTest() ? 1 : 0; }
int ref_pspace_enumerate(const double* start_end_incr, const int* kind, int ndims, char* out, int cap, int* num_sample_points) {
  float fv[16]; double dv[16]; int iv[16];
  if (ndims < 1 || ndims > 16) return -1;
  ParameterSpace ps;
  for (int i = 0; i < ndims; ++i) {
    const double a = start_end_incr[3 * i], b = start_end_incr[3 * i + 1], c = start_end_incr[3 * i + 2];
    const std::string name = "p" + std::to_string(i);
    if (kind[i] == 0) ps.AddParameterDimension(new ParameterRangeFloat(name, &fv[i], (float)a, (float)b, (float)c));
    else if (kind[i] == 1) ps.AddParameterDimension(new ParameterRangeDouble(name, &dv[i], a, b, c));
    else ps.AddParameterDimension(new ParameterRangeInt(name, &iv[i], (int)a, (int)b, (int)c));
  }
  *num_sample_points = ps.GetNumSamplePoints();
  std::string s;
  int visited = 0;
  ps.StartEvaluation();
  do {
    for (int i = 0; i < ndims; ++i) { s += ps.GetDimensionValue(i); s += (i + 1 < ndims) ? "," : "\n"; }
    ++visited;
  } while (ps.IncrEvaluation() && visited < 100000);
  ps.EndEvaluation();
  if ((int)s.size() + 1 > cap) return -2;
  std::memcpy(out, s.c_str(), s.size() + 1);
  return visited;
}

}  // extern "C"
