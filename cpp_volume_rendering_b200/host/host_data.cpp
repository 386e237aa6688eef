// host_data.cpp -- GL-free re-host of the reference's data-side classes (see vrbhost.h for the file:line map).
#include "vrbhost.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <new>
#include <sstream>

namespace vrb {
static thread_local std::string g_err;
const std::string& LastError() { return g_err; }
void SetError(const std::string& s) { g_err = s; }

// glm::lookAt (include/glm/gtc/matrix_transform.inl:403-428)
mat4 lookAt(vec3 eye, vec3 center, vec3 up) {
  vec3 f = normalize(center - eye);
  vec3 s = normalize(cross(f, up));
  vec3 u = cross(s, f);
  mat4 r;
  for (int i = 0; i < 16; ++i) r.m[i] = 0.0f;
  r.m[15] = 1.0f;
  r.m[0] = s.x;  r.m[4] = s.y;  r.m[8] = s.z;
  r.m[1] = u.x;  r.m[5] = u.y;  r.m[9] = u.z;
  r.m[2] = -f.x; r.m[6] = -f.y; r.m[10] = -f.z;
  r.m[12] = -dot(s, eye); r.m[13] = -dot(u, eye); r.m[14] = dot(f, eye);
  return r;
}

Device* Device::Instance() { static Device d; return &d; }
bool Device::Init(int cuda_device) {
  if (m_ctx) return true;
  if (vrb_ctx_create(cuda_device, &m_ctx) != VRB_OK) { SetError(vrb_last_error()); m_ctx = nullptr; return false; }
  return true;
}
void Device::Shutdown() { if (m_ctx) vrb_ctx_destroy(m_ctx); m_ctx = nullptr; }
}  // namespace vrb

namespace vis {

// ---------------------------------------------------------------- transfer function (transferfunction{,1d}.cpp)
TransferControlPoint::TransferControlPoint(double r, double g, double b, int isovalue) {
  m_color.x = (float)r; m_color.y = (float)g; m_color.z = (float)b; m_color.w = 1.0f; m_isoValue = isovalue;
}
TransferControlPoint::TransferControlPoint(double alpha, int isovalue) {
  m_color.x = m_color.y = m_color.z = 0.0f; m_color.w = (float)alpha; m_isoValue = isovalue;
}

TransferFunction1D::TransferFunction1D(int max_value) : m_built(false), max_density(max_value), extinction_coef_type(false) {}
TransferFunction1D::~TransferFunction1D() {}
const char* TransferFunction1D::GetNameClass() { return "TrasnferFunction1D"; }
void TransferFunction1D::SetExtinctionCoefficientInput(bool s) { extinction_coef_type = s; }
void TransferFunction1D::AddRGBControlPoint(TransferControlPoint rgb) { m_cpt_rgb.push_back(rgb); m_built = false; }
void TransferFunction1D::AddAlphaControlPoint(TransferControlPoint alpha) { m_cpt_alpha.push_back(alpha); m_built = false; }
void TransferFunction1D::ClearControlPoints() { m_cpt_rgb.clear(); m_cpt_alpha.clear(); m_built = false; }

void TransferFunction1D::Build() {
  // entries no segment covers are left uninitialised by the reference (transferfunction1d.cpp:124); zero here
  m_transferfunction.assign((size_t)max_density + 1, dvec4());
  BuildLinear();
  m_built = true;
}

// piecewise-linear fill, inclusive at both ends of every segment, later segments overwrite shared end points;
// colour differences in float, interpolation in double (transferfunction1d.cpp:319-358)
void TransferFunction1D::BuildLinear() {
  for (size_t seg = 0; seg + 1 < m_cpt_rgb.size(); ++seg) {
    const TransferControlPoint& p = m_cpt_rgb[seg];
    const TransferControlPoint& q = m_cpt_rgb[seg + 1];
    double dr = (double)(float)(q.m_color.x - p.m_color.x);
    double dg = (double)(float)(q.m_color.y - p.m_color.y);
    double db = (double)(float)(q.m_color.z - p.m_color.z);
    for (int x = p.m_isoValue; x <= q.m_isoValue; ++x) {
      if (x < 0 || x > max_density) continue;
      double k = (double)(x - p.m_isoValue) / (double)(q.m_isoValue - p.m_isoValue);
      m_transferfunction[x].r = p.m_color.x + dr * k;
      m_transferfunction[x].g = p.m_color.y + dg * k;
      m_transferfunction[x].b = p.m_color.z + db * k;
    }
  }
  for (size_t seg = 0; seg + 1 < m_cpt_alpha.size(); ++seg) {
    const TransferControlPoint& p = m_cpt_alpha[seg];
    const TransferControlPoint& q = m_cpt_alpha[seg + 1];
    double da = (double)(float)(q.m_color.w - p.m_color.w);
    for (int x = p.m_isoValue; x <= q.m_isoValue; ++x) {
      if (x < 0 || x > max_density) continue;
      double k = (double)(x - p.m_isoValue) / (double)(q.m_isoValue - p.m_isoValue);
      m_transferfunction[x].a = p.m_color.w + da * k;
    }
  }
}

vec4 TransferFunction1D::Get(double value, double max_data_value) {
  if (!m_built) Build();
  vec4 out;
  if (max_data_value >= 0) value = value * ((double)max_density / max_data_value);
  if (value < 0.0f || value > (float)max_density) return out;
  dvec4 v;
  if (std::fabs(value - (float)max_density) < 0.000001) {
    v = m_transferfunction[max_density];
  } else {
    int iv = (int)value;
    double t = value - iv;
    const dvec4& a = m_transferfunction[iv];
    const dvec4& b = m_transferfunction[iv + 1];
    v.r = (1.0 - t) * a.r + t * b.r; v.g = (1.0 - t) * a.g + t * b.g;
    v.b = (1.0 - t) * a.b + t * b.b; v.a = (1.0 - t) * a.a + t * b.a;
  }
  out.x = (float)v.r; out.y = (float)v.g; out.z = (float)v.b; out.w = (float)v.a;
  return out;
}
float TransferFunction1D::GetOpc(double value, double max_input_value) {
  float val = Get(value, max_input_value).w;
  return extinction_coef_type ? (float)ExtinctionToMaterialOpacity(val) : val;
}
float TransferFunction1D::GetOpcN(double n) {
  float val = Get(n, 1.0).w;
  return extinction_coef_type ? (float)ExtinctionToMaterialOpacity(val) : val;
}
float TransferFunction1D::GetExt(double value, double max_input_value) {
  float val = Get(value, max_input_value).w;
  return !extinction_coef_type ? (float)MaterialOpacityToExtinction(val) : val;
}
float TransferFunction1D::GetExtN(double n) {
  float val = Get(n, 1.0).w;
  return !extinction_coef_type ? (float)MaterialOpacityToExtinction(val) : val;
}
bool TransferFunction1D::GenerateTexture_1D_RGBA(std::vector<float>& out) {
  if (!m_built) Build();
  int n = max_density + 1;
  out.resize((size_t)n * 4);
  for (int i = 0; i < n; ++i) {
    out[4 * i + 0] = (float)m_transferfunction[i].r;
    out[4 * i + 1] = (float)m_transferfunction[i].g;
    out[4 * i + 2] = (float)m_transferfunction[i].b;
    float v4 = (float)m_transferfunction[i].a;
    if (extinction_coef_type) v4 = (float)ExtinctionToMaterialOpacity(v4);
    out[4 * i + 3] = v4;
  }
  return true;
}
bool TransferFunction1D::GenerateTexture_1D_RGBt(std::vector<float>& out) {
  if (!m_built) Build();
  int n = max_density + 1;
  out.resize((size_t)n * 4);
  for (int i = 0; i < n; ++i) {
    out[4 * i + 0] = (float)m_transferfunction[i].r;
    out[4 * i + 1] = (float)m_transferfunction[i].g;
    out[4 * i + 2] = (float)m_transferfunction[i].b;
    float v4 = (float)m_transferfunction[i].a;
    if (!extinction_coef_type) v4 = (float)MaterialOpacityToExtinction(v4);
    out[4 * i + 3] = v4;
  }
  return true;
}

TransferFunction* TransferFunctionReader::ReadTransferFunction(std::string file) {
  size_t dot = file.find_last_of('.');
  if (dot != std::string::npos && file.substr(dot + 1) == "tf1d") return readtf1d(file);
  vrb::SetError("ReadTransferFunction: unsupported extension: " + file);
  return nullptr;
}

// .tf1d (transferfunction1d.h:6-19, reader.cpp:744-814)
TransferFunction* TransferFunctionReader::readtf1d(std::string file) {
  std::ifstream in(file);
  if (!in.is_open()) { vrb::SetError("readtf1d: cannot open " + file); return nullptr; }
  std::string interpolation;
  std::getline(in, interpolation);          // always treated as linear
  int init = 0;
  in >> init;
  TransferFunction1D* tf = nullptr;
  int maxd = 255, extuse = 0;
  if (init == 2) in >> maxd >> extuse;
  else if (init == 1) in >> maxd;
  // the table has max_density + 1 entries: 8- and 16-bit data need 255 / 65535; anything else is a damaged file
  if (in.fail() || maxd < 1 || maxd > 65535) { vrb::SetError("readtf1d: malformed header in " + file); return nullptr; }
  tf = (init == 1 || init == 2) ? new TransferFunction1D(maxd) : new TransferFunction1D();
  if (init == 2) tf->SetExtinctionCoefficientInput(extuse == 1);
  int n = 0;
  in >> n;
  for (int i = 0; i < n && !in.fail(); ++i) {
    double r, g, b; int iso;
    in >> r >> g >> b >> iso;
    if (!in.fail()) tf->AddRGBControlPoint(TransferControlPoint(r, g, b, iso));
  }
  n = 0;
  in >> n;
  for (int i = 0; i < n && !in.fail(); ++i) {
    double a; int iso;
    in >> a >> iso;
    if (!in.fail()) tf->AddAlphaControlPoint(TransferControlPoint(a, iso));
  }
  if (in.fail()) { delete tf; vrb::SetError("readtf1d: malformed file " + file); return nullptr; }
  tf->SetName(file);
  return tf;
}

// ---------------------------------------------------------------- structured volume (structuredgridvolume.cpp)
StructuredGridVolume::StructuredGridVolume(std::string name, unsigned int w, unsigned int h, unsigned int d)
    : m_name(name), m_width(w), m_height(h), m_depth(d), m_scalex(1.0), m_scaley(1.0), m_scalez(1.0),
      m_data_storage_size(UNKNOWN), m_voxel_values(nullptr) {}
StructuredGridVolume::~StructuredGridVolume() {
  if (m_data_storage_size == _8_BITS) delete[] static_cast<unsigned char*>(m_voxel_values);
  else if (m_data_storage_size == _16_BITS) delete[] static_cast<unsigned short*>(m_voxel_values);
}
double StructuredGridVolume::GetDiagonal() {
  double a = GetWidth() * GetScaleX(), b = GetHeight() * GetScaleY(), c = GetDepth() * GetScaleZ();
  return std::sqrt(a * a + b * b + c * c);
}
bool StructuredGridVolume::IsOutOfBoundary(int x, int y, int z) {
  return x < 0 || y < 0 || z < 0 || x >= (int)m_width || y >= (int)m_height || z >= (int)m_depth;
}
void StructuredGridVolume::SetArrayData(void* p, DataStorageSize dss) { m_data_storage_size = dss; m_voxel_values = p; }
double StructuredGridVolume::GetNormalizedSample(int x, int y, int z) {
  if (!m_voxel_values || m_data_storage_size == UNKNOWN || IsOutOfBoundary(x, y, z)) return 0.0;
  size_t id = (size_t)x + (size_t)y * m_width + (size_t)z * m_width * m_height;
  if (m_data_storage_size == _8_BITS) return (double)static_cast<unsigned char*>(m_voxel_values)[id] / (256.0 - 1.0);
  return (double)static_cast<unsigned short*>(m_voxel_values)[id] / (65536.0 - 1.0);
}
unsigned long long StructuredGridVolume::CheckSum() {
  unsigned long long s = 0;
  size_t n = (size_t)m_width * m_height * m_depth;
  if (m_data_storage_size == _8_BITS) for (size_t i = 0; i < n; ++i) s += static_cast<unsigned char*>(m_voxel_values)[i];
  else if (m_data_storage_size == _16_BITS) for (size_t i = 0; i < n; ++i) s += static_cast<unsigned short*>(m_voxel_values)[i];
  return s;
}
double StructuredGridVolume::GetMaxDensity() {
  if (m_data_storage_size == _8_BITS) return 256.0 - 1.0;
  if (m_data_storage_size == _16_BITS) return 65536.0 - 1.0;
  return 0.0;
}

// ---------------------------------------------------------------- readers (reader.cpp:28-60,100-371)
// Sizes come from file names and headers: they are checked before anything is allocated (the reference allocates first and
// crashes on a malformed size).  16384 per axis / 2^36 voxels is far beyond what one GPU holds (2048^3 u16 = 2^34 bytes).
bool vrb_volume_dims_ok(long long w, long long h, long long d) {
  const long long kAxis = 16384, kVoxels = 1ll << 36;
  return w > 0 && h > 0 && d > 0 && w <= kAxis && h <= kAxis && d <= kAxis && w * h * d <= kVoxels;
}
static std::string ext_of(const std::string& f) {
  size_t dot = f.find_last_of('.');
  return dot == std::string::npos ? std::string() : f.substr(dot + 1);
}
StructuredGridVolume* VolumeReader::ReadStructuredVolume(std::string filepath) {
  std::string e = ext_of(filepath);
  if (e == "raw") return readraw(filepath);
  if (e == "syn") return readsyn(filepath);
  if (e == "pvm") return readpvm(filepath);
  vrb::SetError("ReadStructuredVolume: unsupported extension: " + filepath);
  return nullptr;
}

// Name.<bytesPerVoxel>.<W>x<H>x<D>.raw, parsed from the right (reader.cpp:172-205); x fastest; scale 1.
StructuredGridVolume* VolumeReader::readraw(std::string filepath) {
  size_t slash = filepath.find_last_of("/\\");
  std::string name = slash == std::string::npos ? filepath : filepath.substr(slash + 1);
  std::vector<std::string> parts;
  { std::stringstream ss(name); std::string p; while (std::getline(ss, p, '.')) parts.push_back(p); }
  if (parts.size() < 4) { vrb::SetError("readraw: name must be <name>.<bytes>.<W>x<H>x<D>.raw: " + filepath); return nullptr; }
  const std::string& dims = parts[parts.size() - 2];
  int bytes = atoi(parts[parts.size() - 3].c_str());
  int w = 0, h = 0, d = 0;
  if (sscanf(dims.c_str(), "%dx%dx%d", &w, &h, &d) != 3 || w <= 0 || h <= 0 || d <= 0 || (bytes != 1 && bytes != 2)) {
    vrb::SetError("readraw: cannot parse sizes from " + filepath); return nullptr;
  }
  if (!vrb_volume_dims_ok(w, h, d)) { vrb::SetError("readraw: implausible sizes in " + filepath); return nullptr; }
  std::ifstream in(filepath, std::ios::binary);
  if (!in.is_open()) { vrb::SetError("readraw: cannot open " + filepath); return nullptr; }
  size_t n = (size_t)w * h * d;
  in.seekg(0, std::ios::end);
  const std::streamoff file_bytes = in.tellg();
  in.seekg(0, std::ios::beg);
  if (file_bytes < 0 || (unsigned long long)file_bytes < (unsigned long long)n * bytes) {      // before allocating anything
    vrb::SetError("readraw: file shorter than WxHxDxbytes: " + filepath); return nullptr;
  }
  void* data = nullptr;
  try { data = bytes == 1 ? (void*)new unsigned char[n] : (void*)new unsigned short[n]; }
  catch (const std::bad_alloc&) { vrb::SetError("readraw: out of memory for " + filepath); return nullptr; }
  in.read((char*)data, (std::streamsize)(n * bytes));
  if ((size_t)in.gcount() != n * bytes) {
    if (bytes == 1) delete[] (unsigned char*)data; else delete[] (unsigned short*)data;
    vrb::SetError("readraw: file shorter than WxHxDxbytes: " + filepath); return nullptr;
  }
  StructuredGridVolume* v = new StructuredGridVolume(filepath, w, h, d);
  v->SetScale(1.0, 1.0, 1.0);
  v->SetArrayData(data, bytes == 1 ? _8_BITS : _16_BITS);
  return v;
}

// .syn: "W H D" then records "1 x0 y0 z0 x1 y1 z1 v" (half-open box) or "<other> x y z v" (reader.cpp:283-371).
// The reference leaves the buffer uninitialised (:301); zeroed here.
StructuredGridVolume* VolumeReader::readsyn(std::string filepath) {
  std::ifstream in(filepath);
  if (!in.is_open()) { vrb::SetError("readsyn: cannot open " + filepath); return nullptr; }
  int w = 0, h = 0, d = 0;
  in >> w >> h >> d;
  if (in.fail() || w <= 0 || h <= 0 || d <= 0) { vrb::SetError("readsyn: bad header in " + filepath); return nullptr; }
  // a .syn file carries no payload to check the header against: the generator's volumes are small (utils.cpp:373-396)
  if (!vrb_volume_dims_ok(w, h, d) || (long long)w * h * d > (1ll << 33)) { vrb::SetError("readsyn: implausible sizes in " + filepath); return nullptr; }
  size_t n = (size_t)w * h * d;
  unsigned char* data = nullptr;
  try { data = new unsigned char[n]; }
  catch (const std::bad_alloc&) { vrb::SetError("readsyn: out of memory for " + filepath); return nullptr; }
  std::memset(data, 0, n);
  int tag = 0;
  while (in >> tag) {
    if (tag == 1) {
      int x0, y0, z0, x1, y1, z1, v;
      in >> x0 >> y0 >> z0 >> x1 >> y1 >> z1 >> v;
      if (in.fail()) break;
      x0 = std::max(x0, 0); y0 = std::max(y0, 0); z0 = std::max(z0, 0);
      x1 = std::min(x1, w); y1 = std::min(y1, h); z1 = std::min(z1, d);
      for (int z = z0; z < z1; ++z)
        for (int y = y0; y < y1; ++y)
          for (int x = x0; x < x1; ++x) data[(size_t)x + (size_t)w * y + (size_t)w * h * z] = (unsigned char)v;
    } else {
      int x, y, z, v;
      in >> x >> y >> z >> v;
      if (in.fail()) break;
      if (x >= 0 && y >= 0 && z >= 0 && x < w && y < h && z < d) data[(size_t)x + (size_t)w * y + (size_t)w * h * z] = (unsigned char)v;
    }
  }
  StructuredGridVolume* vol = new StructuredGridVolume(filepath, w, h, d);
  vol->SetScale(1.0, 1.0, 1.0);
  vol->SetArrayData(data, _8_BITS);
  return vol;
}

// VolumeReader::readpvm (plain and DDS-compressed .pvm) lives in host_pvm.cpp.

// ---------------------------------------------------------------- camera (libs/vis_utils/camera.cpp)
CameraData::CameraData() : c_type(0), field_of_view_y(45.0f), aspect_ratio(1.0f), z_near(1.0f), z_far(5000.0f) {}
Camera::Camera() : radius(200.0f) {
  c_data.center = vec3(0, 0, 0); c_data.eye = vec3(0, 0, radius); c_data.up = vec3(0, 1, 0);
}
mat4 Camera::LookAt() { return vrb::lookAt(c_data.eye, c_data.center, c_data.up); }
vec3 Camera::GetDir() { return vrb::normalize(c_data.center - c_data.eye); }
vec3 Camera::GetEye() { return c_data.eye; }
void Camera::UpdateAspectRatio(float w, float h) { c_data.aspect_ratio = w / h; }
float Camera::GetAspectRatio() { return c_data.aspect_ratio; }
float Camera::GetFovY() { return c_data.field_of_view_y; }
float Camera::GetTanFovY() { return (float)std::tan(((double)GetFovY() * (3.14159265358979323846264338327950288 / 180.0)) / 2.0); }
void Camera::SetData(CameraData* data) {
  c_data.eye = data->eye; c_data.center = data->center; c_data.up = data->up;
  vec3 dd = c_data.eye - c_data.center;
  radius = std::sqrt(vrb::dot(dd, dd));
}
void Camera::GetCameraVectors(vec3* cforward, vec3* cup, vec3* cright) {
  *cforward = -GetDir();
  *cright = vrb::normalize(vrb::cross(c_data.up, *cforward));
  *cup = vrb::normalize(vrb::cross(*cforward, *cright));
}

// "#list_camera_states": name line, ARCBALL|FLIGHT, eye / center / up triples (camerastatelist.cpp:26-87)
bool CameraStateList::ReadCameraStates(std::string filepath) {
  m_vec_camera_data.clear();
  std::ifstream in(filepath);
  if (!in.is_open()) { vrb::SetError("ReadCameraStates: cannot open " + filepath); return false; }
  std::string line;
  while (std::getline(in, line)) {
    if (line.find_first_not_of(" \t\r\n") == std::string::npos) continue;
    CameraData cd;
    while (!line.empty() && line.back() == '\r') line.pop_back();
    cd.cam_setup_name = line;
    std::string kind;
    if (!std::getline(in, kind)) break;
    while (!kind.empty() && (kind.back() == '\r' || kind.back() == ' ')) kind.pop_back();
    if (kind == "FLIGHT") {
      cd.c_type = Camera::FLIGHT;
    } else if (kind == "ARCBALL") {
      cd.c_type = Camera::ARCBALL;
      in >> cd.eye.x >> cd.eye.y >> cd.eye.z >> cd.center.x >> cd.center.y >> cd.center.z >> cd.up.x >> cd.up.y >> cd.up.z;
      if (in.fail()) { vrb::SetError("ReadCameraStates: malformed state '" + cd.cam_setup_name + "'"); return false; }
      std::getline(in, line);
    }
    m_vec_camera_data.push_back(cd);
  }
  return !m_vec_camera_data.empty();
}
int CameraStateList::NumberOfCameraStates() { return (int)m_vec_camera_data.size(); }
CameraData* CameraStateList::GetCameraState(unsigned int idx) { return idx < m_vec_camera_data.size() ? &m_vec_camera_data[idx] : nullptr; }

// "#list_light_sources": name line, count, per light position / forward / up / right + spot angle (degrees);
// stored z_axis = -forward (lightsourcelist.cpp:81-148)
LightSourceData::LightSourceData()
    : color(1.0f), specular(1.0f), position(0.0f), x_axis(1, 0, 0), y_axis(0, 1, 0), z_axis(0, 0, 1),
      spot_light_angle(4.0f), spot_light_angle_rad(4.0f * 3.14159265358979323846f / 180.0f), energy_density(1.0f) {}
bool LightSourceList::ReadLightSourceLists(std::string filepath) {
  m_vec_lsource_lists.clear();
  std::ifstream in(filepath);
  if (!in.is_open()) { vrb::SetError("ReadLightSourceLists: cannot open " + filepath); return false; }
  std::string line;
  while (std::getline(in, line)) {
    if (line.find_first_not_of(" \t\r\n") == std::string::npos) continue;
    while (!line.empty() && line.back() == '\r') line.pop_back();
    LightSourceListItem item;
    item.l_name = line;
    int n = 0;
    in >> n;
    if (in.fail()) { vrb::SetError("ReadLightSourceLists: malformed list '" + item.l_name + "'"); return false; }
    for (int i = 0; i < n; ++i) {
      LightSourceData l;
      vec3 fwd;
      in >> l.position.x >> l.position.y >> l.position.z >> fwd.x >> fwd.y >> fwd.z >> l.y_axis.x >> l.y_axis.y >> l.y_axis.z >>
          l.x_axis.x >> l.x_axis.y >> l.x_axis.z >> l.spot_light_angle;
      if (in.fail()) { vrb::SetError("ReadLightSourceLists: malformed light in '" + item.l_name + "'"); return false; }
      l.z_axis = -fwd;
      l.spot_light_angle_rad = (l.spot_light_angle * 3.14159265358979323846f) / 180.0f;
      std::getline(in, line);
      item.m_lightsources.push_back(l);
    }
    m_vec_lsource_lists.push_back(item);
  }
  return !m_vec_lsource_lists.empty();
}
int LightSourceList::NumberOfLists() { return (int)m_vec_lsource_lists.size(); }
LightSourceListItem* LightSourceList::GetList(unsigned int idx) { return idx < m_vec_lsource_lists.size() ? &m_vec_lsource_lists[idx] : nullptr; }

// ---------------------------------------------------------------- rendering parameters (renderingparameters.cpp)
RenderingParameters::RenderingParameters()
    : screen_width(768), screen_height(768), m_blinnphong_ka(0.5f), m_blinnphong_kd(0.5f), m_blinnphong_ks(0.8f),
      m_blinnphong_shininess(30.0f), m_current_light_source_id(0) {
  m_vec_light_sources.push_back(LightSourceData());
}
LightSourceData& RenderingParameters::cur() {
  if (m_vec_light_sources.empty()) m_vec_light_sources.push_back(LightSourceData());
  if (m_current_light_source_id >= (int)m_vec_light_sources.size()) m_current_light_source_id = 0;
  return m_vec_light_sources[m_current_light_source_id];
}
void RenderingParameters::SetPhongParameters(float amb, float diff, float spec, float shini) {
  m_blinnphong_ka = amb; m_blinnphong_kd = diff; m_blinnphong_ks = spec; m_blinnphong_shininess = shini;
}
vec3 RenderingParameters::GetLightSourceSpecular() { return cur().specular; }
void RenderingParameters::SetBlinnPhongLightingPosition(vec3 p) { cur().position = p; }
vec3 RenderingParameters::GetBlinnPhongLightingPosition() { return cur().position; }
void RenderingParameters::SetBlinnPhongLightSourceCameraVectors(vec3 f, vec3 u, vec3 r) { cur().z_axis = -f; cur().y_axis = u; cur().x_axis = r; }
vec3 RenderingParameters::GetBlinnPhongLightSourceCameraForward() { return -cur().z_axis; }
vec3 RenderingParameters::GetBlinnPhongLightSourceCameraUp() { return cur().y_axis; }
vec3 RenderingParameters::GetBlinnPhongLightSourceCameraRight() { return cur().x_axis; }
float RenderingParameters::GetSpotLightMaxAngle() { return cur().spot_light_angle; }
void RenderingParameters::SetScreenSize(int w, int h) { screen_width = w; screen_height = h; }
vrb_lighting RenderingParameters::MakeLightingBlock() {
  vrb_lighting L;
  L.ka = m_blinnphong_ka; L.kd = m_blinnphong_kd; L.ks = m_blinnphong_ks; L.shininess = m_blinnphong_shininess;
  vec3 s = GetLightSourceSpecular(), p = GetBlinnPhongLightingPosition();
  vec3 f = GetBlinnPhongLightSourceCameraForward(), u = GetBlinnPhongLightSourceCameraUp(), r = GetBlinnPhongLightSourceCameraRight();
  L.ispecular[0] = s.x; L.ispecular[1] = s.y; L.ispecular[2] = s.z;
  L.light_pos[0] = p.x; L.light_pos[1] = p.y; L.light_pos[2] = p.z;
  L.light_forward[0] = f.x; L.light_forward[1] = f.y; L.light_forward[2] = f.z;
  L.light_up[0] = u.x; L.light_up[1] = u.y; L.light_up[2] = u.z;
  L.light_right[0] = r.x; L.light_right[1] = r.y; L.light_right[2] = r.z;
  L.spot_angle_deg = GetSpotLightMaxAngle();
  return L;
}

// ---------------------------------------------------------------- data manager (datamanager.cpp:69-101,232-330)
DataManager::DataManager() : curr_gradient_comp_model(NONE_GRADIENT), curr_volume_index(0), curr_transferfunction_index(0),
                             curr_vr_volume(nullptr), curr_vr_transferfunction(nullptr) {}
DataManager::~DataManager() { delete curr_vr_volume; delete curr_vr_transferfunction; }

// one "<relative path> <display name>" per line
bool DataManager::ReadList(const char* list_name, std::vector<DataReference>& out) {
  out.clear();
  std::string fn = m_path_to_data + "/" + list_name;
  std::ifstream in(fn);
  if (!in.is_open()) { vrb::SetError(std::string("DataManager: cannot open ") + fn); return false; }
  std::string line;
  while (std::getline(in, line)) {
    size_t a = line.find_first_of('<'), b = line.find_first_of('>');
    size_t c = line.find_last_of('<'), d = line.find_last_of('>');
    if (a == std::string::npos || b == std::string::npos || b <= a) continue;
    DataReference r;
    std::string rel = line.substr(a + 1, b - a - 1);
    r.path = m_path_to_data + "/" + rel;
    r.name = (c != a && d != std::string::npos && d > c) ? line.substr(c + 1, d - c - 1) : rel;
    out.push_back(r);
  }
  return true;
}

bool DataManager::ReadData() {
  if (!ReadList("#list_structured_datasets", stored_structured_datasets)) return false;
  if (!ReadList("#list_transfer_functions", stored_transfer_functions)) return false;
  if (stored_structured_datasets.empty() || stored_transfer_functions.empty()) { vrb::SetError("DataManager: empty list file"); return false; }
  curr_volume_index = 0; curr_transferfunction_index = 0;
  if (!GenerateStructuredVolumeTexture()) return false;
  return SetCurrentTransferFunction(0);
}

bool DataManager::GenerateStructuredVolumeTexture() {
  VolumeReader vr;
  StructuredGridVolume* v = vr.ReadStructuredVolume(stored_structured_datasets[curr_volume_index].path);
  if (!v) return false;
  v->SetName(stored_structured_datasets[curr_volume_index].name);
  return SetStructuredVolume(v);
}

bool DataManager::SetCurrentInputVolume(int id) {
  if (id < 0 || id >= (int)stored_structured_datasets.size()) { vrb::SetError("SetCurrentInputVolume: bad index"); return false; }
  curr_volume_index = id;
  return GenerateStructuredVolumeTexture();
}

bool DataManager::SetCurrentTransferFunction(int id) {
  if (id < 0 || id >= (int)stored_transfer_functions.size()) { vrb::SetError("SetCurrentTransferFunction: bad index"); return false; }
  TransferFunctionReader tfr;
  TransferFunction* tf = tfr.ReadTransferFunction(stored_transfer_functions[id].path);
  if (!tf) return false;
  tf->SetName(stored_transfer_functions[id].name);
  curr_transferfunction_index = id;
  return SetTransferFunction(tf);
}

// vis::GenerateRTexture (libs/volvis_utils/utils.cpp:20-56) -> vrb_volume_upload
bool DataManager::SetStructuredVolume(StructuredGridVolume* vol) {
  if (!vol || !vol->GetArrayData()) { vrb::SetError("SetStructuredVolume: no voxel data"); return false; }
  vrb::Device* dev = vrb::Device::Instance();
  if (!dev->ok()) { vrb::SetError("SetStructuredVolume: device not initialised (RenderingManager::InitGL)"); return false; }
  float scale[3] = {(float)vol->GetScaleX(), (float)vol->GetScaleY(), (float)vol->GetScaleZ()};
  int bpv = vol->GetDataStorageSize() == _8_BITS ? 1 : 2;
  if (vrb_volume_upload(dev->ctx(), vol->GetArrayData(), (int)vol->GetWidth(), (int)vol->GetHeight(), (int)vol->GetDepth(), bpv, scale) != VRB_OK) {
    vrb::SetError(vrb_last_error());
    return false;
  }
  if (curr_vr_volume != vol) delete curr_vr_volume;
  curr_vr_volume = vol;
  curr_tex_volume.w = (int)vol->GetWidth(); curr_tex_volume.h = (int)vol->GetHeight(); curr_tex_volume.d = (int)vol->GetDepth();
  // "Generate gradient, if enabled" (datamanager.cpp:326-327); the upload dropped the previous volume's gradient
  curr_tex_gradient = vrb::DeviceGradientTexture();
  return GenerateStructuredGradientTexture();
}

// DataManager::GenerateStructuredGradientTexture (datamanager.cpp:332-352): the three generators run on the device
bool DataManager::GenerateStructuredGradientTexture() {
  curr_tex_gradient = vrb::DeviceGradientTexture();
  vrb::Device* dev = vrb::Device::Instance();
  if (!dev->ok() || !curr_vr_volume) return false;
  int mode = VRB_GRADIENT_NONE;
  if (curr_gradient_comp_model == SOBEL_FELDMAN_FILTER) mode = VRB_GRADIENT_SOBEL_FELDMAN;
  else if (curr_gradient_comp_model == FINITE_DIFERENCES) mode = VRB_GRADIENT_FINITE_DIFFERENCES;
  else if (curr_gradient_comp_model == COMPUTE_SHADER_SOBEL) mode = VRB_GRADIENT_COMPUTE_SHADER_SOBEL;
  if (vrb_gradient_build(dev->ctx(), mode) != VRB_OK) { vrb::SetError(vrb_last_error()); return false; }
  if (mode != VRB_GRADIENT_NONE) {
    curr_tex_gradient.w = (int)curr_vr_volume->GetWidth(); curr_tex_gradient.h = (int)curr_vr_volume->GetHeight(); curr_tex_gradient.d = (int)curr_vr_volume->GetDepth();
  }
  return true;
}
void DataManager::DeleteGradientData() {
  vrb::Device* dev = vrb::Device::Instance();
  if (dev->ok()) vrb_gradient_build(dev->ctx(), VRB_GRADIENT_NONE);
  curr_tex_gradient = vrb::DeviceGradientTexture();
}
bool DataManager::UpdateStructuredGradientTexture() { DeleteGradientData(); return GenerateStructuredGradientTexture(); }
bool DataManager::SetCurrentGradient(int idx) {                 // datamanager.cpp:545-562: true when the model changed
  STRUCTURED_GRADIENT_TYPE sgt = NONE_GRADIENT;
  if (idx == 0) sgt = SOBEL_FELDMAN_FILTER; else if (idx == 1) sgt = FINITE_DIFERENCES; else if (idx == 2) sgt = COMPUTE_SHADER_SOBEL;
  bool ret = !(sgt == curr_gradient_comp_model);
  if (ret) curr_gradient_comp_model = sgt;
  return ret;
}
std::string DataManager::GetGradientName(STRUCTURED_GRADIENT_TYPE sgt) {
  if (sgt == SOBEL_FELDMAN_FILTER) return "Sobel-Feldman";
  if (sgt == FINITE_DIFERENCES) return "Finite Diferences";
  if (sgt == COMPUTE_SHADER_SOBEL) return "Sobel-Feldman (Compute Shader)";
  return "None";
}
std::string DataManager::CurrentGradientName() { return curr_gradient_comp_model == NONE_GRADIENT ? std::string("NULL") : GetGradientName(curr_gradient_comp_model); }
std::vector<std::string> DataManager::GetGradientGenerationTypeStrList() {
  return {GetGradientName(SOBEL_FELDMAN_FILTER), GetGradientName(FINITE_DIFERENCES), GetGradientName(COMPUTE_SHADER_SOBEL), GetGradientName(NONE_GRADIENT)};
}

bool DataManager::SetTransferFunction(TransferFunction* tf) {
  if (!tf) { vrb::SetError("SetTransferFunction: NULL"); return false; }
  if (curr_vr_transferfunction != tf) delete curr_vr_transferfunction;
  curr_vr_transferfunction = tf;
  return true;
}

std::string DataManager::GetCurrentDataName() { return curr_vr_volume ? curr_vr_volume->GetName() : std::string(); }
std::string DataManager::GetCurrentTransferFunctionName() { return curr_vr_transferfunction ? curr_vr_transferfunction->GetName() : std::string(); }

// ---------------------------------------------------------------- output frame (renderoutputframe.cpp:64-87,187-202)
bool RenderFrameToScreen::UpdateScreenResolution(int s_w, int s_h) {
  vrb::Device* dev = vrb::Device::Instance();
  if (!dev->ok()) { vrb::SetError("UpdateScreenResolution: device not initialised"); return false; }
  if (vrb_frame_resize(dev->ctx(), s_w, s_h) != VRB_OK) { vrb::SetError(vrb_last_error()); return false; }
  m_w = s_w; m_h = s_h; m_sw = s_w; m_sh = s_h; m_filtered = false;
  return true;
}
bool RenderFrameToScreen::UpdateScreenResolutionMultiScaling(int s_w, int s_h, int mw, int mh) {
  vrb::Device* dev = vrb::Device::Instance();
  if (!dev->ok()) { vrb::SetError("UpdateScreenResolutionMultiScaling: device not initialised"); return false; }
  if (mw != 0 && mh != 0) { m_mw = mw; m_mh = mh; } else { mw = m_mw; mh = m_mh; }
  if (mw == 0 || mh == 0) { vrb::SetError("UpdateScreenResolutionMultiScaling: no multipliers set"); return false; }
  if (vrb_frame_resize_multiscaling(dev->ctx(), s_w, s_h, mw, mh) != VRB_OK) { vrb::SetError(vrb_last_error()); return false; }
  m_w = mw < 0 ? s_w / std::abs(mw) : s_w * mw; m_h = mh < 0 ? s_h / std::abs(mh) : s_h * mh;
  m_sw = s_w; m_sh = s_h; m_filtered = true;
  return true;
}
static bool FramePass(int pass, unsigned int kernel) {
  if (vrb_frame_filter(vrb::Device::Instance()->ctx(), pass, (int)kernel) != VRB_OK) { vrb::SetError(vrb_last_error()); return false; }
  return true;
}
bool RenderFrameToScreen::DrawMultiSampleHigherResolutionMode() { return m_filtered ? FramePass(VRB_FILTER_PASS_MULTISAMPLE, m_kernel_filter) : true; }
bool RenderFrameToScreen::DrawHigherResolutionWithDownScale() { return m_filtered ? FramePass(VRB_FILTER_PASS_DOWNSCALE, m_kernel_filter) : true; }
bool RenderFrameToScreen::DrawLowerResolutionWithUpScale() { return m_filtered ? FramePass(VRB_FILTER_PASS_UPSCALE, m_kernel_filter) : true; }
bool RenderFrameToScreen::ClearTexture() {
  if (vrb_frame_clear(vrb::Device::Instance()->ctx()) != VRB_OK) { vrb::SetError(vrb_last_error()); return false; }
  return true;
}
bool RenderFrameToScreen::ReadPixelsRGBA32F(std::vector<float>& out) {
  if (m_filtered) {
    out.resize((size_t)m_sw * m_sh * 4);
    if (vrb_filtered_frame_read_rgba32f(vrb::Device::Instance()->ctx(), out.data()) != VRB_OK) { vrb::SetError(vrb_last_error()); return false; }
    return true;
  }
  out.resize((size_t)m_w * m_h * 4);
  if (vrb_frame_read_rgba32f(vrb::Device::Instance()->ctx(), out.data()) != VRB_OK) { vrb::SetError(vrb_last_error()); return false; }
  return true;
}

}  // namespace vis
