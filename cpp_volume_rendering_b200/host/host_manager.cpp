// host_manager.cpp -- headless RenderingManager (cppvolrend/renderingmanager.cpp, renderer-facing half) and the small
// extern "C" surface (vrbh_*) the Python tests / bench use to drive the C++ host exactly like the reference's main().
#include "vrbhost.h"
#include <cstring>

RenderingManager* RenderingManager::crr_instance = nullptr;
RenderingManager* RenderingManager::Instance() { if (!crr_instance) crr_instance = new RenderingManager(); return crr_instance; }
bool RenderingManager::Exists() { return crr_instance != nullptr; }
void RenderingManager::DestroyInstance() { delete crr_instance; crr_instance = nullptr; }
RenderingManager::RenderingManager()
    : curr_vol_renderer(nullptr), m_current_vr_method_id(0), m_current_camera_state_id(0), m_current_lightsource_data_id(0) {}
RenderingManager::~RenderingManager() {
  for (auto* r : m_vtr_vr_methods) delete r;
  m_vtr_vr_methods.clear();
}

bool RenderingManager::InitGL(int cuda_device) { return vrb::Device::Instance()->Init(cuda_device); }
void RenderingManager::AddVolumeRenderer(BaseVolumeRenderer* v) { m_vtr_vr_methods.push_back(v); }

// renderingmanager.cpp:126-172
bool RenderingManager::InitData(std::string path) {
  m_data_mgr.SetPathToData(path);
  if (!m_data_mgr.ReadData()) return false;
  if (m_camera_state_list.ReadCameraStates(path + "/#list_camera_states")) {
    m_current_camera_state_id = 0;
    curr_rdr_parameters.GetCamera()->SetData(m_camera_state_list.GetCameraState(0));
  }
  if (m_light_source_list.ReadLightSourceLists(path + "/#list_light_sources")) SetLightSourceList(0);
  if (m_vtr_vr_methods.empty()) { vrb::SetError("RenderingManager: No VR method added."); return false; }
  for (auto* r : m_vtr_vr_methods) r->SetExternalResources(&m_data_mgr, &curr_rdr_parameters);
  m_current_vr_method_id = 0;
  if (!SetCurrentVolumeRenderer(0)) return false;
  UpdateLightSourceCameraVectors();
  Reshape(curr_rdr_parameters.GetScreenWidth(), curr_rdr_parameters.GetScreenHeight());
  return true;
}

bool RenderingManager::InitDataInMemory(vis::StructuredGridVolume* vol, vis::TransferFunction* tf) {
  if (vol != m_data_mgr.GetCurrentStructuredVolume() && !m_data_mgr.SetStructuredVolume(vol)) return false;
  if (tf != m_data_mgr.GetCurrentTransferFunction() && !m_data_mgr.SetTransferFunction(tf)) return false;
  if (m_vtr_vr_methods.empty()) { vrb::SetError("RenderingManager: No VR method added."); return false; }
  for (auto* r : m_vtr_vr_methods) r->SetExternalResources(&m_data_mgr, &curr_rdr_parameters);
  return true;
}

bool RenderingManager::SetLightSourceList(int id) {
  vis::LightSourceListItem* l = m_light_source_list.GetList((unsigned)id);
  if (!l) { vrb::SetError("SetLightSourceList: bad index"); return false; }
  m_current_lightsource_data_id = id;
  curr_rdr_parameters.EraseAllLightSources();
  for (auto& ls : l->m_lightsources) curr_rdr_parameters.CreateNewLightSource(ls);
  return true;
}

// renderingmanager.cpp:421-426
void RenderingManager::UpdateLightSourceCameraVectors() {
  vrb::vec3 f, u, r;
  curr_rdr_parameters.GetCamera()->GetCameraVectors(&f, &u, &r);
  curr_rdr_parameters.SetBlinnPhongLightSourceCameraVectors(f, u, r);
}

bool RenderingManager::UpdateDataAndResetCurrentVRMode() {
  return curr_vol_renderer->Init(curr_rdr_parameters.GetScreenWidth(), curr_rdr_parameters.GetScreenHeight());
}

// renderingmanager.cpp:521-540
bool RenderingManager::SetCurrentVolumeRenderer(int id) {
  if (id < 0 || id >= (int)m_vtr_vr_methods.size()) { vrb::SetError("SetCurrentVolumeRenderer: bad index"); return false; }
  if (curr_vol_renderer) curr_vol_renderer->Clean();
  m_current_vr_method_id = id;
  curr_vol_renderer = m_vtr_vr_methods[id];
  if (curr_vol_renderer->GetDataTypeSupport() != m_data_mgr.GetInputVolumeDataType()) { vrb::SetError("renderer does not support the data type"); return false; }
  if (!UpdateDataAndResetCurrentVRMode()) return false;
  ParameterSpace ps;
  curr_vol_renderer->FillParameterSpace(ps);
  return true;
}
bool RenderingManager::SetCurrentVolumeRendererByAbbreviation(const std::string& abbr) {
  for (size_t i = 0; i < m_vtr_vr_methods.size(); ++i)
    if (abbr == m_vtr_vr_methods[i]->GetAbbreviationName()) return SetCurrentVolumeRenderer((int)i);
  vrb::SetError("no renderer with abbreviation " + abbr);
  return false;
}

bool RenderingManager::SetCameraState(int id) {
  vis::CameraData* cd = m_camera_state_list.GetCameraState((unsigned)id);
  if (!cd) { vrb::SetError("SetCameraState: bad index"); return false; }
  m_current_camera_state_id = id;
  SetCamera(cd);
  return true;
}
void RenderingManager::SetCamera(vis::CameraData* data) {
  curr_rdr_parameters.GetCamera()->SetData(data);
  if (curr_vol_renderer) curr_vol_renderer->SetOutdated();
}

// renderingmanager.cpp:321-335
void RenderingManager::Reshape(int w, int h) {
  curr_rdr_parameters.SetScreenSize(w, h);
  curr_rdr_parameters.GetCamera()->UpdateAspectRatio(float(w), float(h));
  if (curr_vol_renderer && curr_vol_renderer->IsBuilt()) { curr_vol_renderer->Reshape(w, h); curr_vol_renderer->SetOutdated(); }
}

// renderingmanager.cpp:174-208 without UI / swap: PrepareRender then the redraw entry for the multiscaling mode
bool RenderingManager::Display() {
  if (!curr_vol_renderer || !curr_vol_renderer->IsBuilt()) { vrb::SetError("Display: no built renderer"); return false; }
  vrb::SetError("");
  if (m_eval_running) curr_vol_renderer->SetOutdated();     // "We always redraw during evaluation" (renderingmanager.cpp:176-180)
  curr_vol_renderer->PrepareRender(curr_rdr_parameters.GetCamera());
  switch (curr_vol_renderer->GetCurrentMultiScalingMode()) {
    case 1: curr_vol_renderer->MultiSampleRedraw(); break;
    case 2: curr_vol_renderer->DownScalingRedraw(); break;
    case 3: curr_vol_renderer->UpScalingRedraw(); break;
    default: curr_vol_renderer->Redraw();
  }
  if (m_eval_running && vrb::LastError().empty()) EvaluationAfterFrame();
  return vrb::LastError().empty();
}

// =================================================================================================================
// extern "C" driver surface
// =================================================================================================================
void vrbh_register_renderers(RenderingManager* m);   // host_register.cpp

extern "C" {

const char* vrbh_last_error(void) { return vrb::LastError().c_str(); }

int vrbh_init(int cuda_device) {
  RenderingManager* m = RenderingManager::Instance();
  if (!m->InitGL(cuda_device)) return 1;
  if (m->GetNumberOfVolumeRenderers() == 0) vrbh_register_renderers(m);
  return 0;
}
void vrbh_shutdown(void) {
  RenderingManager::DestroyInstance();
  vrb::Device::Instance()->Shutdown();
}
void* vrbh_ctx(void) { return vrb::Device::Instance()->ctx(); }

int vrbh_init_data(const char* path) { return RenderingManager::Instance()->InitData(path) ? 0 : 1; }

int vrbh_set_volume(const void* vox, int w, int h, int d, int bpv, double sx, double sy, double sz) {
  if (!vox || (bpv != 1 && bpv != 2) || w <= 0 || h <= 0 || d <= 0) { vrb::SetError("vrbh_set_volume: bad arguments"); return 1; }
  size_t n = (size_t)w * h * d;
  void* copy = bpv == 1 ? (void*)new unsigned char[n] : (void*)new unsigned short[n];
  std::memcpy(copy, vox, n * bpv);
  vis::StructuredGridVolume* v = new vis::StructuredGridVolume("memory", w, h, d);
  v->SetScale(sx, sy, sz);
  v->SetArrayData(copy, bpv == 1 ? vis::_8_BITS : vis::_16_BITS);
  RenderingManager* m = RenderingManager::Instance();
  if (!m->GetDataManager()->SetStructuredVolume(v)) { delete v; return 1; }
  return 0;
}
int vrbh_load_volume(const char* path) {
  vis::VolumeReader vr;
  vis::StructuredGridVolume* v = vr.ReadStructuredVolume(path);
  if (!v) return 1;
  if (!RenderingManager::Instance()->GetDataManager()->SetStructuredVolume(v)) { delete v; return 1; }
  return 0;
}
static vis::TransferFunction1D* make_tf(const double* rgb, int n_rgb, const double* a, int n_a, int maxd, int ext) {
  vis::TransferFunction1D* tf = new vis::TransferFunction1D(maxd);
  tf->SetExtinctionCoefficientInput(ext != 0);
  for (int i = 0; i < n_rgb; ++i) tf->AddRGBControlPoint(vis::TransferControlPoint(rgb[4 * i], rgb[4 * i + 1], rgb[4 * i + 2], (int)rgb[4 * i + 3]));
  for (int i = 0; i < n_a; ++i) tf->AddAlphaControlPoint(vis::TransferControlPoint(a[2 * i], (int)a[2 * i + 1]));
  tf->Build();
  return tf;
}
int vrbh_set_tf_points(const double* rgb, int n_rgb, const double* a, int n_a, int maxd, int ext) {
  return RenderingManager::Instance()->GetDataManager()->SetTransferFunction(make_tf(rgb, n_rgb, a, n_a, maxd, ext)) ? 0 : 1;
}
int vrbh_load_tf(const char* path) {
  vis::TransferFunctionReader r;
  vis::TransferFunction* tf = r.ReadTransferFunction(path);
  if (!tf) return 1;
  return RenderingManager::Instance()->GetDataManager()->SetTransferFunction(tf) ? 0 : 1;
}
// after the volume and TF are set in memory: wire every renderer to the data (InitData without list files)
int vrbh_bind_data(void) {
  RenderingManager* m = RenderingManager::Instance();
  vis::DataManager* dm = m->GetDataManager();
  if (!dm->GetCurrentStructuredVolume() || !dm->GetCurrentTransferFunction()) { vrb::SetError("vrbh_bind_data: set volume and TF first"); return 1; }
  return m->InitDataInMemory(dm->GetCurrentStructuredVolume(), dm->GetCurrentTransferFunction()) ? 0 : 1;
}
// "Gradient" combo of the main window (renderingmanager.cpp:1094-1095: SetCurrentGradient + UpdateStructuredGradientTexture):
// idx 0 Sobel-Feldman, 1 finite differences, 2 compute-shader Sobel, 3 none
int vrbh_set_gradient(int idx) {
  vis::DataManager* dm = RenderingManager::Instance()->GetDataManager();
  if (!dm->SetCurrentGradient(idx)) return 0;
  if (!dm->GetCurrentStructuredVolume()) return 0;             // generated when the volume arrives
  if (!dm->UpdateStructuredGradientTexture()) return 1;
  BaseVolumeRenderer* r = RenderingManager::Instance()->GetCurrentVolumeRenderer();
  if (r) r->SetOutdated();
  return 0;
}
const char* vrbh_gradient_name(void) {
  static std::string s;
  s = RenderingManager::Instance()->GetDataManager()->CurrentGradientName();
  return s.c_str();
}
int vrbh_set_renderer(const char* abbr) { return RenderingManager::Instance()->SetCurrentVolumeRendererByAbbreviation(abbr) ? 0 : 1; }
int vrbh_reinit_renderer(void) { return RenderingManager::Instance()->UpdateDataAndResetCurrentVRMode() ? 0 : 1; }
int vrbh_set_param(const char* name, double value) {
  BaseVolumeRenderer* r = RenderingManager::Instance()->GetCurrentVolumeRenderer();
  if (!r) { vrb::SetError("vrbh_set_param: no renderer"); return 1; }
  if (!r->SetMultiScalingOption(name, value) && !r->SetParameter(name, value)) { vrb::SetError(std::string("unknown parameter ") + name); return 1; }
  return 0;
}
int vrbh_reshape(int w, int h) { RenderingManager::Instance()->Reshape(w, h); return 0; }
int vrbh_set_camera(const float eye[3], const float center[3], const float up[3]) {
  vis::CameraData cd;
  cd.eye = vrb::vec3(eye[0], eye[1], eye[2]); cd.center = vrb::vec3(center[0], center[1], center[2]); cd.up = vrb::vec3(up[0], up[1], up[2]);
  RenderingManager::Instance()->SetCamera(&cd);
  return 0;
}
int vrbh_get_camera_vectors(float fwd[3], float up[3], float right[3]) {   // Camera::GetCameraVectors (camera.cpp:336-341)
  vrb::vec3 f, u, r;
  RenderingManager::Instance()->GetRenderingParameters()->GetCamera()->GetCameraVectors(&f, &u, &r);
  fwd[0] = f.x; fwd[1] = f.y; fwd[2] = f.z; up[0] = u.x; up[1] = u.y; up[2] = u.z; right[0] = r.x; right[1] = r.y; right[2] = r.z;
  return 0;
}
int vrbh_set_camera_state(int id) { return RenderingManager::Instance()->SetCameraState(id) ? 0 : 1; }
int vrbh_num_camera_states(void) { return RenderingManager::Instance()->GetCameraStateList()->NumberOfCameraStates(); }
int vrbh_set_light_list(int id) { return RenderingManager::Instance()->SetLightSourceList(id) ? 0 : 1; }
int vrbh_set_light_position(const float p[3]) {
  RenderingManager::Instance()->GetRenderingParameters()->SetBlinnPhongLightingPosition(vrb::vec3(p[0], p[1], p[2]));
  if (RenderingManager::Instance()->GetCurrentVolumeRenderer()) RenderingManager::Instance()->GetCurrentVolumeRenderer()->SetOutdated();
  return 0;
}
int vrbh_update_light_camera_vectors(void) { RenderingManager::Instance()->UpdateLightSourceCameraVectors(); return 0; }
int vrbh_set_phong(float ka, float kd, float ks, float sh) { RenderingManager::Instance()->GetRenderingParameters()->SetPhongParameters(ka, kd, ks, sh); return 0; }
int vrbh_get_lighting(vrb_lighting* out) { *out = RenderingManager::Instance()->GetRenderingParameters()->MakeLightingBlock(); return 0; }
int vrbh_display(void) { return RenderingManager::Instance()->Display() ? 0 : 1; }
int vrbh_read_rgba(float* out, size_t cap_floats) {
  BaseVolumeRenderer* r = RenderingManager::Instance()->GetCurrentVolumeRenderer();
  if (!r) { vrb::SetError("vrbh_read_rgba: no renderer"); return 1; }
  std::vector<float> px;
  if (!r->ReadOutputRGBA32F(px)) return 1;
  if (px.size() > cap_floats) { vrb::SetError("vrbh_read_rgba: buffer too small"); return 1; }
  std::memcpy(out, px.data(), px.size() * sizeof(float));
  return 0;
}
const char* vrbh_renderer_name(int id, int abbreviation) {
  static std::string s;
  RenderingManager* m = RenderingManager::Instance();
  if (m->GetNumberOfVolumeRenderers() == 0) vrbh_register_renderers(m);
  if (id < 0 || id >= m->GetNumberOfVolumeRenderers()) return nullptr;
  BaseVolumeRenderer* r = m->GetVolumeRenderer(id);
  s = abbreviation ? r->GetAbbreviationName() : r->GetName();
  return s.c_str();
}

// ---- data-side helpers that need no GPU (pinned against oracle/_ref in tests/test_host_cpu.py) ------------------
void* vrbh_tf_create(const double* rgb, int n_rgb, const double* a, int n_a, int maxd, int ext) { return make_tf(rgb, n_rgb, a, n_a, maxd, ext); }
void* vrbh_tf_read(const char* path) { vis::TransferFunctionReader r; return r.ReadTransferFunction(path); }
void vrbh_tf_destroy(void* tf) { delete (vis::TransferFunction*)tf; }
int vrbh_tf_size(void* tf) { return ((vis::TransferFunction*)tf)->GetTextureSize(); }
void vrbh_tf_get(void* tf, double v, double maxv, float out[4]) { vrb::vec4 r = ((vis::TransferFunction*)tf)->Get(v, maxv); out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w; }
float vrbh_tf_get_extn(void* tf, double n) { return ((vis::TransferFunction*)tf)->GetExtN(n); }
float vrbh_tf_get_opcn(void* tf, double n) { return ((vis::TransferFunction*)tf)->GetOpcN(n); }
float vrbh_tf_get_opc(void* tf, double v, double mx) { return ((vis::TransferFunction*)tf)->GetOpc(v, mx); }
int vrbh_tf_textures(void* tf, float* rgbt, float* rgba, int cap_texels) {
  std::vector<float> a, b;
  vis::TransferFunction* t = (vis::TransferFunction*)tf;
  if (!t->GenerateTexture_1D_RGBt(a) || !t->GenerateTexture_1D_RGBA(b)) return -1;
  int n = t->GetTextureSize();
  if (n > cap_texels) return -1;
  std::memcpy(rgbt, a.data(), a.size() * sizeof(float));
  std::memcpy(rgba, b.data(), b.size() * sizeof(float));
  return n;
}
// volume reader: returns a handle; query dims; copy voxels out
void* vrbh_volume_read(const char* path) { vis::VolumeReader r; return r.ReadStructuredVolume(path); }
void vrbh_volume_destroy(void* v) { delete (vis::StructuredGridVolume*)v; }
void vrbh_volume_info(void* v, int dims[3], double scale[3], int* bpv, unsigned long long* checksum) {
  vis::StructuredGridVolume* s = (vis::StructuredGridVolume*)v;
  dims[0] = (int)s->GetWidth(); dims[1] = (int)s->GetHeight(); dims[2] = (int)s->GetDepth();
  scale[0] = s->GetScaleX(); scale[1] = s->GetScaleY(); scale[2] = s->GetScaleZ();
  *bpv = s->GetDataStorageSize() == vis::_8_BITS ? 1 : 2;
  *checksum = s->CheckSum();
}
void vrbh_volume_copy(void* v, void* out) {
  vis::StructuredGridVolume* s = (vis::StructuredGridVolume*)v;
  size_t n = (size_t)s->GetWidth() * s->GetHeight() * s->GetDepth() * (s->GetDataStorageSize() == vis::_8_BITS ? 1 : 2);
  std::memcpy(out, s->GetArrayData(), n);
}
double vrbh_volume_normalized_sample(void* v, int x, int y, int z) { return ((vis::StructuredGridVolume*)v)->GetNormalizedSample(x, y, z); }
// camera / light list parsers
int vrbh_read_camera_states(const char* path, float* out9, int cap) {
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
int vrbh_read_light_lists(const char* path, float* out13, int cap) {
  vis::LightSourceList l;
  if (!l.ReadLightSourceLists(path)) return -1;
  int k = 0;
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
void vrbh_look_at(const float eye[3], const float center[3], const float up[3], float out[16]) {
  vrb::mat4 m = vrb::lookAt(vrb::vec3(eye[0], eye[1], eye[2]), vrb::vec3(center[0], center[1], center[2]), vrb::vec3(up[0], up[1], up[2]));
  std::memcpy(out, m.m, sizeof(m.m));
}
float vrbh_tan_fovy(void) { vis::Camera c; return c.GetTanFovY(); }

}  // extern "C"
