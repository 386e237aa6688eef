// host_renderers.cpp -- BaseVolumeRenderer and its B200 subclasses.  Every GPU action goes through include/vrb200.h.
#include "vrbhost.h"
#include <cstring>

static vrb_ctx* CTX() { return vrb::Device::Instance()->ctx(); }
static bool CK(int rc) { if (rc != VRB_OK) { vrb::SetError(vrb_last_error()); return false; } return true; }

// ------------------------------------------------------------------ BaseVolumeRenderer (volrenderbase.cpp)
BaseVolumeRenderer::BaseVolumeRenderer()
    : vr_built(false), vr_outdated(true), vr_pixel_multiscaling_support(false), vr_pixel_multiscaling_mode(0),
      m_ext_data_manager(nullptr), m_ext_rendering_parameters(nullptr) {}
BaseVolumeRenderer::~BaseVolumeRenderer() {}
void BaseVolumeRenderer::SetExternalResources(vis::DataManager* d, vis::RenderingParameters* r) { m_ext_data_manager = d; m_ext_rendering_parameters = r; }
void BaseVolumeRenderer::Clean() { m_rdr_frame_to_screen.Clean(); SetBuilt(false); }
void BaseVolumeRenderer::ReloadShaders() {}
void BaseVolumeRenderer::Redraw() {}
// Pixel multi-scaling.  The reference's subclasses repeat "clear, dispatch, filter" in each of the three methods
// (rc1prenderer.cpp:153-190, ebsrenderer.cpp:262-299, ...); the dispatch is Redraw(), which renders into whatever size the
// frame has, so the three variants live here once.
void BaseVolumeRenderer::MultiSampleRedraw() { Redraw(); m_rdr_frame_to_screen.DrawMultiSampleHigherResolutionMode(); }
void BaseVolumeRenderer::DownScalingRedraw() { Redraw(); m_rdr_frame_to_screen.DrawHigherResolutionWithDownScale(); }
void BaseVolumeRenderer::UpScalingRedraw() { Redraw(); m_rdr_frame_to_screen.DrawLowerResolutionWithUpScale(); }
#define MULTISAMPLE_NUMBEROFSAMPLES_W 2      /* cppvolrend/defines.h:16-17 */
#define MULTISAMPLE_NUMBEROFSAMPLES_H 2
void BaseVolumeRenderer::Reshape(int w, int h) {            // volrenderbase.cpp:42-68
  if (IsPixelMultiScalingSupported() && GetCurrentMultiScalingMode() > 0) {
    if (GetCurrentMultiScalingMode() == UP_SCALING_RENDER)
      m_rdr_frame_to_screen.UpdateScreenResolutionMultiScaling(w, h, -MULTISAMPLE_NUMBEROFSAMPLES_W, -MULTISAMPLE_NUMBEROFSAMPLES_H);
    else
      m_rdr_frame_to_screen.UpdateScreenResolutionMultiScaling(w, h, MULTISAMPLE_NUMBEROFSAMPLES_W, MULTISAMPLE_NUMBEROFSAMPLES_H);
  } else {
    m_rdr_frame_to_screen.UpdateScreenResolution(w, h);
  }
  SetOutdated();
}
// what the radio buttons and the kernel combo of AddImGuiMultiSampleOptions do (volrenderbase.cpp:121-197)
bool BaseVolumeRenderer::SetMultiScalingOption(const std::string& name, double value) {
  if (!IsPixelMultiScalingSupported()) return false;
  if (name == "MultiScalingMode") {
    int e = (int)value;
    if (e < 0 || e > 3) return false;
    SetCurrentMultiScalingMode(e);
    const int w = m_ext_rendering_parameters->GetScreenWidth(), h = m_ext_rendering_parameters->GetScreenHeight();
    if (e == 0) m_rdr_frame_to_screen.UpdateScreenResolution(w, h);
    else {
      const int s = e == UP_SCALING_RENDER ? -1 : 1;
      m_rdr_frame_to_screen.SetMultiResolutionScreenMultiplier(s * MULTISAMPLE_NUMBEROFSAMPLES_W, s * MULTISAMPLE_NUMBEROFSAMPLES_H);
      m_rdr_frame_to_screen.UpdateScreenResolutionMultiScaling(w, h);
    }
    SetOutdated();
    return true;
  }
  if (name == "ImageKernelFilter") {
    if (value < 0 || value > 5) return false;
    m_rdr_frame_to_screen.SetImageKernelFilter((unsigned int)value);
    return true;
  }
  return false;
}
void BaseVolumeRenderer::SetImGuiComponents() {}
void BaseVolumeRenderer::FillParameterSpace(ParameterSpace& pspace) { pspace.ClearParameterDimensions(); }
void BaseVolumeRenderer::PrepareRender(vis::Camera* camera) { if (IsOutdated()) { Update(camera); vr_outdated = false; } }
void BaseVolumeRenderer::SetOutdated() { vr_outdated = true; }
bool BaseVolumeRenderer::IsOutdated() { return vr_outdated; }
bool BaseVolumeRenderer::IsBuilt() { return vr_built; }
bool BaseVolumeRenderer::IsPixelMultiScalingSupported() { return vr_pixel_multiscaling_support; }
int BaseVolumeRenderer::GetCurrentMultiScalingMode() { return IsPixelMultiScalingSupported() ? vr_pixel_multiscaling_mode : 0; }
void BaseVolumeRenderer::SetCurrentMultiScalingMode(int f) { vr_pixel_multiscaling_mode = f; }
void BaseVolumeRenderer::SetBuilt(bool b) { vr_built = b; }
bool BaseVolumeRenderer::SetParameter(const std::string&, double) { return false; }
void* BaseVolumeRenderer::GetScreenTextureDevicePtr() {
  void* p = nullptr;
  if (vrb_frame_device_ptr(CTX(), &p, nullptr, nullptr) != VRB_OK) return nullptr;
  return p;
}
bool BaseVolumeRenderer::UploadTransferFunction() {
  vis::TransferFunction* tf = m_ext_data_manager->GetCurrentTransferFunction();
  if (!tf) { vrb::SetError("no transfer function"); return false; }
  std::vector<float> rgbt, rgba;
  if (!tf->GenerateTexture_1D_RGBt(rgbt) || !tf->GenerateTexture_1D_RGBA(rgba)) { vrb::SetError("transfer function has no 1-D texture"); return false; }
  return CK(vrb_tf_upload(CTX(), rgbt.data(), rgba.data(), tf->GetTextureSize()));
}
vrb_camera BaseVolumeRenderer::MakeCameraBlock(vis::Camera* camera) {
  vrb_camera c;
  vrb::vec3 e = camera->GetEye();
  c.eye[0] = e.x; c.eye[1] = e.y; c.eye[2] = e.z;
  vrb::mat4 m = camera->LookAt();
  std::memcpy(c.lookat, m.m, sizeof(c.lookat));
  c.tan_fovy = camera->GetTanFovY();
  c.aspect = camera->GetAspectRatio();
  return c;
}

static float DefaultStepSize(vis::StructuredGridVolume* v) {
  // rc1prenderer.cpp:62-63: (0.5f / sqrt(3.0f)) * sqrt(sx^2 + sy^2 + sz^2)
  vrb::dvec3 sv = v->GetScale();
  return float((0.5f / std::sqrt(3.0f)) * std::sqrt(sv.x * sv.x + sv.y * sv.y + sv.z * sv.z));
}

// ------------------------------------------------------------------ RayCasting1Pass (rc1prenderer.cpp)
RayCasting1Pass::RayCasting1Pass() : m_has_tf(false), m_u_step_size(0.5f), m_apply_gradient_shading(false), m_skip_empty(false) {
  std::memset(&m_cam, 0, sizeof(m_cam)); std::memset(&m_light, 0, sizeof(m_light));
  vr_pixel_multiscaling_support = true;
}
RayCasting1Pass::~RayCasting1Pass() { Clean(); }
void RayCasting1Pass::Clean() { m_has_tf = false; BaseVolumeRenderer::Clean(); }
bool RayCasting1Pass::Init(int swidth, int sheight) {
  if (IsBuilt()) Clean();
  if (m_ext_data_manager->GetCurrentVolumeTexture() == nullptr) return false;
  if (!UploadTransferFunction()) return false;
  m_has_tf = true;
  m_u_step_size = DefaultStepSize(m_ext_data_manager->GetCurrentStructuredVolume());
  Reshape(swidth, sheight);
  SetBuilt(true);
  SetOutdated();
  return true;
}
bool RayCasting1Pass::Update(vis::Camera* camera) {
  m_cam = MakeCameraBlock(camera);
  // rc1prenderer.cpp:112-135: ApplyGradientPhongShading + the Blinn-Phong / light uniforms
  m_light = m_ext_rendering_parameters->MakeLightingBlock();
  m_light.apply_phong = (m_apply_gradient_shading && m_ext_data_manager->GetCurrentGradientTexture()) ? 1 : 0;
  return true;
}
void RayCasting1Pass::Redraw() {
  vrb_rc1pass_params p;
  p.step_size = m_u_step_size; p.count_samples = 0; p.skip_empty = m_skip_empty ? 1 : 0;
  CK(vrb_rc1pass_render_lit(CTX(), &m_cam, &p, &m_light));   // ClearTexture + dispatch
}
void RayCasting1Pass::FillParameterSpace(ParameterSpace& pspace) {
  pspace.ClearParameterDimensions();
  pspace.AddParameterDimension(new ParameterRangeFloat("StepSize", &m_u_step_size, 0.2f, 2.0f, 0.1f));
}
bool RayCasting1Pass::SetParameter(const std::string& name, double value) {
  if (name == "StepSize") { m_u_step_size = std::fmax(std::fmin((float)value, 100.0f), 0.01f); SetOutdated(); return true; }
  if (name == "SkipEmptySpace") { m_skip_empty = value != 0.0; SetOutdated(); return true; }
  if (name == "ApplyGradientShading") { m_apply_gradient_shading = value != 0.0; SetOutdated(); return true; }
  return false;
}

// ------------------------------------------------------------------ RayCasting1PassIsoAdapt (rc1pisoadaptrenderer.cpp)
RayCasting1PassIsoAdapt::RayCasting1PassIsoAdapt()
    : m_u_isovalue(0.5f), m_u_step_size_small(0.05f), m_u_step_size_large(1.0f), m_u_step_size_range(0.1f), m_apply_gradient_shading(false) {
  m_u_color[0] = 0.66f; m_u_color[1] = 0.6f; m_u_color[2] = 0.05f; m_u_color[3] = 1.0f;       // :13-22
  std::memset(&m_cam, 0, sizeof(m_cam)); std::memset(&m_light, 0, sizeof(m_light)); std::memset(&m_prm, 0, sizeof(m_prm));
}
RayCasting1PassIsoAdapt::~RayCasting1PassIsoAdapt() { Clean(); }
void RayCasting1PassIsoAdapt::Clean() { BaseVolumeRenderer::Clean(); }
bool RayCasting1PassIsoAdapt::Init(int swidth, int sheight) {      // :61-110: needs the volume texture only
  if (IsBuilt()) Clean();
  if (m_ext_data_manager->GetCurrentVolumeTexture() == nullptr) return false;
  Reshape(swidth, sheight);
  SetBuilt(true);
  SetOutdated();
  return true;
}
bool RayCasting1PassIsoAdapt::Update(vis::Camera* camera) {        // :113-165
  m_cam = MakeCameraBlock(camera);
  m_prm.isovalue = m_u_isovalue; m_prm.step_size_small = m_u_step_size_small; m_prm.step_size_large = m_u_step_size_large;
  m_prm.step_size_range = m_u_step_size_range;
  for (int i = 0; i < 4; ++i) m_prm.color[i] = m_u_color[i];
  m_prm.count_samples = 0;
  m_light = m_ext_rendering_parameters->MakeLightingBlock();
  m_light.apply_phong = (m_apply_gradient_shading && m_ext_data_manager->GetCurrentGradientTexture()) ? 1 : 0;
  return true;
}
void RayCasting1PassIsoAdapt::Redraw() { CK(vrb_iso_render(CTX(), &m_cam, &m_light, &m_prm)); }   // ClearTexture + dispatch (:168-179)
void RayCasting1PassIsoAdapt::FillParameterSpace(ParameterSpace& pspace) {                        // :182-188
  pspace.ClearParameterDimensions();
  pspace.AddParameterDimension(new ParameterRangeFloat("StepSizeSmall", &m_u_step_size_small, 0.01f, 0.25f, 0.05f));
  pspace.AddParameterDimension(new ParameterRangeFloat("StepSizeLarge", &m_u_step_size_large, 0.25f, 2.0f, 0.25f));
  pspace.AddParameterDimension(new ParameterRangeFloat("StepSizeRange", &m_u_step_size_range, 0.05f, 0.26f, 0.05f));
}
bool RayCasting1PassIsoAdapt::SetParameter(const std::string& name, double v) {                   // the sliders of SetImGuiComponents (:191-250)
  if (name == "Isovalue") m_u_isovalue = (float)v;
  else if (name == "StepSizeSmall") m_u_step_size_small = std::fmax((float)v, 1e-4f);
  else if (name == "StepSizeLarge") m_u_step_size_large = std::fmax((float)v, 1e-4f);
  else if (name == "StepSizeRange") m_u_step_size_range = (float)v;
  else if (name == "ColorR") m_u_color[0] = (float)v;
  else if (name == "ColorG") m_u_color[1] = (float)v;
  else if (name == "ColorB") m_u_color[2] = (float)v;
  else if (name == "ColorA") m_u_color[3] = (float)v;
  else if (name == "ApplyGradientShading") m_apply_gradient_shading = v != 0.0;
  else return false;
  SetOutdated();
  return true;
}

// ------------------------------------------------------------------ RC1PExtinctionBasedShading (ebsrenderer.cpp)
RC1PExtinctionBasedShading::RC1PExtinctionBasedShading()
    : m_has_tf(false), m_has_sat(false), m_u_step_size(0.5f),
      apply_ambient_occlusion(true), ambient_occlusion_shells(15), ambient_occlusion_radius(1.0f),
      apply_directional_shadows(true), dir_shadow_cone_samples(120), dir_shadow_cone_angle(1.0f),
      dir_shadow_sample_interval(2.0f), dir_shadow_initial_step(2.0f), dir_shadow_user_interface_weight(1.0f),
      dir_cone_max_distance(0.0f), type_of_shadow(0) {
  m_pre_illum_str_vol.SetActive(false);                        // ebsrenderer.cpp:46-47
  m_pre_illum_str_vol.SetLightCacheResolution(32, 32, 32);
  std::memset(&m_cam, 0, sizeof(m_cam)); std::memset(&m_light, 0, sizeof(m_light)); std::memset(&m_prm, 0, sizeof(m_prm));
  vr_pixel_multiscaling_support = true;
}
RC1PExtinctionBasedShading::~RC1PExtinctionBasedShading() { Clean(); }
void RC1PExtinctionBasedShading::Clean() { m_has_tf = false; m_has_sat = false; BaseVolumeRenderer::Clean(); }

// ebsrenderer.cpp:624-723.  The per-voxel extinction tf->GetExtN(vol->GetNormalizedSample(x,y,z)) only depends on the
// voxel VALUE, so the host evaluates it once per possible value and the device does fill + three scan passes.
bool RC1PExtinctionBasedShading::GenerateExtinctionSAT3DTex(vis::StructuredGridVolume* vol, vis::TransferFunction* tf) {
  int n = vol->GetDataStorageSize() == vis::_8_BITS ? 256 : 65536;
  double maxv = vol->GetMaxDensity();
  std::vector<float> lut((size_t)n);
  for (int v = 0; v < n; ++v) lut[v] = tf->GetExtN((double)v / maxv);
  return CK(vrb_sat_build(CTX(), lut.data(), n));
}

bool RC1PExtinctionBasedShading::Init(int swidth, int sheight) {
  if (IsBuilt()) Clean();
  if (m_ext_data_manager->GetCurrentVolumeTexture() == nullptr) return false;
  if (!UploadTransferFunction()) return false;
  m_has_tf = true;
  vis::StructuredGridVolume* vold = m_ext_data_manager->GetCurrentStructuredVolume();
  if (!GenerateExtinctionSAT3DTex(vold, m_ext_data_manager->GetCurrentTransferFunction())) return false;
  m_has_sat = true;
  float v_w = vold->GetWidth() * vold->GetScaleX(), v_h = vold->GetHeight() * vold->GetScaleY(), v_d = vold->GetDepth() * vold->GetScaleZ();
  float Dv = std::sqrt(v_w * v_w + v_h * v_h + v_d * v_d);
  dir_cone_max_distance = 0.75f * Dv;            // ebsrenderer.cpp:98-105
  m_u_step_size = DefaultStepSize(vold);
  Reshape(swidth, sheight);
  SetBuilt(true);
  SetOutdated();
  return true;
}

bool RC1PExtinctionBasedShading::Update(vis::Camera* camera) {
  m_cam = MakeCameraBlock(camera);
  m_light = m_ext_rendering_parameters->MakeLightingBlock();
  m_light.apply_phong = (m_apply_gradient_shading && m_ext_data_manager->GetCurrentGradientTexture()) ? 1 : 0;
  m_prm.step_size = m_u_step_size;
  m_prm.apply_occlusion = apply_ambient_occlusion ? 1 : 0;
  m_prm.apply_shadow = apply_directional_shadows ? 1 : 0;
  m_prm.amb_occ_shells = ambient_occlusion_shells;
  m_prm.amb_occ_radius = ambient_occlusion_radius;
  m_prm.sdw_cone_angle_rad = (float)(dir_shadow_cone_angle * 3.14159265358979323846264338327950288 / 180.0);
  m_prm.sdw_sample_interval = dir_shadow_sample_interval;
  m_prm.sdw_initial_step = dir_shadow_initial_step;
  m_prm.sdw_ui_weight = dir_shadow_user_interface_weight;
  m_prm.sdw_cone_max_distance = dir_cone_max_distance;
  m_prm.type_of_shadow = type_of_shadow;
  m_prm.count_samples = 0;
  if (m_pre_illum_str_vol.IsActive()) {                         // PreComputeLightCache on every Update (ebsrenderer.cpp:127,441-555)
    const int* res = m_pre_illum_str_vol.GetLightCacheResolution();
    if (!CK(vrb_ebs_light_cache_build(CTX(), &m_light, &m_prm, res[0], res[1], res[2]))) return false;
  }
  return true;
}
void RC1PExtinctionBasedShading::Redraw() {
  if (m_pre_illum_str_vol.IsActive()) {                         // rendering shader = obj_ray_marching.comp (ebsrenderer.cpp:571)
    vrb_obj_params op;
    op.step_size = m_u_step_size; op.apply_occlusion = m_prm.apply_occlusion; op.apply_shadow = m_prm.apply_shadow; op.count_samples = 0;
    CK(vrb_obj_march_render(CTX(), &m_cam, &m_light, &op));
    return;
  }
  CK(vrb_ebs_render(CTX(), &m_cam, &m_light, &m_prm));
}
void RC1PExtinctionBasedShading::FillParameterSpace(ParameterSpace& pspace) {     // ebsrenderer.cpp:430-435
  pspace.ClearParameterDimensions();
  pspace.AddParameterDimension(new ParameterRangeInt("AmbientOccShells", &ambient_occlusion_shells, 1, 20, 1));
  pspace.AddParameterDimension(new ParameterRangeFloat("AmbientOccRadius", &ambient_occlusion_radius, 0.1f, 1.5f, 0.1f));
}
bool RC1PExtinctionBasedShading::SetParameter(const std::string& name, double v) {
  if (name == "StepSize") m_u_step_size = (float)v;
  else if (name == "ApplyGradientShading") m_apply_gradient_shading = v != 0.0;
  else if (name == "ApplyOcclusion") apply_ambient_occlusion = v != 0.0;
  else if (name == "ApplyShadow") apply_directional_shadows = v != 0.0;
  else if (name == "UsePreIllumination") m_pre_illum_str_vol.SetActive(v != 0.0);
  else if (name == "LightCacheResolution") m_pre_illum_str_vol.SetLightCacheResolution((int)v, (int)v, (int)v);
  else if (name == "AmbOccShells") ambient_occlusion_shells = (int)v;
  else if (name == "AmbOccRadius") ambient_occlusion_radius = (float)v;
  else if (name == "DirSdwConeAngle") dir_shadow_cone_angle = (float)v;          // degrees, as in the UI
  else if (name == "DirSdwSampleInterval") dir_shadow_sample_interval = (float)v;
  else if (name == "DirSdwInitialStep") dir_shadow_initial_step = (float)v;
  else if (name == "DirSdwUserInterfaceWeight") dir_shadow_user_interface_weight = (float)v;
  else if (name == "DirSdwConeMaxDistance") dir_cone_max_distance = (float)v;
  else if (name == "TypeOfShadow") type_of_shadow = (int)v;
  else return false;
  SetOutdated();
  return true;
}
