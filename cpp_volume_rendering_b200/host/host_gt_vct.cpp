// host_gt_vct.cpp -- host side of the cone ground-truth renderer (rc1pcrtgt/crtgtrenderer.{h,cpp}) and of the
// voxel-cone-tracing renderer (rc1pvctsg/vctrenderer.{h,cpp}, preprocessingstages.{h,cpp}).
#include "vrbhost.h"
#include <cstring>
#include <random>

static vrb_ctx* CTX() { return vrb::Device::Instance()->ctx(); }
static bool CK(int rc) { if (rc != VRB_OK) { vrb::SetError(vrb_last_error()); return false; } return true; }
static const float kPIf = 3.14159265358979323846264338327950288f;

static float StepFromScale(vis::StructuredGridVolume* v) {
  vrb::dvec3 sv = v->GetScale();
  return float((0.5 / std::sqrt(3.0)) * std::sqrt(sv.x * sv.x + sv.y * sv.y + sv.z * sv.z));   // crtgtrenderer.cpp:106-107 (double here)
}

// ------------------------------------------------------------------ RC1PConeLightGroundTruthSteps
RC1PConeLightGroundTruthSteps::RC1PConeLightGroundTruthSteps()
    : m_u_step_size(0.5f), m_u_light_ray_initial_step(1.0f), m_u_light_ray_step_size(0.5f), m_light_parameters_outdated(true),
      m_apply_occlusion(false), m_occ_num_rays_sampled(1), m_occ_cone_aperture_angle(90.f), m_occ_cone_distance_eval(100.0f),
      m_apply_shadows(false), m_sdw_num_rays_sampled(1), m_sdw_cone_aperture_angle(1.0f), m_sdw_cone_distance_eval(100.0f),
      m_shadow_type(0) {
  std::memset(&m_cam, 0, sizeof(m_cam)); std::memset(&m_light, 0, sizeof(m_light)); std::memset(&m_prm, 0, sizeof(m_prm));
  vr_pixel_multiscaling_support = true;
}
RC1PConeLightGroundTruthSteps::~RC1PConeLightGroundTruthSteps() { Clean(); }
void RC1PConeLightGroundTruthSteps::Clean() { BaseVolumeRenderer::Clean(); }

bool RC1PConeLightGroundTruthSteps::Init(int shader_width, int shader_height) {
  if (IsBuilt()) Clean();
  if (m_ext_data_manager->GetCurrentVolumeTexture() == nullptr) return false;
  if (!UploadTransferFunction()) return false;
  Reshape(shader_width, shader_height);
  vis::StructuredGridVolume* vol = m_ext_data_manager->GetCurrentStructuredVolume();
  m_u_step_size = StepFromScale(vol);
  m_occ_cone_distance_eval = (float)(vol->GetDiagonal() * 0.50f);     // crtgtrenderer.cpp:111-113
  m_sdw_cone_distance_eval = (float)(vol->GetDiagonal() * 0.75f);
  m_light_parameters_outdated = true;
  SetBuilt(true);
  SetOutdated();
  return true;
}

// crtgtrenderer.cpp:131-187: a default-constructed engine + uniform_real_distribution<float>(0,1) created inside every
// regeneration (so the sequence restarts), draws consumed as (theta, phi) per ray, all occlusion rays first.
// theta is uniform in angle, not in solid angle -- kept.  std::default_random_engine is implementation-defined
// (minstd_rand0 in libstdc++), which is why the C ABI takes the tables as explicit inputs.
void RC1PConeLightGroundTruthSteps::GenerateRayTables(int n_occ, float occ_ap, int n_sdw, float sdw_ap, std::vector<float>& occ, std::vector<float>& sdw) {
  std::default_random_engine generator;
  std::uniform_real_distribution<float> distribution(0.0, 1.0);
  const vrb::vec3 gtray(0, 0, 1);
  auto fill = [&](int n, float aperture, std::vector<float>& out) {
    out.resize((size_t)n * 3);
    for (int i = 0; i < n; ++i) {
      float theta = distribution(generator) * (kPIf * (aperture / 180.f));
      float phi = distribution(generator) * (kPIf * (360.f / 180.f));
      vrb::vec3 v = RodriguesRotation(gtray, theta, vrb::vec3(0, 1, 0));
      v = RodriguesRotation(v, phi, gtray);
      out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
    }
  };
  fill(n_occ, occ_ap, occ);
  fill(n_sdw, sdw_ap, sdw);
}

bool RC1PConeLightGroundTruthSteps::Update(vis::Camera* camera) {
  if (m_light_parameters_outdated) {
    std::vector<float> occ, sdw;
    GenerateRayTables(m_occ_num_rays_sampled, m_occ_cone_aperture_angle, m_sdw_num_rays_sampled, m_sdw_cone_aperture_angle, occ, sdw);
    if (!CK(vrb_gt_set_rays(CTX(), occ.data(), m_occ_num_rays_sampled, sdw.data(), m_sdw_num_rays_sampled))) return false;
    m_light_parameters_outdated = false;
  }
  m_cam = MakeCameraBlock(camera);
  m_light = m_ext_rendering_parameters->MakeLightingBlock();
  m_light.apply_phong = (m_apply_gradient_shading && m_ext_data_manager->GetCurrentGradientTexture()) ? 1 : 0;
  // this shader's LightCamForward is -GetBlinnPhongLightSourceCameraForward() (crtgtrenderer.cpp:226-236)
  for (int i = 0; i < 3; ++i) m_light.light_forward[i] = -m_light.light_forward[i];
  m_prm.step_size = m_u_step_size;
  m_prm.light_ray_initial_gap = m_u_light_ray_initial_step;
  m_prm.light_ray_step_size = m_u_light_ray_step_size;
  m_prm.apply_occlusion = m_apply_occlusion ? 1 : 0; m_prm.occ_num_rays = m_occ_num_rays_sampled; m_prm.occ_cone_distance = m_occ_cone_distance_eval;
  m_prm.apply_shadow = m_apply_shadows ? 1 : 0; m_prm.sdw_num_rays = m_sdw_num_rays_sampled; m_prm.sdw_cone_distance = m_sdw_cone_distance_eval;
  m_prm.shadow_type = m_shadow_type;
  m_prm.count_samples = 0;
  return true;
}
// crtgtrenderer.cpp:340-350: the ground-truth frame when m_show_frame_texture, else RedrawCube (:327-338).  In the reference
// the flag is the "Show Generated Frame Texture" checkbox, off by default and reset by every camera / parameter change; a
// headless host has no checkbox, so here it defaults to ON (SetParameter("ShowFrameTexture", 0) gives the placeholder).
void RC1PConeLightGroundTruthSteps::Redraw() {
  if (m_show_frame_texture) CK(vrb_gt_render(CTX(), &m_cam, &m_light, &m_prm));
  else CK(vrb_gt_cube_render(CTX(), &m_cam));
}
bool RC1PConeLightGroundTruthSteps::SetParameter(const std::string& name, double v) {
  if (name == "StepSize") m_u_step_size = (float)v;
  else if (name == "ApplyGradientShading") m_apply_gradient_shading = v != 0.0;
  else if (name == "LightRayInitialGap") m_u_light_ray_initial_step = (float)v;
  else if (name == "LightRayStepSize") m_u_light_ray_step_size = (float)v;
  else if (name == "ApplyConeOcclusion") m_apply_occlusion = v != 0.0;
  else if (name == "OccNumberOfSampledRays") { m_occ_num_rays_sampled = (int)v; m_light_parameters_outdated = true; }
  else if (name == "OccConeApertureAngle") { m_occ_cone_aperture_angle = (float)v; m_light_parameters_outdated = true; }
  else if (name == "OccConeDistanceEvaluation") m_occ_cone_distance_eval = (float)v;
  else if (name == "ApplyConeShadow") m_apply_shadows = v != 0.0;
  else if (name == "SdwNumberOfSampledRays") { m_sdw_num_rays_sampled = (int)v; m_light_parameters_outdated = true; }
  else if (name == "SdwConeApertureAngle") { m_sdw_cone_aperture_angle = (float)v; m_light_parameters_outdated = true; }
  else if (name == "SdwConeDistanceEvaluation") m_sdw_cone_distance_eval = (float)v;
  else if (name == "SdwShadowType") m_shadow_type = (int)v;
  else if (name == "ShowFrameTexture") m_show_frame_texture = v != 0.0;
  else return false;
  SetOutdated();
  return true;
}

// ------------------------------------------------------------------ VCTPreProcessing / RC1PVoxelConeTracingSGPU
bool VCTPreProcessing::PreProcess(vis::StructuredGridVolume* vol, vis::TransferFunction* tf) {
  int dens_val = (int)vol->GetMaxDensity();                        // OpacityGaussianEvaluation: int dens_val (:149)
  std::vector<float> opc((size_t)dens_val + 1);
  for (int i = 0; i <= dens_val; ++i) opc[i] = tf->GetOpc((double)i, (double)dens_val);
  if (!CK(vrb_vct_build(CTX(), opc.data(), dens_val + 1))) return false;
  float ms = 0.0f;
  if (!CK(vrb_vct_info(CTX(), nullptr, nullptr, 0, nullptr, nullptr, &ms))) return false;
  maximum_standard_deviation = ms;
  return true;
}

RC1PVoxelConeTracingSGPU::RC1PVoxelConeTracingSGPU()
    : m_u_step_size(0.5f), apply_ambient_occlusion(true), apply_voxel_cone_tracing(true), cone_step_size(2.0f),
      cone_step_size_increase_rate(1.0f), cone_initial_step(2.0f), cone_apex_angle(2.0f), apply_correction_factor(true),
      opacity_correction_factor(2.0f), cone_number_of_samples(50) {
  m_pre_illum_str_vol.SetActive(false);                        // vctrenderer.cpp:47-48
  m_pre_illum_str_vol.SetLightCacheResolution(32, 32, 32);
  std::memset(&m_cam, 0, sizeof(m_cam)); std::memset(&m_light, 0, sizeof(m_light)); std::memset(&m_prm, 0, sizeof(m_prm));
  vr_pixel_multiscaling_support = true;
}
RC1PVoxelConeTracingSGPU::~RC1PVoxelConeTracingSGPU() { Clean(); }
void RC1PVoxelConeTracingSGPU::Clean() { BaseVolumeRenderer::Clean(); }
bool RC1PVoxelConeTracingSGPU::Init(int swidth, int sheight) {
  if (IsBuilt()) Clean();
  if (m_ext_data_manager->GetCurrentVolumeTexture() == nullptr) return false;
  if (!UploadTransferFunction()) return false;
  vis::StructuredGridVolume* vol = m_ext_data_manager->GetCurrentStructuredVolume();
  if (!pre_processing.PreProcess(vol, m_ext_data_manager->GetCurrentTransferFunction())) return false;   // vctrenderer.cpp:97-101
  vrb::dvec3 sv = vol->GetScale();
  m_u_step_size = float((0.5f / std::sqrt(3.0f)) * std::sqrt(sv.x * sv.x + sv.y * sv.y + sv.z * sv.z));
  Reshape(swidth, sheight);
  SetBuilt(true);
  SetOutdated();
  return true;
}
bool RC1PVoxelConeTracingSGPU::Update(vis::Camera* camera) {
  m_cam = MakeCameraBlock(camera);
  m_light = m_ext_rendering_parameters->MakeLightingBlock();
  m_light.apply_phong = (m_apply_gradient_shading && m_ext_data_manager->GetCurrentGradientTexture()) ? 1 : 0;
  m_prm.step_size = m_u_step_size;
  m_prm.apply_occlusion = apply_ambient_occlusion ? 1 : 0;
  m_prm.apply_shadow = apply_voxel_cone_tracing ? 1 : 0;
  m_prm.tan_cone_apex_angle = std::tan(cone_apex_angle * kPIf / 180.0f);      // vctrenderer.cpp:145
  m_prm.cone_step_size = cone_step_size;
  m_prm.cone_step_increase_rate = cone_step_size_increase_rate;
  m_prm.cone_initial_step = cone_initial_step;
  m_prm.opacity_correction_factor = opacity_correction_factor;
  m_prm.apply_opacity_correction = apply_correction_factor ? 1 : 0;
  m_prm.cone_number_of_samples = cone_number_of_samples;
  m_prm.volume_max_density = (float)m_ext_data_manager->GetCurrentStructuredVolume()->GetMaxDensity();
  m_prm.volume_max_stddev = (float)pre_processing.maximum_standard_deviation;
  m_prm.count_samples = 0;
  if (m_pre_illum_str_vol.IsActive()) {                         // PreComputeLightCache on every Update (vctrenderer.cpp:126,393-515)
    const int* res = m_pre_illum_str_vol.GetLightCacheResolution();
    if (!CK(vrb_vct_light_cache_build(CTX(), &m_light, &m_prm, res[0], res[1], res[2]))) return false;
  }
  return true;
}
void RC1PVoxelConeTracingSGPU::Redraw() {
  if (m_pre_illum_str_vol.IsActive()) {                         // rendering shader = obj_ray_marching.comp
    vrb_obj_params op;
    op.step_size = m_u_step_size; op.apply_occlusion = m_prm.apply_occlusion; op.apply_shadow = m_prm.apply_shadow; op.count_samples = 0;
    CK(vrb_obj_march_render(CTX(), &m_cam, &m_light, &op));
    return;
  }
  CK(vrb_vct_render(CTX(), &m_cam, &m_light, &m_prm));
}
// the reference leaves this renderer without a parameter space (BaseVolumeRenderer::FillParameterSpace clears it);
// the step-size sweep of rc1prenderer.cpp:225-229 is offered here too so that the evaluation harness has something to vary
void RC1PVoxelConeTracingSGPU::FillParameterSpace(ParameterSpace& pspace) {
  pspace.ClearParameterDimensions();
  pspace.AddParameterDimension(new ParameterRangeFloat("StepSize", &m_u_step_size, 0.2f, 2.0f, 0.1f));
}
bool RC1PVoxelConeTracingSGPU::SetParameter(const std::string& name, double v) {
  if (name == "StepSize") m_u_step_size = (float)v;
  else if (name == "ApplyGradientShading") m_apply_gradient_shading = v != 0.0;
  else if (name == "ApplyOcclusion") apply_ambient_occlusion = v != 0.0;
  else if (name == "ApplyShadow") apply_voxel_cone_tracing = v != 0.0;
  else if (name == "UsePreIllumination") m_pre_illum_str_vol.SetActive(v != 0.0);
  else if (name == "LightCacheResolution") m_pre_illum_str_vol.SetLightCacheResolution((int)v, (int)v, (int)v);
  else if (name == "ConeStepSize") cone_step_size = (float)v;
  else if (name == "ConeStepIncreaseRate") cone_step_size_increase_rate = (float)v;
  else if (name == "ConeInitialStep") cone_initial_step = (float)v;
  else if (name == "ConeApexAngle") cone_apex_angle = (float)v;
  else if (name == "ApplyOpacityCorrectionFactor") apply_correction_factor = v != 0.0;
  else if (name == "OpacityCorrectionFactor") opacity_correction_factor = (float)v;
  else if (name == "ConeNumberOfSamples") cone_number_of_samples = (int)v;
  else return false;
  SetOutdated();
  return true;
}

extern "C" void vrbh_gt_ray_tables(int n_occ, float occ_aperture_deg, int n_sdw, float sdw_aperture_deg, float* occ_out, float* sdw_out) {
  std::vector<float> o, s;
  RC1PConeLightGroundTruthSteps::GenerateRayTables(n_occ, occ_aperture_deg, n_sdw, sdw_aperture_deg, o, s);
  std::memcpy(occ_out, o.data(), o.size() * sizeof(float));
  std::memcpy(sdw_out, s.data(), s.size() * sizeof(float));
}
