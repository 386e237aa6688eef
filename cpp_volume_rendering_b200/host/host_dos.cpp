// host_dos.cpp -- host side of the directional-occlusion renderer: cone section schedule (double precision, stays on
// the host like in the reference), pyramid generator front end, and the BaseVolumeRenderer subclass.
//   ConeGaussianSampler          rc1pdosct/conegaussiansampler.{h,cpp}
//   ExtinctionCoefficientVolume  rc1pdosct/extcoefvolumegenerator.{h,cpp}
//   RC1PConeTracingDirOcclusionShading  rc1pdosct/dosrcrenderer.{h,cpp}
#include "vrbhost.h"
#include <cstring>

static vrb_ctx* CTX() { return vrb::Device::Instance()->ctx(); }
static bool CK(int rc) { if (rc != VRB_OK) { vrb::SetError(vrb_last_error()); return false; } return true; }
static const double kPI = 3.14159265358979323846264338327950288;
static const double kDiv3 = 1.0 + (2.0 / std::sqrt(3.0));   // D_HEMISPHERE_CONE_DIV_3 (conegaussiansampler.h:18)
static const double kDiv7 = 3.010000;                       // D_HEMISPHERE_CONE_DIV_7 (:19)

// v cos t + (k x v) sin t + k (k.v)(1 - cos t), normalised; float overload (see oracle/oracle_dos.cpp on why)
vrb::vec3 RodriguesRotation(vrb::vec3 v, float teta, vrb::vec3 k) {
  float c = std::cos(teta), s = std::sin(teta);
  vrb::vec3 r = v * c + vrb::cross(k, v) * s + (k * vrb::dot(k, v)) * (1.0f - c);
  return vrb::normalize(r);
}

static float clampf(float v, float lo, float hi) { return std::fmin(std::fmax(v, lo), hi); }

ConeGaussianSampler::ConeGaussianSampler()
    : gaussian_samples_1(0), gaussian_samples_3(0), gaussian_samples_7(0), ray3_adj_weight(0), ray7_adj_weight(0),
      cone_half_angle(30.0f), initial_step(3.0f), max_gaussian_packing(_7), covered_distance(100.0f),
      ui_weight_percentage(1.0f), d_sigma(1.25f), r_sigma(2.0f) {}
void ConeGaussianSampler::SetConeHalfAngle(float a) { cone_half_angle = clampf(a, 0.5f, 89.5f); }
void ConeGaussianSampler::SetInitialStep(float s) { initial_step = s > 0.0f ? s : 0.0f; }
void ConeGaussianSampler::SetMaxGaussianPacking(int p) { if (p >= 0 && p <= 2) max_gaussian_packing = (CONEPACKING)p; }
void ConeGaussianSampler::SetCoveredDistance(float d) { covered_distance = d > 10.0f ? d : 10.0f; }
void ConeGaussianSampler::SetIntegrationHalfStepMultiplier(float v) { d_sigma = clampf(v, 1.0f, 3.0f); }
void ConeGaussianSampler::SetGaussianSigmaLimitMultiplier(float s) { r_sigma = clampf(s, 0.5f, 3.0f); }

// AddGaussianSampleStep / With3 / With7 (conegaussiansampler.cpp:418-498) as one escalation loop over the packing
// stages 1 -> 3 -> 7: a section is accepted at the first stage whose sub-cone radius fits r_sigma * sigma.
bool ConeGaussianSampler::AddGaussianSampleStep(double curr_pos, double sg, int* n_gaussians) {
  static const int count[3] = {1, 3, 7};
  const double divisor[3] = {1.0, kDiv3, kDiv7};
  int stage = *n_gaussians > 3 ? 2 : (*n_gaussians > 1 ? 1 : 0);
  while (true) {
    *n_gaussians = count[stage];
    double half = stage == 0 ? (double)GetConeHalfAngle() : ((double)GetConeHalfAngle() / divisor[stage]);
    double cone_radius = curr_pos * std::tan(half * kPI / 180.0);
    if (cone_radius > (double)GetGaussianSigmaLimitMultiplier() * sg) {
      if (stage < 2 && (int)max_gaussian_packing > stage) { ++stage; continue; }
      return false;
    }
    SectionInfo si;
    si.number_of_gaussians = count[stage]; si.distance_from_origin = curr_pos; si.cone_radius = cone_radius;
    si.sampled_gaussian_sigma = sg; si.d_integral = si.mip_map_level = si.amplitude = 0.0;
    data_cone_sectionsinfo.push_back(si);
    return true;
  }
}

// midpoint rule over [-R, R] in segments of ~0.05 (:393-414)
double ConeGaussianSampler::IntegrateGaussian(double sdev, double cone_radius) {
  int nt = (int)std::ceil((2.0 * cone_radius) / 0.05);
  double segment = (2.0 * cone_radius) / double(nt);
  double s0 = -cone_radius + segment * 0.5;
  double S = 0.0;
  for (int i = 0; i < nt; i++) {
    double x = s0 + segment * double(i);
    S += ((1.0 / (std::sqrt(2.0 * kPI) * sdev)) * std::exp(-(x * x) / (2.0 * sdev * sdev))) * segment;
  }
  return S;
}

bool ConeGaussianSampler::ComputeConeIntegrationSteps(double min_sg_gaussian) {
  data_cone_sectionsinfo.clear();
  data_cone_intervalsinfo.clear();
  gaussian_samples_1 = gaussian_samples_3 = gaussian_samples_7 = 0;
  const double A = (double)GetConeHalfAngle();
  const vrb::vec3 Z(0, 0, 1), Y(0, 1, 0);
  // circle packings: 3 rays at 120 degrees, 7 rays = centre + 6 at 60 degrees (:221-252)
  {
    double t1 = (A - A / kDiv3) * kPI / 180.0;
    ray3_axis[0] = RodriguesRotation(Z, (float)t1, Y);
    ray3_axis[0] = RodriguesRotation(ray3_axis[0], (float)(30.0 * kPI / 180.0), Z);
    ray3_axis[1] = RodriguesRotation(ray3_axis[0], (float)(120.0 * kPI / 180.0), Z);
    ray3_axis[2] = RodriguesRotation(ray3_axis[1], (float)(120.0 * kPI / 180.0), Z);
  }
  {
    ray7_axis[0] = Z;
    double t1 = (A - A / kDiv7) * kPI / 180.0;
    ray7_axis[1] = RodriguesRotation(ray7_axis[0], (float)t1, Y);
    for (int i = 2; i < 7; ++i) ray7_axis[i] = RodriguesRotation(ray7_axis[i - 1], (float)(60.0 * kPI / 180.0), Z);
  }
  int n_gaussians = 1;
  double curr_pos = initial_step, sigma = min_sg_gaussian;
  while (!AddGaussianSampleStep(curr_pos, sigma, &n_gaussians)) sigma *= 2.0;
  const double dsg = (double)GetIntegrationHalfStepMultiplier();
  while (curr_pos < (double)GetCoveredDistance()) {
    double si = dsg * sigma;                                   // first half of the interval at the current sigma
    while (!AddGaussianSampleStep(curr_pos + si + (dsg * sigma), sigma, &n_gaussians)) sigma *= 2.0;
    si += dsg * sigma;                                         // second half at the (possibly doubled) sigma
    IntervalsInfo iv; iv.s_position = curr_pos; iv.s_distance = si;
    data_cone_intervalsinfo.push_back(iv);
    curr_pos += si;
  }
  // ComputeAdditionalInfo (:318-386)
  ray3_adj_weight = vrb::dot(Z, ray3_axis[0]);
  ray7_adj_weight = vrb::dot(Z, ray7_axis[1]);
  for (size_t i = 0; i + 1 < data_cone_sectionsinfo.size(); ++i)
    if (data_cone_sectionsinfo[i].number_of_gaussians > data_cone_sectionsinfo[i + 1].number_of_gaussians) {
      vrb::SetError("ConeGaussianSampler: Wrong number of gaussians per section"); return false;
    }
  if (data_cone_sectionsinfo.size() != data_cone_intervalsinfo.size() + 1) { vrb::SetError("ConeGaussianSampler: Wrong number of sections and intervals"); return false; }
  IntervalsInfo last; last.s_position = data_cone_intervalsinfo.back().s_position + data_cone_intervalsinfo.back().s_distance; last.s_distance = 0.0;
  data_cone_intervalsinfo.push_back(last);
  for (size_t i = 0; i < data_cone_sectionsinfo.size(); ++i) {
    SectionInfo& s = data_cone_sectionsinfo[i];
    if (s.number_of_gaussians == 1) gaussian_samples_1++;
    else if (s.number_of_gaussians == 3) gaussian_samples_3++;
    else gaussian_samples_7++;
    s.d_integral = i == 0 ? s.sampled_gaussian_sigma * std::sqrt(2.0 * kPI) * 0.5 : data_cone_intervalsinfo[i - 1].s_distance * 0.5;
    double pr = IntegrateGaussian(s.sampled_gaussian_sigma, s.cone_radius);
    double Ac = kPI * s.cone_radius * s.cone_radius;
    double Ig = s.sampled_gaussian_sigma * std::sqrt(2.0 * kPI);
    s.amplitude = ((pr * pr) * (Ig * Ig)) / Ac;
    s.mip_map_level = std::log2(s.sampled_gaussian_sigma / min_sg_gaussian);
  }
  return true;
}

bool ConeGaussianSampler::GetConeSectionsInfoTex(std::vector<float>& out) {
  int n = GetNumberOfComputedConeSections();
  if (n <= 0) return false;
  out.resize((size_t)n * 4);
  for (int i = 0; i < n; ++i) {
    out[4 * i + 0] = (float)data_cone_intervalsinfo[i].s_distance;
    out[4 * i + 1] = (float)data_cone_sectionsinfo[i].mip_map_level;
    out[4 * i + 2] = (float)data_cone_sectionsinfo[i].d_integral;
    out[4 * i + 3] = (float)data_cone_sectionsinfo[i].amplitude;
  }
  return true;
}

vrb_cone_sampler ConeGaussianSampler::MakeUniformBlock(std::vector<float>& storage) {
  vrb_cone_sampler u;
  std::memset(&u, 0, sizeof(u));
  GetConeSectionsInfoTex(storage);
  u.sections = storage.data();
  u.n_sections = GetNumberOfComputedConeSections();
  u.integration_samples[0] = gaussian_samples_1; u.integration_samples[1] = gaussian_samples_3; u.integration_samples[2] = gaussian_samples_7;
  u.initial_step = (float)GetInitialStep();
  u.ray7_adj_weight = (float)GetRay7AdjacentWeight();
  u.ui_weight = (float)ui_weight_percentage;
  for (int i = 0; i < 3; ++i) { vrb::vec3 a = Get3ConeRayID(i); u.ray_axes[i][0] = a.x; u.ray_axes[i][1] = a.y; u.ray_axes[i][2] = a.z; }
  for (int i = 0; i < 7; ++i) { vrb::vec3 a = Get7ConeRayID(i); u.ray_axes[3 + i][0] = a.x; u.ray_axes[3 + i][1] = a.y; u.ray_axes[3 + i][2] = a.z; }
  return u;
}

// ------------------------------------------------------------------ ExtinctionCoefficientVolume
ExtinctionCoefficientVolume::ExtinctionCoefficientVolume() : base_level_sigma0(1.0f), map_specific_volume_resolution(true) {
  res[0] = res[1] = res[2] = 128;     // extcoefvolumegenerator.cpp:10-15
}
bool ExtinctionCoefficientVolume::BuildMipMappedTexture() {
  if (IsUsingCustomExtCoefVolumeResolution()) return CK(vrb_extcoef_build(CTX(), base_level_sigma0, res[0], res[1], res[2]));
  return CK(vrb_extcoef_build(CTX(), base_level_sigma0, 0, 0, 0));
}

// ------------------------------------------------------------------ RC1PConeTracingDirOcclusionShading
RC1PConeTracingDirOcclusionShading::RC1PConeTracingDirOcclusionShading()
    : m_u_step_size(0.5f), glsl_apply_occlusion(true), glsl_apply_shadow(false), type_of_shadow(0),
      m_cones_outdated(true), m_pyramid_outdated(true) {
  sampler_occlusion.SetUIWeightPercentage(0.350f);            // dosrcrenderer.cpp:47-49
  sampler_occlusion.SetConeHalfAngle(20.0f);
  sampler_occlusion.SetMaxGaussianPacking(ConeGaussianSampler::_3);
  sampler_shadow.SetUIWeightPercentage(1.0f);                 // :56-59
  sampler_shadow.SetConeHalfAngle(0.5f);
  sampler_shadow.SetMaxGaussianPacking(ConeGaussianSampler::_1);
  m_pre_illum_str_vol.SetActive(false);                        // :63-64
  m_pre_illum_str_vol.SetLightCacheResolution(32, 32, 32);
  std::memset(&m_cam, 0, sizeof(m_cam)); std::memset(&m_light, 0, sizeof(m_light)); std::memset(&m_prm, 0, sizeof(m_prm));
  vr_pixel_multiscaling_support = true;
}
RC1PConeTracingDirOcclusionShading::~RC1PConeTracingDirOcclusionShading() { Clean(); }
void RC1PConeTracingDirOcclusionShading::Clean() { BaseVolumeRenderer::Clean(); }

bool RC1PConeTracingDirOcclusionShading::GenerateExtCoefVolume() {
  if (!ext_coef_vol_gen.BuildMipMappedTexture()) return false;
  m_pyramid_outdated = false;
  SetOutdated();
  return true;
}
bool RC1PConeTracingDirOcclusionShading::GenerateConeSamples() {
  if (!sampler_occlusion.ComputeConeIntegrationSteps(ext_coef_vol_gen.GetBaseLevelGaussianSigma0())) return false;
  if (!sampler_shadow.ComputeConeIntegrationSteps(ext_coef_vol_gen.GetBaseLevelGaussianSigma0())) return false;
  std::vector<float> so, ss;
  vrb_cone_sampler uo = sampler_occlusion.MakeUniformBlock(so), us = sampler_shadow.MakeUniformBlock(ss);
  if (!CK(vrb_dos_set_cones(CTX(), &uo, &us))) return false;
  m_cones_outdated = false;
  SetOutdated();
  return true;
}

bool RC1PConeTracingDirOcclusionShading::Init(int swidth, int sheight) {
  if (IsBuilt()) Clean();
  if (m_ext_data_manager->GetCurrentVolumeTexture() == nullptr) return false;
  if (!UploadTransferFunction()) return false;     // RGBt for the march, RGBA for the pyramid
  vis::StructuredGridVolume* vol = m_ext_data_manager->GetCurrentStructuredVolume();
  sampler_occlusion.SetCoveredDistance((float)(vol->GetDiagonal() * 0.50f));   // dosrcrenderer.cpp:111-113
  sampler_shadow.SetCoveredDistance((float)(vol->GetDiagonal() * 0.75f));
  if (!GenerateExtCoefVolume()) return false;
  if (!GenerateConeSamples()) return false;
  vrb::dvec3 sv = vol->GetScale();
  m_u_step_size = float((0.5f / std::sqrt(3.0f)) * std::sqrt(sv.x * sv.x + sv.y * sv.y + sv.z * sv.z));
  Reshape(swidth, sheight);
  SetBuilt(true);
  SetOutdated();
  return true;
}

bool RC1PConeTracingDirOcclusionShading::Update(vis::Camera* camera) {
  if (m_pyramid_outdated && !GenerateExtCoefVolume()) return false;
  if (m_cones_outdated && !GenerateConeSamples()) return false;
  m_cam = MakeCameraBlock(camera);
  m_light = m_ext_rendering_parameters->MakeLightingBlock();
  m_light.apply_phong = (m_apply_gradient_shading && m_ext_data_manager->GetCurrentGradientTexture()) ? 1 : 0;
  m_prm.step_size = m_u_step_size;
  m_prm.apply_occlusion = glsl_apply_occlusion ? 1 : 0;
  m_prm.apply_shadow = glsl_apply_shadow ? 1 : 0;
  m_prm.type_of_shadow = type_of_shadow;
  // glm::cos(glm::pi<float>() * angle / 180.f) (dosrcrenderer.cpp:159)
  m_prm.spot_cos = std::cos(3.14159265358979323846264338327950288f * m_ext_rendering_parameters->GetSpotLightMaxAngle() / 180.f);
  m_prm.count_samples = 0;
  if (m_pre_illum_str_vol.IsActive()) {
    // PreComputeLightCache(camera) on every Update (dosrcrenderer.cpp:134-144,555-657)
    vrb::vec3 eye = camera->GetEye(), f, u, r;
    camera->GetCameraVectors(&f, &u, &r);
    float e3[3] = {eye.x, eye.y, eye.z}, u3[3] = {u.x, u.y, u.z};
    const int* res = m_pre_illum_str_vol.GetLightCacheResolution();
    if (!CK(vrb_dos_light_cache_build(CTX(), e3, u3, &m_light, &m_prm, res[0], res[1], res[2]))) return false;
  }
  return true;
}
void RC1PConeTracingDirOcclusionShading::Redraw() {
  if (m_pre_illum_str_vol.IsActive()) {
    // the rendering shader is _common_shaders/obj_ray_marching.comp while the cache is active (:674)
    vrb_obj_params op;
    op.step_size = m_u_step_size; op.apply_occlusion = m_prm.apply_occlusion; op.apply_shadow = m_prm.apply_shadow; op.count_samples = 0;
    CK(vrb_obj_march_render(CTX(), &m_cam, &m_light, &op));
    return;
  }
  CK(vrb_dos_render(CTX(), &m_cam, &m_light, &m_prm));
}
// the reference leaves this renderer without a parameter space (BaseVolumeRenderer::FillParameterSpace clears it);
// the step-size sweep of rc1prenderer.cpp:225-229 is offered here too so that the evaluation harness has something to vary
void RC1PConeTracingDirOcclusionShading::FillParameterSpace(ParameterSpace& pspace) {
  pspace.ClearParameterDimensions();
  pspace.AddParameterDimension(new ParameterRangeFloat("StepSize", &m_u_step_size, 0.2f, 2.0f, 0.1f));
}
bool RC1PConeTracingDirOcclusionShading::SetParameter(const std::string& name, double v) {
  if (name == "StepSize") m_u_step_size = (float)v;
  else if (name == "ApplyGradientShading") m_apply_gradient_shading = v != 0.0;
  else if (name == "ApplyOcclusion") glsl_apply_occlusion = v != 0.0;
  else if (name == "ApplyShadow") glsl_apply_shadow = v != 0.0;
  else if (name == "TypeOfShadow") type_of_shadow = (int)v;
  else if (name == "UsePreIllumination") m_pre_illum_str_vol.SetActive(v != 0.0);
  else if (name == "LightCacheResolution") m_pre_illum_str_vol.SetLightCacheResolution((int)v, (int)v, (int)v);
  else if (name == "OccConeHalfAngle") { sampler_occlusion.SetConeHalfAngle((float)v); m_cones_outdated = true; }
  else if (name == "OccMaxGaussianPacking") { sampler_occlusion.SetMaxGaussianPacking((int)v); m_cones_outdated = true; }
  else if (name == "OccUIWeight") { sampler_occlusion.SetUIWeightPercentage((float)v); m_cones_outdated = true; }
  else if (name == "OccCoveredDistance") { sampler_occlusion.SetCoveredDistance((float)v); m_cones_outdated = true; }
  else if (name == "SdwConeHalfAngle") { sampler_shadow.SetConeHalfAngle((float)v); m_cones_outdated = true; }
  else if (name == "SdwMaxGaussianPacking") { sampler_shadow.SetMaxGaussianPacking((int)v); m_cones_outdated = true; }
  else if (name == "SdwUIWeight") { sampler_shadow.SetUIWeightPercentage((float)v); m_cones_outdated = true; }
  else if (name == "SdwCoveredDistance") { sampler_shadow.SetCoveredDistance((float)v); m_cones_outdated = true; }
  else if (name == "BaseLevelGaussianSigma0") { ext_coef_vol_gen.SetBaseLevelGaussianSigma0((float)v); m_pyramid_outdated = m_cones_outdated = true; }
  else if (name == "UseCustomExtCoefVolumeResolution") { ext_coef_vol_gen.UseCustomExtCoefVolumeResolution(v != 0.0); m_pyramid_outdated = true; }
  else if (name == "CustomExtCoefVolumeResolution") { ext_coef_vol_gen.SetCustomExtCoefVolumeResolution((int)v, (int)v, (int)v); m_pyramid_outdated = true; }
  else return false;
  SetOutdated();
  return true;
}

// ---- extern "C": cone schedule for the CPU pinning tests (no GPU needed) --------------------------------------------
extern "C" int vrbh_cone_sampler_compute(float half_angle, float initial_step, int max_packing, float covered_distance,
                                         float d_sigma, float r_sigma, float ui_weight, double min_sigma,
                                         float* sections_out, int cap_sections, int counts[3], float axes[30], float adj[2]) {
  ConeGaussianSampler s;
  s.SetConeHalfAngle(half_angle); s.SetInitialStep(initial_step); s.SetMaxGaussianPacking(max_packing);
  s.SetCoveredDistance(covered_distance); s.SetIntegrationHalfStepMultiplier(d_sigma); s.SetGaussianSigmaLimitMultiplier(r_sigma);
  s.SetUIWeightPercentage(ui_weight);
  if (!s.ComputeConeIntegrationSteps(min_sigma)) return -2;
  std::vector<float> st;
  vrb_cone_sampler u = s.MakeUniformBlock(st);
  if (u.n_sections > cap_sections) return -1;
  std::memcpy(sections_out, st.data(), st.size() * sizeof(float));
  for (int i = 0; i < 3; ++i) counts[i] = u.integration_samples[i];
  std::memcpy(axes, u.ray_axes, sizeof(float) * 30);
  adj[0] = (float)s.GetRay3AdjacentWeight(); adj[1] = u.ray7_adj_weight;
  return u.n_sections;
}
