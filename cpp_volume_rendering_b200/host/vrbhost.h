// vrbhost.h -- GL-free C++ host side of vrb200: the reference's plugin-facing classes re-hosted headless.
//
// Same class names, method names, argument meaning and call order as lquatrin/cpp_volume_rendering so that a
// renderer written against the reference reads the same here; GL objects are replaced by handles into the C ABI
// (include/vrb200.h).  Citations are relative to the reference tree:
//   BaseVolumeRenderer          cppvolrend/volrenderbase.h:25-98, volrenderbase.cpp
//   RenderingManager            cppvolrend/renderingmanager.h:29-167 (renderer-facing calls only; no UI)
//   vis::DataManager            libs/volvis_utils/datamanager.{h,cpp}
//   vis::RenderingParameters    libs/volvis_utils/renderingparameters.{h,cpp}
//   vis::TransferFunction1D     libs/volvis_utils/transferfunction1d.{h,cpp}
//   vis::StructuredGridVolume   libs/volvis_utils/structuredgridvolume.{h,cpp}
//   vis::VolumeReader / TransferFunctionReader   libs/volvis_utils/reader.cpp
//   vis::Camera / CameraStateList / LightSourceList   libs/vis_utils/camera.cpp, libs/volvis_utils/*list.cpp
// Nothing in here calls exit(): errors surface as false / nullptr + vrb::LastError().
#pragma once
#include <cmath>
#include <cstdint>
#include <string>
#include <stdexcept>
#include <cmath>
#include <fstream>
#include <vector>
#include "../../include/vrb200.h"

namespace vrb {
struct vec3 { float x = 0, y = 0, z = 0; vec3() {} vec3(float a, float b, float c) : x(a), y(b), z(c) {} explicit vec3(float a) : x(a), y(a), z(a) {} };
struct dvec3 { double x = 0, y = 0, z = 0; dvec3() {} dvec3(double a, double b, double c) : x(a), y(b), z(c) {} };
struct vec4 { float x = 0, y = 0, z = 0, w = 0; };
struct dvec4 { double r = 0, g = 0, b = 0, a = 0; };
struct mat4 { float m[16]; };   // column major, m[4*col + row]
inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline vec3 normalize(vec3 a) { float r = 1.0f / std::sqrt(dot(a, a)); return a * r; }
mat4 lookAt(vec3 eye, vec3 center, vec3 up);
const std::string& LastError();
void SetError(const std::string& s);

// The analogue of "the GL context": one vrb_ctx per process/GPU, created by RenderingManager::InitGL().
class Device {
 public:
  static Device* Instance();
  bool Init(int cuda_device);
  void Shutdown();
  vrb_ctx* ctx() { return m_ctx; }
  bool ok() const { return m_ctx != nullptr; }
 private:
  vrb_ctx* m_ctx = nullptr;
};
// Opaque stand-ins for gl::Texture3D* / gl::Texture1D* return values (non-null == resident on the device).
struct DeviceVolumeTexture { int w = 0, h = 0, d = 0; };
struct DeviceTransferFunctionTexture { int n = 0; };
struct DeviceGradientTexture { int w = 0, h = 0, d = 0; };
}  // namespace vrb

// -------------------------------------------------------------------------------------------------------------
namespace vis {
using vrb::vec3; using vrb::dvec3; using vrb::vec4; using vrb::dvec4; using vrb::mat4;

enum GRID_VOLUME_DATA_TYPE { STRUCTURED = 0, UNSTRUCTURED = 1, NONE_GRID = 2 };

class TransferControlPoint {
 public:
  TransferControlPoint(double r, double g, double b, int isovalue);
  TransferControlPoint(double alpha, int isovalue);
  vec4 m_color;
  int m_isoValue;
};

class TransferFunction {
 public:
  virtual ~TransferFunction() {}
  virtual const char* GetNameClass() = 0;
  virtual vec4 Get(double value, double max_input_value = -1.0) = 0;
  virtual float GetOpc(double, double = -1.0) { return -1.0f; }
  virtual float GetOpcN(double) { return -1.0f; }
  virtual float GetExt(double, double = -1.0) { return -1.0f; }
  virtual float GetExtN(double) { return -1.0f; }
  // GenerateTexture_1D_RGBA / _RGBt: fill the GL_FLOAT client arrays (n x 4) the reference hands to glTexImage1D.
  virtual bool GenerateTexture_1D_RGBA(std::vector<float>& out) { (void)out; return false; }
  virtual bool GenerateTexture_1D_RGBt(std::vector<float>& out) { (void)out; return false; }
  virtual int GetTextureSize() { return 0; }
  std::string GetName() { return m_name; }
  void SetName(std::string n) { m_name = n; }
  double ExtinctionToMaterialOpacity(float extinction) { return 1.0 - (double)std::exp(-extinction); }   // fp32 exp (glm::exp(float)), double subtraction
  double MaterialOpacityToExtinction(float opacity) { return std::log(1.0 / (1.0 - (double)opacity)); }
 protected:
  std::string m_name;
};

class TransferFunction1D : public TransferFunction {
 public:
  explicit TransferFunction1D(int max_value = 255);
  ~TransferFunction1D() override;
  const char* GetNameClass() override;
  vec4 Get(double value, double max_data_value = -1.0) override;
  float GetOpc(double value, double max_input_value = -1.0) override;
  float GetOpcN(double normalized_value) override;
  float GetExt(double value, double max_input_value = -1.0) override;
  float GetExtN(double normalized_value) override;
  bool GenerateTexture_1D_RGBA(std::vector<float>& out) override;
  bool GenerateTexture_1D_RGBt(std::vector<float>& out) override;
  int GetTextureSize() override { return max_density + 1; }
  void SetExtinctionCoefficientInput(bool s);
  void AddRGBControlPoint(TransferControlPoint rgb);
  void AddAlphaControlPoint(TransferControlPoint alpha);
  void ClearControlPoints();
  void Build();
  bool m_built;
 private:
  void BuildLinear();
  std::vector<TransferControlPoint> m_cpt_rgb, m_cpt_alpha;
  std::vector<dvec4> m_transferfunction;
  int max_density;
  bool extinction_coef_type;
};

class TransferFunctionReader {
 public:
  TransferFunction* ReadTransferFunction(std::string file);
 private:
  TransferFunction* readtf1d(std::string file);
};

enum DataStorageSize : unsigned int { UNKNOWN = 0, _8_BITS = 1, _16_BITS = 2 };

class StructuredGridVolume {
 public:
  StructuredGridVolume(std::string name = "Unknown", unsigned int width = 0, unsigned int height = 0, unsigned int depth = 0);
  ~StructuredGridVolume();
  std::string GetName() { return m_name; }
  void SetName(std::string n) { m_name = n; }
  unsigned int GetWidth() { return m_width; }
  unsigned int GetHeight() { return m_height; }
  unsigned int GetDepth() { return m_depth; }
  double GetScaleX() { return m_scalex; }
  double GetScaleY() { return m_scaley; }
  double GetScaleZ() { return m_scalez; }
  dvec3 GetScale() { return dvec3(m_scalex, m_scaley, m_scalez); }
  void SetScale(double sx, double sy, double sz) { m_scalex = sx; m_scaley = sy; m_scalez = sz; }
  double GetDiagonal();
  bool IsOutOfBoundary(int x, int y, int z);
  // takes ownership of input_vol_data (allocated with new unsigned char[] / new unsigned short[])
  void SetArrayData(void* input_vol_data, DataStorageSize dss);
  void* GetArrayData() { return m_voxel_values; }
  DataStorageSize GetDataStorageSize() { return m_data_storage_size; }
  double GetNormalizedSample(int x, int y, int z);
  unsigned long long CheckSum();
  double GetMaxDensity();
 private:
  std::string m_name;
  unsigned int m_width, m_height, m_depth;
  double m_scalex, m_scaley, m_scalez;
  DataStorageSize m_data_storage_size;
  void* m_voxel_values;
};

class VolumeReader {
 public:
  // dispatch on extension: .raw (name.<bytes>.<W>x<H>x<D>.raw), .syn, .pvm (PVM/PVM2/PVM3, plain or DDS v3d/v3e)
  StructuredGridVolume* ReadStructuredVolume(std::string filepath);
 private:
  StructuredGridVolume* readraw(std::string filepath);
  StructuredGridVolume* readsyn(std::string filepath);
  StructuredGridVolume* readpvm(std::string filepath);
};
// host_pvm.cpp: PVM3 writer, dds_version 0 = plain, 1 = "DDS v3d", 2 = "DDS v3e" (the reference has no active writer)
bool WritePvm(const std::string& path, const void* voxels, int w, int h, int d, int bytes_per_voxel, const double scale[3], int dds_version);

class CameraData {
 public:
  CameraData();
  std::string cam_setup_name;
  int c_type;
  vec3 eye, center, up;
  float field_of_view_y, aspect_ratio, z_near, z_far;
};

class Camera {
 public:
  enum CAMERA_BEHAVIOUR { FLIGHT = 0, ARCBALL = 1 };
  Camera();
  mat4 LookAt();
  vec3 GetDir();
  vec3 GetEye();
  void UpdateAspectRatio(float w, float h);
  float GetAspectRatio();
  float GetFovY();
  float GetTanFovY();
  void SetData(CameraData* data);
  void GetCameraVectors(vec3* cforward, vec3* cup, vec3* cright);
 private:
  CameraData c_data;
  float radius;
};

class CameraStateList {
 public:
  bool ReadCameraStates(std::string filepath);
  int NumberOfCameraStates();
  CameraData* GetCameraState(unsigned int idx);
 private:
  std::vector<CameraData> m_vec_camera_data;
};

class LightSourceData {
 public:
  LightSourceData();
  vec3 color, specular, position, x_axis, y_axis, z_axis;
  float spot_light_angle, spot_light_angle_rad, energy_density;
};
class LightSourceListItem {
 public:
  std::string l_name;
  std::vector<LightSourceData> m_lightsources;
};
class LightSourceList {
 public:
  bool ReadLightSourceLists(std::string filepath);
  int NumberOfLists();
  LightSourceListItem* GetList(unsigned int idx);
 private:
  std::vector<LightSourceListItem> m_vec_lsource_lists;
};

class RenderingParameters {
 public:
  RenderingParameters();
  Camera* GetCamera() { return &s_camera; }
  void SetPhongParameters(float amb, float diff, float spec, float shini);
  float GetBlinnPhongKambient() { return m_blinnphong_ka; }
  float GetBlinnPhongKdiffuse() { return m_blinnphong_kd; }
  float GetBlinnPhongKspecular() { return m_blinnphong_ks; }
  float GetBlinnPhongNshininess() { return m_blinnphong_shininess; }
  void EraseAllLightSources() { m_vec_light_sources.clear(); }
  void CreateNewLightSource(LightSourceData lsd) { m_vec_light_sources.push_back(lsd); }
  int GetNumberOfLightSources() { return (int)m_vec_light_sources.size(); }
  vec3 GetLightSourceSpecular();
  void SetBlinnPhongLightingPosition(vec3 lightpos);
  vec3 GetBlinnPhongLightingPosition();
  void SetBlinnPhongLightSourceCameraVectors(vec3 lcamforward, vec3 lcamup, vec3 lcamright);
  vec3 GetBlinnPhongLightSourceCameraForward();
  vec3 GetBlinnPhongLightSourceCameraUp();
  vec3 GetBlinnPhongLightSourceCameraRight();
  float GetSpotLightMaxAngle();
  void SetScreenSize(int width, int height);
  int GetScreenWidth() { return screen_width; }
  int GetScreenHeight() { return screen_height; }
  // the lighting block of the C ABI filled from the accessors above
  vrb_lighting MakeLightingBlock();
 private:
  LightSourceData& cur();
  Camera s_camera;
  int screen_width, screen_height;
  float m_blinnphong_ka, m_blinnphong_kd, m_blinnphong_ks, m_blinnphong_shininess;
  std::vector<LightSourceData> m_vec_light_sources;
  int m_current_light_source_id;
};

struct DataReference { std::string path, name; };

class DataManager {
 public:
  // libs/volvis_utils/datamanager.h:64-69
  enum STRUCTURED_GRADIENT_TYPE : unsigned int { SOBEL_FELDMAN_FILTER = 0, FINITE_DIFERENCES = 1, COMPUTE_SHADER_SOBEL = 2, NONE_GRADIENT = 3 };
  DataManager();
  ~DataManager();
  void SetPathToData(std::string s_path_to_data) { m_path_to_data = s_path_to_data; }
  // reads #list_structured_datasets / #list_transfer_functions, loads entry 0 of each and uploads them
  bool ReadData();
  int GetNumberOfStructuredDatasets() { return (int)stored_structured_datasets.size(); }
  int GetNumberOfTransferFunctions() { return (int)stored_transfer_functions.size(); }
  bool SetCurrentInputVolume(int id);
  bool SetCurrentTransferFunction(int id);
  // headless additions: adopt in-memory data (ownership passes to the manager)
  bool SetStructuredVolume(StructuredGridVolume* vol);
  bool SetTransferFunction(TransferFunction* tf);
  GRID_VOLUME_DATA_TYPE GetInputVolumeDataType() { return STRUCTURED; }
  StructuredGridVolume* GetCurrentStructuredVolume() { return curr_vr_volume; }
  TransferFunction* GetCurrentTransferFunction() { return curr_vr_transferfunction; }
  vrb::DeviceVolumeTexture* GetCurrentVolumeTexture() { return curr_tex_volume.w ? &curr_tex_volume : nullptr; }
  // gradient texture (datamanager.cpp:191-222,326-352,499-612): NONE_GRADIENT by default (:27); the generators run on
  // the device (vrb_gradient_build) and the texture lives in the vrb_ctx
  vrb::DeviceGradientTexture* GetCurrentGradientTexture() { return curr_tex_gradient.w ? &curr_tex_gradient : nullptr; }
  void DeleteGradientData();
  bool UpdateStructuredGradientTexture();
  int GetCurrentGradientGenerationTypeID() { return (int)curr_gradient_comp_model; }
  int GetGradientIndex(STRUCTURED_GRADIENT_TYPE sgt) { return (int)sgt <= 2 ? (int)sgt : 3; }
  bool SetCurrentGradient(int idx);
  std::string GetGradientName(STRUCTURED_GRADIENT_TYPE sgt);
  std::string CurrentGradientName();
  std::vector<std::string> GetGradientGenerationTypeStrList();
  std::string GetCurrentDataName();
  std::string GetCurrentTransferFunctionName();
 private:
  bool ReadList(const char* list_name, std::vector<DataReference>& out);
  bool GenerateStructuredVolumeTexture();
  bool GenerateStructuredGradientTexture();
  STRUCTURED_GRADIENT_TYPE curr_gradient_comp_model;
  vrb::DeviceGradientTexture curr_tex_gradient;
  std::string m_path_to_data;
  std::vector<DataReference> stored_structured_datasets, stored_transfer_functions;
  int curr_volume_index, curr_transferfunction_index;
  StructuredGridVolume* curr_vr_volume;
  TransferFunction* curr_vr_transferfunction;
  vrb::DeviceVolumeTexture curr_tex_volume;
};

// Output image of a renderer (libs/vis_utils/renderoutputframe.{h,cpp}); lives in the vrb_ctx.
enum IMAGE_FILTER_KERNEL : unsigned int {            // libs/vis_utils/filters/utils.hpp:8-15
  K1_BOX = 0, K2_HAT = 1, K4_CATMULL_ROM = 2, K4_MITCHELL_NETRAVALI = 3, K4_CARDINAL_BSPLINE_3 = 4, K4_CARDINAL_OMOMS3 = 5 };
class RenderFrameToScreen {
 public:
  void Clean() { m_mw = m_mh = 0; m_filtered = false; }
  bool UpdateScreenResolution(int s_w, int s_h);
  // renderoutputframe.cpp:89-145: multipliers 0 keep the ones set with SetMultiResolutionScreenMultiplier
  bool UpdateScreenResolutionMultiScaling(int s_w, int s_h, int mw = 0, int mh = 0);
  void SetMultiResolutionScreenMultiplier(int mw, int mh) { m_mw = mw; m_mh = mh; }
  bool ClearTexture();
  int GetWidth() { return m_w; }        // of the RENDERED frame (m_screen_output)
  int GetHeight() { return m_h; }
  // the image-space passes of MultiSampleRedraw / DownScalingRedraw / UpScalingRedraw (:265-540), minus the GL draw
  bool DrawMultiSampleHigherResolutionMode();
  bool DrawHigherResolutionWithDownScale();
  bool DrawLowerResolutionWithUpScale();
  void SetImageKernelFilter(unsigned int k_i) { m_kernel_filter = k_i; }
  unsigned int GetImageKernelFilter() { return m_kernel_filter; }
  // glGetTexImage(GL_RGBA, GL_FLOAT) (renderingmanager.cpp:637-640) of what Draw() puts on screen: the filtered frame in
  // the multi-scaling modes, the rendered frame otherwise
  bool ReadPixelsRGBA32F(std::vector<float>& out);
 private:
  int m_w = 0, m_h = 0, m_sw = 0, m_sh = 0;
  int m_mw = 0, m_mh = 0;
  bool m_filtered = false;
  unsigned int m_kernel_filter = K2_HAT;
};
}  // namespace vis

// -------------------------------------------------------------------------------------------------------------
// ParameterSpace (cppvolrend/utils/parameterspace.{h,cpp}): the parameter sweep of the evaluation harness (Evaluation.md).
class ParameterRangeBase {
 public:
  explicit ParameterRangeBase(const std::string& name) : m_name(name) {}
  virtual ~ParameterRangeBase() {}
  virtual void Start() = 0;
  virtual bool End() const = 0;
  virtual void Incr() = 0;
  virtual void SaveCurrentValue() = 0;
  virtual void RestoreCurrentValue() = 0;
  virtual int NumSteps() const = 0;
  virtual std::string GetValueStr() const = 0;
  const std::string& GetName() const { return m_name; }
 protected:
  std::string m_name;
};
// parameterspace.h:63-160: the range drives the variable `param` points at; inverted or non-advancing ranges are refused
// (the reference `throw;`s, i.e. terminates; here std::invalid_argument)
template <typename T>
class ParameterRangeNumeric : public ParameterRangeBase {
 public:
  ParameterRangeNumeric(const std::string& name, T* param, const T start, const T end, const T incr)
      : ParameterRangeBase(name), m_start(0), m_end(2), m_incr(1), m_previousvalue(0), m_curr(nullptr) { Set(start, end, incr, param); }
  void Set(const T start, const T end, const T incr, T* param) {
    if (start > end || incr <= 0 || !param) throw std::invalid_argument("ParameterRangeNumeric: inverted range, non-positive increment or null parameter");
    m_start = start; m_end = end; m_incr = incr; m_curr = param;
  }
  void Start() override { *m_curr = m_start; }
  bool End() const override { return (*m_curr > m_end); }
  void Incr() override { *m_curr += m_incr; }
  void SaveCurrentValue() override { m_previousvalue = *m_curr; }
  void RestoreCurrentValue() override { *m_curr = m_previousvalue; }
  int NumSteps() const override { return 1 + (int)std::ceil((m_end - m_start) / m_incr); }
  std::string GetValueStr() const override { return std::to_string(*m_curr); }
 protected:
  T m_start, m_end, m_incr, m_previousvalue;
  T* m_curr;
};
using ParameterRangeFloat = ParameterRangeNumeric<float>;
using ParameterRangeDouble = ParameterRangeNumeric<double>;
using ParameterRangeInt = ParameterRangeNumeric<int>;
class ParameterSpace {
 public:
  ParameterSpace() : m_numsamples_cached(0) {}
  virtual ~ParameterSpace() { ClearParameterDimensions(); }
  void AddParameterDimension(ParameterRangeBase* param) { m_dimensions.push_back(param); ComputeNumSamplePoints(); }   // takes ownership
  void ClearParameterDimensions() { for (auto* p : m_dimensions) delete p; m_dimensions.clear(); ComputeNumSamplePoints(); }
  int GetNumDimensions() const { return (int)m_dimensions.size(); }
  const std::string& GetDimensionName(const int idx) const { return m_dimensions[idx]->GetName(); }
  const std::string GetDimensionValue(const int idx) const { return m_dimensions[idx]->GetValueStr(); }
  int GetNumSamplePoints() const { return m_numsamples_cached; }
  void StartEvaluation() { for (auto* p : m_dimensions) { p->SaveCurrentValue(); p->Start(); } }
  void EndEvaluation() { for (auto* p : m_dimensions) p->RestoreCurrentValue(); }
  bool IncrEvaluation();                                   // last dimension first; false at the end of the space
 protected:
  int ComputeNumSamplePoints();
  std::vector<ParameterRangeBase*> m_dimensions;
 private:
  int m_numsamples_cached;
};
bool ParameterSpaceTest();                                 // parameterspace.cpp:103-149

class BaseVolumeRenderer {
 public:
  enum MULTISCALING { SINGLE_RAY_PER_PIXEL = 0, MULTIPLE_RAYS_PER_PIXEL = 1, DOWN_SCALING_RENDER = 2, UP_SCALING_RENDER = 3 };
  BaseVolumeRenderer();
  virtual ~BaseVolumeRenderer();
  void SetExternalResources(vis::DataManager* data_mgr, vis::RenderingParameters* rdr_prm);
  virtual const char* GetName() = 0;
  virtual const char* GetAbbreviationName() = 0;
  virtual void Clean();
  virtual void ReloadShaders();
  virtual bool Init(int shader_width, int shader_height) = 0;
  virtual bool Update(vis::Camera* camera) = 0;
  virtual void Redraw();
  virtual void MultiSampleRedraw();
  virtual void DownScalingRedraw();
  virtual void UpScalingRedraw();
  virtual void Reshape(int w, int h);
  virtual void SetImGuiComponents();
  virtual vis::GRID_VOLUME_DATA_TYPE GetDataTypeSupport() = 0;
  virtual void FillParameterSpace(ParameterSpace& pspace);
  void PrepareRender(vis::Camera* camera);
  virtual void SetOutdated();
  bool IsOutdated();
  bool IsBuilt();
  bool IsPixelMultiScalingSupported();
  int GetCurrentMultiScalingMode();
  void SetCurrentMultiScalingMode(int f);
  // GetScreenTextureID() analogue: device pointer of the RGBA16F image
  virtual void* GetScreenTextureDevicePtr();
  // float RGBA read-back of the output texture, row 0 = bottom
  bool ReadOutputRGBA32F(std::vector<float>& out) { return m_rdr_frame_to_screen.ReadPixelsRGBA32F(out); }
  // headless addition: set a named parameter (what the ImGui widgets do to the members); false if unknown
  virtual bool SetParameter(const std::string& name, double value);
  // headless stand-in for AddImGuiMultiSampleOptions (volrenderbase.cpp:121-197): "MultiScalingMode" 0..3, "ImageKernelFilter" 0..5
  bool SetMultiScalingOption(const std::string& name, double value);
 protected:
  void SetBuilt(bool b_built);
  bool UploadTransferFunction();   // GenerateTexture_1D_RGBt + _RGBA -> vrb_tf_upload
  vrb_camera MakeCameraBlock(vis::Camera* camera);
  bool vr_built, vr_outdated, vr_pixel_multiscaling_support;
  int vr_pixel_multiscaling_mode;
  vis::DataManager* m_ext_data_manager;
  vis::RenderingParameters* m_ext_rendering_parameters;
  vis::RenderFrameToScreen m_rdr_frame_to_screen;
};

// cppvolrend/structured/rc1pass/rc1prenderer.{h,cpp}
class RayCasting1Pass : public BaseVolumeRenderer {
 public:
  RayCasting1Pass();
  ~RayCasting1Pass() override;
  const char* GetName() override { return "1-Pass - Ray Casting"; }
  const char* GetAbbreviationName() override { return "s_1rc"; }
  vis::GRID_VOLUME_DATA_TYPE GetDataTypeSupport() override { return vis::STRUCTURED; }
  void Clean() override;
  bool Init(int swidth, int sheight) override;
  bool Update(vis::Camera* camera) override;
  void Redraw() override;
  void FillParameterSpace(ParameterSpace& pspace) override;
  bool SetParameter(const std::string& name, double value) override;
 private:
  bool m_has_tf;
  float m_u_step_size;
  bool m_apply_gradient_shading;
  bool m_skip_empty;
  vrb_lighting m_light;
  vrb_camera m_cam;
};

// cppvolrend/structured/rc1pisoadapt/rc1pisoadaptrenderer.{h,cpp}: adaptive-step isosurface ray caster
class RayCasting1PassIsoAdapt : public BaseVolumeRenderer {
 public:
  RayCasting1PassIsoAdapt();
  ~RayCasting1PassIsoAdapt() override;
  const char* GetName() override { return "1-Pass - Isosurface Raycaster Adaptive"; }
  const char* GetAbbreviationName() override { return "iso"; }
  vis::GRID_VOLUME_DATA_TYPE GetDataTypeSupport() override { return vis::STRUCTURED; }
  void Clean() override;
  bool Init(int swidth, int sheight) override;
  bool Update(vis::Camera* camera) override;
  void Redraw() override;
  void FillParameterSpace(ParameterSpace& pspace) override;
  bool SetParameter(const std::string& name, double value) override;
 protected:
  float m_u_isovalue, m_u_step_size_small, m_u_step_size_large, m_u_step_size_range;
  float m_u_color[4];
  bool m_apply_gradient_shading;
 private:
  vrb_camera m_cam; vrb_lighting m_light; vrb_iso_params m_prm;
};

// cppvolrend/utils/preillumination.{h,cpp}: the optional object-space light cache (inactive by default, 32^3 in the
// renderers).  The texture itself lives in the context (vrb_*_light_cache_build); this class keeps the reference's state.
class PreIlluminationStructuredVolume {
 public:
  explicit PreIlluminationStructuredVolume(int n_channels = 2) : m_active(false), m_n_channels(n_channels) { m_res[0] = m_res[1] = m_res[2] = 8; }
  bool IsActive() const { return m_active; }
  void SetActive(bool f) { m_active = f; }
  void SetLightCacheResolution(int w, int h, int d) { m_res[0] = w; m_res[1] = h; m_res[2] = d; }
  const int* GetLightCacheResolution() const { return m_res; }
 private:
  bool m_active; int m_n_channels; int m_res[3];
};

// cppvolrend/structured/rc1pextbsd/ebsrenderer.{h,cpp}
class RC1PExtinctionBasedShading : public BaseVolumeRenderer {
 public:
  RC1PExtinctionBasedShading();
  ~RC1PExtinctionBasedShading() override;
  const char* GetName() override { return "1-Pass - Extinction-based Shading"; }
  const char* GetAbbreviationName() override { return "s_1rc_eb"; }
  vis::GRID_VOLUME_DATA_TYPE GetDataTypeSupport() override { return vis::STRUCTURED; }
  void Clean() override;
  bool Init(int swidth, int sheight) override;
  bool Update(vis::Camera* camera) override;
  void Redraw() override;
  void FillParameterSpace(ParameterSpace& pspace) override;
  bool SetParameter(const std::string& name, double value) override;
 private:
  bool GenerateExtinctionSAT3DTex(vis::StructuredGridVolume* vol, vis::TransferFunction* tf);
  bool m_has_tf, m_has_sat;
  float m_u_step_size;
  bool apply_ambient_occlusion; int ambient_occlusion_shells; float ambient_occlusion_radius;
  bool apply_directional_shadows; int dir_shadow_cone_samples; float dir_shadow_cone_angle;
  float dir_shadow_sample_interval, dir_shadow_initial_step, dir_shadow_user_interface_weight, dir_cone_max_distance;
  int type_of_shadow;
  PreIlluminationStructuredVolume m_pre_illum_str_vol;
  vrb_camera m_cam; vrb_lighting m_light; vrb_ebs_params m_prm;
  bool m_apply_gradient_shading = false;   // "Apply Gradient Shading" checkbox; ApplyPhongShading = this && gradient texture
};

// cppvolrend/structured/rc1pdosct/conegaussiansampler.{h,cpp}: section schedule of one cone (host-side doubles).
class ConeGaussianSampler {
 public:
  struct SectionInfo { int number_of_gaussians; double distance_from_origin, cone_radius, sampled_gaussian_sigma, d_integral, mip_map_level, amplitude; };
  struct IntervalsInfo { double s_position, s_distance; };
  enum CONEPACKING { _1 = 0, _3 = 1, _7 = 2 };
  ConeGaussianSampler();
  float GetConeHalfAngle() { return cone_half_angle; }
  void SetConeHalfAngle(float angle);
  float GetInitialStep() { return initial_step; }
  void SetInitialStep(float istep);
  int GetMaxGaussianPackingInt() { return (int)max_gaussian_packing; }
  void SetMaxGaussianPacking(int p);
  float GetCoveredDistance() { return covered_distance; }
  void SetCoveredDistance(float midist);
  float GetIntegrationHalfStepMultiplier() { return d_sigma; }
  void SetIntegrationHalfStepMultiplier(float v);
  float GetGaussianSigmaLimitMultiplier() { return r_sigma; }
  void SetGaussianSigmaLimitMultiplier(float s);
  void SetUIWeightPercentage(float a) { ui_weight_percentage = a; }
  int GetNumberOfComputedConeSections() { return (int)data_cone_sectionsinfo.size(); }
  // the GL_FLOAT client array of GetConeSectionsInfoTex: n x [interval distance, mip level, d_integral, amplitude]
  bool GetConeSectionsInfoTex(std::vector<float>& out);
  std::vector<SectionInfo> GetConeSectionsInfoVec() { return data_cone_sectionsinfo; }
  bool ComputeConeIntegrationSteps(double min_sg_gaussian);   // false on the invariant violations the reference exit()s on
  double GetRay3AdjacentWeight() { return ray3_adj_weight; }
  double GetRay7AdjacentWeight() { return ray7_adj_weight; }
  vrb::vec3 Get3ConeRayID(int i) { return (i < 0 || i > 2) ? vrb::vec3(0.0f) : ray3_axis[i]; }
  vrb::vec3 Get7ConeRayID(int i) { return (i < 0 || i > 6) ? vrb::vec3(0.0f) : ray7_axis[i]; }
  // the uniform block BindCone*Uniforms uploads (dosrcrenderer.cpp:805-985); `storage` keeps the section array alive
  vrb_cone_sampler MakeUniformBlock(std::vector<float>& storage);
  int gaussian_samples_1, gaussian_samples_3, gaussian_samples_7;
  vrb::vec3 ray3_axis[3], ray7_axis[7];
  double ray3_adj_weight, ray7_adj_weight;
  float cone_half_angle, initial_step;
  CONEPACKING max_gaussian_packing;
  float covered_distance, ui_weight_percentage, d_sigma, r_sigma;
 private:
  bool AddGaussianSampleStep(double curr_pos, double sg_gaussian, int* n_gaussians);
  double IntegrateGaussian(double sdev_gaussian, double cone_radius);
  std::vector<SectionInfo> data_cone_sectionsinfo;
  std::vector<IntervalsInfo> data_cone_intervalsinfo;
};
vrb::vec3 RodriguesRotation(vrb::vec3 v, float teta, vrb::vec3 k);   // libs/math_utils/utils.cpp:149-156

// cppvolrend/structured/rc1pdosct/extcoefvolumegenerator.{h,cpp}
class ExtinctionCoefficientVolume {
 public:
  ExtinctionCoefficientVolume();
  // BuildMipMappedTexture: the pyramid is built on the device from the resident volume + RGBA transfer function
  bool BuildMipMappedTexture();
  float GetBaseLevelGaussianSigma0() { return base_level_sigma0; }
  void SetBaseLevelGaussianSigma0(float s) { base_level_sigma0 = s; }
  bool IsUsingCustomExtCoefVolumeResolution() { return map_specific_volume_resolution; }
  void UseCustomExtCoefVolumeResolution(bool b) { map_specific_volume_resolution = b; }
  void SetCustomExtCoefVolumeResolution(int w, int h, int d) { res[0] = w; res[1] = h; res[2] = d; }
  void GetCustomExtCoefVolumeResolution(int out[3]) { out[0] = res[0]; out[1] = res[1]; out[2] = res[2]; }
 private:
  float base_level_sigma0;
  bool map_specific_volume_resolution;
  int res[3];
};

// cppvolrend/structured/rc1pdosct/dosrcrenderer.{h,cpp}
class RC1PConeTracingDirOcclusionShading : public BaseVolumeRenderer {
 public:
  RC1PConeTracingDirOcclusionShading();
  ~RC1PConeTracingDirOcclusionShading() override;
  const char* GetName() override { return "1-Pass - Directional Occlusion Ray Casting - Cone Tracing"; }
  const char* GetAbbreviationName() override { return "s_1rc_dos"; }
  vis::GRID_VOLUME_DATA_TYPE GetDataTypeSupport() override { return vis::STRUCTURED; }
  void Clean() override;
  bool Init(int swidth, int sheight) override;
  bool Update(vis::Camera* camera) override;
  void Redraw() override;
  void FillParameterSpace(ParameterSpace& pspace) override;
  bool SetParameter(const std::string& name, double value) override;
 private:
  bool GenerateExtCoefVolume();
  bool GenerateConeSamples();
  float m_u_step_size;
  bool glsl_apply_occlusion, glsl_apply_shadow;
  int type_of_shadow;
  bool m_cones_outdated, m_pyramid_outdated;
  ConeGaussianSampler sampler_occlusion, sampler_shadow;
  ExtinctionCoefficientVolume ext_coef_vol_gen;
  PreIlluminationStructuredVolume m_pre_illum_str_vol;
  vrb_camera m_cam; vrb_lighting m_light; vrb_dos_params m_prm;
  bool m_apply_gradient_shading = false;   // "Apply Gradient Shading" checkbox; ApplyPhongShading = this && gradient texture
};

// cppvolrend/structured/rc1pcrtgt/crtgtrenderer.{h,cpp}
class RC1PConeLightGroundTruthSteps : public BaseVolumeRenderer {
 public:
  RC1PConeLightGroundTruthSteps();
  ~RC1PConeLightGroundTruthSteps() override;
  const char* GetName() override { return "1-Pass - Cone Ray Tracing - Ground Truth - Steps"; }
  const char* GetAbbreviationName() override { return "s_1rc_gt_c"; }
  vis::GRID_VOLUME_DATA_TYPE GetDataTypeSupport() override { return vis::STRUCTURED; }
  void Clean() override;
  bool Init(int shader_width, int shader_height) override;
  bool Update(vis::Camera* camera) override;
  void Redraw() override;     // one call = the converged frame of the reference's progressive RedrawFrameTexture loop
  bool SetParameter(const std::string& name, double value) override;
  // the ray tables Update() draws (crtgtrenderer.cpp:131-187), n x 3 floats, before the RGB16F rounding
  static void GenerateRayTables(int n_occ, float occ_aperture_deg, int n_sdw, float sdw_aperture_deg,
                                std::vector<float>& occ, std::vector<float>& sdw);
 private:
  float m_u_step_size, m_u_light_ray_initial_step, m_u_light_ray_step_size;
  bool m_light_parameters_outdated;
  bool m_apply_occlusion; int m_occ_num_rays_sampled; float m_occ_cone_aperture_angle, m_occ_cone_distance_eval;
  bool m_apply_shadows; int m_sdw_num_rays_sampled; float m_sdw_cone_aperture_angle, m_sdw_cone_distance_eval;
  int m_shadow_type;
  vrb_camera m_cam; vrb_lighting m_light; vrb_gt_params m_prm;
  bool m_apply_gradient_shading = false;   // "Apply Gradient Shading" checkbox; ApplyPhongShading = this && gradient texture
  bool m_show_frame_texture = true;        // crtgtrenderer.cpp:31 has false (a checkbox); see Redraw()
};

// cppvolrend/structured/rc1pvctsg/preprocessingstages.{h,cpp}: front end of the device pre-passes
class VCTPreProcessing {
 public:
  VCTPreProcessing() : maximum_standard_deviation(0.0) {}
  // PreProcessSuperVoxels + PreProcessPreIntegrationTable in one device call
  bool PreProcess(vis::StructuredGridVolume* vol, vis::TransferFunction* tf);
  double maximum_standard_deviation;
};

// cppvolrend/structured/rc1pvctsg/vctrenderer.{h,cpp}
class RC1PVoxelConeTracingSGPU : public BaseVolumeRenderer {
 public:
  RC1PVoxelConeTracingSGPU();
  ~RC1PVoxelConeTracingSGPU() override;
  const char* GetName() override { return "1-Pass - Voxel Cone Tracing - Single GPU"; }
  const char* GetAbbreviationName() override { return "s_1rc_vct"; }
  vis::GRID_VOLUME_DATA_TYPE GetDataTypeSupport() override { return vis::STRUCTURED; }
  void Clean() override;
  bool Init(int swidth, int sheight) override;
  bool Update(vis::Camera* camera) override;
  void Redraw() override;
  void FillParameterSpace(ParameterSpace& pspace) override;
  bool SetParameter(const std::string& name, double value) override;
 private:
  float m_u_step_size;
  bool apply_ambient_occlusion, apply_voxel_cone_tracing;
  float cone_step_size, cone_step_size_increase_rate, cone_initial_step, cone_apex_angle;
  bool apply_correction_factor; float opacity_correction_factor; int cone_number_of_samples;
  VCTPreProcessing pre_processing;
  PreIlluminationStructuredVolume m_pre_illum_str_vol;
  vrb_camera m_cam; vrb_lighting m_light; vrb_vct_params m_prm;
  bool m_apply_gradient_shading = false;   // "Apply Gradient Shading" checkbox; ApplyPhongShading = this && gradient texture
};

// cppvolrend/renderingmanager.{h,cpp}: headless re-host of the renderer-facing half.
class RenderingManager {
 public:
  static RenderingManager* Instance();
  static bool Exists();
  static void DestroyInstance();
  bool InitGL(int cuda_device = 0);                      // creates the device context instead of a GL context
  void AddVolumeRenderer(BaseVolumeRenderer* volrend);   // takes ownership
  bool InitData(std::string path_to_data);               // ReadData + camera/light lists + first renderer
  bool InitDataInMemory(vis::StructuredGridVolume* vol, vis::TransferFunction* tf);   // headless: no list files
  bool Display();                                        // PrepareRender + Redraw of the current renderer
  void Reshape(int w, int h);
  bool SetCurrentVolumeRenderer(int id);
  bool SetCurrentVolumeRendererByAbbreviation(const std::string& abbr);
  BaseVolumeRenderer* GetCurrentVolumeRenderer() { return curr_vol_renderer; }
  int GetNumberOfVolumeRenderers() { return (int)m_vtr_vr_methods.size(); }
  BaseVolumeRenderer* GetVolumeRenderer(int id) { return (id >= 0 && id < (int)m_vtr_vr_methods.size()) ? m_vtr_vr_methods[id] : nullptr; }
  bool SetCameraState(int id);
  void SetCamera(vis::CameraData* data);
  bool SetLightSourceList(int id);
  void UpdateLightSourceCameraVectors();
  vis::DataManager* GetDataManager() { return &m_data_mgr; }
  vis::RenderingParameters* GetRenderingParameters() { return &curr_rdr_parameters; }
  vis::CameraStateList* GetCameraStateList() { return &m_camera_state_list; }
  vis::LightSourceList* GetLightSourceList() { return &m_light_source_list; }
  bool UpdateDataAndResetCurrentVRMode();
  // ---- evaluation harness (renderingmanager.cpp:174-181,261-317,409-419,476-492,803-860; Evaluation.md) -----------
  // "Start Evaluation": FillParameterSpace of the current renderer, <base>/eval_DD-MM-YYYY_HH-MM-SS/{eval.csv,img/};
  // Display() then redraws every frame, and after every m_eval_numframes frames writes one CSV row + one screenshot
  // and moves to the next sample point.  RunEvaluation = StartEvaluation + Display() until the sweep is over.
  bool StartEvaluation(const std::string& base_directory, int frames_per_sample);
  bool RunEvaluation(const std::string& base_directory, int frames_per_sample);
  bool IsEvaluationRunning() const { return m_eval_running; }
  const std::string& GetEvaluationDirectory() const { return m_eval_basedirectory; }
  // GetFrontBufferPixelData(false) + GenerateImgFile(..., "PNG"): what glReadPixels returns after the frame was blended
  // (GL_SRC_ALPHA, GL_ONE_MINUS_SRC_ALPHA) over the white clear colour, 8 bits per channel, written top row first
  bool GetFrontBufferPixelData(std::vector<unsigned char>& rgb, int* w, int* h);
  bool SaveScreenshot(std::string filename);
  // "Set Reference" / "Generate Diff" buttons (:600-716): CIEDE2000 of the current frame against a stored one, mapped
  // through the white -> red transfer function of the reference and written as PNG
  bool StoreReferenceImage();
  bool GenerateDiffImage(const std::string& filename, double* max_delta_e = nullptr);
 private:
  RenderingManager();
  ~RenderingManager();
  static RenderingManager* crr_instance;
  vis::DataManager m_data_mgr;
  vis::RenderingParameters curr_rdr_parameters;
  vis::CameraStateList m_camera_state_list;
  vis::LightSourceList m_light_source_list;
  std::vector<BaseVolumeRenderer*> m_vtr_vr_methods;
  BaseVolumeRenderer* curr_vol_renderer;
  int m_current_vr_method_id, m_current_camera_state_id, m_current_lightsource_data_id;
  void EvaluationAfterFrame();
  ParameterSpace m_eval_paramspace;
  bool m_eval_running = false;
  int m_eval_numframes = 100, m_eval_currframe = 0, m_eval_currsample = 0;     // renderingmanager.cpp:1242-1245
  double m_eval_lasttime = 0.0;
  std::string m_eval_basedirectory, m_eval_imgdirectory;
  std::ofstream m_eval_csvfile;
  std::vector<float> s_ref_image; int s_ref_w = 0, s_ref_h = 0;
};
// libs/vis_utils/colorutils.cpp:221-311 (CIEDE2000 on 8-bit-range sRGB triplets) and :148-196 (ColorSpaces::RGBtoLAB)
double Cie2000Comparison(const double* rgb_a, const double* rgb_b);
bool WritePNG(const std::string& path, int w, int h, const unsigned char* rgb_top_row_first);
