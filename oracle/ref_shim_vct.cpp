// oracle/ref_shim_vct.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// extern "C" wrapper around the reference's VCT CPU pre-passes (cppvolrend/structured/rc1pvctsg/preprocessingstages.cpp,
// compiled in place from /root/reference; see oracle/Makefile target `ref`).  Its own translation unit because the
// class header pulls in three headers g++ rejects, shadowed by the stand-ins under ref_stubs/vct (first on the include
// path for this file and for preprocessingstages.cpp only).  GL calls are recorded, not executed.
#include <structured/rc1pvctsg/preprocessingstages.h>
#include <volvis_utils/transferfunction1d.h>
#include <cstring>
#include <vector>

const std::vector<float>& ref_last_tex2d(int* w, int* h);     // ref_shim.cpp: the last Texture2D::SetData client array

// glTexImage3D as VCTPreProcessing::PreProcessSuperVoxels calls it (one GL_RG / GL_FLOAT upload per mip level,
// preprocessingstages.cpp:131-146): the GLEW entry point is a function pointer, pointed at a recorder here.
struct RecordedLevel { int level, w, h, d; std::vector<float> rg; };
static std::vector<RecordedLevel> g_tex_image3d;
static void GLAPIENTRY record_tex_image3d(GLenum, GLint level, GLint, GLsizei w, GLsizei h, GLsizei d, GLint, GLenum format, GLenum type, const void* pixels) {
  RecordedLevel r; r.level = level; r.w = w; r.h = h; r.d = d;
  if (format == GL_RG && type == GL_FLOAT && pixels) r.rg.assign((const float*)pixels, (const float*)pixels + (size_t)w * h * d * 2);
  g_tex_image3d.push_back(r);
}
PFNGLTEXIMAGE3DPROC __glewTexImage3D = record_tex_image3d;

extern "C" {

// ---- VCTPreProcessing::PreProcessSuperVoxels + PreProcessPreIntegrationTable (preprocessingstages.cpp:35-202) run as they are.
// levels_rg: the GL_RG / GL_FLOAT arrays handed to glTexImage3D, concatenated level by level (before the RG16F rounding);
// dims: 3 ints per level; lut: the GL_FLOAT array handed to Texture2D::SetData (before the R16F rounding), lut_wh its size.
// tf = handle of ref_tf_create.  Returns the number of levels, or -1 when a capacity is too small.
int ref_vct_preprocess(const void* vox, int w, int h, int d, int bpv, void* tf, float* levels_rg, unsigned long long cap_floats, int* dims, int cap_levels,
                       double* max_stddev, float* lut, unsigned long long lut_cap, int* lut_wh) {
  vis::StructuredGridVolume vol("v", w, h, d);
  vol.SetArrayData(const_cast<void*>(vox), bpv == 1 ? vis::DataStorageSize::_8_BITS : vis::DataStorageSize::_16_BITS);
  g_tex_image3d.clear();
  VCTPreProcessing pre;
  pre.PreProcessSuperVoxels(&vol);
  *max_stddev = pre.maximum_standard_deviation;
  int n = (int)g_tex_image3d.size();
  unsigned long long off = 0;
  bool ok = n <= cap_levels;
  for (int i = 0; ok && i < n; ++i) {
    const RecordedLevel& r = g_tex_image3d[i];
    dims[3 * i] = r.w; dims[3 * i + 1] = r.h; dims[3 * i + 2] = r.d;
    if (r.level != i || off + r.rg.size() > cap_floats) { ok = false; break; }
    std::memcpy(levels_rg + off, r.rg.data(), r.rg.size() * sizeof(float));
    off += r.rg.size();
  }
  if (ok && lut) {
    pre.PreProcessPreIntegrationTable(&vol, (vis::TransferFunction*)tf);
    const std::vector<float>& t = ref_last_tex2d(&lut_wh[0], &lut_wh[1]);
    if (t.size() > lut_cap) ok = false;
    else std::memcpy(lut, t.data(), t.size() * sizeof(float));
  }
  pre.Destroy();
  vol.SetArrayData(nullptr, vis::DataStorageSize::UNKNOWN);
  return ok ? n : -1;
}

}  // extern "C"
