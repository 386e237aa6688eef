// march_vct_common.cuh -- types shared by both filter modes of the rc1pvctsg marcher (march_vct.cu, hwf_vct.cu).
#ifndef VRB_MARCH_VCT_COMMON
#define VRB_MARCH_VCT_COMMON
struct v3f { float x, y, z; };
__device__ __forceinline__ v3f vm(float x, float y, float z) { v3f r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ v3f operator+(v3f a, v3f b) { return vm(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ v3f operator-(v3f a, v3f b) { return vm(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ v3f operator*(v3f a, float s) { return vm(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ v3f operator/(v3f a, v3f b) { return vm(a.x / b.x, a.y / b.y, a.z / b.z); }
__device__ __forceinline__ v3f vnrm(v3f a) { float d = a.x * a.x + a.y * a.y + a.z * a.z; float r = 1.0f / sqrtf(d); return a * r; }

struct SvLevel { const __half2* tex; int w, h, d; };
struct VctConst {
  SvLevel lev[VRB_MAX_LEVELS];
  int n_levels;
  const __half* lut; int lut_w, lut_h;
  v3f VSS, light_pos;
  vrb_vct_params P;
  float ka, kd, corr_fact;
  cudaTextureObject_t sv_tex, lut_tex;   // VCT_HW: RG16F mipmapped 3-D texture and R16F 2-D LUT texture
  v3f inv_VSS;
};

int vrb_vct_launch_hw(vrb_ctx* c, const vrb_camera* cam, const VctConst& C, int count_samples);   // hwf_vct.cu
int vrb_vct_light_cache_launch_hw(vrb_ctx* c, const VctConst& C, int rw, int rh, int rd);          // hwf_vct.cu
int vrb_sv_tex_prepare(vrb_ctx* c);                                                                // vct_prepass.cu
#endif
