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

// one pyramid level: padded array of w x h x d texels.  Brick contexts hold a WINDOW of the whole volume's level: gw/gh/gd
// are the level's dims in the whole volume and (ox,oy,oz) the window's first texel (whole-volume builds: g == local, o == 0).
struct SvLevel { const __half2* tex; int w, h, d; int gw, gh, gd; int ox, oy, oz; };
// brick of a sort-last partition as the VCT marcher sees it (same meaning as BrickView in sort_last.cu)
struct VctBrick {
  float offx, offy, offz;    // origin - ghost_lo: global voxel index of local texel 0
  int lo[3], hi[3];          // owned cells [lo, hi)
  int nx, ny, nz;            // global resolution
  float kx, ky, kz;          // global N / G
};
#define VRB_VCT_MAX_FRONT 64
struct VctFront { const float* p[VRB_VCT_MAX_FRONT]; int n; };
struct VctConst {
  SvLevel lev[VRB_MAX_LEVELS];
  int n_levels;
  const __half* lut; int lut_w, lut_h;
  v3f VSS, light_pos;
  vrb_vct_params P;
  float ka, kd, corr_fact;
  cudaTextureObject_t sv_tex, lut_tex;   // VCT_HW: RG16F mipmapped 3-D texture and R16F 2-D LUT texture
  v3f inv_VSS;
  PhongView ph;                          // ApplyPhongShading (vct_ray_bbox_marching.comp:163-182)
  v3f sv_scale, sv_bias;                 // VCT_HW: normalised pyramid coordinate = wpos * sv_scale + sv_bias (window of a brick)
};

int vrb_vct_launch_hw(vrb_ctx* c, const vrb_camera* cam, const VctConst& C, int count_samples);   // hwf_vct.cu
int vrb_vct_brick_launch_hw(vrb_ctx* c, const vrb_camera* cam, const VctConst& C, const VctBrick& B, const VctFront& front, int mode,
                             int count_samples);                                                   // hwf_vct.cu
int vrb_vct_light_cache_launch_hw(vrb_ctx* c, const VctConst& C, int rw, int rh, int rd);          // hwf_vct.cu
int vrb_sv_tex_prepare(vrb_ctx* c);                                                                // vct_prepass.cu
#endif
