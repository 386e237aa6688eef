// march_gt_common.cuh -- types shared by both filter modes of the rc1pcrtgt marcher (march_gt.cu, hwf_gt.cu).
#ifndef VRB_MARCH_GT_COMMON
#define VRB_MARCH_GT_COMMON
struct g3 { float x, y, z; };
__device__ __forceinline__ g3 gm(float x, float y, float z) { g3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ g3 operator+(g3 a, g3 b) { return gm(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ g3 operator-(g3 a, g3 b) { return gm(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ g3 operator-(g3 a) { return gm(-a.x, -a.y, -a.z); }
__device__ __forceinline__ g3 operator*(g3 a, float s) { return gm(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float gdot(g3 a, g3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ g3 gcross(g3 a, g3 b) { return gm(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
__device__ __forceinline__ g3 gnrm(g3 a) { float r = 1.0f / sqrtf(gdot(a, a)); return a * r; }

struct GtConst {
  g3 G;
  vrb_gt_params P;
  float ka, kd;
  g3 light_pos, light_fwd, light_up, light_right;
  const float* occ_rays; const float* sdw_rays;    // n x 3, fp16-rounded
  PhongView ph;                                    // ApplyGradientPhongShading (gt_ray_marching.comp:277-295)
};

int vrb_gt_launch_hw(vrb_ctx* c, const vrb_camera* cam, const GtConst& C, int count_samples);   // hwf_gt.cu
#endif
