// march_dos_common.cuh -- types shared by both filter modes of the rc1pdosct marcher (march_dos.cu, hwf_dos.cu).
#ifndef VRB_MARCH_DOS_COMMON
#define VRB_MARCH_DOS_COMMON
struct d3 { float x, y, z; };
__device__ __forceinline__ d3 m3(float x, float y, float z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ d3 operator+(d3 a, d3 b) { return m3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ d3 operator-(d3 a, d3 b) { return m3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ d3 operator-(d3 a) { return m3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ d3 operator*(d3 a, float s) { return m3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ d3 operator/(d3 a, d3 b) { return m3(a.x / b.x, a.y / b.y, a.z / b.z); }
__device__ __forceinline__ float dot3(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ d3 cross3(d3 a, d3 b) { return m3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
__device__ __forceinline__ d3 nrm3(d3 a) { float r = 1.0f / sqrtf(dot3(a, a)); return a * r; }

struct DosConst {
  LevelView lev[VRB_MAX_LEVELS];
  int n_levels;
  d3 VSS;
  ConeView occ, sdw;
  vrb_dos_params P;
  float ka, kd;
  d3 eye, light_pos, light_fwd, light_up, light_right;
  cudaTextureObject_t pyr_tex;   // DOS_HW: mipmapped 3-D texture of the same levels (trilinear + linear between levels)
  d3 inv_VSS;
  PhongView ph;                  // ApplyPhongShading (rc1pdosct/ray_bbox_marching.comp:629-648)
};

int vrb_dos_launch_hw(vrb_ctx* c, const vrb_camera* cam, const DosConst& C, int count_samples);   // hwf_dos.cu
int vrb_dos_light_cache_launch_hw(vrb_ctx* c, const DosConst& C, const float eye_up[3], int rw, int rh, int rd);   // hwf_dos.cu
int vrb_pyr_tex_prepare(vrb_ctx* c);                                                                // extcoef_pyramid.cu
#endif  // VRB_MARCH_DOS_COMMON
