// empty_space.cu -- occupancy cells for result-preserving empty-space skipping (the reference has no
// skipping: ray_marching_1p.comp:118-171 samples every step).
//
// A primary sample with padded floor index (ix,iy,iz) (see vrb_sample_volume) reads the padded texels ix..ix+1 per
// axis.  Cell (cx,cy,cz) collects the floor indices [8c, 8c+7] per axis, i.e. the padded texels [8c, 8c+8].  If no TF
// texel that a density in [min, max] of those texels (widened by one TF texel on both sides against blend rounding) can
// select has alpha != 0, every sample in the cell has src.a == 0 exactly and contributes nothing: the marcher may skip
// the fetches.  The ray parameter is still advanced by the same fp32 additions, so pixels and loop counts do not change.
#include "vrb_internal.cuh"

#define CELL_SHIFT 3
#define CELL 8

// one warp per cell
__global__ void __launch_bounds__(256)
k_cell_minmax(VolView v, int cw, int ch, int cd, __half2* __restrict__ mm) {
  const long long cell = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (cell >= (long long)cw * ch * cd) return;
  const int cx = (int)(cell % cw), cy = (int)((cell / cw) % ch), cz = (int)(cell / ((long long)cw * ch));
  const int x0 = cx * CELL, y0 = cy * CELL, z0 = cz * CELL;
  const int nx = min(CELL + 1, v.pw - x0), ny = min(CELL + 1, v.ph - y0), nz = min(CELL + 1, v.pd - z0);
  float lo = 3.0e38f, hi = -3.0e38f;
  const int n = nx * ny * nz;
  for (int i = lane; i < n; i += 32) {
    int x = i % nx, y = (i / nx) % ny, z = i / (nx * ny);
    float t = __half2float(__ldg(v.tex + (long long)(z0 + z) * v.slice + (long long)(y0 + y) * v.pw + (x0 + x)));
    lo = fminf(lo, t); hi = fmaxf(hi, t);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) mm[cell] = __floats2half2_rn(lo, hi);     // texels are fp16: exact
}

// nz[i] = number of padded TF texels j < i whose alpha (extinction) is not exactly 0; nz has n + 3 entries
__global__ void k_tf_nonzero_prefix(const float4* __restrict__ tf, int n, int* __restrict__ nz) {
  __shared__ int carry;
  __shared__ int warp_tot[32];
  if (threadIdx.x == 0) { carry = 0; nz[0] = 0; }
  __syncthreads();
  const int total = n + 2;
  for (int base = 0; base < total; base += 1024) {
    int i = base + threadIdx.x;
    int f = (i < total && tf[i].w != 0.0f) ? 1 : 0;
    int x = f;
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = warp_tot[threadIdx.x];
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += y; }
      warp_tot[threadIdx.x] = w;
    }
    __syncthreads();
    int incl = x + ((threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0) + carry;
    if (i < total) nz[i + 1] = incl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
}

__global__ void k_cell_flags(const __half2* __restrict__ mm, long long ncells, const int* __restrict__ nz, int n,
                             unsigned char* __restrict__ flags, unsigned* __restrict__ n_empty) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = i < ncells;
  bool empty = false;
  if (in) {
    float2 r = __half22float2(mm[i]);
    // vrb_sample_tf: up = s*n + 0.5 clamped to [0, n+0.5], texels floor(up), floor(up)+1 of the padded table
    float ulo = fminf(fmaxf(r.x * (float)n + 0.5f, 0.0f), (float)n + 0.5f);
    float uhi = fminf(fmaxf(r.y * (float)n + 0.5f, 0.0f), (float)n + 0.5f);
    int lo = max((int)floorf(ulo) - 1, 0), hi = min((int)floorf(uhi) + 2, n + 1);
    empty = (nz[hi + 1] - nz[lo]) == 0;
    flags[i] = empty ? 0 : 1;
  }
  const unsigned m = __ballot_sync(0xffffffffu, empty);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_empty, (unsigned)__popc(m));
}

void vrb_free_cells(vrb_ctx* c) {
  if (c->d_cell_mm) cudaFree(c->d_cell_mm);
  if (c->d_cell_flags) cudaFree(c->d_cell_flags);
  if (c->d_tf_nz) cudaFree(c->d_tf_nz);
  if (c->d_cell_count) cudaFree(c->d_cell_count);
  c->d_cell_count = nullptr;
  c->d_cell_mm = nullptr; c->d_cell_flags = nullptr; c->d_tf_nz = nullptr;
  c->cell_mm_valid = c->cell_flags_valid = false;
  c->cell_dims[0] = c->cell_dims[1] = c->cell_dims[2] = 0;
}

int vrb_cells_prepare(vrb_ctx* c) {
  if (c->cell_mm_valid && c->cell_flags_valid) return VRB_OK;
  VolView v = c->vol_view();
  // floor indices run over [0, w] per axis
  const int cw = (c->vw >> CELL_SHIFT) + 1, ch = (c->vh >> CELL_SHIFT) + 1, cd = (c->vd >> CELL_SHIFT) + 1;
  const long long ncells = (long long)cw * ch * cd;
  if (!c->cell_mm_valid) {
    if (c->d_cell_mm) { VRB_CUDA(cudaFree(c->d_cell_mm)); c->d_cell_mm = nullptr; }
    if (c->d_cell_flags) { VRB_CUDA(cudaFree(c->d_cell_flags)); c->d_cell_flags = nullptr; }
    VRB_CUDA(cudaMalloc(&c->d_cell_mm, (size_t)ncells * sizeof(__half2)));
    VRB_CUDA(cudaMalloc(&c->d_cell_flags, (size_t)ncells));
    const long long threads = ncells * 32;
    k_cell_minmax<<<(unsigned)((threads + 255) / 256), 256, 0, c->stream>>>(v, cw, ch, cd, c->d_cell_mm);
    VRB_CUDA(cudaGetLastError());
    c->launches++;
    c->cell_dims[0] = cw; c->cell_dims[1] = ch; c->cell_dims[2] = cd;
    c->cell_mm_valid = true;
    c->cell_flags_valid = false;
  }
  if (!c->cell_flags_valid) {
    if (c->d_tf_nz) { VRB_CUDA(cudaFree(c->d_tf_nz)); c->d_tf_nz = nullptr; }
    VRB_CUDA(cudaMalloc(&c->d_tf_nz, (size_t)(c->tf_n + 3) * sizeof(int)));
    k_tf_nonzero_prefix<<<1, 1024, 0, c->stream>>>(c->d_tf_rgbt, c->tf_n, c->d_tf_nz);
    VRB_CUDA(cudaGetLastError());
    if (!c->d_cell_count) VRB_CUDA(cudaMalloc(&c->d_cell_count, sizeof(unsigned)));
    VRB_CUDA(cudaMemsetAsync(c->d_cell_count, 0, sizeof(unsigned), c->stream));
    k_cell_flags<<<(unsigned)((ncells + 255) / 256), 256, 0, c->stream>>>(c->d_cell_mm, ncells, c->d_tf_nz, c->tf_n, c->d_cell_flags, c->d_cell_count);
    VRB_CUDA(cudaGetLastError());
    unsigned n_empty = 0;
    VRB_CUDA(cudaMemcpyAsync(&n_empty, c->d_cell_count, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    VRB_CUDA(cudaStreamSynchronize(c->stream));
    c->cell_empty_fraction = (float)((double)n_empty / (double)ncells);
    c->launches += 2;
    c->cell_flags_valid = true;
  }
  return VRB_OK;
}
