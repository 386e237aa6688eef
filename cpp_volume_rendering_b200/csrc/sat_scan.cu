// sat_scan.cu -- 3-D summed-area table as three sm_100a prefix-scan passes.
// Replaces RC1PExtinctionBasedShading::GenerateExtinctionSAT3DTex (rc1pextbsd/ebsrenderer.cpp:624-723) and
// vis::SummedAreaTable3D<T>::BuildSAT (libs/vis_utils/summedareatable.h:218-278), which the reference runs
// single-threaded on the CPU.
//
//   pass X  fill + scan along x : raw voxel -> LUT (extinction per voxel value) -> inclusive scan of each row,
//           one warp per row, lanes stride the row so loads and stores are coalesced.        reads b_v, writes 8 B
//   pass Y  running sum along y : one thread per (x,z) column, 8 independent loads in flight. reads 8 B, writes 8 B
//   pass Z  running sum along z : one thread per (x,y) column, writes the fp32 texel the marcher samples.
//                                                                                              reads 8 B, writes 4 B
// => b_v + 36 algorithmic bytes per cell of the bordered (W+2)(H+2)(D+2) grid.
// fp64 accumulation like the reference; the association differs from its 7-term recurrence, so the float result is
// compared with a relative tolerance (tests/test_sat_gpu.py); the integer instantiation is bit-exact.
#include "vrb_internal.cuh"
#include <vector>
#include <cstring>
#include <cstdlib>
#include <algorithm>

template <typename T> __device__ __forceinline__ T shfl_up_t(T v, int o);
template <> __device__ __forceinline__ double shfl_up_t<double>(double v, int o) { return __shfl_up_sync(0xffffffffu, v, o); }
template <> __device__ __forceinline__ unsigned long long shfl_up_t<unsigned long long>(unsigned long long v, int o) { return __shfl_up_sync(0xffffffffu, v, o); }
template <typename T> __device__ __forceinline__ T shfl_idx_t(T v, int l) { return __shfl_sync(0xffffffffu, v, l); }

// B = border width (1 for the extinction SAT, 0 for the integer mode).
template <typename T, typename VoxT, typename LutT, int B>
__global__ void __launch_bounds__(256)
k_sat_fill_scan_x(const VoxT* __restrict__ raw, const LutT* __restrict__ lut, T* __restrict__ S, int vw, int vh, int vd, int z0, int nz) {
  // z0, nz: the slices [z0, z0 + nz) of the bordered grid this launch fills; S points at slice z0 (whole grid: 0, d)
  const int w = vw + 2 * B, h = vh + 2 * B, d = nz;
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)h * d;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = warp0; row < rows; row += nwarps) {
    const int y = (int)(row % h), z = z0 + (int)(row / h);
    T* out = S + (size_t)w * (size_t)row;
    const bool interior = (B == 0) || (y >= 1 && y <= vh && z >= 1 && z <= vd);
    if (!interior) {
      for (int x = lane; x < w; x += 32) out[x] = T(0);
      continue;
    }
    const VoxT* in = raw + (size_t)vw * ((size_t)(y - B) + (size_t)vh * (size_t)(z - B));
    T carry = T(0);
    for (int x0 = 0; x0 < w; x0 += 32) {
      const int x = x0 + lane;
      T v = T(0);
      if (x >= B && x < vw + B) v = (T)__ldg(lut + in[x - B]);
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        T t = shfl_up_t<T>(v, o);
        if (lane >= o) v += t;
      }
      v += carry;
      if (x < w) out[x] = v;
      carry = shfl_idx_t<T>(v, 31);
    }
  }
}

// running sum along y, in place; thread per (x,z)
template <typename T>
__global__ void __launch_bounds__(128) k_sat_scan_y(T* __restrict__ S, int w, int h, int d) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)w * d) return;
  int x = (int)(i % w), z = (int)(i / w);
  T* p = S + (size_t)x + (size_t)w * h * (size_t)z;
  T acc = T(0);
  int y = 0;
  for (; y + 8 <= h; y += 8) {
    T v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = p[(size_t)(y + k) * w];
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc += v[k]; p[(size_t)(y + k) * w] = acc; }
  }
  for (; y < h; ++y) { acc += p[(size_t)y * w]; p[(size_t)y * w] = acc; }
}

// running sum along z; thread per (x,y); writes OutT (fp32 texel or u64)
template <typename T, typename OutT>
__global__ void __launch_bounds__(128) k_sat_scan_z(const T* __restrict__ S, OutT* __restrict__ out, int w, int h, int d) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t plane = (size_t)w * h;
  if (i >= (long long)plane) return;
  const T* p = S + i;
  OutT* q = out + i;
  T acc = T(0);
  int z = 0;
  for (; z + 8 <= d; z += 8) {
    T v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = p[(size_t)(z + k) * plane];
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc += v[k]; q[(size_t)(z + k) * plane] = (OutT)acc; }
  }
  for (; z < d; ++z) { acc += p[(size_t)z * plane]; q[(size_t)z * plane] = (OutT)acc; }
}

// Packed copies of the float SAT for the marcher: every texel carries the neighbours its GL_LINEAR footprint needs,
// already clamped to the edge, so one 64-bit (pairs) or 128-bit (quads) load replaces 2 or 4 scalar loads.
__global__ void __launch_bounds__(256) k_sat_pack2(const float* __restrict__ S, float2* __restrict__ P, int w, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int x = (int)(i % w);
    P[i] = make_float2(S[i], S[x + 1 < w ? i + 1 : i]);
  }
}
__global__ void __launch_bounds__(256) k_sat_pack4(const float* __restrict__ S, float4* __restrict__ P, int w, int h, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int x = (int)(i % w), y = (int)((i / w) % h);
    long long ix = x + 1 < w ? i + 1 : i;
    long long dy = y + 1 < h ? w : 0;
    P[i] = make_float4(S[i], S[ix], S[i + dy], S[ix + dy]);
  }
}

// Gather atlas: slice z of the SAT becomes tile (z % T, z / T) of a 2-D image, each tile (w+2)x(h+2) with a replicated
// one-texel gutter, so that tex2Dgather (which only exists for 2-D arrays) sees clamp-to-edge inside every tile.
__global__ void __launch_bounds__(256)
k_sat_atlas(const float* __restrict__ S, float* __restrict__ A, int w, int h, int d, int T, long long pitch_f) {
  const int tw = w + 2, th = h + 2;
  const long long n = (long long)tw * th * d;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int px = (int)(i % tw), py = (int)((i / tw) % th), z = (int)(i / ((long long)tw * th));
    int sx = min(max(px - 1, 0), w - 1), sy = min(max(py - 1, 0), h - 1);
    long long ax = (long long)(z % T) * tw + px, ay = (long long)(z / T) * th + py;
    A[ay * pitch_f + ax] = S[(long long)sx + (long long)w * (sy + (long long)h * z)];
  }
}

void vrb_free_sat_atlas(vrb_ctx* c) {
  if (c->sat_tex) cudaDestroyTextureObject(c->sat_tex);
  if (c->sat_array) cudaFreeArray(c->sat_array);
  c->sat_tex = 0; c->sat_array = nullptr; c->atlas_tiles_x = 0;
}

static int build_sat_atlas(vrb_ctx* c, int w, int h, int d) {
  const int tw = w + 2, th = h + 2;
  int T = 1;
  while ((long long)T * T < d) ++T;                       // ~square atlas
  T = std::min(T, 131072 / tw);
  const int rows = (d + T - 1) / T;
  if (T < 1 || (long long)rows * th > 65536) return VRB_ERR_UNSUPPORTED;     // does not fit a 2-D texture: caller falls back
  const size_t aw = (size_t)T * tw, ah = (size_t)rows * th;
  cudaChannelFormatDesc fd = cudaCreateChannelDesc<float>();
  VRB_CUDA(cudaMallocArray(&c->sat_array, &fd, aw, ah, cudaArrayTextureGather));
  float* tmp = nullptr;
  VRB_CUDA(cudaMalloc(&tmp, aw * ah * sizeof(float)));
  VRB_CUDA(cudaMemsetAsync(tmp, 0, aw * ah * sizeof(float), c->stream));
  const long long n = (long long)tw * th * d;
  k_sat_atlas<<<(int)std::min<long long>((n + 255) / 256, 148 * 32), 256, 0, c->stream>>>(c->d_sat, tmp, w, h, d, T, (long long)aw);
  c->launches++;
  cudaError_t e = cudaMemcpy2DToArrayAsync(c->sat_array, 0, 0, tmp, aw * sizeof(float), aw * sizeof(float), ah, cudaMemcpyDeviceToDevice, c->stream);
  cudaError_t e2 = cudaStreamSynchronize(c->stream);
  cudaFree(tmp);
  VRB_REQUIRE(e == cudaSuccess && e2 == cudaSuccess, VRB_ERR_CUDA, "vrb_sat_build: atlas copy failed: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
  cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
  rd.resType = cudaResourceTypeArray; rd.res.array.array = c->sat_array;
  cudaTextureDesc td; memset(&td, 0, sizeof(td));
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
  VRB_CUDA(cudaCreateTextureObject(&c->sat_tex, &rd, &td, nullptr));
  c->atlas_tiles_x = T;
  return VRB_OK;
}

static int pack_sat(vrb_ctx* c, size_t n, int w, int h) {
  if (c->d_sat_packed) { VRB_CUDA(cudaFree(c->d_sat_packed)); c->d_sat_packed = nullptr; }
  vrb_free_sat_atlas(c);
  const char* env = getenv("VRB_SAT_PACK");
  int pack = env ? atoi(env) : 8;
  if (pack != 1 && pack != 2 && pack != 4 && pack != 8) pack = 8;
  if (pack == 8) {
    int rc = build_sat_atlas(c, w, h, (int)(n / ((size_t)w * h)));
    if (rc == VRB_OK) { c->sat_pack = 8; return VRB_OK; }
    if (rc != VRB_ERR_UNSUPPORTED) return rc;
    vrb_free_sat_atlas(c);
    pack = 1;                                             // too large for a 2-D texture (e.g. 2048^3): linear loads
  }
  c->sat_pack = pack;
  if (pack == 1) return VRB_OK;
  VRB_CUDA(cudaMalloc(&c->d_sat_packed, n * sizeof(float) * pack));
  int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 32);
  if (pack == 2) k_sat_pack2<<<blocks, 256, 0, c->stream>>>(c->d_sat, (float2*)c->d_sat_packed, w, (long long)n);
  else           k_sat_pack4<<<blocks, 256, 0, c->stream>>>(c->d_sat, (float4*)c->d_sat_packed, w, h, (long long)n);
  VRB_CUDA(cudaGetLastError());
  c->launches++;
  return VRB_OK;
}

template <typename T, typename LutT, typename OutT, int B>
static int run_sat(vrb_ctx* c, const LutT* d_lut, T* d_tmp, OutT* d_out) {
  const int w = c->vw + 2 * B, h = c->vh + 2 * B, d = c->vd + 2 * B;
  const long long rows = (long long)h * d;
  int blocks_x = (int)std::min<long long>((rows + 7) / 8, 148LL * 64);
  // device time of the three scan passes alone (allocation of the fp64 scratch excluded): vrb_last_prepass_ms()
  cudaEvent_t e0, e1;
  VRB_CUDA(cudaEventCreate(&e0)); VRB_CUDA(cudaEventCreate(&e1));
  VRB_CUDA(cudaEventRecord(e0, c->stream));
  if (c->bpv == 1)
    k_sat_fill_scan_x<T, uint8_t, LutT, B><<<blocks_x, 256, 0, c->stream>>>((const uint8_t*)c->d_raw, d_lut, d_tmp, c->vw, c->vh, c->vd, 0, d);
  else
    k_sat_fill_scan_x<T, uint16_t, LutT, B><<<blocks_x, 256, 0, c->stream>>>((const uint16_t*)c->d_raw, d_lut, d_tmp, c->vw, c->vh, c->vd, 0, d);
  VRB_CUDA(cudaGetLastError());
  long long ny = (long long)w * d, nz = (long long)w * h;
  k_sat_scan_y<T><<<(unsigned)((ny + 127) / 128), 128, 0, c->stream>>>(d_tmp, w, h, d);
  VRB_CUDA(cudaGetLastError());
  k_sat_scan_z<T, OutT><<<(unsigned)((nz + 127) / 128), 128, 0, c->stream>>>(d_tmp, d_out, w, h, d);
  VRB_CUDA(cudaGetLastError());
  VRB_CUDA(cudaEventRecord(e1, c->stream));
  VRB_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  VRB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  c->last_prepass_ms = ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  c->launches += 3;
  return VRB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Reference-order build.  SummedAreaTable3D<double>::BuildSAT (summedareatable.h:218-278) evaluates
//   S(x,y,z) = v + S(x-1,y-1,z-1) + S(x,y,z-1) + S(x,y-1,z) + S(x-1,y,z) - S(x-1,y-1,z) - S(x,y-1,z-1) - S(x-1,y,z-1)
// left to right in fp64.  The sums are not exactly representable, so only this evaluation order reproduces the
// reference's doubles, and hence its floats, bit for bit (three scans land on the other side of a float rounding
// boundary for ~1e-6 of the texels, by one ulp, which the marcher's cancelling box queries amplify).  Every cell
// depends on the three previous anti-diagonal planes x+y+z = k-1..k-3 only: one launch per plane, the fp64 state is a
// rolling window of four (y,z) planes (8 MB at 512^3, L2-resident), all plane reads coalesced.
// ---------------------------------------------------------------------------------------------------------
template <typename VoxT>
__global__ void __launch_bounds__(256)
k_sat_wavefront(const VoxT* __restrict__ raw, const float* __restrict__ lut, int W, int H, int D, int k,
                const double* __restrict__ p1, const double* __restrict__ p2, const double* __restrict__ p3,
                double* __restrict__ p0, float* __restrict__ sat) {
  const int sw = W + 2, sh = H + 2, sd = D + 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sh * sd) return;
  const int y = i % sh, z = i / sh;
  const int x = k - y - z;
  if (x < 0 || x >= sw) return;
  double val = 0.0;
  if (x >= 1 && y >= 1 && z >= 1) {
    float e = 0.0f;
    if (x <= W && y <= H && z <= D) e = __ldg(lut + raw[(size_t)(x - 1) + (size_t)W * ((size_t)(y - 1) + (size_t)H * (size_t)(z - 1))]);
    const int c = z * sh + y, cz = (z - 1) * sh + y;
    val = (double)e;
    val = __dadd_rn(val, p3[cz - 1]);     // S(x-1, y-1, z-1)
    val = __dadd_rn(val, p1[cz]);         // S(x,   y,   z-1)
    val = __dadd_rn(val, p1[c - 1]);      // S(x,   y-1, z)
    val = __dadd_rn(val, p1[c]);          // S(x-1, y,   z)
    val = __dadd_rn(val, -p2[c - 1]);     // S(x-1, y-1, z)
    val = __dadd_rn(val, -p2[cz - 1]);    // S(x,   y-1, z-1)
    val = __dadd_rn(val, -p2[cz]);        // S(x-1, y,   z-1)
  }
  p0[z * sh + y] = val;
  sat[(size_t)x + (size_t)sw * ((size_t)y + (size_t)sh * (size_t)z)] = (float)val;
}

// The same recurrence in ONE launch: the bordered grid is cut into 8x8x8 tiles; a tile needs the three tiles before it
// (x-1, y-1, z-1) and is itself a small wavefront of 22 anti-diagonals kept in shared memory (__syncthreads between them).
// CTAs of 64 threads take tiles from a ticket counter in order of the tile diagonal X+Y+Z, wait for the three
// predecessors' done-flags, load the tile's one-cell halo of fp64 values from L2, run the 22 steps, store the fp64 values
// (for the successors' halos) and the float texels, and raise their flag.  A waiting CTA only ever waits for tiles with
// smaller tickets, i.e. tiles held by CTAs that are already running: no deadlock whatever the grid size.  The critical
// path drops from 1540 kernel launches to 193 tile steps; every cell is still evaluated by the reference's 7-term
// expression, left to right, in fp64 (bit-identical floats).
#define SAT_TX 8                  // tile extent along x (cells a thread walks); y and z extents are 8 (64 threads = the (y, z) columns).  32 was measured too: 6.36 ms against 6.08 ms
#define SAT_T 8
#define SAT_PX (SAT_TX + 1)       // padded extents of the shared-memory block (one halo cell below on every axis)
#define SAT_PY (SAT_T + 1)
template <typename VoxT>
__global__ void __launch_bounds__(64)
k_sat_tiles(const VoxT* __restrict__ raw, const float* __restrict__ lut, int W, int H, int D, const unsigned* __restrict__ order, unsigned n_tiles,
            int tx_n, int ty_n, unsigned* __restrict__ ticket, int* __restrict__ done, double* __restrict__ S64, float* __restrict__ sat) {
  __shared__ double sm[SAT_PX * SAT_PY * SAT_PY];
  __shared__ unsigned s_tile;
  constexpr int SLICE = SAT_PX * SAT_PY;
  const int sw = W + 2, sh = H + 2, sd = D + 2;
  const int tid = threadIdx.x;
  const int ly = tid & 7, lz = tid >> 3;                 // this thread's (y, z) column of the tile
  for (;;) {
    __syncthreads();                                     // the previous tile's shared memory is no longer read
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned tk = s_tile;
    if (tk >= n_tiles) break;
    const unsigned packed = __ldg(order + tk);
    const int TX = (int)(packed & 1023u), TY = (int)((packed >> 10) & 1023u), TZ = (int)(packed >> 20);
    const int x0 = TX * SAT_TX, y0 = TY * SAT_T, z0 = TZ * SAT_T;
    // this thread's extinction values along x, fetched while the predecessors finish
    const int gy = y0 + ly, gz = z0 + lz;
    float e[SAT_TX];
#pragma unroll
    for (int k = 0; k < SAT_TX; ++k) {
      const int gx = x0 + k;
      e[k] = 0.0f;
      if (gx >= 1 && gy >= 1 && gz >= 1 && gx <= W && gy <= H && gz <= D)
        e[k] = __ldg(lut + raw[(size_t)(gx - 1) + (size_t)W * ((size_t)(gy - 1) + (size_t)H * (size_t)(gz - 1))]);
    }
    // wait for the predecessors (threads 0..2, one each)
    if (tid < 3) {
      const int px = TX - (tid == 0), py = TY - (tid == 1), pz = TZ - (tid == 2);
      if (px >= 0 && py >= 0 && pz >= 0) {
        const volatile int* f = done + ((size_t)pz * ty_n + py) * tx_n + px;
        // (bounded: a logic error must show up as a failed test, not as a hung GPU)
        for (long long spins = 0; *f == 0 && spins < (1ll << 27); ++spins) { }
      }
      __threadfence();
    }
    __syncthreads();
    // halo (the x0-1 / y0-1 / z0-1 planes, edges and corner) from the fp64 array: L2 loads, the values come from other SMs
    for (int i = tid; i < SAT_PX * SAT_PY * SAT_PY; i += 64) {
      const int hx = i % SAT_PX, hy = (i / SAT_PX) % SAT_PY, hz = i / SLICE;
      if (hx && hy && hz) continue;                      // interior: computed below
      const int gx = x0 + hx - 1, hgy = y0 + hy - 1, hgz = z0 + hz - 1;
      double v = 0.0;
      if (gx >= 0 && hgy >= 0 && hgz >= 0 && gx < sw && hgy < sh && hgz < sd) v = __ldcg(S64 + (size_t)gx + (size_t)sw * ((size_t)hgy + (size_t)sh * (size_t)hgz));
      sm[i] = v;
    }
    __syncthreads();
    double* c = sm + (lz + 1) * SLICE + (ly + 1) * SAT_PX + 1;   // cell (0, ly, lz) of the tile
#pragma unroll
    for (int d = 0; d < SAT_TX + 2 * SAT_T - 2; ++d) {
      const int lx = d - ly - lz;
      if (lx >= 0 && lx < SAT_TX) {
        const int gx = x0 + lx;
        double val = 0.0;
        if (gx >= 1 && gy >= 1 && gz >= 1) {
          const double* q = c + lx;
          float ev = 0.0f;
#pragma unroll
          for (int k = 0; k < SAT_TX; ++k) if (k == lx) ev = e[k];
          val = (double)ev;
          val = __dadd_rn(val, q[-SLICE - SAT_PX - 1]);   // S(x-1, y-1, z-1)
          val = __dadd_rn(val, q[-SLICE]);                // S(x,   y,   z-1)
          val = __dadd_rn(val, q[-SAT_PX]);               // S(x,   y-1, z)
          val = __dadd_rn(val, q[-1]);                    // S(x-1, y,   z)
          val = __dadd_rn(val, -q[-SAT_PX - 1]);          // S(x-1, y-1, z)
          val = __dadd_rn(val, -q[-SLICE - SAT_PX]);      // S(x,   y-1, z-1)
          val = __dadd_rn(val, -q[-SLICE - 1]);           // S(x-1, y,   z-1)
        }
        c[lx] = val;
      }
      __syncthreads();
    }
    if (gy < sh && gz < sd) {
      const size_t row = (size_t)sw * ((size_t)gy + (size_t)sh * (size_t)gz);
#pragma unroll 8
      for (int k = 0; k < SAT_TX; ++k) {
        const int gx = x0 + k;
        if (gx < sw) { const double v = c[k]; __stcg(S64 + row + gx, v); sat[row + gx] = (float)v; }
      }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicExch(done + ((size_t)TZ * ty_n + TY) * tx_n + TX, 1);
  }
}

static int run_sat_reference_tiles(vrb_ctx* c, const float* d_lut, float* d_sat) {
  const int sw = c->vw + 2, sh = c->vh + 2, sd = c->vd + 2;
  const int tx_n = (sw + SAT_TX - 1) / SAT_TX, ty_n = (sh + SAT_T - 1) / SAT_T, tz_n = (sd + SAT_T - 1) / SAT_T;
  VRB_REQUIRE(tx_n <= 1023 && ty_n <= 1023 && tz_n <= 1023, VRB_ERR_UNSUPPORTED, "vrb_sat_build: volume too large for the tiled reference-order build");
  const size_t n_tiles = (size_t)tx_n * ty_n * tz_n;
  // tiles in order of their diagonal (counting sort over X+Y+Z)
  std::vector<unsigned> order(n_tiles);
  {
    const int nd = tx_n + ty_n + tz_n - 2;
    std::vector<size_t> start((size_t)nd + 1, 0);
    for (int z = 0; z < tz_n; ++z) for (int y = 0; y < ty_n; ++y) for (int x = 0; x < tx_n; ++x) start[(size_t)(x + y + z) + 1]++;
    for (int d = 0; d < nd; ++d) start[(size_t)d + 1] += start[d];
    for (int z = 0; z < tz_n; ++z) for (int y = 0; y < ty_n; ++y) for (int x = 0; x < tx_n; ++x)
      order[start[(size_t)(x + y + z)]++] = (unsigned)x | ((unsigned)y << 10) | ((unsigned)z << 20);
  }
  unsigned* d_order = nullptr; unsigned* d_ticket = nullptr; int* d_done = nullptr; double* d_s64 = nullptr;
  const size_t n = (size_t)sw * sh * sd;
  cudaError_t e = cudaMalloc(&d_order, n_tiles * sizeof(unsigned));
  if (e == cudaSuccess) e = cudaMalloc(&d_ticket, sizeof(unsigned));
  if (e == cudaSuccess) e = cudaMalloc(&d_done, n_tiles * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&d_s64, n * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_order, order.data(), n_tiles * sizeof(unsigned), cudaMemcpyHostToDevice, c->stream);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  float ms = 0.f;
  if (e == cudaSuccess) {
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, c->stream);
    cudaMemsetAsync(d_ticket, 0, sizeof(unsigned), c->stream);
    cudaMemsetAsync(d_done, 0, n_tiles * sizeof(int), c->stream);
    const unsigned blocks = (unsigned)std::min<size_t>(n_tiles, (size_t)148 * 24);
    if (c->bpv == 1) k_sat_tiles<uint8_t><<<blocks, 64, 0, c->stream>>>((const uint8_t*)c->d_raw, d_lut, c->vw, c->vh, c->vd, d_order, (unsigned)n_tiles, tx_n, ty_n, d_ticket, d_done, d_s64, d_sat);
    else             k_sat_tiles<uint16_t><<<blocks, 64, 0, c->stream>>>((const uint16_t*)c->d_raw, d_lut, c->vw, c->vh, c->vd, d_order, (unsigned)n_tiles, tx_n, ty_n, d_ticket, d_done, d_s64, d_sat);
    e = cudaGetLastError();
    cudaEventRecord(e1, c->stream);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  c->last_prepass_ms = ms;
  if (d_order) cudaFree(d_order);
  if (d_ticket) cudaFree(d_ticket);
  if (d_done) cudaFree(d_done);
  if (d_s64) cudaFree(d_s64);
  VRB_CUDA(e);
  c->launches += 1;
  return VRB_OK;
}

static int run_sat_reference_order(vrb_ctx* c, const float* d_lut, float* d_sat) {
  const int sw = c->vw + 2, sh = c->vh + 2, sd = c->vd + 2;
  const size_t plane = (size_t)sh * sd;
  double* d_planes = nullptr;
  VRB_CUDA(cudaMalloc(&d_planes, 4 * plane * sizeof(double)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, c->stream);
  cudaMemsetAsync(d_planes, 0, 4 * plane * sizeof(double), c->stream);
  const unsigned blocks = (unsigned)((plane + 255) / 256);
  const int nplanes = sw + sh + sd - 2;
  for (int k = 0; k < nplanes; ++k) {
    double* p0 = d_planes + (size_t)(k & 3) * plane;
    const double* p1 = d_planes + (size_t)((k + 3) & 3) * plane;
    const double* p2 = d_planes + (size_t)((k + 2) & 3) * plane;
    const double* p3 = d_planes + (size_t)((k + 1) & 3) * plane;
    if (c->bpv == 1) k_sat_wavefront<uint8_t><<<blocks, 256, 0, c->stream>>>((const uint8_t*)c->d_raw, d_lut, c->vw, c->vh, c->vd, k, p1, p2, p3, p0, d_sat);
    else             k_sat_wavefront<uint16_t><<<blocks, 256, 0, c->stream>>>((const uint16_t*)c->d_raw, d_lut, c->vw, c->vh, c->vd, k, p1, p2, p3, p0, d_sat);
  }
  cudaError_t e = cudaGetLastError();
  cudaEventRecord(e1, c->stream);
  if (e == cudaSuccess) e = cudaEventSynchronize(e1);
  float ms = 0.f;
  if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
  c->last_prepass_ms = ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d_planes);
  VRB_CUDA(e);
  c->launches += (uint64_t)nplanes;
  return VRB_OK;
}

extern "C" int vrb_sat_set_order(vrb_ctx* c, int order) {
  VRB_REQUIRE(c, VRB_ERR_INVALID, "vrb_sat_set_order: NULL context");
  VRB_REQUIRE(order == VRB_SAT_ORDER_REFERENCE || order == VRB_SAT_ORDER_SCAN, VRB_ERR_INVALID, "vrb_sat_set_order: order %d", order);
  c->sat_order = order;
  return VRB_OK;
}
extern "C" int vrb_sat_get_order(const vrb_ctx* c) { return c ? c->sat_order : -1; }

extern "C" int vrb_sat_build(vrb_ctx* c, const float* ext_lut, int n_lut) {
  VRB_REQUIRE(c && ext_lut, VRB_ERR_INVALID, "vrb_sat_build: NULL argument");
  VRB_REQUIRE(c->d_raw, VRB_ERR_STATE, "vrb_sat_build: no volume uploaded");
  VRB_REQUIRE(n_lut == (c->bpv == 1 ? 256 : 65536), VRB_ERR_INVALID, "vrb_sat_build: LUT must have %d entries, got %d",
              c->bpv == 1 ? 256 : 65536, n_lut);
  VRB_CUDA(cudaSetDevice(c->device));
  const int w = c->vw + 2, h = c->vh + 2, d = c->vd + 2;
  const size_t n = (size_t)w * h * d;
  if (c->d_sat) { VRB_CUDA(cudaFree(c->d_sat)); c->d_sat = nullptr; }
  VRB_CUDA(cudaMalloc(&c->d_sat, n * sizeof(float)));
  float* d_lut = nullptr; double* d_tmp = nullptr;
  VRB_CUDA(cudaMalloc(&d_lut, (size_t)n_lut * sizeof(float)));
  const bool scan = c->sat_order == VRB_SAT_ORDER_SCAN;
  cudaError_t e = scan ? cudaMalloc(&d_tmp, n * sizeof(double)) : cudaSuccess;
  if (e != cudaSuccess) { cudaFree(d_lut); vrb_set_error("vrb_sat_build: cudaMalloc(%zu): %s", n * sizeof(double), cudaGetErrorString(e)); return VRB_ERR_CUDA; }
  int rc = VRB_OK;
  e = cudaMemcpyAsync(d_lut, ext_lut, (size_t)n_lut * sizeof(float), cudaMemcpyHostToDevice, c->stream);
  if (e != cudaSuccess) { vrb_set_error("vrb_sat_build: H2D: %s", cudaGetErrorString(e)); rc = VRB_ERR_CUDA; }
  if (rc == VRB_OK) {
    const char* how = getenv("VRB_SAT_REFERENCE");             // "planes": round 1's one launch per anti-diagonal plane (A/B)
    if (scan) rc = run_sat<double, float, float, 1>(c, d_lut, d_tmp, c->d_sat);
    else if (how && !strcmp(how, "planes")) rc = run_sat_reference_order(c, d_lut, c->d_sat);
    else rc = run_sat_reference_tiles(c, d_lut, c->d_sat);
  }
  cudaError_t es = cudaStreamSynchronize(c->stream);
  if (rc == VRB_OK && es != cudaSuccess) { vrb_set_error("vrb_sat_build: %s", cudaGetErrorString(es)); rc = VRB_ERR_CUDA; }
  cudaFree(d_lut); if (d_tmp) cudaFree(d_tmp);      // release the fp64 scratch before allocating the packed copy
  if (rc == VRB_OK) rc = pack_sat(c, n, w, h);
  if (rc == VRB_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) { vrb_set_error("vrb_sat_build: packing failed"); rc = VRB_ERR_CUDA; }
  if (rc == VRB_OK) { c->sat_w = w; c->sat_h = h; c->sat_d = d; }
  return rc;
}

// ---------------------------------------------------------------------------------------------------------
// Sharded build (SURVEY.md section 8e, row "SAT build"): the bordered grid is cut into z-slabs, one per GPU.  A rank scans its
// slab (the same three passes, z restarted at the slab's first slice), hands out its LAST fp64 plane, receives the sum of the
// last planes of the slabs below it (one all-gather of (W+2)(H+2) doubles per rank, cpp_volume_rendering_b200/dist.py), adds
// that plane to every slice of its slab and stores the float texels.  The float slabs are then exchanged so that every rank
// of a sort-first run holds the whole table, and vrb_sat_commit builds the marcher's gather atlas.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_sat_slab_finish(const double* __restrict__ S, const double* __restrict__ prefix, float* __restrict__ out, size_t plane, int nz) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (size_t)gridDim.x * blockDim.x) {
    const double p = prefix ? prefix[i] : 0.0;
    for (int z = 0; z < nz; ++z) out[(size_t)z * plane + i] = (float)(S[(size_t)z * plane + i] + p);
  }
}

static void free_slab(vrb_ctx* c) { if (c->d_sat_slab64) cudaFree(c->d_sat_slab64); c->d_sat_slab64 = nullptr; c->sat_slab_lo = c->sat_slab_hi = 0; }

extern "C" int vrb_sat_build_slab(vrb_ctx* c, const float* ext_lut, int n_lut, int z_lo, int z_hi) {
  VRB_REQUIRE(c && ext_lut, VRB_ERR_INVALID, "vrb_sat_build_slab: NULL argument");
  VRB_REQUIRE(c->d_raw, VRB_ERR_STATE, "vrb_sat_build_slab: no volume uploaded");
  VRB_REQUIRE(n_lut == (c->bpv == 1 ? 256 : 65536), VRB_ERR_INVALID, "vrb_sat_build_slab: LUT must have %d entries, got %d", c->bpv == 1 ? 256 : 65536, n_lut);
  const int w = c->vw + 2, h = c->vh + 2, d = c->vd + 2;
  VRB_REQUIRE(z_lo >= 0 && z_lo < z_hi && z_hi <= d, VRB_ERR_INVALID, "vrb_sat_build_slab: slab [%d, %d) of %d slices", z_lo, z_hi, d);
  VRB_CUDA(cudaSetDevice(c->device));
  const size_t plane = (size_t)w * h, n = plane * d;
  if (!c->d_sat || c->sat_w != w || c->sat_h != h || c->sat_d != d) {
    if (c->d_sat) { VRB_CUDA(cudaFree(c->d_sat)); c->d_sat = nullptr; }
    VRB_CUDA(cudaMalloc(&c->d_sat, n * sizeof(float)));
    c->sat_w = w; c->sat_h = h; c->sat_d = d;
  }
  free_slab(c);
  const int nz = z_hi - z_lo;
  VRB_CUDA(cudaMalloc(&c->d_sat_slab64, plane * nz * sizeof(double)));
  c->sat_slab_lo = z_lo; c->sat_slab_hi = z_hi;
  float* d_lut = nullptr;
  VRB_CUDA(cudaMalloc(&d_lut, (size_t)n_lut * sizeof(float)));
  cudaError_t e = cudaMemcpyAsync(d_lut, ext_lut, (size_t)n_lut * sizeof(float), cudaMemcpyHostToDevice, c->stream);
  double* S = c->d_sat_slab64;
  const long long rows = (long long)h * nz;
  const int blocks_x = (int)std::min<long long>((rows + 7) / 8, 148LL * 64);
  if (e == cudaSuccess) {
    if (c->bpv == 1) k_sat_fill_scan_x<double, uint8_t, float, 1><<<blocks_x, 256, 0, c->stream>>>((const uint8_t*)c->d_raw, d_lut, S, c->vw, c->vh, c->vd, z_lo, nz);
    else             k_sat_fill_scan_x<double, uint16_t, float, 1><<<blocks_x, 256, 0, c->stream>>>((const uint16_t*)c->d_raw, d_lut, S, c->vw, c->vh, c->vd, z_lo, nz);
    const long long ny = (long long)w * nz, nzc = (long long)w * h;
    k_sat_scan_y<double><<<(unsigned)((ny + 127) / 128), 128, 0, c->stream>>>(S, w, h, nz);
    k_sat_scan_z<double, double><<<(unsigned)((nzc + 127) / 128), 128, 0, c->stream>>>(S, S, w, h, nz);
    e = cudaGetLastError();
    c->launches += 3;
  }
  cudaError_t es = cudaStreamSynchronize(c->stream);
  cudaFree(d_lut);
  VRB_CUDA(e); VRB_CUDA(es);
  return VRB_OK;
}

extern "C" int vrb_sat_slab_plane(vrb_ctx* c, void** dev_plane, size_t* count) {
  VRB_REQUIRE(c && dev_plane && count, VRB_ERR_INVALID, "vrb_sat_slab_plane: NULL argument");
  VRB_REQUIRE(c->d_sat_slab64, VRB_ERR_STATE, "vrb_sat_slab_plane: no slab built (vrb_sat_build_slab)");
  const size_t plane = (size_t)c->sat_w * c->sat_h;
  *dev_plane = c->d_sat_slab64 + plane * (size_t)(c->sat_slab_hi - c->sat_slab_lo - 1);
  *count = plane;
  return VRB_OK;
}

extern "C" int vrb_sat_finish_slab(vrb_ctx* c, const void* dev_prefix_plane) {
  VRB_REQUIRE(c, VRB_ERR_INVALID, "vrb_sat_finish_slab: ctx is NULL");
  VRB_REQUIRE(c->d_sat_slab64 && c->d_sat, VRB_ERR_STATE, "vrb_sat_finish_slab: no slab built (vrb_sat_build_slab)");
  VRB_CUDA(cudaSetDevice(c->device));
  const size_t plane = (size_t)c->sat_w * c->sat_h;
  const int nz = c->sat_slab_hi - c->sat_slab_lo;
  k_sat_slab_finish<<<(unsigned)std::min<size_t>((plane + 255) / 256, 148 * 16), 256, 0, c->stream>>>(
      c->d_sat_slab64, (const double*)dev_prefix_plane, c->d_sat + plane * (size_t)c->sat_slab_lo, plane, nz);
  VRB_CUDA(cudaGetLastError());
  c->launches++;
  VRB_CUDA(cudaStreamSynchronize(c->stream));
  free_slab(c);
  return VRB_OK;
}

extern "C" int vrb_sat_device_ptr(vrb_ctx* c, void** dev, int dims[3]) {
  VRB_REQUIRE(c && dev, VRB_ERR_INVALID, "vrb_sat_device_ptr: NULL argument");
  VRB_REQUIRE(c->d_sat, VRB_ERR_STATE, "vrb_sat_device_ptr: no SAT");
  *dev = c->d_sat;
  if (dims) { dims[0] = c->sat_w; dims[1] = c->sat_h; dims[2] = c->sat_d; }
  return VRB_OK;
}

extern "C" int vrb_sat_commit(vrb_ctx* c) {
  VRB_REQUIRE(c, VRB_ERR_INVALID, "vrb_sat_commit: ctx is NULL");
  VRB_REQUIRE(c->d_sat, VRB_ERR_STATE, "vrb_sat_commit: no SAT");
  VRB_CUDA(cudaSetDevice(c->device));
  int rc = pack_sat(c, (size_t)c->sat_w * c->sat_h * c->sat_d, c->sat_w, c->sat_h);
  if (rc == VRB_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) { vrb_set_error("vrb_sat_commit: packing failed"); rc = VRB_ERR_CUDA; }
  return rc;
}

extern "C" int vrb_sat_build_u64(vrb_ctx* c, const uint32_t* lut_u32, int n_lut, uint64_t* host_out) {
  VRB_REQUIRE(c && lut_u32 && host_out, VRB_ERR_INVALID, "vrb_sat_build_u64: NULL argument");
  VRB_REQUIRE(c->d_raw, VRB_ERR_STATE, "vrb_sat_build_u64: no volume uploaded");
  VRB_REQUIRE(n_lut == (c->bpv == 1 ? 256 : 65536), VRB_ERR_INVALID, "vrb_sat_build_u64: LUT must have %d entries", c->bpv == 1 ? 256 : 65536);
  VRB_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)c->vw * c->vh * c->vd;
  uint32_t* d_lut = nullptr; unsigned long long *d_tmp = nullptr, *d_out = nullptr;
  VRB_CUDA(cudaMalloc(&d_lut, (size_t)n_lut * sizeof(uint32_t)));
  cudaError_t e1 = cudaMalloc(&d_tmp, n * 8), e2 = cudaMalloc(&d_out, n * 8);
  int rc = VRB_OK;
  if (e1 != cudaSuccess || e2 != cudaSuccess) { vrb_set_error("vrb_sat_build_u64: cudaMalloc failed"); rc = VRB_ERR_CUDA; }
  if (rc == VRB_OK && cudaMemcpyAsync(d_lut, lut_u32, (size_t)n_lut * 4, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { vrb_set_error("vrb_sat_build_u64: H2D failed"); rc = VRB_ERR_CUDA; }
  if (rc == VRB_OK) rc = run_sat<unsigned long long, uint32_t, unsigned long long, 0>(c, d_lut, d_tmp, d_out);
  if (rc == VRB_OK && cudaMemcpyAsync(host_out, d_out, n * 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { vrb_set_error("vrb_sat_build_u64: D2H failed"); rc = VRB_ERR_CUDA; }
  cudaError_t es = cudaStreamSynchronize(c->stream);
  if (rc == VRB_OK && es != cudaSuccess) { vrb_set_error("vrb_sat_build_u64: %s", cudaGetErrorString(es)); rc = VRB_ERR_CUDA; }
  cudaFree(d_lut); if (d_tmp) cudaFree(d_tmp); if (d_out) cudaFree(d_out);
  return rc;
}

extern "C" int vrb_sat_read(vrb_ctx* c, float* host_out) {
  VRB_REQUIRE(c && host_out, VRB_ERR_INVALID, "vrb_sat_read: NULL argument");
  VRB_REQUIRE(c->d_sat, VRB_ERR_STATE, "vrb_sat_read: no SAT built");
  VRB_CUDA(cudaSetDevice(c->device));
  size_t n = (size_t)c->sat_w * c->sat_h * c->sat_d;
  VRB_CUDA(cudaMemcpyAsync(host_out, c->d_sat, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  VRB_CUDA(cudaStreamSynchronize(c->stream));
  return VRB_OK;
}
