// cta_order.cu -- longest-first CTA schedule from the previous frame's measured CTA durations.
// The EBS marcher's CTAs differ in cost by orders of magnitude (a ray that crosses semi-transparent material shades a
// hundred samples, most rays four), and a sort-first rank that owns 1/8 of the frame has too few CTAs to hide the long
// ones: ncu shows 16 % resident warps instead of 22 % and a ~0.5 ms tail on a 1.1 ms kernel.  Consecutive frames of an
// interactive session are nearly identical, so the durations measured in frame i give the launch order of frame i+1
// (longest processing time first).  The image does not depend on the order; only the tail does.
#include "vrb_internal.cuh"
#include <cub/device/device_radix_sort.cuh>

__global__ void k_iota(unsigned int* p, unsigned int n) {
  unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

void vrb_free_cta_order(vrb_ctx* c) {
  if (c->d_cta_cost) cudaFree(c->d_cta_cost);
  if (c->d_cta_keys) cudaFree(c->d_cta_keys);
  if (c->d_cta_keys_sorted) cudaFree(c->d_cta_keys_sorted);
  if (c->d_cta_iota) cudaFree(c->d_cta_iota);
  if (c->d_cta_order) cudaFree(c->d_cta_order);
  if (c->d_cta_sort_tmp) cudaFree(c->d_cta_sort_tmp);
  c->d_cta_cost = c->d_cta_keys = c->d_cta_keys_sorted = c->d_cta_iota = c->d_cta_order = nullptr;
  c->d_cta_sort_tmp = nullptr; c->cta_sort_tmp_bytes = 0; c->cta_n = 0; c->cta_cost_valid = false;
}

int vrb_cta_order_prepare(vrb_ctx* c, unsigned n_ctas, unsigned long long sig, const unsigned int** order, unsigned int** cost) {
  static const int enabled = getenv("VRB_CTA_ORDER") ? atoi(getenv("VRB_CTA_ORDER")) : 1;
  *order = nullptr; *cost = nullptr;
  if (!enabled || n_ctas < 2) return VRB_OK;
  if (c->cta_n != n_ctas) {
    vrb_free_cta_order(c);
    const size_t b = (size_t)n_ctas * sizeof(unsigned int);
    VRB_CUDA(cudaMalloc(&c->d_cta_cost, b)); VRB_CUDA(cudaMalloc(&c->d_cta_keys, b)); VRB_CUDA(cudaMalloc(&c->d_cta_keys_sorted, b));
    VRB_CUDA(cudaMalloc(&c->d_cta_iota, b)); VRB_CUDA(cudaMalloc(&c->d_cta_order, b));
    k_iota<<<(n_ctas + 255) / 256, 256, 0, c->stream>>>(c->d_cta_iota, n_ctas);
    size_t tmp = 0;
    VRB_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp, c->d_cta_keys, c->d_cta_keys_sorted, c->d_cta_iota, c->d_cta_order, (int)n_ctas, 0, 32, c->stream));
    VRB_CUDA(cudaMalloc(&c->d_cta_sort_tmp, tmp));
    c->cta_sort_tmp_bytes = tmp;
    c->cta_n = n_ctas;
    c->cta_cost_valid = false;
  }
  if (c->cta_sig != sig) { c->cta_cost_valid = false; c->cta_sig = sig; }
  if (c->cta_cost_valid) {
    // previous frame's durations -> this frame's order (the cost buffer is about to be overwritten)
    VRB_CUDA(cudaMemcpyAsync(c->d_cta_keys, c->d_cta_cost, (size_t)n_ctas * sizeof(unsigned int), cudaMemcpyDeviceToDevice, c->stream));
    size_t tmp = c->cta_sort_tmp_bytes;
    VRB_CUDA(cub::DeviceRadixSort::SortPairsDescending(c->d_cta_sort_tmp, tmp, c->d_cta_keys, c->d_cta_keys_sorted, c->d_cta_iota, c->d_cta_order,
                                                       (int)n_ctas, 0, 32, c->stream));
    *order = c->d_cta_order;
    c->launches += 6;     // cub radix sort of 32-bit keys: histogram, scan, four onesweep passes
  }
  VRB_CUDA(cudaMemsetAsync(c->d_cta_cost, 0, (size_t)n_ctas * sizeof(unsigned int), c->stream));
  *cost = c->d_cta_cost;
  c->cta_cost_valid = true;     // valid once the launch that follows has run
  return VRB_OK;
}
