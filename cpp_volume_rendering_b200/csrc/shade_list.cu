// shade_list.cu -- storage of the deferred shading list (shade_list.cuh): entry arrays, chunk headers, per-warp heads.
#include "vrb_internal.cuh"
#include <cstdlib>
#include <algorithm>

void vrb_free_shade_list(vrb_ctx* c) {
  if (c->d_sl_a) cudaFree(c->d_sl_a);
  if (c->d_sl_b) cudaFree(c->d_sl_b);
  if (c->d_sl_next) cudaFree(c->d_sl_next);
  if (c->d_sl_head) cudaFree(c->d_sl_head);
  if (c->d_sl_counters) cudaFree(c->d_sl_counters);
  if (c->h_sl_counters) cudaFreeHost(c->h_sl_counters);
  c->d_sl_a = c->d_sl_b = nullptr; c->d_sl_next = nullptr; c->d_sl_head = c->d_sl_counters = c->h_sl_counters = nullptr;
  c->sl_capacity = c->sl_heads = 0;
}

static int sl_alloc_entries(vrb_ctx* c, unsigned capacity) {
  if (c->d_sl_a) { VRB_CUDA(cudaFree(c->d_sl_a)); c->d_sl_a = nullptr; }
  if (c->d_sl_b) { VRB_CUDA(cudaFree(c->d_sl_b)); c->d_sl_b = nullptr; }
  if (c->d_sl_next) { VRB_CUDA(cudaFree(c->d_sl_next)); c->d_sl_next = nullptr; }
  c->sl_capacity = 0;
  VRB_CUDA(cudaMalloc(&c->d_sl_a, (size_t)capacity * sizeof(float4)));
  VRB_CUDA(cudaMalloc(&c->d_sl_b, (size_t)capacity * sizeof(float4)));
  VRB_CUDA(cudaMalloc(&c->d_sl_next, (size_t)capacity * sizeof(unsigned)));
  c->sl_capacity = capacity;
  return VRB_OK;
}

int vrb_sl_begin(vrb_ctx* c, unsigned n_lanes, ShadeListView* out) {
  if (!c->d_sl_counters) {
    VRB_CUDA(cudaMalloc(&c->d_sl_counters, 4 * sizeof(unsigned)));
    VRB_CUDA(cudaMallocHost(&c->h_sl_counters, 2 * sizeof(unsigned)));
  }
  if (c->sl_heads < n_lanes) {
    if (c->d_sl_head) { VRB_CUDA(cudaFree(c->d_sl_head)); c->d_sl_head = nullptr; }
    VRB_CUDA(cudaMalloc(&c->d_sl_head, (size_t)n_lanes * sizeof(unsigned)));
    c->sl_heads = n_lanes;
  }
  if (!c->sl_capacity) {
    // first guess: six visible samples per pixel of the frame (config 3 needs 3.6); a frame that needs more is marched
    // twice once (vrb_sl_counts grows the list to what the march asked for).  Multiple of the chunk size.
    unsigned long long cap = 6ull * (unsigned long long)c->fw * (unsigned long long)c->fh / (unsigned long long)std::max(1, c->part.nranks);
    if (const char* e = getenv("VRB_SL_CAPACITY")) cap = strtoull(e, nullptr, 10);
    cap = std::min<unsigned long long>(std::max<unsigned long long>(cap, 1024ull), 0xffffff00ull) & ~31ull;
    int rc = sl_alloc_entries(c, (unsigned)cap);
    if (rc != VRB_OK) return rc;
  }
  VRB_CUDA(cudaMemsetAsync(c->d_sl_counters, 0, 4 * sizeof(unsigned), c->stream));
  // unused slots are recognised by pixel == -1: all bits set in a[]
  VRB_CUDA(cudaMemsetAsync(c->d_sl_a, 0xff, (size_t)c->sl_capacity * sizeof(float4), c->stream));
  out->a = c->d_sl_a; out->b = c->d_sl_b; out->next = c->d_sl_next; out->head = c->d_sl_head; out->counters = c->d_sl_counters;
  out->capacity = c->sl_capacity;
  return VRB_OK;
}

int vrb_sl_counts(vrb_ctx* c, unsigned* entries, bool* overflow) {
  VRB_CUDA(cudaMemcpyAsync(c->h_sl_counters, c->d_sl_counters, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  VRB_CUDA(cudaStreamSynchronize(c->stream));
  const unsigned n = c->h_sl_counters[0];
  *entries = n; *overflow = false;
  c->sl_last_entries = n; c->sl_last_chunks = c->h_sl_counters[1];
  if (n > c->sl_capacity) {
    VRB_REQUIRE(n < 0xf0000000u, VRB_ERR_UNSUPPORTED, "deferred shading list: %u visible samples in one frame", n);
    const unsigned long long want = std::min<unsigned long long>((unsigned long long)n + n / 4 + 1024, 0xffffff00ull) & ~31ull;
    int rc = sl_alloc_entries(c, (unsigned)want);
    if (rc != VRB_OK) return rc;
    *overflow = true;
  }
  return VRB_OK;
}
