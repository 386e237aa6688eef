// shade_list.cuh -- deferred shading of the lit marchers (rc1pdosct, rc1pextbsd, rc1pcrtgt, rc1pvctsg), sm_100a.
//
// In every lit shader of the reference the opacity a ray accumulates, and therefore where it terminates, depends only on
// the transfer function (a = 1 - exp(-tau h); e.g. rc1pdosct/ray_bbox_marching.comp:700-716), never on the lighting
// terms.  A frame is therefore split into three kernels instead of one:
//
//   march      one lane per ray: volume tap + TF per step, opacity accumulation and the 0.99 cut exactly as in the shader;
//              every sample with alpha > 0 is APPENDED to a list (position, pixel, TF colour, alpha).  Cheap and divergent.
//   shade      one lane per list entry: the expensive lighting term (cone taps, SAT boxes, secondary rays).  Every entry
//              costs the same, all 32 lanes work, no ray is longer than another: no tail, no longest-CTA floor.
//   composite  one lane per ray again: walks the ray's entries in march order and applies the shader's front-to-back
//              arithmetic to the shaded colours, operation by operation.
//
// The list is appended warp by warp: the 8x4 rays of a warp each park on their next visible sample, then the warp takes
// one contiguous CHUNK of the entry array (one atomicAdd) and one chunk header {lane mask, first entry, next chunk};
// headers of a warp form a singly linked list starting at head[warp].  Entries of a chunk belong to neighbouring rays at
// similar depth, so the shade kernel's warps (32 consecutive entries) fetch neighbouring texels.
#pragma once

#define VRB_SL_NONE 0xffffffffu

struct ShadeListView {
  float4* a;            // per entry, written by march: sample position (texture space) xyz, pixel index (int bits)
  float4* b;            // per entry: march writes TF rgb + alpha of the step; shade overwrites rgb with the lit colour
  uint4* hdr;           // per chunk: x = lane mask, y = first entry, z = next chunk of the same warp (VRB_SL_NONE = last)
  unsigned* head;       // per marching warp: first chunk (VRB_SL_NONE = the warp appended nothing)
  unsigned* counters;   // [0] entries appended, [1] chunks appended (both keep counting past the capacity)
  unsigned capacity;    // entries (and chunks) the arrays hold
};

#ifdef __CUDACC__
// Warp-collective append.  `has`: this lane parks a sample.  Returns the lane's entry index, or VRB_SL_NONE when the lane
// has nothing or the list is full (the counters still advance, so the host learns the size it needs and repeats the frame).
__device__ __forceinline__ unsigned sl_append(const ShadeListView& L, bool has, unsigned& last_chunk, unsigned warp_id, unsigned lane) {
  const unsigned mask = __ballot_sync(0xffffffffu, has);
  if (!mask) return VRB_SL_NONE;
  const unsigned n = __popc(mask);
  unsigned base = 0, h = 0;
  if (lane == 0) { base = atomicAdd(&L.counters[0], n); h = atomicAdd(&L.counters[1], 1u); }
  base = __shfl_sync(0xffffffffu, base, 0);
  h = __shfl_sync(0xffffffffu, h, 0);
  if (base + n > L.capacity || base + n < base) return VRB_SL_NONE;
  if (lane == 0) {
    L.hdr[h] = make_uint4(mask, base, VRB_SL_NONE, 0u);
    if (last_chunk == VRB_SL_NONE) L.head[warp_id] = h; else L.hdr[last_chunk].z = h;
  }
  last_chunk = h;
  return has ? base + __popc(mask & ((1u << lane) - 1u)) : VRB_SL_NONE;
}
#endif
