// shade_list.cuh -- deferred shading of the lit marchers (rc1pdosct, rc1pextbsd, rc1pcrtgt, rc1pvctsg), sm_100a.
//
// In every lit shader of the reference the opacity a ray accumulates, and therefore where it terminates, depends only on
// the transfer function (a = 1 - exp(-tau h); e.g. rc1pdosct/ray_bbox_marching.comp:700-716), never on the lighting
// terms.  A frame is therefore split into three kernels instead of one:
//
//   march      one lane per ray: volume tap + TF per step, opacity accumulation and the 0.99 cut exactly as in the shader;
//              every sample with alpha > 0 is APPENDED to a list (position, pixel, TF colour, alpha).  Cheap and divergent.
//   shade      one lane (or one warp task) per list entry: the expensive lighting term (cone taps, SAT boxes, secondary
//              rays).  Every entry costs about the same, all 32 lanes work, no ray is longer than another: no tail, no
//              longest-CTA floor, and a sort-first rank's share of the frame is balanced by construction.
//   composite  one lane per ray again: walks the ray's entries in march order and applies the shader's front-to-back
//              arithmetic to the shaded colours, operation by operation.
//
// Appending never makes a lane wait for rays that have nothing to append: the warp owns a current CHUNK of VRB_SL_CHUNK
// consecutive slots (one 64-bit word {first slot, slots used} in shared memory); the lanes that append in the same loop
// iteration take their slots with one compare-and-swap on that word, and a full chunk is replaced with ONE atomicAdd on the
// global cursor.  Slots of a
// chunk therefore belong to the 8x4 neighbouring rays of one warp at similar depth, so the shade kernel's warps (32
// consecutive slots) fetch neighbouring texels.  The entries of one ray are chained through next[]; slots that were
// reserved but never written (the tail of a warp's last chunk) keep pixel == -1 and are skipped by the shade kernels.
#pragma once

#define VRB_SL_NONE 0xffffffffu
#define VRB_SL_CHUNK 32u

struct ShadeListView {
  float4* a;            // per slot, written by march: sample position (texture space) xyz, pixel index (int bits; -1 = unused slot)
  float4* b;            // per slot: march writes TF rgb + alpha of the step; shade overwrites rgb with the lit colour
  unsigned* next;       // per slot: the same ray's next entry (VRB_SL_NONE = last)
  unsigned* head;       // per pixel of the frame: the ray's first entry (VRB_SL_NONE = none); written for every ray that hits the box
  unsigned* counters;   // [0] slots reserved (keeps counting past the capacity), [1] entries written, [2] work cursor of a shade kernel
  unsigned capacity;    // slots the arrays hold
};

#ifdef __CUDACC__
// Append for the lanes of a warp that reach this call together (any subset; they are found with __activemask()): the
// lowest of them reserves their n slots at once and hands them out by rank.  `wc`: the warp's chunk word in shared memory
// {first slot : 32, slots used : 32}, initialised to VRB_SL_CHUNK used (= "full": the first push reserves a chunk); a word
// with more than VRB_SL_CHUNK used slots is being replaced by another group of the same warp (possible with independent
// thread scheduling) and is polled.  A request that does not fit takes the rest of the current chunk and the head of a new
// one.  Returns the lane's slot, or VRB_SL_NONE when the list is full (the cursor still advances, so the host learns the
// size it needs and repeats the frame).
__device__ __forceinline__ unsigned sl_push(const ShadeListView& L, unsigned long long* wc, unsigned lane) {
  const unsigned m = __activemask();
  const unsigned n = (unsigned)__popc(m), rank = (unsigned)__popc(m & ((1u << lane) - 1u));
  const int leader = __ffs((int)m) - 1;
  unsigned b0 = 0, b1 = 0, k = 0;
  if ((int)lane == leader) {
    for (;;) {
      const unsigned long long w = *reinterpret_cast<volatile unsigned long long*>(wc);
      const unsigned used = (unsigned)w, base = (unsigned)(w >> 32);
      if (used > VRB_SL_CHUNK) continue;                                   // being replaced
      k = min(n, VRB_SL_CHUNK - used);
      if (k == n) {
        if (atomicCAS(wc, w, w + (unsigned long long)n) == w) { b0 = base + used; break; }
      } else if (atomicCAS(wc, w, ((unsigned long long)base << 32) | (unsigned long long)(VRB_SL_CHUNK + 1u)) == w) {
        const unsigned g = atomicAdd(&L.counters[0], VRB_SL_CHUNK);
        b0 = base + used; b1 = g;
        atomicExch(wc, ((unsigned long long)g << 32) | (unsigned long long)(n - k));
        break;
      }
    }
  }
  b0 = __shfl_sync(m, b0, leader); b1 = __shfl_sync(m, b1, leader); k = __shfl_sync(m, k, leader);
  const unsigned slot = rank < k ? b0 + rank : b1 + (rank - k);
  return slot < L.capacity ? slot : VRB_SL_NONE;
}
#endif
