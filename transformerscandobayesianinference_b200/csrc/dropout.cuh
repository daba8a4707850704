// Counter-based dropout masks (nothing is stored: forward and backward regenerate the same bits).
//
// The reference applies torch dropout at four sites of every encoder layer (torch:nn/modules/transformer.py:961-982 and
// torch:nn/functional.py multi_head_attention_forward): on the softmax probabilities, on the attention block's output, after
// the GELU and on the MLP block's output.  Here the keep decision of element (row, col) of a site is one byte of a 32-bit
// integer hash of (site seed, row, col / 4):   keep  <=>  byte >= thr,   thr = round(256 p)  in [0, 255]
// so the effective drop probability is thr / 256 (p = 0.2 -> 51/256 = 0.1992, p = 0.5 exact) and kept values are scaled
// by 256 / (256 - thr).  Four neighbouring columns share one hash evaluation.
#pragma once
#include <stdint.h>

namespace pfn {

__host__ __device__ __forceinline__ uint32_t drop_hash(uint32_t seed, uint32_t row, uint32_t col4) {
  uint32_t x = (row * 0x9E3779B1u) ^ (col4 * 0x85EBCA77u) ^ seed;
  x ^= x >> 16; x *= 0x7FEB352Du;
  x ^= x >> 15; x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ bool drop_keep_byte(uint32_t h, int k, int thr) { return static_cast<int>((h >> (8 * k)) & 0xFFu) >= thr; }
__host__ __device__ __forceinline__ bool drop_keep(uint32_t seed, uint32_t row, uint32_t col, int thr) {
  return drop_keep_byte(drop_hash(seed, row, col >> 2), static_cast<int>(col & 3u), thr);
}
__host__ __device__ __forceinline__ float drop_scale(int thr) { return 256.0f / static_cast<float>(256 - thr); }

}  // namespace pfn
