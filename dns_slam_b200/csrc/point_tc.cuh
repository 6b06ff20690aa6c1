// Shared pieces of the tcgen05 point kernels (point_tc.cu: weight preparation; point_tc2.cu: the kernels).
#pragma once
#include "tc_common.cuh"

namespace dns {


// shared-memory carve (bytes)
constexpr int kXTile = 10 * 2048;       // one half of the X tile [10 chunks][128][16]
constexpr int kW1Tile = 10 * 64 * 16;   // one half of the combined W1 tile [10 chunks][64 rows][16]
constexpr int kW2Tile = 4 * 48 * 16;    // one half of a W2 tile [4 chunks][48 rows][16]
constexpr int kDOTile = 6 * 2048;       // one half of a dOut tile [6 chunks][128][16]

// combined layer-1 tile (the forward kernel fetches the layer-2 tiles later, over this one)
__device__ __forceinline__ void load_w1_tc(unsigned char* W1_hi, unsigned char* W1_lo, const uint4* __restrict__ wc,
                                           const uint4* __restrict__ we, bool fine) {
  const uint4 z4 = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < 640; i += blockDim.x) {  // rows 0..31 coarse, 32..63 expert
    int c = i >> 6, j = i & 63;
    uint4 h, l;
    if (j < 32) {
      h = wc[c * 32 + j];
      l = wc[320 + c * 32 + j];
    } else if (fine) {
      h = we[c * 32 + j - 32];
      l = we[320 + c * 32 + j - 32];
    } else {
      h = l = z4;
    }
    reinterpret_cast<uint4*>(W1_hi)[i] = h;
    reinterpret_cast<uint4*>(W1_lo)[i] = l;
  }
}


}  // namespace dns
