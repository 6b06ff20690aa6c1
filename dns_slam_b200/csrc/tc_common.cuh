// Helpers shared by the tcgen05 point / ray kernels: unrolled 16-bin OneBlob in registers and the
// bf16 hi/lo chunk writer of the canonical UMMA operand layout  [chunk][point][8 x bf16].
#pragma once
#include "render.cuh"
#include "umma.cuh"

namespace dns {

// The quartic kernel has support +-1/16, so only the bins floor(16x)-1 .. floor(16x)+1 (periodic) are
// non-zero: those three are evaluated with the very formula of oneblob_fwd (identical values), the other
// thirteen are exact zeros there too (saturated CDFs cancel).  x far outside [0,1] takes the dense path.
__device__ __forceinline__ void oneblob16(float x, float (&o)[16]) {
  if (x < -0.5f || x > 1.5f) {
    float prev = cdf3(0.0f - x, 16.0f);
#pragma unroll
    for (int b = 0; b < 16; ++b) {
      float cur = cdf3((float)(b + 1) / 16.0f - x, 16.0f);
      o[b] = cur - prev;
      prev = cur;
    }
    return;
  }
  const int bx = (int)floorf(x * 16.0f);
  int bb[3];
  float vv[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int b = (bx + k - 1) & 15;
    bb[k] = b;
    vv[k] = cdf3((float)(b + 1) / 16.0f - x, 16.0f) - cdf3((float)b / 16.0f - x, 16.0f);
  }
#pragma unroll
  for (int b = 0; b < 16; ++b) o[b] = b == bb[0] ? vv[0] : (b == bb[1] ? vv[1] : (b == bb[2] ? vv[2] : 0.0f));
}
__device__ __forceinline__ float oneblob16_bwd(float x, const float (&d)[16]) {
  if (x < -0.5f || x > 1.5f) {
    float prev = pdf3(0.0f - x, 16.0f), acc = 0.f;
#pragma unroll
    for (int b = 0; b < 16; ++b) {
      float cur = pdf3((float)(b + 1) / 16.0f - x, 16.0f);
      acc += d[b] * (cur - prev);
      prev = cur;
    }
    return -16.0f * acc;
  }
  const int bx = (int)floorf(x * 16.0f);
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int b = (bx + k - 1) & 15;
    float db = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) db = j == b ? d[j] : db;
    acc += db * (pdf3((float)(b + 1) / 16.0f - x, 16.0f) - pdf3((float)b / 16.0f - x, 16.0f));
  }
  return -16.0f * acc;
}
__device__ __forceinline__ void put_chunk(unsigned char* hi_tile, unsigned char* lo_tile, int chunk, int cs, int point,
                                          const float* v8) {
  uint4 h, l;
  split8(make_float4(v8[0], v8[1], v8[2], v8[3]), make_float4(v8[4], v8[5], v8[6], v8[7]), h, l);
  *reinterpret_cast<uint4*>(hi_tile + chunk * cs + point * 16) = h;
  *reinterpret_cast<uint4*>(lo_tile + chunk * cs + point * 16) = l;
}

}  // namespace dns
