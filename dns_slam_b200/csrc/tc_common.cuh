// Helpers shared by the tcgen05 point / ray kernels: unrolled 16-bin OneBlob in registers and the
// bf16 hi/lo chunk writer of the canonical UMMA operand layout  [chunk][point][8 x bf16].
#pragma once
#include "render.cuh"
#include "umma.cuh"

namespace dns {

__device__ __forceinline__ void oneblob16(float x, float (&o)[16]) {
  float prev = cdf3(0.0f - x, 16.0f);
#pragma unroll
  for (int b = 0; b < 16; ++b) {
    float cur = cdf3((float)(b + 1) / 16.0f - x, 16.0f);
    o[b] = cur - prev;
    prev = cur;
  }
}
__device__ __forceinline__ float oneblob16_bwd(float x, const float (&d)[16]) {
  float prev = pdf3(0.0f - x, 16.0f), acc = 0.f;
#pragma unroll
  for (int b = 0; b < 16; ++b) {
    float cur = pdf3((float)(b + 1) / 16.0f - x, 16.0f);
    acc += d[b] * (cur - prev);
    prev = cur;
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
