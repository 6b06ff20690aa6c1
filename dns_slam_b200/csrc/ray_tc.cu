// Weight preparation and launcher of the ray kernel with the two wide per-point GEMMs on tcgen05 (sm_100a); the kernel
// itself is k_ray_tc2 (ray_tc2.cu).
//
// Same mathematics as k_ray (render.cu; slams/tracking.py:188-214, slams/mapping.py:603-635,
// utils/common.py:506-537): colour + logit layer 1 per point, colour head, logit layer 2 on the
// composited hidden state, occupancy compositing, p/d/l losses and the full backward.  The difference is
// WHERE the two 112x64 contractions per point run:
//
//   forward   H[p][0..63]  = X[p][0..111] . W1^T      -> tcgen05.mma, M = 128 points, N = 64, K = 112
//   backward  dX[p][0..111] = dH[p][0..63] . W1       -> tcgen05.mma, M = 128 points, N = 112, K = 64
//
// Operands are bf16 hi + lo halves (three products hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM), so
// the results stay inside the 1e-3 parity bar.  Each thread owns one sample point = one TMEM lane: it
// writes its row of X / dH as 16-byte feature chunks  [chunk][point][8 x bf16]  (K-major canonical UMMA
// layout without swizzle: SBO = 128 B between 8-point groups, LBO = T*16 B between feature chunks) and
// reads its accumulator row back with tcgen05.ld.32x32b.  The SAME shared-memory copy of W1
// ([feature chunk][hidden row][8 features]) is the K-major B operand of the forward GEMM and the
// MN-major B operand of the backward GEMM, so no transposed copy exists.
#include "tc_common.cuh"

namespace dns {


// colour | logit layer-1 weights -> bf16 hi / lo chunk tiles [14 feature chunks][64 hidden rows][8 features]
__global__ void k_prep_w1o_tc(const float* __restrict__ color, const float* __restrict__ logit, uint4* __restrict__ hi,
                              uint4* __restrict__ lo) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 14 * 64) return;
  int c = i >> 6, j = i & 63;
  const float* src = (j < 32 ? color + j * kIn2 : logit + (j - 32) * kIn2) + 8 * c;
  float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
  uint4 h, l;
  split8(a, b, h, l);
  hi[i] = h;
  lo[i] = l;
}

void pick_ray_block_any(int S, int& T, int& RPC) { pick_ray_block_tc2(S, T, RPC); }

int launch_ray_tc(const RayArgs& ra, const float* color, const float* logit, uint4* w1_hi, uint4* w1_lo, bool prep,
                  int64_t n_rays_chunk, cudaStream_t st) {
  if (prep) k_prep_w1o_tc<<<(14 * 64 + 127) / 128, 128, 0, st>>>(color, logit, w1_hi, w1_lo);
  return launch_ray_tc2(ra, w1_hi, w1_lo, n_rays_chunk, st);
}

}  // namespace dns
