// Weight preparation for the point kernels with the MLP GEMMs on tcgen05 (sm_100a); the kernels are in point_tc2.cu.
//
// Same mathematics as k_point_fwd / k_point_bwd (render.cu): OneBlob + hash-grid encode, the coarse MLP
// 80 -> 32 -> 33 (models/decoder.py:80-94), the class-expert MLP of the tile (slams/mapping.py:590-601),
// the latent / free-space / opacity loss terms (mapping.py:123-126, utils/common.py:769-802) and their
// backward down to the hash-table scatter.  One CTA = one 128-slot tile = one 128-row MMA tile, one thread
// per slot = one TMEM lane.  GEMMs (bf16 hi + lo halves, 3 products, fp32 accumulation in TMEM):
//
//   fwd   H[128 x 64]    = X[128 x 80]  . [W1 coarse ; W1 expert]^T            K = 80
//         Oc[128 x 48]   = Hc[128 x 32] . W2 coarse^T,  Of likewise            K = 32
//   bwd   dHc[128 x 32]  = dOc[128 x 48] . W2 coarse,   dHf likewise           K = 48
//         dX[128 x 80]   = [dHc | dHf][128 x 64] . [W1 coarse ; W1 expert]     K = 64  (sums both nets)
//
// Operand tiles use the no-swizzle canonical UMMA layout  [chunk of 8 features][row][16 B]; the same
// shared-memory weight copy serves as K-major B operand in the forward and MN-major B operand in the
// backward GEMMs (see ray_tc.cu).
#include "point_tc.cuh"

namespace dns {


// params [n][4096] (W1[32][80] | W2[48][32]) -> bf16 hi/lo chunk tiles
__global__ void k_prep_net80_tc(const float* __restrict__ params, uint4* __restrict__ out) {
  const float* p = params + (int64_t)blockIdx.x * 4096;
  uint4* o = out + (int64_t)blockIdx.x * kNetTc;
  for (int i = threadIdx.x; i < 320 + 192; i += blockDim.x) {
    float4 a, b;
    if (i < 320) {
      int c = i >> 5, j = i & 31;
      const float* src = p + j * kIn1 + 8 * c;
      a = *reinterpret_cast<const float4*>(src);
      b = *reinterpret_cast<const float4*>(src + 4);
    } else {
      int k = i - 320, c = k / 48, r = k - c * 48;
      if (r < DNS_LATENT) {
        const float* src = p + 2560 + r * 32 + 8 * c;
        a = *reinterpret_cast<const float4*>(src);
        b = *reinterpret_cast<const float4*>(src + 4);
      } else {
        a = b = make_float4(0.f, 0.f, 0.f, 0.f);  // padded output rows of the tcnn layout stay out of the maths
      }
    }
    uint4 h, l;
    split8(a, b, h, l);
    if (i < 320) {
      o[i] = h;
      o[320 + i] = l;
    } else {
      o[640 + (i - 320)] = h;
      o[832 + (i - 320)] = l;
    }
  }
}

int prep_nets_tc(const float* coarse, const float* experts, int n_experts, uint4* wc, uint4* we, cudaStream_t st) {
  k_prep_net80_tc<<<1, 128, 0, st>>>(coarse, wc);
  if (experts && n_experts > 0) k_prep_net80_tc<<<n_experts, 128, 0, st>>>(experts, we);
  return check_launch("prep_nets_tc");
}

}  // namespace dns
