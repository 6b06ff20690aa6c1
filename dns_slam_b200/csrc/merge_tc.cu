// Merge branch of the pixel features on tcgen05: models/decoder.py:67-77 in two kernels.
//
//   forward   per (view r, point p) row:  x = (refer_p - lo) / (hi - lo)  ->  OneBlob(48) || feature(64)
//             -> H = relu(X . W1^T) (K = 112, N = 32) -> O = H . W2^T (K = 32, N = 32) -> out[p] += O / R
//   backward  dO = d_out[p] / R -> dH = (dO . W2) * [H > 0] -> dX[:, 0..47] = dH . W1[:, 0..47]
//             -> OneBlob backward -> d_refer_p;  the gathered features carry no gradient (the reference rounds the
//             pixel coordinates, utils/common.py:657);  dW1 = X^T dH, dW2 = dO^T H through k_dw_img (tc.cu) from
//             the bf16 hi/lo tile images both kernels leave behind.
//
// One CTA = one 128-row tile, one thread per row = one TMEM lane; operands bf16 hi + lo halves in the canonical
// no-swizzle UMMA layout (see point_tc.cu / ray_tc.cu).  Replaces the operator chain OneBlob -> concat -> MLP ->
// mean of the drop-in modules (5 launches + 3 weight-gradient GEMMs + torch glue per target frame).
#include <string.h>

#include "tc_common.cuh"

namespace dns {

constexpr int kMW1 = 14 * 32;   // uint4 per half of the W1 tile [14 feature chunks][32 hidden rows]
constexpr int kMW2 = 4 * 32;    // uint4 per half of the W2 tile [4 hidden chunks][32 output rows]
constexpr int kMergeW = 2 * kMW1 + 2 * kMW2;   // W1 hi | W1 lo | W2 hi | W2 lo

// params = W1[32][112] | W2[32][32] (tinycudann layout) -> bf16 hi/lo chunk tiles
__global__ void k_prep_merge_tc(const float* __restrict__ params, uint4* __restrict__ out) {
  for (int i = threadIdx.x; i < kMW1 + kMW2; i += blockDim.x) {
    const float* src;
    int dst_hi, dst_lo;
    if (i < kMW1) {
      int c = i >> 5, j = i & 31;
      src = params + j * kIn2 + 8 * c;
      dst_hi = i;
      dst_lo = kMW1 + i;
    } else {
      int k = i - kMW1, c = k >> 5, j = k & 31;
      src = params + 32 * kIn2 + j * 32 + 8 * c;
      dst_hi = 2 * kMW1 + k;
      dst_lo = 2 * kMW1 + kMW2 + k;
    }
    uint4 h, l;
    split8(*reinterpret_cast<const float4*>(src), *reinterpret_cast<const float4*>(src + 4), h, l);
    out[dst_hi] = h;
    out[dst_lo] = l;
  }
}

struct MergeArgs {
  const float* refer_p;   // [R*P][3]
  const float* code;      // [R*P][64]
  const uint4* w;         // prepared weights (kMergeW)
  int64_t n_rows, P;
  int R;
  Bound B;
  float* out;             // [P][32] (+= O / R)
  const float* d_out;     // [P][32]
  float* d_refer_p;       // [R*P][3]
  uint4* Ximg;            // [tile][2][14][128]
  uint4* Himg;            // [tile][2][4][128]
  uint4* dHimg;           // [tile][2][4][128]
  uint4* dOimg;           // [tile][2][4][128]
  int keep;               // forward: write the X / H images (a backward will follow)
};

__device__ __forceinline__ void load_merge_weights(unsigned char* W, const uint4* __restrict__ w) {
  for (int i = threadIdx.x; i < kMergeW; i += blockDim.x) reinterpret_cast<uint4*>(W)[i] = w[i];
}

__global__ void __launch_bounds__(kTile) k_merge_fwd_tc(MergeArgs a) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  unsigned char* X_hi = sm;
  unsigned char* X_lo = sm + 14 * 2048;
  unsigned char* H_hi = sm;                  // aliases the X tile after the first GEMM
  unsigned char* H_lo = sm + 4 * 2048;
  unsigned char* W = sm + 28 * 2048;
  unsigned char* W1_hi = W;
  unsigned char* W1_lo = W + kMW1 * 16;
  unsigned char* W2_hi = W + 2 * kMW1 * 16;
  unsigned char* W2_lo = W2_hi + kMW2 * 16;
  const int tile = blockIdx.x, tid = threadIdx.x, warp = tid >> 5;
  load_merge_weights(W, a.w);
  if (warp == 0) tmem_alloc(&tmem_base_s, 64);
  if (tid == 0) mbar_init(&bar, 1);
  const int64_t row = (int64_t)tile * kTile + tid;
  const bool valid = row < a.n_rows;
  uint4* ximg = a.keep ? a.Ximg + (int64_t)tile * (28 * kTile) + tid : nullptr;
#define XIMG(c) (ximg ? ximg + (c) * kTile : nullptr), (ximg ? ximg + (14 + (c)) * kTile : nullptr)
  if (valid) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x = (float)(((double)a.refer_p[3 * row + c] - a.B.lo[c]) / a.B.ext[c]);
      float pe[16];
      oneblob16(x, pe);
      put_chunk_img(X_hi, X_lo, 2 * c, 2048, tid, pe, XIMG(2 * c));
      put_chunk_img(X_hi, X_lo, 2 * c + 1, 2048, tid, pe + 8, XIMG(2 * c + 1));
    }
    const float4* s4 = reinterpret_cast<const float4*>(a.code + row * 64);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 u = s4[2 * c], v = s4[2 * c + 1];
      const float f[8] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
      put_chunk_img(X_hi, X_lo, 6 + c, 2048, tid, f, XIMG(6 + c));
    }
  } else {
    const uint4 z4 = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int c = 0; c < 14; ++c) {
      *reinterpret_cast<uint4*>(X_hi + c * 2048 + tid * 16) = z4;
      *reinterpret_cast<uint4*>(X_lo + c * 2048 + tid * 16) = z4;
      if (ximg) ximg[c * kTile] = ximg[(14 + c) * kTile] = z4;
    }
  }
#undef XIMG
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  if (tid == 0) {   // H = X . W1^T
    const uint32_t idesc = umma_idesc_bf16(128, 32, 0, 0);
#pragma unroll 1
    for (int ks = 0; ks < 7; ++ks) {
      const uint32_t aoff = ks * 4096, boff = ks * 1024;
      const uint64_t a_hi = umma_desc(smem_u32(X_hi) + aoff, 2048, 128), a_lo = umma_desc(smem_u32(X_lo) + aoff, 2048, 128);
      const uint64_t b_hi = umma_desc(smem_u32(W1_hi) + boff, 512, 128), b_lo = umma_desc(smem_u32(W1_lo) + boff, 512, 128);
      umma_bf16(tmem_d, a_hi, b_hi, idesc, ks > 0 ? 1u : 0u);
      umma_bf16(tmem_d, a_lo, b_hi, idesc, 1u);
      umma_bf16(tmem_d, a_hi, b_lo, idesc, 1u);
    }
    umma_commit(&bar);
  }
  mbar_wait_cta(&bar, 0);
  tc_fence_after();
  const uint32_t lane_addr = tmem_d + ((uint32_t)(warp * 32) << 16);
  {
    uint4* himg = a.keep ? a.Himg + (int64_t)tile * (8 * kTile) + tid : nullptr;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      float v[16];
      tmem_ld16(lane_addr + 16 * g, v);
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = fmaxf(v[k], 0.f);
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const int c = 2 * g + c2;
        put_chunk_img(H_hi, H_lo, c, 2048, tid, v + 8 * c2, himg ? himg + c * kTile : nullptr,
                      himg ? himg + (4 + c) * kTile : nullptr);
      }
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {   // O = H . W2^T
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, 32, 0, 0);
#pragma unroll 1
    for (int ks = 0; ks < 2; ++ks) {
      const uint32_t aoff = ks * 4096, boff = ks * 1024;
      const uint64_t a_hi = umma_desc(smem_u32(H_hi) + aoff, 2048, 128), a_lo = umma_desc(smem_u32(H_lo) + aoff, 2048, 128);
      const uint64_t b_hi = umma_desc(smem_u32(W2_hi) + boff, 512, 128), b_lo = umma_desc(smem_u32(W2_lo) + boff, 512, 128);
      umma_bf16(tmem_d + 32, a_hi, b_hi, idesc, ks > 0 ? 1u : 0u);
      umma_bf16(tmem_d + 32, a_lo, b_hi, idesc, 1u);
      umma_bf16(tmem_d + 32, a_hi, b_lo, idesc, 1u);
    }
    umma_commit(&bar);
  }
  mbar_wait_cta(&bar, 1);
  tc_fence_after();
  {
    const float inv_r = 1.f / (float)a.R;
    float* dst = valid ? a.out + (row % a.P) * 32 : nullptr;   // rows are view-major: row = r * P + p
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      float v[16];
      tmem_ld16(lane_addr + 32 + 16 * g, v);
      if (valid) {
#pragma unroll
        for (int k = 0; k < 16; ++k) atomicAdd(dst + 16 * g + k, v[k] * inv_r);   // mean over the R views
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 64);
}

__global__ void __launch_bounds__(kTile) k_merge_bwd_tc(MergeArgs a) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  unsigned char* DO_hi = sm;                 // dOut tile [4 chunks]
  unsigned char* DO_lo = sm + 4 * 2048;
  unsigned char* DH_hi = sm + 8 * 2048;      // dH tile [4 chunks]
  unsigned char* DH_lo = sm + 12 * 2048;
  unsigned char* W = sm + 16 * 2048;
  unsigned char* W1_hi = W;
  unsigned char* W1_lo = W + kMW1 * 16;
  unsigned char* W2_hi = W + 2 * kMW1 * 16;
  unsigned char* W2_lo = W2_hi + kMW2 * 16;
  const int tile = blockIdx.x, tid = threadIdx.x, warp = tid >> 5;
  load_merge_weights(W, a.w);
  if (warp == 0) tmem_alloc(&tmem_base_s, 128);
  if (tid == 0) mbar_init(&bar, 1);
  const int64_t row = (int64_t)tile * kTile + tid;
  const bool valid = row < a.n_rows;
  {
    uint4* doimg = a.dOimg + (int64_t)tile * (8 * kTile) + tid;
    const float inv_r = 1.f / (float)a.R;
    const float4* s4 = valid ? reinterpret_cast<const float4*>(a.d_out + (row % a.P) * 32) : nullptr;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (valid) {
        const float4 u = s4[2 * c], v = s4[2 * c + 1];
        f[0] = u.x * inv_r; f[1] = u.y * inv_r; f[2] = u.z * inv_r; f[3] = u.w * inv_r;
        f[4] = v.x * inv_r; f[5] = v.y * inv_r; f[6] = v.z * inv_r; f[7] = v.w * inv_r;
      }
      put_chunk_img(DO_hi, DO_lo, c, 2048, tid, f, doimg + c * kTile, doimg + (4 + c) * kTile);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  if (tid == 0) {   // dH = dO . W2   (B = W2 tile MN-major: hidden contiguous; LBO 128 over output rows, SBO 512 over hidden chunks)
    const uint32_t idesc = umma_idesc_bf16(128, 32, 0, 1);
#pragma unroll 1
    for (int ks = 0; ks < 2; ++ks) {
      const uint32_t aoff = ks * 4096, boff = ks * 256;
      const uint64_t a_hi = umma_desc(smem_u32(DO_hi) + aoff, 2048, 128), a_lo = umma_desc(smem_u32(DO_lo) + aoff, 2048, 128);
      const uint64_t b_hi = umma_desc(smem_u32(W2_hi) + boff, 128, 512), b_lo = umma_desc(smem_u32(W2_lo) + boff, 128, 512);
      umma_bf16(tmem_d, a_hi, b_hi, idesc, ks > 0 ? 1u : 0u);
      umma_bf16(tmem_d, a_lo, b_hi, idesc, 1u);
      umma_bf16(tmem_d, a_hi, b_lo, idesc, 1u);
    }
    umma_commit(&bar);
  }
  mbar_wait_cta(&bar, 0);
  tc_fence_after();
  const uint32_t lane_addr = tmem_d + ((uint32_t)(warp * 32) << 16);
  {
    // ReLU mask: the bf16 hi half of the stashed activation is non-zero exactly where the activation was positive
    const uint4* himg = a.Himg + (int64_t)tile * (8 * kTile) + tid;
    uint4* dhimg = a.dHimg + (int64_t)tile * (8 * kTile) + tid;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      float v[16];
      tmem_ld16(lane_addr + 16 * g, v);
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const int c = 2 * g + c2;
        const uint4 hv = himg[c * kTile];
        const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
        float dh[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) dh[e] = ((hw[e >> 1] >> (16 * (e & 1))) & 0x7fffu) ? v[8 * c2 + e] : 0.f;
        put_chunk_img(DH_hi, DH_lo, c, 2048, tid, dh, dhimg + c * kTile, dhimg + (4 + c) * kTile);
      }
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {   // dX[:, 0..47] = dH . W1[:, 0..47]   (B = W1 tile MN-major: LBO 128 over hidden rows, SBO 512 over feature chunks)
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, 48, 0, 1);
#pragma unroll 1
    for (int ks = 0; ks < 2; ++ks) {
      const uint32_t aoff = ks * 4096, boff = ks * 256;
      const uint64_t a_hi = umma_desc(smem_u32(DH_hi) + aoff, 2048, 128), a_lo = umma_desc(smem_u32(DH_lo) + aoff, 2048, 128);
      const uint64_t b_hi = umma_desc(smem_u32(W1_hi) + boff, 128, 512), b_lo = umma_desc(smem_u32(W1_lo) + boff, 128, 512);
      umma_bf16(tmem_d + 32, a_hi, b_hi, idesc, ks > 0 ? 1u : 0u);
      umma_bf16(tmem_d + 32, a_lo, b_hi, idesc, 1u);
      umma_bf16(tmem_d + 32, a_hi, b_lo, idesc, 1u);
    }
    umma_commit(&bar);
  }
  mbar_wait_cta(&bar, 1);
  tc_fence_after();
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v[16];
    tmem_ld16(lane_addr + 32 + 16 * c, v);
    if (valid) {
      const float x = (float)(((double)a.refer_p[3 * row + c] - a.B.lo[c]) / a.B.ext[c]);
      a.d_refer_p[3 * row + c] = oneblob16_bwd(x, v) / (float)a.B.ext[c];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 128);
}

}  // namespace dns

using namespace dns;

extern "C" {

// images (X 448 B/row, H / dH / dOut 128 B/row each, rows padded to whole tiles) + prepared weights
int64_t dns_merge_workspace_bytes(int64_t n_rows) {
  const int64_t tiles = (n_rows + kTile - 1) / kTile;
  return tiles * kTile * (448 + 3 * 128) + kMergeW * 16 + 4 * 256;
}

static void merge_carve(MergeArgs& m, void* workspace, int64_t n_rows) {
  const int64_t tiles = (n_rows + kTile - 1) / kTile;
  char* p = (char*)workspace;
  auto take = [&](int64_t bytes) {
    char* q = p;
    p += (bytes + 255) & ~(int64_t)255;
    return q;
  };
  m.w = (const uint4*)take(kMergeW * 16);
  m.Ximg = (uint4*)take(tiles * kTile * 448);
  m.Himg = (uint4*)take(tiles * kTile * 128);
  m.dHimg = (uint4*)take(tiles * kTile * 128);
  m.dOimg = (uint4*)take(tiles * kTile * 128);
}

static void merge_attrs() {
  static unsigned long long seen = 0;
  if (!first_call_on_device(seen)) return;
  cudaFuncSetAttribute(k_merge_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 28 * 2048 + kMergeW * 16);
  cudaFuncSetAttribute(k_merge_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 2048 + kMergeW * 16);
}

int dns_merge_fwd(const float* refer_p, const float* code, const float* params, int64_t P, int R, const double bound[3][2],
                  float* out, int keep_for_backward, void* workspace, int64_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_rows = P * R;
  if (n_rows <= 0) return DNS_OK;
  if (!workspace || workspace_bytes < dns_merge_workspace_bytes(n_rows)) {
    set_error("merge: workspace too small");
    return DNS_ERR_ARG;
  }
  if (((uintptr_t)code & 15) || ((uintptr_t)params & 15)) {
    set_error("merge: code / params must be 16-byte aligned");
    return DNS_ERR_ARG;
  }
  MergeArgs m;
  memset(&m, 0, sizeof(m));
  merge_carve(m, workspace, n_rows);
  m.refer_p = refer_p; m.code = code; m.n_rows = n_rows; m.P = P; m.R = R; m.out = out; m.keep = keep_for_backward;
  for (int c = 0; c < 3; ++c) {
    m.B.lo[c] = bound[c][0];
    m.B.ext[c] = bound[c][1] - bound[c][0];
  }
  merge_attrs();
  PhaseScope ph(phFeature, st, 3);
  cudaMemsetAsync(out, 0, sizeof(float) * P * 32, st);
  k_prep_merge_tc<<<1, 128, 0, st>>>(params, (uint4*)m.w);
  const int tiles = (int)((n_rows + kTile - 1) / kTile);
  k_merge_fwd_tc<<<tiles, kTile, 28 * 2048 + kMergeW * 16, st>>>(m);
  return check_launch("merge_fwd");
}

int dns_merge_bwd(const float* refer_p, const float* d_out, int64_t P, int R, const double bound[3][2], float* d_refer_p,
                  float* d_params, void* workspace, int64_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_rows = P * R;
  if (n_rows <= 0) return DNS_OK;
  if (!workspace || workspace_bytes < dns_merge_workspace_bytes(n_rows)) {
    set_error("merge: workspace too small");
    return DNS_ERR_ARG;
  }
  MergeArgs m;
  memset(&m, 0, sizeof(m));
  merge_carve(m, workspace, n_rows);   // weights, X and H images are those of the forward call
  m.refer_p = refer_p; m.d_out = d_out; m.n_rows = n_rows; m.P = P; m.R = R; m.d_refer_p = d_refer_p;
  for (int c = 0; c < 3; ++c) {
    m.B.lo[c] = bound[c][0];
    m.B.ext[c] = bound[c][1] - bound[c][0];
  }
  merge_attrs();
  PhaseScope ph(phFeature, st, 3);
  const int tiles = (int)((n_rows + kTile - 1) / kTile);
  k_merge_bwd_tc<<<tiles, kTile, 16 * 2048 + kMergeW * 16, st>>>(m);
  if (int e = check_launch("merge_bwd")) return e;
  if (d_params) {   // accumulated: dW1[32][112] = dH^T X, dW2[32][32] = dO^T H
    DwImgArgs g;
    memset(&g, 0, sizeof(g));
    g.L = DwImg{m.Ximg, 14, 0, 14, kIn2}; g.Cc = DwImg{m.dHimg, 4, 0, 4, 32};
    g.RS = kTile; g.subs_per_tile = 1; g.n_tiles_host = tiles;
    g.out0 = d_params; g.split = 32; g.sl0 = 1; g.sc0 = kIn2;
    int e = launch_dw_img(g, st);
    g.L = DwImg{m.dOimg, 4, 0, 4, 32}; g.Cc = DwImg{m.Himg, 4, 0, 4, 32};
    g.out0 = d_params + 32 * kIn2; g.sl0 = 32; g.sc0 = 1;
    e |= launch_dw_img(g, st);
    if (e) return DNS_ERR_CUDA;
  }
  return DNS_OK;
}

}  // extern "C"
