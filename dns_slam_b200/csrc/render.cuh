// Shared argument structs / enums of the fused render kernels (render.cu, point_tc.cu, ray_tc.cu).
#pragma once
#include "common.cuh"

namespace dns {

enum { kTrack = 0, kMap = 1, kTv = 2 };
// counts[] slots
enum { cMask = 0, cDpos = 1, cFront = 2, cBand = 3, cTiles = 4, cErr = 5 };
// raw loss sums
enum { rP = 0, rD = 1, rL = 2, rLt = 3, rFs = 4, rOp = 5 };

struct RayArgs {
  unsigned long long* phase_clk;   // -DDNS_ABLATE builds only (DNS_PHASE_CLK_RAY): [24] summed clock64 deltas of thread 0
  int feat_band;                   // dns_render_args::features_band_only
  int w2l_smem;                    // layer 2 of the logit head copied into shared memory (set by launch_ray_tc2 when it fits)
  int mode;
  int S, T, RPC, C, C4;
  int64_t N_total, ray0, Nc;  // chunk of rays [ray0, ray0 + Nc)
  Bound B;
  const float* rays_o;
  const float* rays_d;
  const float* z;
  const float* gt_color;
  const float* gt_depth;
  const int64_t* gt_label;
  const uint8_t* mask;
  const float* features;
  const float* fine36;
  const float* W1T2;
  const float* W2cT;
  const float* logit;  // tcnn layout; W2l = logit + 32*112, [Cpad][32]
  const uint4* w16_hi; // fp16 hi / lo halves of the [64 x 112] colour|logit layer-1 weights (forward GEMM of the tcgen05 path)
  const uint4* w16_lo;
  const int* counts;
  int* err;       // counts[cErr]: 3 = label outside [0, n_class)
  float lam_p, lam_d, lam_l;
  float* pred_color;
  float* pred_depth;
  float* pred_var;
  float* pred_logits;
  float* raw;
  float* dfine36;
  // tcgen05 MAP path: d(latent) rows leave the kernel in SLOT order (row inv[point - ray0*S] of dfine36s), so that the
  // point backward kernel streams them instead of chasing the class permutation
  const int* inv;
  float* dfine36s;
  float* d_features;
  float* d_rays_o;
  float* d_rays_d;
  // stashes (point / ray order of the chunk)
  float* X2;     // [Pc][112]
  float* dH2;    // [Pc][64]
  float* Hcol;   // [Pc][32]
  float* dpre;   // [Pc][4]
  float* dlogit; // [Nc][C4]
  float* Hbar;   // [Nc][32]
  // tcgen05 path: per-CTA bf16 hi/lo tile images [cta][sub][half][chunk][RS rows] over the fp32 regions above
  uint4* X2img;    // 14 chunks
  uint4* dH2img;   // 8 chunks: colour | logit hidden gradients
  uint4* Hcolimg;  // 4 chunks
  uint4* dpreimg;  // 2 chunks (3 used columns)
  int RS;          // rows per sub-tile (T or T/2)
  int fwd_only;    // inference: stop after the predictions
  int need_dparams, need_drays, need_dfeat;
};

struct PointArgs {
  // geometry
  const float* rays_o;
  const float* rays_d;
  const float* z;
  const float* gt_depth;
  int S;
  int64_t N_total, P_total;  // whole batch (loss denominators, class quirk)
  int64_t p0, Pc;            // this chunk: global offset and number of points
  Bound B;
  dns_grid G;
  const float2* table;
  // TV lattice
  int n;
  double voxel, jit[3], off[3];
  const double* tv_oj;   // device override: off[3] | jit[3]
  int tv_agg_levels;      // TV backward: levels [0, tv_agg_levels) are pre-reduced per cell inside the warp (hashgrid_bwd_rows)
  // slots
  const int* perm;        // slot -> chunk-local point, -1 = padding; NULL = identity
  const int* tile_class;  // tile -> expert row (MAP)
  const int* counts;      // device counts (n_tiles at cTiles when perm != NULL)
  int n_tiles_host;
  // weights (k-major blocks)
  const float* WTc;
  const float* WTe;
  // latents in point order, padded rows of 36
  float* fine36;
  float* coarse36;
  float* dfine36;
  // tcgen05 MAP path, SLOT-order latent IMAGES [tile][9 chunks of 4 channels][128 slots] (slot_img): diff36s = coarse - fine
  // (channels 0..32) | fine[32] (channel 33), written by the forward kernel; dfine36s written by the ray kernel.  A warp of
  // the backward reads 512 contiguous bytes per chunk (rows of 36 floats put every lane into its own cache line).
  float* diff36s;
  float* dfine36s;
  int want_coarse_pt;   // also store the coarse latents in point order (coarse_out requested)
  float* occ;   // TV: [n^3]
  float* docc;  // TV
  // stashes in slot order
  float* Jst;   // tcgen05 path, only when ray gradients are wanted: Jacobian image [tile][24 float4 chunks][128 slots] of
                // d(grid features) / d(x) written by the forward kernel (chunks 0..11: levels 0..7, 12..23: levels 8..15)
  float* Xst;   // [Q][80]
  float* Hc;    // [Q][32]
  float* Hf;
  float* dHc;
  float* dHf;
  float* dOc;   // [Q][36]
  float* dOf;
  // tcgen05 path: the same stashes as bf16 hi/lo tile images  [tile][half][chunk][128 slots][8 x bf16]
  // (they overlay the fp32 regions above; chunks per half: X 10, H / dH 8 (4 without experts), dOut 10 (5))
  uint4* Ximg;
  uint4* Himg;
  uint4* dHimg;
  uint4* dOimg;
  // losses / gradients
  float lam_lt, lam_fs, lam_op, trunc, sigma;
  float* raw;
  float2* d_table;
  // Privatised gradient copies of the SMALL leading levels (tcgen05 backward): levels 0 .. priv_levels-1 (entries
  // [0, priv_end)) are reduced into copy (blockIdx.x % priv_copies) of d_priv [priv_copies][priv_end] instead of d_table;
  // k_priv_reduce folds the copies back.  Every level receives the same number of reductions, so the 32 KB of level 0
  // would otherwise take 1/16 of them on 256 cache lines -- ncu showed the busiest L2 slice at 94 % of its tag rate while
  // the average slice sat at 59 % (profiles/r02m_ncu_summary.md).
  float2* d_priv;
  uint32_t priv_end;
  int priv_levels, priv_copies;
  float* d_rays_o;
  float* d_rays_d;
  int need_dparams, need_drays;
  int dbg;   // -DDNS_ABLATE builds only (DNS_DBG env): 2 no stash stores, 4 no table atomics, 8 no regather
  unsigned long long* phase_clk;   // -DDNS_ABLATE builds only: [16] summed clock64 deltas of thread 0 per kernel phase
};

// float4 index of channel chunk `chunk` (4 channels) of slot q in a slot-order latent image
__device__ __forceinline__ int64_t slot_img(int64_t q, int chunk) { return ((q >> 7) * 9 + chunk) * 128 + (q & 127); }

template <int MODE>
__device__ __forceinline__ bool slot_point(const PointArgs& a, int64_t q, int64_t& i, int64_t& r, float& zv, float x[3]) {
  if (MODE == kTv) {
    int64_t n = a.n, n3 = n * n * n;
    if (q >= n3) {
      x[0] = x[1] = x[2] = 0.f;   // the lane still takes part in the warp shuffles of hashgrid_bwd_rows
      return false;
    }
    i = q;
    // x runs fastest over the slots: the lanes of a warp are consecutive lattice points of one x row, i.e. neighbours in the
    // dense levels' memory order (index = x + y res + z res^2) and, at hashed levels, within one aligned block (the x prime
    // is 1), so the forward gathers of a warp coalesce and the backward can pre-reduce per cell (hashgrid_bwd_rows).  occ /
    // docc are therefore stored [z][y][x]; the stencil (k_tv_stencil) is symmetric in the three axes.
    const uint32_t qq = (uint32_t)q, nn = (uint32_t)a.n, qn = qq / nn;   // n <= 512: 32-bit divisions
    int64_t idx[3] = {qq - qn * nn, qn % nn, qn / nn};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double jit = a.tv_oj ? a.tv_oj[3 + c] : a.jit[c], off = a.tv_oj ? a.tv_oj[c] : a.off[c];
      double pt = (((double)idx[c] + jit) * a.voxel + a.B.lo[c]) + off;
      x[c] = (float)((pt - a.B.lo[c]) / a.B.ext[c]);
    }
    r = 0;
    zv = 0.f;
    return true;
  } else {
    i = a.perm ? (int64_t)a.perm[q] : q;
    if (i < 0 || i >= a.Pc) return false;
    int64_t p = a.p0 + i;
    r = p < 0x7fffffffLL ? (int64_t)((uint32_t)p / (uint32_t)a.S) : p / a.S;   // 32-bit division whenever the index allows
    zv = a.z[p];
    point_from_ray(a.rays_o + 3 * r, a.rays_d + 3 * r, zv, a.B, x);
    return true;
  }
}

__device__ __forceinline__ void load_block(float* dst, const float* __restrict__ src, int n4) {
  const float4* s = reinterpret_cast<const float4*>(src);
  float4* d = reinterpret_cast<float4*>(dst);
  for (int i = threadIdx.x; i < n4; i += blockDim.x) d[i] = s[i];
}

// 80 -> 32 (ReLU) -> 36: h and out in registers
__device__ __forceinline__ void net80_fwd(const float* xrow, const float* W1T, const float* W2T, float (&h)[32],
                                          float (&out)[kOutP]) {
  zero(h);
  accum_layer<32>(xrow, 1, kIn1, W1T, 32, h);
#pragma unroll
  for (int j = 0; j < 32; ++j) h[j] = fmaxf(h[j], 0.f);
  zero(out);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float4* w = reinterpret_cast<const float4*>(W2T + j * kOutP);
#pragma unroll
    for (int q = 0; q < kOutP / 4; ++q) {
      float4 v = w[q];
      out[4 * q + 0] = fmaf(h[j], v.x, out[4 * q + 0]);
      out[4 * q + 1] = fmaf(h[j], v.y, out[4 * q + 1]);
      out[4 * q + 2] = fmaf(h[j], v.z, out[4 * q + 2]);
      out[4 * q + 3] = fmaf(h[j], v.w, out[4 * q + 3]);
    }
  }
}
template <int N>
__device__ __forceinline__ void store_row(float* dst, const float (&v)[N]) {
  float4* d = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int q = 0; q < N / 4; ++q) d[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
template <int N>
__device__ __forceinline__ void load_row(const float* src, float (&v)[N]) {
  const float4* s = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int q = 0; q < N / 4; ++q) {
    float4 t = s[q];
    v[4 * q] = t.x;
    v[4 * q + 1] = t.y;
    v[4 * q + 2] = t.z;
    v[4 * q + 3] = t.w;
  }
}

// fs / opacity masks of one sample (utils/common.py:786-792)
__device__ __forceinline__ void opacity_masks(float zv, float d, float trunc, float& front, float& band, float& valid) {
  bool f = zv < __fsub_rn(d, trunc), b = zv > __fadd_rn(d, trunc);
  valid = d > 0.f ? 1.f : 0.f;
  front = f ? 1.f : 0.f;
  band = (!f && !b) ? valid : 0.f;
}


// tcgen05 point kernels (point_tc.cu); wc / we: prepared weight tiles, kNetTc uint4 per net:
//   W1 hi [10][32] | W1 lo | W2 hi [4][48] | W2 lo  as bf16 halves (backward GEMMs), then  W1 hi | W1 lo  as fp16 halves
//   (the forward layer-1 GEMM, whose result decides the ReLU: tc_common.cuh put_chunk_f16_img)
constexpr int kNetTc = 1664;
constexpr int kNetW1F16 = 1024;
int prep_nets_tc(const float* coarse, const float* experts, int n_experts, uint4* wc, uint4* we, cudaStream_t st);
int launch_point_fwd_tc(int mode, const PointArgs& pa, int tiles, const uint4* wc, const uint4* we, cudaStream_t st);
int launch_point_bwd_tc(int mode, const PointArgs& pa, int tiles, const uint4* wc, const uint4* we, cudaStream_t st);

// tcgen05 ray kernel, two threads per point (ray_tc.cu)
void pick_ray_block_tc2(int S, int& T, int& RPC);
size_t ray_tc2_smem_bytes(int T, int RPC, int C4);
int launch_ray_tc2(const RayArgs& ra, uint4* w1_hi, uint4* w1_lo, int64_t n_rays_chunk, cudaStream_t st);
void pick_ray_block_any(int S, int& T, int& RPC);
int launch_ray_tc(const RayArgs& ra, const float* color, const float* logit, uint4* w1_hi, uint4* w1_lo, bool prep,
                  int64_t n_rays_chunk, cudaStream_t st);

}  // namespace dns
