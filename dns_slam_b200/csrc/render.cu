// Fused render + loss + backward of the DNS-SLAM inner loop (sm_100a).
//
// Restates, as three point/ray kernels plus weight-gradient GEMMs:
//   Tracker.renderer + losses   slams/tracking.py:188-214, 85-96, 326-338
//   Mapper.renderer + fine_fn   slams/mapping.py:590-635   (class(p) = label[p mod N] quirk)
//   mapping losses              slams/mapping.py:110-126, 891-907; utils/common.py:769-802
//   occupancy compositing       utils/common.py:506-537
//   Mapper.smoothness           slams/mapping.py:129-159 (TV of the coarse occupancy)
//
// Pipeline for one chunk of rays (slots = points regrouped so that every 128-slot tile has
// ONE semantic class -> one expert MLP per tile, weights broadcast from shared memory):
//   k_point_fwd   encode (OneBlob + hash grid) -> coarse MLP [-> expert MLP] -> latents, lt/fs/op
//   k_ray         out_fn (colour + logit layer 1 per point, logit layer 2 per RAY: it is linear,
//                 so it commutes with compositing) -> compositing -> p/d/l losses -> backward
//                 down to d(latents), d(features), d(rays)
//   k_point_bwd   lt/fs/op gradients + d(latents) -> MLP backward -> hash-table scatter, d(rays)
//   k_dw_gemm     dW = X^T dH for every MLP from stashed activations (ops.cu)
#include <stdlib.h>
#include <string.h>

#include "render.cuh"

namespace dns {

int launch_dw_gemm(const float* A, int lda, int M, const float* B, int ldb, int N, int64_t n_rows,
                   const int* n_tiles_dev, int n_tiles_host, const int* tile_class, float* C, int ldc,
                   int64_t c_stride, cudaStream_t st, bool tensor_cores);


// ---------------------------------------------------------------------------------------
// weight re-layout (once per call): tcnn row-major -> k-major blocks read with broadcast LDS.128
// ---------------------------------------------------------------------------------------
// params [n][4096] = W1[32][80] | W2[48][32]  ->  WT [n][kNetT] = W1T[80][32] | W2T[32][36]
__global__ void k_transpose_net80(const float* __restrict__ params, float* __restrict__ WT) {
  const float* p = params + (int64_t)blockIdx.x * 4096;
  float* w = WT + (int64_t)blockIdx.x * kNetT;
  for (int i = threadIdx.x; i < 2560; i += blockDim.x) {
    int k = i >> 5, j = i & 31;
    w[i] = p[j * 80 + k];
  }
  for (int i = threadIdx.x; i < 32 * kOutP; i += blockDim.x) {
    int j = i / kOutP, c = i - j * kOutP;
    w[2560 + i] = c < DNS_LATENT ? p[2560 + c * 32 + j] : 0.f;
  }
}
// colour / logit nets: W1T2 [112][64] (cols 0..31 colour hidden, 32..63 logit hidden), W2cT [32][4]
__global__ void k_transpose_out(const float* __restrict__ color, const float* __restrict__ logit,
                                float* __restrict__ W1T2, float* __restrict__ W2cT) {
  for (int i = threadIdx.x; i < kIn2 * 64; i += blockDim.x) {
    int k = i >> 6, j = i & 63;
    W1T2[i] = j < 32 ? color[j * kIn2 + k] : logit[(j - 32) * kIn2 + k];
  }
  for (int i = threadIdx.x; i < 128; i += blockDim.x) {
    int j = i >> 2, c = i & 3;
    W2cT[i] = c < 3 ? color[32 * kIn2 + c * 32 + j] : 0.f;
  }
}

// ---------------------------------------------------------------------------------------
// batch-global counts (loss denominators and the count_nonzero guards of common.py:794)
// ---------------------------------------------------------------------------------------
__global__ void k_counts(const float* __restrict__ gt_depth, const float* __restrict__ z, const uint8_t* __restrict__ mask,
                         int64_t N, int S, float trunc, int* counts) {
  int n_mask = 0, n_dpos = 0, n_front = 0, n_band = 0;
  int64_t P = N * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / S;
    float d = gt_depth[r], zv = z[i];
    bool front = zv < __fsub_rn(d, trunc), back = zv > __fadd_rn(d, trunc);
    n_front += front;
    n_band += (!front && !back && d > 0.f);
    if (i - r * S == 0) {
      n_dpos += d > 0.f;
      n_mask += mask ? (mask[r] != 0) : 1;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    n_mask += __shfl_xor_sync(0xffffffffu, n_mask, o);
    n_dpos += __shfl_xor_sync(0xffffffffu, n_dpos, o);
    n_front += __shfl_xor_sync(0xffffffffu, n_front, o);
    n_band += __shfl_xor_sync(0xffffffffu, n_band, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (n_mask) atomicAdd(counts + cMask, n_mask);
    if (n_dpos) atomicAdd(counts + cDpos, n_dpos);
    if (n_front) atomicAdd(counts + cFront, n_front);
    if (n_band) atomicAdd(counts + cBand, n_band);
  }
}

// ---------------------------------------------------------------------------------------
// class-homogeneous slot tiles (MAP): counting sort of the chunk's points by label[p mod N]
// ---------------------------------------------------------------------------------------
// Both passes aggregate per block in shared memory (kSortPer points per thread, classes kept in registers): one
// global atomic per class and block instead of one per warp-level group of equal labels.
constexpr int kSortPer = 8;
__global__ void __launch_bounds__(256) k_class_hist(const int64_t* __restrict__ label, int64_t N, int64_t p0, int64_t Pc,
                                                    int nci, int* hist, int* counts) {
  extern __shared__ int sh[];   // [nci]
  for (int c = threadIdx.x; c < nci; c += blockDim.x) sh[c] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * (blockDim.x * kSortPer);
#pragma unroll
  for (int k = 0; k < kSortPer; ++k) {
    const int64_t i = base + (int64_t)k * blockDim.x + threadIdx.x;
    if (i < Pc) {
      int64_t c = label[(p0 + i) % N];
      if (c < 0 || c >= nci) {
        counts[cErr] = 1;
        c = 0;
      }
      atomicAdd(sh + (int)c, 1);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < nci; c += blockDim.x)
    if (sh[c]) atomicAdd(hist + c, sh[c]);
}
// one block: slot_start[c] (in slots), number of tiles, tile -> expert row
__global__ void k_class_scan(const int* __restrict__ hist, int nci, const int* __restrict__ class_to_expert,
                             int* slot_start, int* cursor, int* tile_class, int* counts) {
  if (threadIdx.x == 0) {
    int tiles = 0;
    for (int c = 0; c < nci; ++c) {
      slot_start[c] = tiles * kTile;
      cursor[c] = 0;
      int nt = (hist[c] + kTile - 1) / kTile;
      if (nt > 0 && class_to_expert[c] < 0) counts[cErr] = 2;  // "Fine decoders does NOT have class" (mapping.py:595)
      tiles += nt;
    }
    slot_start[nci] = tiles * kTile;
    counts[cTiles] = tiles;
  }
  __syncthreads();
  for (int c = 0; c < nci; ++c) {
    int t0 = slot_start[c] / kTile, t1 = slot_start[c + 1] / kTile, e = class_to_expert[c];
    for (int t = t0 + threadIdx.x; t < t1; t += blockDim.x) tile_class[t] = e;
  }
}
__global__ void __launch_bounds__(256) k_class_scatter(const int64_t* __restrict__ label, int64_t N, int64_t p0, int64_t Pc,
                                                       int nci, const int* __restrict__ slot_start, int* cursor, int* perm,
                                                       int* inv) {
  extern __shared__ int sh[];   // [nci] block counts, then running ranks | [nci] block bases
  int* cnt = sh;
  int* bas = sh + nci;
  for (int c = threadIdx.x; c < nci; c += blockDim.x) cnt[c] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * (blockDim.x * kSortPer);
  int cls[kSortPer];
#pragma unroll
  for (int k = 0; k < kSortPer; ++k) {
    const int64_t i = base + (int64_t)k * blockDim.x + threadIdx.x;
    cls[k] = -1;
    if (i < Pc) {
      int64_t c64 = label[(p0 + i) % N];
      cls[k] = (c64 < 0 || c64 >= nci) ? 0 : (int)c64;
      atomicAdd(cnt + cls[k], 1);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < nci; c += blockDim.x) {
    bas[c] = cnt[c] ? slot_start[c] + atomicAdd(cursor + c, cnt[c]) : 0;   // this block's range in the class region
    cnt[c] = 0;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kSortPer; ++k) {
    if (cls[k] >= 0) {
      const int64_t i = base + (int64_t)k * blockDim.x + threadIdx.x;
      const int slot = bas[cls[k]] + atomicAdd(cnt + cls[k], 1);
      perm[slot] = (int)i;
      if (inv) inv[i] = slot;
    }
  }
}

// Whole-batch fast path: when the call covers the complete, unsharded batch every ray index occurs exactly S
// times among {p mod N}, so the class-c points are {r + k N : r in rays_c, 0 <= k < S}.  A counting sort of the
// N RAYS (N atomics instead of N*S) plus a closed-form slot -> point map replaces the per-point sort.
__global__ void k_ray_hist(const int64_t* __restrict__ label, int64_t N, int nci, int* ray_cnt, int* counts) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < N; r += (int64_t)gridDim.x * blockDim.x) {
    int64_t c = label[r];
    if (c < 0 || c >= nci) {
      counts[cErr] = 1;
      c = 0;
    }
    unsigned m = __match_any_sync(__activemask(), (int)c);
    if ((threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(ray_cnt + c, __popc(m));
  }
}
__global__ void k_class_scan_rays(const int* __restrict__ ray_cnt, int nci, int S, const int* __restrict__ class_to_expert,
                                  int* slot_start, int* ray_start, int* cursor, int* tile_class, int* counts) {
  if (threadIdx.x == 0) {
    int tiles = 0, rays = 0;
    for (int c = 0; c < nci; ++c) {
      slot_start[c] = tiles * kTile;
      ray_start[c] = rays;
      cursor[c] = 0;
      int nt = (int)(((int64_t)ray_cnt[c] * S + kTile - 1) / kTile);
      if (nt > 0 && class_to_expert[c] < 0) counts[cErr] = 2;
      tiles += nt;
      rays += ray_cnt[c];
    }
    slot_start[nci] = tiles * kTile;
    ray_start[nci] = rays;
    counts[cTiles] = tiles;
  }
  __syncthreads();
  for (int c = 0; c < nci; ++c) {
    int t0 = slot_start[c] / kTile, t1 = slot_start[c + 1] / kTile, e = class_to_expert[c];
    for (int t = t0 + threadIdx.x; t < t1; t += blockDim.x) tile_class[t] = e;
  }
}
__global__ void k_ray_scatter(const int64_t* __restrict__ label, int64_t N, int nci, const int* __restrict__ ray_start,
                              int* cursor, int* rays_sorted) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < N; r += (int64_t)gridDim.x * blockDim.x) {
    int64_t c64 = label[r];
    int c = (c64 < 0 || c64 >= nci) ? 0 : (int)c64;
    unsigned m = __match_any_sync(__activemask(), c);
    int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(cursor + c, __popc(m));
    base = __shfl_sync(m, base, leader);
    rays_sorted[ray_start[c] + base + __popc(m & ((1u << lane) - 1u))] = (int)r;
  }
}
// one CTA per tile: slot q -> point r + k N of the tile's class, -1 for the padding tail of the class
__global__ void __launch_bounds__(kTile) k_perm_fill(const int* __restrict__ slot_start, const int* __restrict__ ray_start,
                                                     const int* __restrict__ ray_cnt, const int* __restrict__ rays_sorted,
                                                     int nci, int S, int64_t N, const int* __restrict__ counts, int* perm,
                                                     int* inv) {
  __shared__ int cls;
  const int tile = blockIdx.x;
  if (tile >= counts[cTiles]) return;
  const int q0 = tile * kTile;
  if (threadIdx.x == 0) {
    int c = 0;
    while (c + 1 < nci && slot_start[c + 1] <= q0) ++c;
    cls = c;
  }
  __syncthreads();
  const int c = cls, nc = ray_cnt[c];
  const int64_t t = (int64_t)q0 + threadIdx.x - slot_start[c];
  int p = -1;
  if (nc > 0 && t < (int64_t)nc * S) p = (int)(rays_sorted[ray_start[c] + (int)(t % nc)] + (t / nc) * N);
  perm[q0 + threadIdx.x] = p;
  if (inv && p >= 0) inv[p] = q0 + threadIdx.x;
}

// ---------------------------------------------------------------------------------------
// point kernels
// ---------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kTile) k_point_fwd(PointArgs a) {
  extern __shared__ float sm[];
  float* XS = sm;                    // [128][81]
  float* Wc = XS + kTile * kXld;     // coarse block
  float* Wf = Wc + kNetT;            // expert block (MAP)
  __shared__ float red[32];
  const int tile = blockIdx.x, tid = threadIdx.x;
  const int n_tiles = a.perm ? a.counts[cTiles] : a.n_tiles_host;
  if (tile >= n_tiles) return;
  int expert = -1;
  load_block(Wc, a.WTc, kNetT / 4);
  if (MODE == kMap) {
    expert = a.tile_class[tile];
    if (expert >= 0) load_block(Wf, a.WTe + (int64_t)expert * kNetT, kNetT / 4);
  }
  const int64_t q = (int64_t)tile * kTile + tid;
  int64_t i, r;
  float zv, x[3];
  const bool valid = slot_point<MODE>(a, q, i, r, zv, x);
  float* xrow = XS + tid * kXld;
  if (valid) {
#pragma unroll
    for (int c = 0; c < 3; ++c) oneblob_fwd(x[c], 16, xrow + 16 * c, 1);
    hashgrid_fwd(a.G, a.table, x, xrow + DNS_PE_DIM, 1);
  } else {
    for (int k = 0; k < kIn1; ++k) xrow[k] = 0.f;
  }
  __syncthreads();
  float h[32], out[kOutP];
  net80_fwd(xrow, Wc, Wc + 2560, h, out);
  store_row(a.Hc + q * 32, h);
  float lt = 0.f, fs = 0.f, op = 0.f;
  if (MODE == kTv) {
    if (valid) a.occ[q] = out[0];
  } else if (MODE == kTrack) {
    if (valid) store_row(a.fine36 + (a.p0 + i) * kOutP, out);
  } else {
    float fo[kOutP];
    if (valid) store_row(a.coarse36 + (a.p0 + i) * kOutP, out);
    if (expert >= 0) {
      net80_fwd(xrow, Wf, Wf + 2560, h, fo);
    } else {
      zero(h);
      zero(fo);
    }
    store_row(a.Hf + q * 32, h);
    if (valid) {
      store_row(a.fine36 + (a.p0 + i) * kOutP, fo);
#pragma unroll
      for (int c = 0; c < DNS_LATENT; ++c) {
        float d = out[c] - fo[c];
        lt = fmaf(d, d, lt);
      }
      float front, band, vd, d = a.gt_depth[r];
      opacity_masks(zv, d, a.trunc, front, band, vd);
      float o = sigmoidf_(10.f * fo[32]);
      float t = o * front * vd;
      fs = t * t;
      float u = (zv - d) / a.sigma;
      float ps = 0.5f * __expf(-0.5f * u * u);
      float e = o * band - ps * band;
      op = e * e;
    }
  }
  if (a.need_dparams) {  // X stash: the tile is one contiguous [128][80] block
    __syncthreads();
    float* dst = a.Xst + (int64_t)tile * kTile * kIn1;
    for (int e = tid; e < kTile * kIn1; e += kTile) {
      int rr = e / kIn1, k = e - rr * kIn1;
      dst[e] = XS[rr * kXld + k];
    }
  }
  if (MODE == kMap) {
    lt = block_reduce_sum(lt, red);
    fs = block_reduce_sum(fs, red);
    op = block_reduce_sum(op, red);
    if (tid == 0) {
      atomicAdd(a.raw + rLt, lt);
      atomicAdd(a.raw + rFs, fs);
      atomicAdd(a.raw + rOp, op);
    }
  }
}

// dh = relu'(h) * (W2^T d_out);  d_in[k] (+)= sum_j dh[j] W1T[k][j]
__device__ __forceinline__ void net80_bwd(const float (&d_out)[kOutP], const float* hrow, const float* W1T,
                                          const float* W2T, float* xrow, bool accumulate, float* dh_st, float* do_st) {
  float h[32], dh[32];
  load_row(hrow, h);
#pragma unroll
  for (int j = 0; j < 32; ++j) dh[j] = h[j] > 0.f ? dot_row<kOutP>(d_out, W2T + j * kOutP) : 0.f;
  if (dh_st) store_row(dh_st, dh);
  if (do_st) store_row(do_st, d_out);
#pragma unroll 4
  for (int k = 0; k < kIn1; ++k) {
    float v = dot_row<32>(dh, W1T + k * 32);
    xrow[k] = accumulate ? xrow[k] + v : v;
  }
}

template <int MODE>
__global__ void __launch_bounds__(kTile) k_point_bwd(PointArgs a) {
  extern __shared__ float sm[];
  float* XS = sm;
  float* Wc = XS + kTile * kXld;
  float* Wf = Wc + kNetT;
  const int tile = blockIdx.x, tid = threadIdx.x;
  const int n_tiles = a.perm ? a.counts[cTiles] : a.n_tiles_host;
  if (tile >= n_tiles) return;
  int expert = -1;
  load_block(Wc, a.WTc, kNetT / 4);
  if (MODE == kMap) {
    expert = a.tile_class[tile];
    if (expert >= 0) load_block(Wf, a.WTe + (int64_t)expert * kNetT, kNetT / 4);
  }
  __syncthreads();
  const int64_t q = (int64_t)tile * kTile + tid;
  int64_t i, r;
  float zv, x[3];
  const bool valid = slot_point<MODE>(a, q, i, r, zv, x);
  float* xrow = XS + tid * kXld;
  float d_out[kOutP];
  zero(d_out);
  float* dHc = a.need_dparams ? a.dHc + q * 64 : nullptr;
  float* dOc = a.need_dparams ? a.dOc + q * kOutP : nullptr;
  if (MODE == kTv) {
    if (valid) d_out[0] = a.docc[q];
    net80_bwd(d_out, a.Hc + q * 32, Wc, Wc + 2560, xrow, false, dHc, dOc);
  } else if (MODE == kTrack) {
    if (valid) load_row(a.dfine36 + (a.p0 + i) * kOutP, d_out);
    net80_bwd(d_out, a.Hc + q * 32, Wc, Wc + 2560, xrow, false, dHc, dOc);
  } else {
    float g_lt = 2.f * a.lam_lt / (33.f * (float)a.P_total);
    float co[kOutP], fo[kOutP];
    zero(co);
    zero(fo);
    if (valid) {
      const int64_t p = a.p0 + i;
      load_row(a.dfine36 + p * kOutP, d_out);
      load_row(a.coarse36 + p * kOutP, co);
      load_row(a.fine36 + p * kOutP, fo);
#pragma unroll
      for (int c = 0; c < DNS_LATENT; ++c) d_out[c] -= g_lt * (co[c] - fo[c]);
      if (a.counts[cFront] > 0 && a.counts[cBand] > 0) {
        float front, band, vd, d = a.gt_depth[r];
        opacity_masks(zv, d, a.trunc, front, band, vd);
        float o = sigmoidf_(10.f * fo[32]);
        float u = (zv - d) / a.sigma;
        float ps = 0.5f * __expf(-0.5f * u * u);
        float inv_p = 1.f / (float)a.P_total;
        float d_o = 2.f * a.lam_fs * inv_p * o * front * vd + 2.f * a.lam_op * inv_p * (o - ps) * band;
        d_out[32] += d_o * 10.f * o * (1.f - o);
      }
    }
    if (expert >= 0) {
      net80_bwd(d_out, a.Hf + q * 32, Wf, Wf + 2560, xrow, false, a.need_dparams ? a.dHc + q * 64 + 32 : nullptr,
                a.need_dparams ? a.dOf + q * kOutP : nullptr);
    } else {
      for (int k = 0; k < kIn1; ++k) xrow[k] = 0.f;
    }
    // coarse net: only the latent loss reaches it in mapping (mapping.py:624-626 render from fine)
#pragma unroll
    for (int c = 0; c < kOutP; ++c) d_out[c] = (valid && c < DNS_LATENT) ? g_lt * (co[c] - fo[c]) : 0.f;
    net80_bwd(d_out, a.Hc + q * 32, Wc, Wc + 2560, xrow, true, dHc, dOc);
  }
  if (!valid) return;
  float dx[3] = {0.f, 0.f, 0.f}, dxg[3];
  if (a.need_drays) {
#pragma unroll
    for (int c = 0; c < 3; ++c) dx[c] = oneblob_bwd(x[c], 16, xrow + 16 * c, 1);
  }
  hashgrid_bwd(a.G, a.table, a.need_dparams ? a.d_table : nullptr, x, xrow + DNS_PE_DIM, 1, a.need_drays != 0, dxg);
  if (a.need_drays && MODE != kTv) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float g = (dx[c] + dxg[c]) / (float)a.B.ext[c];
      atomicAdd(a.d_rays_o + 3 * r + c, g);
      atomicAdd(a.d_rays_d + 3 * r + c, g * zv);
    }
  }
}

// ---------------------------------------------------------------------------------------
// ray kernel: out_fn + compositing + losses + backward (thread per sample point)
// ---------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) k_ray(RayArgs a) {
  extern __shared__ float sm[];
  const int T = a.T, S = a.S, RPC = a.RPC, C = a.C, C4 = a.C4, ld = T + 1;
  float* W1 = sm;                       // [112][64]
  float* W2c = W1 + kIn2 * 64;          // [32][4]
  float* XC = W2c + 128;                // [48][T+1] per-thread columns / staging
  float* bs = XC + 48 * ld;             // [T]
  float* us = bs + T;                   // [T]
  float* ws = us + T;                   // [T]
  float* HB = ws + T;                   // [RPC][32]
  float* QV = HB + RPC * 32;            // [RPC][32]
  float* RO = QV + RPC * 32;            // [RPC][8] ray outputs
  float* RG = RO + RPC * 8;             // [RPC][8] ray gradients
  float* LG = RG + RPC * 8;             // [RPC][C4] logits, then d_logits
  float* LS = LG + RPC * C4;            // [4] CTA loss partial sums
  const int t = threadIdx.x;
  load_block(W1, a.W1T2, kIn2 * 64 / 4);
  load_block(W2c, a.W2cT, 32);
  if (t < 4) LS[t] = 0.f;
  const int lr = t / S, s = t - lr * S;
  const int64_t rl = (int64_t)blockIdx.x * RPC + lr;      // chunk-local ray
  const bool valid = lr < RPC && rl < a.Nc;
  const int64_t r = a.ray0 + rl;                          // global ray
  const int64_t pl = rl * S + s, p = r * S + s;           // chunk-local / global point
  const int rb = lr * S;                                  // first thread of my ray
  float* xc = XC + t;
  float x[3] = {0.f, 0.f, 0.f}, zv = 0.f, occ = 0.f;
  float h[64];
  zero(h);
  __syncthreads();
  if (valid) {
    zv = a.z[p];
    point_from_ray(a.rays_o + 3 * r, a.rays_d + 3 * r, zv, a.B, x);
    // segment 1: OneBlob(x) -> 48
#pragma unroll
    for (int c = 0; c < 3; ++c) oneblob_fwd(x[c], 16, xc + 16 * c * ld, ld);
    if (a.need_dparams)
      for (int k = 0; k < 48; ++k) a.X2[pl * kIn2 + k] = xc[k * ld];
    accum_layer<64>(xc, ld, 48, W1, 64, h);
    // segment 2: latent channels 1..32 of the fine (TRACK: coarse) output
    {
      float row[kOutP];
      load_row(a.fine36 + p * kOutP, row);
      occ = row[0];
#pragma unroll
      for (int k = 0; k < 32; ++k) xc[k * ld] = row[1 + k];
      if (a.need_dparams)
        for (int k = 0; k < 32; ++k) a.X2[pl * kIn2 + 48 + k] = row[1 + k];
    }
    accum_layer<64>(xc, ld, 32, W1 + 48 * 64, 64, h);
    // segment 3: merged pixel feature
    {
      float row[32];
      if (a.features) load_row(a.features + p * 32, row);
      else zero(row);
#pragma unroll
      for (int k = 0; k < 32; ++k) xc[k * ld] = row[k];
      if (a.need_dparams) store_row(a.X2 + pl * kIn2 + 80, row);
    }
    accum_layer<64>(xc, ld, 32, W1 + 80 * 64, 64, h);
#pragma unroll
    for (int j = 0; j < 64; ++j) h[j] = fmaxf(h[j], 0.f);
  }
  // colour head: 32 -> 3, sigmoid (decoder.py:123)
  float rgb[3] = {0.f, 0.f, 0.f};
  {
    float pre[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float4 w = *reinterpret_cast<const float4*>(W2c + 4 * j);
      pre[0] = fmaf(h[j], w.x, pre[0]);
      pre[1] = fmaf(h[j], w.y, pre[1]);
      pre[2] = fmaf(h[j], w.z, pre[2]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) rgb[c] = sigmoidf_(pre[c]);
  }
  // occupancy compositing (common.py:524-532)
  const float alpha = valid ? sigmoidf_(10.f * occ) : 0.f;
  const float b = __fadd_rn(1.f - alpha, 1e-10f);
  bs[t] = b;
  __syncthreads();
  float Ts = 1.f;
  if (valid)
    for (int j = 0; j < s; ++j) Ts *= bs[rb + j];
  const float u = alpha * Ts;
  us[t] = u;
  __syncthreads();
  float sumu = 0.f;
  if (valid)
    for (int j = 0; j < S; ++j) sumu += us[rb + j];
  const float w = valid ? (a.fwd_only == 2 ? 1.f : u / sumu) : 0.f;   // 2: free-point query, no compositing
  __syncthreads();
  // stage w * {logit hidden (32), rgb (3), z} and reduce per ray
  if (valid) {
#pragma unroll
    for (int j = 0; j < 32; ++j) xc[j * ld] = w * h[32 + j];
#pragma unroll
    for (int c = 0; c < 3; ++c) xc[(32 + c) * ld] = w * rgb[c];
    xc[35 * ld] = w * zv;
  }
  __syncthreads();
  for (int e = t; e < RPC * 36; e += T) {
    int l2 = e / 36, o = e - l2 * 36;
    if ((int64_t)blockIdx.x * RPC + l2 < a.Nc) {
      const float* col = XC + o * ld + l2 * S;
      float acc = 0.f;
      for (int j = 0; j < S; ++j) acc += col[j];
      if (o < 32) HB[l2 * 32 + o] = acc;
      else RO[l2 * 8 + (o - 32)] = acc;  // 0..2 rgb, 3 depth
    }
  }
  __syncthreads();
  const float dz = valid ? zv - RO[lr * 8 + 3] : 0.f;
  bs[t] = w * dz * dz;
  us[t] = w * dz;
  // semantic head layer 2 on the composited hidden state (linear => commutes with compositing)
  const float* W2l = a.logit + 32 * kIn2;
  for (int e = t; e < RPC * C; e += T) {
    int l2 = e / C, c = e - l2 * C;
    if ((int64_t)blockIdx.x * RPC + l2 < a.Nc) {
      const float4* wr = reinterpret_cast<const float4*>(W2l + c * 32);
      const float* hb = HB + l2 * 32;
      float acc = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 v = __ldg(wr + q);
        acc = fmaf(hb[4 * q], v.x, acc);
        acc = fmaf(hb[4 * q + 1], v.y, acc);
        acc = fmaf(hb[4 * q + 2], v.z, acc);
        acc = fmaf(hb[4 * q + 3], v.w, acc);
      }
      LG[l2 * C4 + c] = acc;
    }
  }
  __syncthreads();
  // per-ray losses and their gradients (one thread per ray)
  if (valid && s == 0) {
    float var = 0.f, swdz = 0.f;
    for (int j = 0; j < S; ++j) {
      var += bs[rb + j];
      swdz += us[rb + j];
    }
    float* ro = RO + lr * 8;
    float* rg = RG + lr * 8;
    float* lg = LG + lr * C4;
    const float dhat = ro[3];
    a.pred_color[3 * r] = ro[0];
    a.pred_color[3 * r + 1] = ro[1];
    a.pred_color[3 * r + 2] = ro[2];
    a.pred_depth[r] = dhat;
    a.pred_var[r] = var;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) {
      a.pred_logits[r * C + c] = lg[c];
      mx = fmaxf(mx, lg[c]);
    }
    const float gd = a.gt_depth[r];
    int64_t lab = a.gt_label[r];
    if (lab < 0 || lab >= C) {   // torch's cross_entropy raises here; the flag surfaces as losses[7] = -3
      *a.err = 3;
      lab = 0;
    }
    const bool track = a.mode == kTrack;
    const bool m = track ? (a.mask ? a.mask[r] != 0 : true) : true;
    const float n_ray = track ? (float)a.counts[cMask] : (float)a.N_total;
    float lp = 0.f, ldp = 0.f, ll = 0.f;
    float g_rgb[3] = {0.f, 0.f, 0.f}, g_d = 0.f, g_var = 0.f, g_ce = 0.f;
    if (m) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float e = ro[c] - a.gt_color[3 * r + c];
        lp = fmaf(e, e, lp);
        g_rgb[c] = a.lam_p * 2.f * e / (3.f * n_ray);
      }
      float diff = dhat - gd;
      float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
      if (track) {  // tracking.py:89-92: |d - d^| / sqrt(var + 1e-10)
        float vv = var + 1e-10f, inv = rsqrtf(vv);
        ldp = fabsf(diff) * inv;
        g_d = a.lam_d * sgn * inv / n_ray;
        g_var = a.lam_d * fabsf(diff) * (-0.5f) * inv / vv / n_ray;
        g_d += g_var * (-2.f) * swdz;
      } else if (gd > 0.f) {  // mapping.py:114-117
        ldp = fabsf(diff);
        g_d = a.lam_d * sgn / (float)a.counts[cDpos];
      }
      float se = 0.f;
      for (int c = 0; c < C; ++c) se += __expf(lg[c] - mx);
      float lse = logf(se) + mx;
      ll = lse - lg[lab];
      g_ce = a.lam_l / n_ray;
      for (int c = 0; c < C; ++c) lg[c] = g_ce * (__expf(lg[c] - lse) - (c == lab ? 1.f : 0.f));
    } else {
      for (int c = 0; c < C; ++c) lg[c] = 0.f;
    }
    for (int c = C; c < C4; ++c) lg[c] = 0.f;
    rg[0] = g_rgb[0];
    rg[1] = g_rgb[1];
    rg[2] = g_rgb[2];
    rg[3] = g_d;
    rg[4] = g_var;
    atomicAdd(LS + 0, lp);
    atomicAdd(LS + 1, ldp);
    atomicAdd(LS + 2, ll);
    if (a.need_dparams) {
      for (int c = 0; c < C4; ++c) a.dlogit[rl * C4 + c] = lg[c];
      for (int j = 0; j < 32; ++j) a.Hbar[rl * 32 + j] = HB[lr * 32 + j];
    }
  }
  __syncthreads();
  if (t == 0) {
    atomicAdd(a.raw + rP, LS[0]);
    atomicAdd(a.raw + rD, LS[1]);
    atomicAdd(a.raw + rL, LS[2]);
  }
  if (a.fwd_only) return;   // inference (uniform over the CTA)
  // QV[ray][j] = sum_c d_logit[c] * W2l[c][j]
  for (int e = t; e < RPC * 32; e += T) {
    int l2 = e >> 5, j = e & 31;
    if ((int64_t)blockIdx.x * RPC + l2 < a.Nc) {
      const float* lg = LG + l2 * C4;
      float acc = 0.f;
      for (int c = 0; c < C; ++c) acc = fmaf(lg[c], __ldg(W2l + c * 32 + j), acc);
      QV[e] = acc;
    }
  }
  __syncthreads();
  // ---- per-point backward
  float d_w = 0.f;
  if (valid) {
    const float* rg = RG + lr * 8;
    const float* qv = QV + lr * 32;
    d_w = rg[0] * rgb[0] + rg[1] * rgb[1] + rg[2] * rgb[2] + rg[3] * zv + rg[4] * dz * dz;
#pragma unroll
    for (int j = 0; j < 32; ++j) d_w = fmaf(h[32 + j], qv[j], d_w);
  }
  bs[t] = w * d_w;
  ws[t] = b;
  __syncthreads();
  float G = 0.f;
  if (valid)
    for (int j = 0; j < S; ++j) G += bs[rb + j];
  const float d_u = valid ? (d_w - G) / sumu : 0.f;
  us[t] = d_u * u;
  __syncthreads();
  float d_occ = 0.f;
  if (valid) {
    float suf = 0.f;
    for (int j = s + 1; j < S; ++j) suf += us[rb + j];
    float d_alpha = d_u * Ts - suf / b;
    d_occ = d_alpha * 10.f * alpha * (1.f - alpha);
  }
  if (valid) {
    const float* rg = RG + lr * 8;
    const float* qv = QV + lr * 32;
    float dp[4];
#pragma unroll
    for (int c = 0; c < 3; ++c) dp[c] = w * rg[c] * rgb[c] * (1.f - rgb[c]);
    dp[3] = 0.f;
    if (a.need_dparams) {
      float hc[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) hc[j] = h[j];
      store_row(a.Hcol + pl * 32, hc);
      store_row(a.dpre + pl * 4, dp);
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float4 wv = *reinterpret_cast<const float4*>(W2c + 4 * j);
      h[j] = h[j] > 0.f ? dp[0] * wv.x + dp[1] * wv.y + dp[2] * wv.z : 0.f;
      h[32 + j] = h[32 + j] > 0.f ? w * qv[j] : 0.f;
    }
    if (a.need_dparams) store_row(a.dH2 + pl * 64, h);
    // d(latent) -> dfine row, channel 0 = d(occupancy)
    {
      float row[kOutP];
      row[0] = d_occ;
#pragma unroll
      for (int k = 0; k < 32; ++k) row[1 + k] = dot_row<64>(h, W1 + (48 + k) * 64);
      row[33] = row[34] = row[35] = 0.f;
      store_row(a.dfine36 + p * kOutP, row);
    }
    if (a.need_dfeat) {
      float row[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) row[k] = dot_row<64>(h, W1 + (80 + k) * 64);
      store_row(a.d_features + p * 32, row);
    }
  }
  if (a.need_drays) {
    __syncthreads();
    if (valid) {
      for (int k = 0; k < 48; ++k) xc[k * ld] = dot_row<64>(h, W1 + k * 64);
      float g[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) g[c] = oneblob_bwd(x[c], 16, xc + 16 * c * ld, ld) / (float)a.B.ext[c];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        xc[c * ld] = g[c];
        xc[(3 + c) * ld] = g[c] * zv;
      }
    }
    __syncthreads();
    for (int e = t; e < RPC * 6; e += T) {
      int l2 = e / 6, o = e - l2 * 6;
      int64_t rr = (int64_t)blockIdx.x * RPC + l2;
      if (rr < a.Nc) {
        const float* col = XC + o * ld + l2 * S;
        float acc = 0.f;
        for (int j = 0; j < S; ++j) acc += col[j];
        if (o < 3) a.d_rays_o[3 * (a.ray0 + rr) + o] = acc;
        else a.d_rays_d[3 * (a.ray0 + rr) + o - 3] = acc;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------
__global__ void k_finalize(int mode, const float* raw, const int* counts, int64_t N, int64_t P, float lp, float ld,
                           float ll, float llt, float lfs, float lop, float* losses) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float p, d, l, lt = 0.f, fs = 0.f, op = 0.f, nv;
  if (mode == kTrack) {
    nv = (float)counts[cMask];
    p = raw[rP] / (3.f * nv);
    d = raw[rD] / nv;
    l = raw[rL] / nv;
  } else {
    nv = (float)N;
    p = raw[rP] / (3.f * nv);
    d = raw[rD] / (float)counts[cDpos];
    l = raw[rL] / nv;
    lt = raw[rLt] / (33.f * (float)P);
    if (counts[cFront] > 0 && counts[cBand] > 0) {
      fs = raw[rFs] / (float)P;
      op = raw[rOp] / (float)P;
    }
  }
  losses[0] = p;
  losses[1] = d;
  losses[2] = l;
  losses[3] = lt;
  losses[4] = fs;
  losses[5] = op;
  losses[6] = lp * p + ld * d + ll * l + llt * lt + lfs * fs + lop * op;
  losses[7] = counts[cErr] ? -(float)counts[cErr] : nv;
}
__global__ void k_unpad33(const float* __restrict__ src36, float* __restrict__ dst33, int64_t P) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * DNS_LATENT) return;
  int64_t p = i / DNS_LATENT;
  dst33[i] = src36[p * kOutP + (i - p * DNS_LATENT)];
}
// TV stencil (mapping.py:153-157): loss and d(loss)/d(occ)
__global__ void k_tv_stencil(const float* __restrict__ occ, int n, float inv_norm, float lambda, float* docc,
                             float* loss) {
  __shared__ float red[32];
  int64_t n3 = (int64_t)n * n * n;
  int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  if (q < n3) {
    int ix = (int)(q / ((int64_t)n * n)), iy = (int)((q / n) % n), iz = (int)(q % n);
    float v = occ[q], g = 0.f;
    const int64_t st[3] = {(int64_t)n * n, n, 1};
    const int id[3] = {ix, iy, iz};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (id[c] + 1 < n) {
        float d = occ[q + st[c]] - v;
        acc = fmaf(d, d, acc);
        g -= 2.f * d;
      }
      if (id[c] > 0) g += 2.f * (v - occ[q - st[c]]);
    }
    docc[q] = g * inv_norm * lambda;
  }
  acc = block_reduce_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(loss, acc * inv_norm);
}

// ---------------------------------------------------------------------------------------
// workspace carving
// ---------------------------------------------------------------------------------------
struct Carver {
  char* base;
  int64_t off;
  template <typename Tp>
  Tp* take(int64_t n) {
    off = (off + 255) & ~(int64_t)255;
    Tp* p = base ? reinterpret_cast<Tp*>(base + off) : nullptr;
    off += n * (int64_t)sizeof(Tp);
    return p;
  }
};
struct RenderWs {
  int* counts;
  float* raw;
  int *hist, *slot_start, *cursor, *ray_start, *rays_sorted;
  float *WTc, *WTe, *W1T2, *W2cT;
  uint4 *W1o_hi, *W1o_lo;   // bf16 hi / lo chunk tiles of the colour|logit layer-1 weights (tcgen05 path)
  uint4 *W1o16_hi, *W1o16_lo;   // the same weights as fp16 halves (forward GEMM)
  uint4 *wc_tc, *we_tc;     // prepared tiles of the coarse net and of every class expert (kNetTc uint4 each)
  int *perm, *inv, *tile_class;
  float *fine36, *coarse36, *dfine36, *diff36s, *dfine36s;
  float *Xst, *Hc, *Hf, *dHc, *dHf, *dOc, *dOf, *Jst;
  float *X2, *dH2, *Hcol, *dpre, *dlogit, *Hbar;
  float2* dpriv;   // privatised gradient copies of the small hash-grid levels (kPrivBytes)
  int64_t Q, tiles;
};
// Privatised copies of the small leading levels of the gradient table (PointArgs::d_priv): levels of fewer than 2^16
// entries, at most kPrivLevelBytes together, as many copies (<= kPrivCopies) as fit kPrivBytes.
constexpr int64_t kPrivBytes = 8 << 20, kPrivLevelBytes = 2 << 20;
constexpr int kPrivCopies = 8;
static void priv_plan(const dns_grid& G, int& levels, uint32_t& end, int& copies) {
  levels = 0;
  end = 0;
  while (levels < G.n_levels && G.size[levels] < 65536u && (int64_t)(G.offset[levels + 1]) * 8 <= kPrivLevelBytes) {
    ++levels;
    end = G.offset[levels];
  }
  copies = 0;
  if (levels > 0) {
    int64_t k = kPrivBytes / ((int64_t)end * 8);
    copies = (int)(k < kPrivCopies ? k : kPrivCopies);
    if (copies < 2) levels = copies = 0, end = 0;
  }
}
// d_table[i] += sum_k d_priv[k][i]; the copies are cleared for the next call
__global__ void k_priv_reduce(float2* __restrict__ d_table, float2* __restrict__ d_priv, uint32_t n, int copies) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float sx = 0.f, sy = 0.f;
  for (int k = 0; k < copies; ++k) {
    const float2 v = d_priv[(size_t)k * n + i];
    sx += v.x;
    sy += v.y;
  }
  if (sx != 0.f || sy != 0.f) {
    float2 t = d_table[i];
    t.x += sx;
    t.y += sy;
    d_table[i] = t;
  }
}
static int64_t carve(RenderWs& w, char* base, int mode, int64_t Nc, int S, int C, int nci) {
  Carver c{base, 0};
  const int64_t Pc = Nc * S;
  const bool map = mode == DNS_MODE_MAP;
  w.tiles = (Pc + kTile - 1) / kTile + (map ? nci : 0);
  w.Q = w.tiles * kTile;
  const int C4 = (C + 3) & ~3;
  w.counts = c.take<int>(32);                          // counts [16] | raw [16]: one block, one memset per call
  w.raw = reinterpret_cast<float*>(w.counts ? w.counts + 16 : nullptr);
  w.hist = c.take<int>(nci + 1);
  w.slot_start = c.take<int>(nci + 2);
  w.cursor = c.take<int>(nci + 1);
  w.ray_start = c.take<int>(nci + 2);
  w.rays_sorted = c.take<int>(map ? Nc : 4);
  w.WTc = c.take<float>(kNetT);
  w.WTe = c.take<float>((int64_t)kNetT * (map ? nci : 0) + 4);
  w.W1T2 = c.take<float>(kIn2 * 64);
  w.W2cT = c.take<float>(128);
  w.W1o_hi = c.take<uint4>(14 * 64);
  w.W1o_lo = c.take<uint4>(14 * 64);
  w.W1o16_hi = c.take<uint4>(14 * 64);
  w.W1o16_lo = c.take<uint4>(14 * 64);
  w.wc_tc = c.take<uint4>(kNetTc);
  w.we_tc = c.take<uint4>((int64_t)kNetTc * (map ? nci : 0) + 4);
  w.perm = c.take<int>(map ? w.Q : 4);
  w.inv = c.take<int>(map ? Pc : 4);
  w.diff36s = c.take<float>(map ? w.Q * kOutP : 4);
  w.dfine36s = c.take<float>(map ? w.Q * kOutP : 4);
  w.tile_class = c.take<int>(w.tiles);
  w.fine36 = c.take<float>(Pc * kOutP);
  w.coarse36 = c.take<float>(map ? Pc * kOutP : 4);
  w.dfine36 = c.take<float>(Pc * kOutP);
  // Jacobian image of the grid features (PointArgs::Jst): 24 float4 per slot, written by the forward point kernel when
  // ray gradients are wanted.  (Round 1 tried the same values as [slot][96] ROWS -- every lane its own cache line, slower
  // than re-reading the corners; the image is lane-contiguous like the other tile images.)
  w.Jst = c.take<float>(w.Q * 96);
  // Stashes for the weight gradients.  SIMT path: fp32 rows.  tcgen05 path: the same regions hold bf16 hi/lo
  // tile images (32 B per value pair = the fp32 footprint; dOut rows padded 36 -> 40, ray-side rows padded to
  // whole CTAs of T rows).
  w.Xst = c.take<float>(w.Q * kIn1);
  w.Hc = c.take<float>(w.Q * 32 * (map ? 2 : 1));
  w.Hf = w.Hc ? w.Hc + w.Q * 32 : nullptr;
  w.dHc = c.take<float>(w.Q * 64);   // [Q][64]: columns 0..31 coarse, 32..63 class expert
  w.dHf = w.dHc ? w.dHc + 32 : nullptr;
  w.dOc = c.take<float>(w.Q * 40 * (map ? 2 : 1));
  w.dOf = w.dOc ? w.dOc + w.Q * 40 : nullptr;
  int Tt, Rt;
  pick_ray_block_any(S, Tt, Rt);
  const int64_t img_rows = ((Nc + Rt - 1) / Rt) * Tt;
  const int64_t rows2 = img_rows > Pc ? img_rows : Pc;
  w.X2 = c.take<float>(rows2 * kIn2);
  w.dH2 = c.take<float>(rows2 * 64);
  w.Hcol = c.take<float>(rows2 * 32);
  w.dpre = c.take<float>(rows2 * 8);
  w.dlogit = c.take<float>(Nc * C4);
  w.Hbar = c.take<float>(Nc * 32);
  w.dpriv = c.take<float2>(kPrivBytes / 8);
  return c.off + 256;
}
constexpr int64_t kMaxChunkRays = 1 << 17;

static void pick_ray_block(int S, int& T, int& RPC) {
  int best_r = 1;
  double best = 0.0;
  int rmax = 256 / S;
  if (rmax < 1) rmax = 1;
  if (rmax > 16) rmax = 16;
  for (int rr = 1; rr <= rmax; ++rr) {
    int tt = (rr * S + 31) & ~31;
    double eff = (double)(rr * S) / tt;
    if (eff >= best - 1e-9) {
      best = eff;
      best_r = rr;
    }
  }
  RPC = best_r;
  T = (RPC * S + 31) & ~31;
}

}  // namespace dns

using namespace dns;

extern "C" {

int64_t dns_render_workspace_bytes(int mode, int n_rays, int n_samples, int n_class, int n_class_ids) {
  RenderWs w;
  int64_t nc = n_rays < kMaxChunkRays ? n_rays : kMaxChunkRays;
  if (nc < 1) nc = 1;
  return carve(w, nullptr, mode, nc, n_samples, n_class, n_class_ids < 1 ? 1 : n_class_ids);
}

int dns_render_fwd_bwd(const dns_render_args* a, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int mode = a->mode, S = a->n_samples, C = a->n_class;
  const int64_t N = a->n_rays;
  if (mode != DNS_MODE_TRACK && mode != DNS_MODE_MAP) {
    set_error("render: bad mode %d", mode);
    return DNS_ERR_ARG;
  }
  if (N <= 0 || S <= 0 || S > 256 || C < 1 || C > 128) {
    set_error("render: need n_rays > 0, 1 <= n_samples <= 256, 1 <= n_class <= 128 (got %lld, %d, %d)", (long long)N, S, C);
    return DNS_ERR_UNSUPPORTED;
  }
  if (a->grid.n_levels != 16 || a->grid.n_features != 2) {
    set_error("render: the fused path is built for 16 levels x 2 features");
    return DNS_ERR_UNSUPPORTED;
  }
  if (((uintptr_t)a->table & 15) || ((uintptr_t)a->d_table & 15)) {
    set_error("render: table / d_table must be 16-byte aligned (paired corner accesses)");
    return DNS_ERR_ARG;
  }
  if (a->features_band_only && a->use_simt) {
    set_error("render: features_band_only is a contract of the tcgen05 path");
    return DNS_ERR_UNSUPPORTED;
  }
  const bool map = mode == DNS_MODE_MAP;
  const int nci = map ? (a->n_class_ids < 1 ? 1 : a->n_class_ids) : 1;
  if (nci > 4096) {
    set_error("render: at most 4096 class ids (got %d)", nci);
    return DNS_ERR_UNSUPPORTED;
  }
  if (map && (!a->experts || !a->class_to_expert || a->n_experts < 1 || a->n_experts > nci)) {
    set_error("render: MAP mode needs experts, class_to_expert and 1 <= n_experts <= n_class_ids");
    return DNS_ERR_ARG;
  }
  // largest chunk whose scratch fits the workspace
  RenderWs w;
  int64_t Nc = N < kMaxChunkRays ? N : kMaxChunkRays;
  while (Nc > 1 && carve(w, nullptr, mode, Nc, S, C, nci) > a->workspace_bytes) Nc = (Nc + 1) / 2;
  if (!a->workspace || carve(w, nullptr, mode, Nc, S, C, nci) > a->workspace_bytes) {
    set_error("render: workspace too small (%lld bytes)", (long long)a->workspace_bytes);
    return DNS_ERR_ARG;
  }
  carve(w, (char*)a->workspace, mode, Nc, S, C, nci);
  const int C4 = (C + 3) & ~3;
  const bool sharded = a->n_rays_total > 0;
  const int64_t Ntot = sharded ? a->n_rays_total : N;           // global batch (denominators, class rule)
  const int64_t goff = sharded ? a->ray_offset * S : 0;         // global id of local point 0
  const int64_t* lab_all = (sharded && a->gt_label_all) ? a->gt_label_all : a->gt_label;
  const int64_t P = Ntot * S;
  Bound B;
  for (int c = 0; c < 3; ++c) {
    B.lo[c] = a->bound[c][0];
    B.ext[c] = a->bound[c][1] - a->bound[c][0];
  }
  const bool tc = !a->use_simt;
  // weight preparation: bf16 hi/lo chunk tiles (tcgen05 path) or k-major fp32 copies (SIMT path)
  PhaseScope* ph = new PhaseScope(phPrep, st, 1 + (a->global_counts ? 0 : 1) + (map ? 2 : 1));
  cudaMemsetAsync(w.counts, 0, 32 * sizeof(int), st);   // counts and raw loss partials
  if (!tc) {
    k_transpose_net80<<<1, 256, 0, st>>>(a->coarse, w.WTc);
    if (map) k_transpose_net80<<<a->n_experts, 256, 0, st>>>(a->experts, w.WTe);
  } else if (int e = prep_nets_tc(a->coarse, map ? a->experts : nullptr, map ? a->n_experts : 0, w.wc_tc, w.we_tc, st)) {
    delete ph;
    return e;
  }
  if (!tc) k_transpose_out<<<1, 256, 0, st>>>(a->color, a->logit, w.W1T2, w.W2cT);   // tcgen05 path: k_prep_w1o_tc writes W2cT
  {
    if (a->global_counts) {
      cudaMemcpyAsync(w.counts, a->global_counts, 4 * sizeof(int), cudaMemcpyDeviceToDevice, st);
    } else {
      int64_t blocks = (N * S + 255) / 256;
      k_counts<<<(int)(blocks < 1184 ? blocks : 1184), 256, 0, st>>>(a->gt_depth, a->z_vals, map ? nullptr : a->mask, N,
                                                                    S, a->opacity_trunc, w.counts);
    }
  }
  delete ph;
  if (int e = check_launch("render prep")) return e;

  static unsigned long long seen = 0;
  if (first_call_on_device(seen)) {
    cudaFuncSetAttribute(k_point_fwd<kTrack>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_point_fwd<kMap>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_point_fwd<kTv>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_point_bwd<kTrack>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_point_bwd<kMap>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_point_bwd<kTv>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_ray, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  }
  const size_t smem_pt = sizeof(float) * (kTile * kXld + kNetT * (map ? 2 : 1));
  int T, RPC;
  if (tc) pick_ray_block_any(S, T, RPC);
  else pick_ray_block(S, T, RPC);
  const size_t smem_ray =
      sizeof(float) * (kIn2 * 64 + 128 + 48 * (T + 1) + 3 * T + RPC * (32 + 32 + 8 + 8 + C4) + 8);

  for (int64_t ray0 = 0; ray0 < N; ray0 += Nc) {
    const int64_t nc = (N - ray0) < Nc ? (N - ray0) : Nc;
    const int64_t p0 = ray0 * S, Pc = nc * S;
    const int tiles_max = (int)((Pc + kTile - 1) / kTile + (map ? nci : 0));
    PointArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.rays_o = a->rays_o; pa.rays_d = a->rays_d; pa.z = a->z_vals; pa.gt_depth = a->gt_depth;
    pa.S = S; pa.N_total = Ntot; pa.P_total = P; pa.p0 = p0; pa.Pc = Pc; pa.B = B; pa.G = a->grid;
    pa.table = (const float2*)a->table;
    pa.counts = w.counts; pa.n_tiles_host = tiles_max;
    pa.WTc = w.WTc; pa.WTe = w.WTe;
    // point-order buffers are indexed with GLOBAL point ids: shift the chunk-local base
    pa.fine36 = w.fine36 - p0 * kOutP; pa.coarse36 = w.coarse36 - p0 * kOutP; pa.dfine36 = w.dfine36 - p0 * kOutP;
    pa.Jst = w.Jst; pa.Xst = w.Xst; pa.Hc = w.Hc; pa.Hf = w.Hf; pa.dHc = w.dHc; pa.dHf = w.dHf; pa.dOc = w.dOc; pa.dOf = w.dOf;
    pa.Ximg = (uint4*)w.Xst; pa.Himg = (uint4*)w.Hc; pa.dHimg = (uint4*)w.dHc; pa.dOimg = (uint4*)w.dOc;
    pa.lam_lt = a->lambda_lt; pa.lam_fs = a->lambda_fs; pa.lam_op = a->lambda_op;
    pa.trunc = a->opacity_trunc; pa.sigma = a->opacity_sigma;
    pa.raw = w.raw; pa.d_table = (float2*)a->d_table; pa.d_rays_o = a->d_rays_o; pa.d_rays_d = a->d_rays_d;
    pa.need_dparams = a->need_dparams && !a->forward_only;
    pa.need_drays = a->need_drays && a->d_rays_o && a->d_rays_d && !a->forward_only;
    // (tracking keeps the corner re-read: its slots are in ray order, neighbouring lanes share cells and the re-read hits
    // L1 -- the image would cost the forward more than it saves the backward: 0.229 -> 0.248 ms at 1024 x 96)
    if (!tc || !pa.need_drays || !map) pa.Jst = nullptr;
    if (tc && pa.need_dparams && a->d_table) {
      priv_plan(a->grid, pa.priv_levels, pa.priv_end, pa.priv_copies);
      pa.d_priv = pa.priv_copies ? w.dpriv : nullptr;
    }
#ifdef DNS_ABLATE
    { const char* e = getenv("DNS_DBG"); pa.dbg = e ? atoi(e) : 0; }
    if (getenv("DNS_NO_PRIV")) pa.d_priv = nullptr;
    if (getenv("DNS_NO_JIMG")) pa.Jst = nullptr;
    if (const char* e = getenv("DNS_PHASE_CLK")) pa.phase_clk = (unsigned long long*)strtoull(e, nullptr, 0);   // device pointer
    if (const char* e = getenv("DNS_PRIV_COPIES")) { if (pa.d_priv && atoi(e) >= 1 && atoi(e) <= pa.priv_copies) pa.priv_copies = atoi(e); }
    if (const char* e = getenv("DNS_PRIV_LEVELS")) { if (pa.d_priv && atoi(e) >= 1 && atoi(e) <= pa.priv_levels) { pa.priv_levels = atoi(e); pa.priv_end = a->grid.offset[pa.priv_levels]; } }
#endif
    if (map) {
      bool whole = !sharded && ray0 == 0 && nc == N && (int64_t)N * S < 2147483647LL;
#ifdef DNS_ABLATE
      if (getenv("DNS_GENERIC_PREP")) whole = false;
#endif
      PhaseScope phc(phClassPrep, st, whole ? 4 : 3);
      cudaMemsetAsync(w.hist, 0, (nci + 1) * sizeof(int), st);
      if (whole) {
        int rblocks = (int)((N + 255) / 256);
        rblocks = rblocks < 592 ? rblocks : 592;
        k_ray_hist<<<rblocks, 256, 0, st>>>(a->gt_label, N, nci, w.hist, w.counts);
        k_class_scan_rays<<<1, 256, 0, st>>>(w.hist, nci, S, a->class_to_expert, w.slot_start, w.ray_start, w.cursor,
                                            w.tile_class, w.counts);
        k_ray_scatter<<<rblocks, 256, 0, st>>>(a->gt_label, N, nci, w.ray_start, w.cursor, w.rays_sorted);
        k_perm_fill<<<tiles_max, kTile, 0, st>>>(w.slot_start, w.ray_start, w.hist, w.rays_sorted, nci, S, N, w.counts, w.perm, w.inv);
      } else {
      cudaMemsetAsync(w.perm, 0xFF, (size_t)tiles_max * kTile * sizeof(int), st);
      const int grid = (int)((Pc + 256 * kSortPer - 1) / (256 * kSortPer));
      k_class_hist<<<grid, 256, nci * sizeof(int), st>>>(lab_all, Ntot, goff + p0, Pc, nci, w.hist, w.counts);
      k_class_scan<<<1, 256, 0, st>>>(w.hist, nci, a->class_to_expert, w.slot_start, w.cursor, w.tile_class, w.counts);
      k_class_scatter<<<grid, 256, 2 * nci * sizeof(int), st>>>(lab_all, Ntot, goff + p0, Pc, nci, w.slot_start, w.cursor,
                                                                w.perm, w.inv);
      }
      pa.perm = w.perm; pa.tile_class = w.tile_class;
      if (tc && !a->forward_only) {   // slot-order hand-over of the latent rows between the three kernels
        pa.diff36s = w.diff36s;
        pa.dfine36s = w.dfine36s;
      }
      pa.want_coarse_pt = a->coarse_out != nullptr || !pa.diff36s;
    }
    {
      PhaseScope php(phPointFwd, st, 1);
      if (tc) {
        if (int e = launch_point_fwd_tc(map ? kMap : kTrack, pa, tiles_max, w.wc_tc, w.we_tc, st)) return e;
      } else if (map) k_point_fwd<kMap><<<tiles_max, kTile, smem_pt, st>>>(pa);
      else k_point_fwd<kTrack><<<tiles_max, kTile, smem_pt, st>>>(pa);
    }
    if (int e = check_launch("point_fwd")) return e;

    RayArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.mode = mode; ra.S = S; ra.T = T; ra.RPC = RPC; ra.C = C; ra.C4 = C4;
    ra.N_total = Ntot; ra.ray0 = ray0; ra.Nc = nc; ra.B = B;
    ra.rays_o = a->rays_o; ra.rays_d = a->rays_d; ra.z = a->z_vals; ra.gt_color = a->gt_color;
    ra.gt_depth = a->gt_depth; ra.gt_label = a->gt_label; ra.mask = map ? nullptr : a->mask; ra.features = a->features;
    ra.fine36 = pa.fine36; ra.w16_hi = w.W1o16_hi; ra.w16_lo = w.W1o16_lo; ra.W1T2 = w.W1T2; ra.W2cT = w.W2cT; ra.logit = a->logit; ra.counts = w.counts;
    ra.lam_p = a->lambda_p; ra.lam_d = a->lambda_d; ra.lam_l = a->lambda_l;
    ra.pred_color = a->pred_color; ra.pred_depth = a->pred_depth; ra.pred_var = a->pred_var;
    ra.pred_logits = a->pred_logits; ra.raw = w.raw; ra.err = w.counts + cErr; ra.dfine36 = pa.dfine36; ra.d_features = a->d_features;
    ra.d_rays_o = a->d_rays_o; ra.d_rays_d = a->d_rays_d;
    if (pa.dfine36s) { ra.inv = w.inv; ra.dfine36s = w.dfine36s; }
    ra.X2 = w.X2; ra.dH2 = w.dH2; ra.Hcol = w.Hcol; ra.dpre = w.dpre; ra.dlogit = w.dlogit; ra.Hbar = w.Hbar;
    ra.X2img = (uint4*)w.X2; ra.dH2img = (uint4*)w.dH2; ra.Hcolimg = (uint4*)w.Hcol; ra.dpreimg = (uint4*)w.dpre;
    ra.RS = T <= 96 ? T : (T == 160 ? 80 : (T == 256 ? 128 : 64));   // sub-tile rows of the ray-side images (divides T)
#ifdef DNS_ABLATE
    if (const char* e = getenv("DNS_PHASE_CLK_RAY")) ra.phase_clk = (unsigned long long*)strtoull(e, nullptr, 0);   // device pointer
    if (const char* e = getenv("DNS_RAY_RS")) ra.RS = atoi(e);
#endif
    const bool fwd_only = a->forward_only != 0;
    ra.fwd_only = a->forward_only;
    ra.need_dparams = a->need_dparams && !fwd_only; ra.need_drays = a->need_drays && a->d_rays_o && a->d_rays_d && !fwd_only;
    ra.need_dfeat = a->need_dfeat && a->d_features && !fwd_only;
    ra.feat_band = a->features_band_only;
    pa.need_drays = ra.need_drays;
    {
      PhaseScope phr(phRay, st, 1 + ((tc && ray0 == 0) ? 1 : 0));
      if (tc) {
        if (int e = launch_ray_tc(ra, a->color, a->logit, w.W1o_hi, w.W1o_lo, ray0 == 0, nc, st)) return e;
      } else {
        k_ray<<<(int)((nc + RPC - 1) / RPC), T, smem_ray, st>>>(ra);
      }
    }
    if (int e = check_launch("ray")) return e;
    if (!fwd_only) {
      PhaseScope phb(phPointBwd, st, pa.d_priv ? 2 : 1);
      if (pa.d_priv) cudaMemsetAsync(pa.d_priv, 0, (size_t)pa.priv_copies * pa.priv_end * sizeof(float2), st);
      if (tc) {
        if (int e = launch_point_bwd_tc(map ? kMap : kTrack, pa, tiles_max, w.wc_tc, w.we_tc, st)) return e;
      } else if (map) k_point_bwd<kMap><<<tiles_max, kTile, smem_pt, st>>>(pa);
      else k_point_bwd<kTrack><<<tiles_max, kTile, smem_pt, st>>>(pa);
      if (pa.d_priv)
        k_priv_reduce<<<(pa.priv_end + 255) / 256, 256, 0, st>>>((float2*)a->d_table, pa.d_priv, pa.priv_end, pa.priv_copies);
    }
    if (int e = check_launch("point_bwd")) return e;

    if (a->fine) k_unpad33<<<(int)((Pc * DNS_LATENT + 255) / 256), 256, 0, st>>>(w.fine36, a->fine + p0 * DNS_LATENT, Pc);
    if (map && a->coarse_out)
      k_unpad33<<<(int)((Pc * DNS_LATENT + 255) / 256), 256, 0, st>>>(w.coarse36, a->coarse_out + p0 * DNS_LATENT, Pc);

    if (a->need_dparams && !fwd_only) {
      PhaseScope phg(phDwGemm, st, tc ? 4 : (map ? 8 : 6));
      const int64_t Qrows = (int64_t)tiles_max * kTile;
      const int* ntd = map ? w.counts + cTiles : nullptr;
      int e = 0;
      const int pt_tiles = (int)((Pc + kTile - 1) / kTile), ray_tiles = (int)((nc + kTile - 1) / kTile);
      if (tc) {
        // tcgen05 path: every operand is a bf16 hi/lo tile image streamed by bulk copies (tc.cu: k_dw_img).
        //   X80^T [dHc | dHf]            -> coarse W1 (+ class-expert W1), the shared input read once
        //   dOut^T H  per net            -> coarse W2, class-expert W2
        //   X112^T [dH colour | dH logit] -> colour / logit W1;   dpre^T H colour -> colour W2
        const int hch = map ? 8 : 4, doch = map ? 10 : 5;
        DwImgArgs g;
        memset(&g, 0, sizeof(g));
        g.L = DwImg{pa.Ximg, 10, 0, 10, kIn1}; g.Cc = DwImg{pa.dHimg, hch, 0, hch, map ? 64 : 32};
        g.RS = kTile; g.subs_per_tile = 1; g.n_tiles_dev = ntd; g.n_tiles_host = tiles_max;
        g.tile_class = map ? w.tile_class : nullptr;
        g.out0 = a->d_coarse; g.split = 32; g.sl0 = 1; g.sc0 = kIn1; g.cls0 = 0;
        g.out1 = map ? a->d_experts : nullptr; g.sl1 = 1; g.sc1 = kIn1; g.cls1 = 4096;
        e |= launch_dw_img(g, st);
        if (map) {   // coarse and class-expert layer 2 in ONE pass: block-diagonal [dOc | dOf]^T [Hc | Hf]
          g.L = DwImg{pa.dOimg, doch, 0, 10, DNS_LATENT}; g.Cc = DwImg{pa.Himg, hch, 0, 8, 32};
          g.diag_l = 40; g.diag_c = 32;
          g.tile_class = w.tile_class;
          g.out0 = a->d_coarse + 2560; g.sl0 = 32; g.sc0 = 1; g.cls0 = 0;
          g.out1 = a->d_experts + 2560; g.sl1 = 32; g.sc1 = 1; g.cls1 = 4096;
          e |= launch_dw_img(g, st);
        } else {
          g.L = DwImg{pa.dOimg, doch, 0, 5, DNS_LATENT}; g.Cc = DwImg{pa.Himg, hch, 0, 4, 32};
          g.tile_class = nullptr;
          g.out0 = a->d_coarse + 2560; g.split = 32; g.sl0 = 32; g.sc0 = 1; g.cls0 = 0; g.out1 = nullptr;
          e |= launch_dw_img(g, st);
        }
        memset(&g, 0, sizeof(g));
        g.L = DwImg{ra.X2img, 14, 0, 14, kIn2}; g.Cc = DwImg{ra.dH2img, 8, 0, 8, 64};
        g.RS = ra.RS; g.subs_per_tile = T / ra.RS; g.n_tiles_host = (int)((nc + RPC - 1) / RPC);
        g.out0 = a->d_color; g.split = 32; g.sl0 = 1; g.sc0 = kIn2; g.out1 = a->d_logit; g.sl1 = 1; g.sc1 = kIn2;
        // ... and, in the same pass, the colour head's layer 2: dpre^T H colour (second operand pair of k_dw_img)
        g.L2 = DwImg{ra.dpreimg, 1, 0, 1, 3}; g.C2 = DwImg{ra.Hcolimg, 4, 0, 4, 32};
        g.out2 = a->d_color + 32 * kIn2; g.sl2 = 32; g.sc2 = 1;
        e |= launch_dw_img(g, st);
        e |= launch_dw_gemm(w.dlogit, C4, C, w.Hbar, 32, 32, nc, nullptr, ray_tiles, nullptr, a->d_logit + 32 * kIn2, 32, 0, st, tc);
        if (e) return DNS_ERR_CUDA;
        continue;
      } else {
        e |= launch_dw_gemm(w.dHc, 64, 32, w.Xst, kIn1, kIn1, Qrows, ntd, tiles_max, nullptr, a->d_coarse, kIn1, 0, st, tc);
        if (map)
          e |= launch_dw_gemm(w.dHf, 64, 32, w.Xst, kIn1, kIn1, Qrows, ntd, tiles_max, w.tile_class, a->d_experts, kIn1, 4096, st, tc);
        e |= launch_dw_gemm(w.dH2, 64, 32, w.X2, kIn2, kIn2, Pc, nullptr, pt_tiles, nullptr, a->d_color, kIn2, 0, st, tc);
        e |= launch_dw_gemm(w.dH2 + 32, 64, 32, w.X2, kIn2, kIn2, Pc, nullptr, pt_tiles, nullptr, a->d_logit, kIn2, 0, st, tc);
      }
      // layer-2 weight gradients: dOut^T H
      e |= launch_dw_gemm(w.dOc, kOutP, DNS_LATENT, w.Hc, 32, 32, Qrows, ntd, tiles_max, nullptr, a->d_coarse + 2560, 32, 0, st, tc);
      if (map)
        e |= launch_dw_gemm(w.dOf, kOutP, DNS_LATENT, w.Hf, 32, 32, Qrows, ntd, tiles_max, w.tile_class, a->d_experts + 2560, 32, 4096, st, tc);
      e |= launch_dw_gemm(w.dpre, 4, 3, w.Hcol, 32, 32, Pc, nullptr, pt_tiles, nullptr, a->d_color + 32 * kIn2, 32, 0, st, tc);
      e |= launch_dw_gemm(w.dlogit, C4, C, w.Hbar, 32, 32, nc, nullptr, ray_tiles, nullptr, a->d_logit + 32 * kIn2, 32, 0, st, tc);
      if (e) return DNS_ERR_CUDA;
    }
  }
  PhaseScope phf(phFinalize, st, 1);
  k_finalize<<<1, 32, 0, st>>>(mode, w.raw, w.counts, Ntot, P, a->lambda_p, a->lambda_d, a->lambda_l, a->lambda_lt,
                               a->lambda_fs, a->lambda_op, a->losses);
  return check_launch("finalize");
}

int dns_render_counts(const dns_render_args* a, int32_t* counts4, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t N = a->n_rays, S = a->n_samples;
  if (N <= 0 || S <= 0 || !counts4) {
    set_error("render_counts: bad arguments");
    return DNS_ERR_ARG;
  }
  PhaseScope ph(phPrep, st, 2);
  cudaMemsetAsync(counts4, 0, 4 * sizeof(int), st);
  int64_t blocks = (N * S + 255) / 256;
  k_counts<<<(int)(blocks < 1184 ? blocks : 1184), 256, 0, st>>>(
      a->gt_depth, a->z_vals, a->mode == DNS_MODE_MAP ? nullptr : a->mask, N, (int)S, a->opacity_trunc, counts4);
  return check_launch("render_counts");
}

int64_t dns_tv_workspace_bytes(int n) {
  int64_t n3 = (int64_t)n * n * n;
  int64_t Q = ((n3 + kTile - 1) / kTile) * kTile;
  return 4096 + 32768 + sizeof(float) * (kNetT + 2 * n3 + Q * (kIn1 + 32 + 64 + 40)) + 10 * 256 + kPrivBytes + 256;
}

int dns_tv_fwd_bwd(const dns_tv_args* a, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int n = a->n;
  if (n < 2 || n > 512) {
    set_error("tv: lattice size out of range (%d)", n);
    return DNS_ERR_ARG;
  }
  if (!a->workspace || a->workspace_bytes < dns_tv_workspace_bytes(n)) {
    set_error("tv: workspace too small");
    return DNS_ERR_ARG;
  }
  if (((uintptr_t)a->table & 15) || ((uintptr_t)a->d_table & 15)) {
    set_error("tv: table / d_table must be 16-byte aligned (paired corner accesses)");
    return DNS_ERR_ARG;
  }
  const int64_t n3 = (int64_t)n * n * n;
  const int tiles = (int)((n3 + kTile - 1) / kTile);
  const int64_t Q = (int64_t)tiles * kTile;
  Carver c{(char*)a->workspace, 0};
  float* WTc = c.take<float>(kNetT);
  float* occ = c.take<float>(n3);
  float* docc = c.take<float>(n3);
  float* Xst = c.take<float>(Q * kIn1);
  float* Hc = c.take<float>(Q * 32);
  float* dHc = c.take<float>(Q * 64);
  float* dOc = c.take<float>(Q * 40);   // fp32 rows of 36, or the 5-chunk tile image of the tcgen05 path
  uint4* wc_tc = c.take<uint4>(kNetTc);
  float2* dpriv = c.take<float2>(kPrivBytes / 8);
  const bool tc = !a->use_simt;
  static unsigned long long seen = 0;
  if (first_call_on_device(seen)) {
    cudaFuncSetAttribute(k_point_fwd<kTv>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_point_bwd<kTv>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  }
  if (a->use_simt) k_transpose_net80<<<1, 256, 0, st>>>(a->coarse, WTc);   // k-major fp32 copy: SIMT kernels only
  cudaMemsetAsync(a->loss, 0, sizeof(float), st);
  PointArgs pa;
  memset(&pa, 0, sizeof(pa));
  for (int k = 0; k < 3; ++k) {
    pa.B.lo[k] = a->bound[k][0];
    pa.B.ext[k] = a->bound[k][1] - a->bound[k][0];
    pa.jit[k] = a->jitter[k];
    pa.off[k] = a->offset[k];
  }
  pa.tv_oj = a->offset_jitter_dev;
  pa.n = n; pa.voxel = a->voxel; pa.G = a->grid; pa.table = (const float2*)a->table;
  pa.n_tiles_host = tiles; pa.WTc = WTc; pa.occ = occ; pa.docc = docc;
  pa.Xst = Xst; pa.Hc = Hc; pa.dHc = dHc; pa.dOc = dOc;
  pa.Ximg = (uint4*)Xst; pa.Himg = (uint4*)Hc; pa.dHimg = (uint4*)dHc; pa.dOimg = (uint4*)dOc;
  pa.d_table = (float2*)a->d_table; pa.need_dparams = a->need_dparams; pa.need_drays = 0;
  if (tc && a->need_dparams && a->d_table) {   // the lattice is spatially coherent: neighbouring points reduce into the SAME
    priv_plan(a->grid, pa.priv_levels, pa.priv_end, pa.priv_copies);   // coarse cells, the contention case of PointArgs::d_priv
    pa.d_priv = pa.priv_copies ? dpriv : nullptr;
#ifdef DNS_ABLATE
    if (getenv("DNS_NO_PRIV")) pa.d_priv = nullptr;
#endif
  }
  if (tc && a->need_dparams && a->d_table) {
    // levels whose cells are wider than the lattice spacing along x (the slots' fastest axis): several lanes of a warp share
    // a cell there and hashgrid_bwd_rows reduces them before the table sees them
    const double step = a->voxel / (a->bound[0][1] - a->bound[0][0]);
    double limit = 1.5;   // cells at least 2/3 of a lattice step wide: runs of 1-2 lanes still save the handed-over x+1 plane
#ifdef DNS_ABLATE
    if (const char* e = getenv("DNS_TV_AGG")) limit = atof(e);
#endif
    int lv = 0;
    while (lv < a->grid.n_levels && step * a->grid.scale[lv] < limit) ++lv;
    pa.tv_agg_levels = lv;
    // pre-reduced, the small levels no longer contend: no privatised copies (measured 0.910 -> 0.896 ms at 127^3, two launches fewer)
    if (lv >= pa.priv_levels) pa.d_priv = nullptr;
  }
  const size_t smem_pt = sizeof(float) * (kTile * kXld + kNetT);
  PhaseScope* pht = new PhaseScope(phTvFwd, st, 4);
  if (tc) {
    prep_nets_tc(a->coarse, nullptr, 0, wc_tc, nullptr, st);
    launch_point_fwd_tc(kTv, pa, tiles, wc_tc, wc_tc, st);
  } else {
    k_point_fwd<kTv><<<tiles, kTile, smem_pt, st>>>(pa);
  }
  const float inv_norm = 1.0f / ((float)a->smooth_pts * (float)a->smooth_pts * (float)a->smooth_pts);
  k_tv_stencil<<<(int)((n3 + 255) / 256), 256, 0, st>>>(occ, n, inv_norm, a->lambda_sm, docc, a->loss);
  delete pht;
  if (int e = check_launch("tv fwd")) return e;
  if (a->need_dparams) {
    PhaseScope phtb(phTvBwd, st, pa.d_priv ? 3 : 2);
    if (pa.d_priv) cudaMemsetAsync(pa.d_priv, 0, (size_t)pa.priv_copies * pa.priv_end * sizeof(float2), st);
    if (tc) launch_point_bwd_tc(kTv, pa, tiles, wc_tc, wc_tc, st);
    else k_point_bwd<kTv><<<tiles, kTile, smem_pt, st>>>(pa);
    if (pa.d_priv)
      k_priv_reduce<<<(pa.priv_end + 255) / 256, 256, 0, st>>>((float2*)a->d_table, pa.d_priv, pa.priv_end, pa.priv_copies);
    if (int e = check_launch("tv bwd")) return e;
    int e = 0;
    if (tc) {
      DwImgArgs g;
      memset(&g, 0, sizeof(g));
      g.L = DwImg{pa.Ximg, 10, 0, 10, kIn1}; g.Cc = DwImg{pa.dHimg, 4, 0, 4, 32};
      g.RS = kTile; g.subs_per_tile = 1; g.n_tiles_host = tiles;
      g.out0 = a->d_coarse; g.split = 32; g.sl0 = 1; g.sc0 = kIn1;
      // layer 2 (only the occupancy channel carries a gradient) in the same pass: second operand pair of k_dw_img
      g.L2 = DwImg{pa.dOimg, 5, 0, 5, 1}; g.C2 = DwImg{pa.Himg, 4, 0, 4, 32};
      g.out2 = a->d_coarse + 2560; g.sl2 = 32; g.sc2 = 1;
      e |= launch_dw_img(g, st);
    } else {
      e |= launch_dw_gemm(dHc, 64, 32, Xst, kIn1, kIn1, Q, nullptr, tiles, nullptr, a->d_coarse, kIn1, 0, st, tc);
      e |= launch_dw_gemm(dOc, kOutP, 1, Hc, 32, 32, Q, nullptr, tiles, nullptr, a->d_coarse + 2560, 32, 0, st, tc);
    }
    if (e) return DNS_ERR_CUDA;
  }
  return DNS_OK;
}

}  // extern "C"
