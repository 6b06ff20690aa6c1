// Point kernels on tcgen05, TWO threads per slot (256-thread CTAs, one 128-slot tile per CTA).
//
// Same mathematics as k_point_fwd / k_point_bwd (render.cu): OneBlob + hash-grid encode, the coarse MLP
// 80 -> 32 -> 33 (models/decoder.py:80-94), the class-expert MLP of the tile (slams/mapping.py:590-601), the latent /
// free-space / opacity loss terms (mapping.py:123-126, utils/common.py:769-802) and their backward down to the
// hash-table scatter.  One CTA = one 128-slot tile = one 128-row MMA tile.  GEMMs (bf16 hi + lo halves, 3 products,
// fp32 accumulation in TMEM):
//
//   fwd   H[128 x 64]    = X[128 x 80]  . [W1 coarse ; W1 expert]^T            K = 80
//         Oc[128 x 48]   = Hc[128 x 32] . W2 coarse^T,  Of likewise            K = 32
//   bwd   dHc[128 x 32]  = dOc[128 x 48] . W2 coarse,   dHf likewise           K = 48
//         dX[128 x 80]   = [dHc | dHf][128 x 64] . [W1 coarse ; W1 expert]     K = 64  (sums both nets)
//
// A first version ran one thread per slot (12 warps per SM) and ncu showed it waiting on the 128 hash-table gathers
// (forward) / 128 vector atomics (backward) each thread issues.  Here the two threads of a slot (tid and tid + 128:
// same TMEM lane quarter, so both may read the slot's accumulator row) split that work:
//
//   forward   group 0: OneBlob of the 3 coordinates + hash-grid levels 0..7;   group 1: levels 8..15
//             hidden epilogue: group g owns net g (coarse / class expert) or half of the single net
//             output epilogue: group 0 coarse row + latent loss, group 1 fine row + free-space / opacity terms
//   backward  group 0: dOut channels 0..23 of both nets, group 1: channels 24..35 (+ the opacity gradient)
//             dH epilogue per net as above;  dX: group 0 OneBlob + levels 0..7, group 1 levels 8..15
//
// so 24 warps per SM are resident with the same shared-memory budget.
#include <stdio.h>

#include "point_tc.cuh"

namespace dns {

constexpr int kTile2 = 2 * kTile;

// params [n][4096] (W1[32][80] | W2[48][32]) -> bf16 hi/lo chunk tiles
// block 0: the coarse net -> wc; block 1 + e: class expert e -> we (one launch for both)
__global__ void k_prep_net80_tc(const float* __restrict__ coarse, uint4* __restrict__ wc, const float* __restrict__ experts,
                                uint4* __restrict__ we) {
  const int e = (int)blockIdx.x - 1;
  const float* p = e < 0 ? coarse : experts + (int64_t)e * 4096;
  uint4* o = e < 0 ? wc : we + (int64_t)e * kNetTc;
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
      split8_f16(a, b, h, l);      // the forward layer-1 GEMM runs on fp16 halves
      o[kNetW1F16 + i] = h;
      o[kNetW1F16 + 320 + i] = l;
    } else {
      o[640 + (i - 320)] = h;
      o[832 + (i - 320)] = l;
    }
  }
}

int prep_nets_tc(const float* coarse, const float* experts, int n_experts, uint4* wc, uint4* we, cudaStream_t st) {
  k_prep_net80_tc<<<1 + ((experts && n_experts > 0) ? n_experts : 0), 128, 0, st>>>(coarse, wc, experts, we);
  return check_launch("prep_nets_tc");
}


template <int MODE>
__global__ void __launch_bounds__(kTile2, 3) k_point_fwd_tc2(PointArgs a, const uint4* __restrict__ wc_all,
                                                             const uint4* __restrict__ we_all) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar, wbar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float red[32];
  // 60 KB: the operand region holds X, then H; the weight region holds W1, then (bulk-copied while the threads
  // run the hidden-layer epilogue) the two W2 tiles.  Two CTAs fit a 132 KB carve-out, which leaves 124 KB of L1
  // to the hash-table gathers.
  unsigned char* X_hi = sm;
  unsigned char* X_lo = sm + kXTile;
  unsigned char* H_hi = sm;                 // aliases the X tile once the first GEMM has completed
  unsigned char* H_lo = sm + 8 * 2048;
  unsigned char* W1_hi = sm + 2 * kXTile;
  unsigned char* W1_lo = W1_hi + kW1Tile;
  unsigned char* W2c_hi = W1_hi;            // hi | lo; aliases W1 once the first GEMM has completed
  unsigned char* W2f_hi = W2c_hi + 2 * kW2Tile;
  const int tile = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, row = tid & (kTile - 1), grp = tid >> 7;
  const int n_tiles = a.perm ? a.counts[cTiles] : a.n_tiles_host;
  if (tile >= n_tiles) return;
  int expert = -1;
  if (MODE == kMap) expert = a.tile_class[tile];
  const bool fine = MODE == kMap && expert >= 0;
  const uint4* we_net = we_all + (int64_t)(fine ? expert : 0) * kNetTc;
  load_w1_tc(W1_hi, W1_lo, wc_all + kNetW1F16, we_net + kNetW1F16, fine);   // fp16 halves
  if (warp == 0) tmem_alloc(&tmem_base_s, 128);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init(&wbar, 1);
  }

  const int64_t q = (int64_t)tile * kTile + row;
  int64_t i, r;
  float zv, x[3];
  const bool valid = slot_point<MODE>(a, q, i, r, zv, x);
  uint4* ximg = (a.need_dparams && !(DNS_DBG(a) & 2)) ? a.Ximg + (int64_t)tile * (20 * kTile) + row : nullptr;
  float4* jimg = (MODE != kTv && a.Jst && a.need_drays)
                     ? reinterpret_cast<float4*>(a.Jst) + ((int64_t)tile * 24 + 12 * grp) * kTile + row : nullptr;
#define XIMG(c) (ximg ? ximg + (c) * kTile : nullptr), (ximg ? ximg + (10 + (c)) * kTile : nullptr)
  if (valid) {
    if (grp == 0) {
#pragma unroll 1   // one copy of the OneBlob + operand-split code instead of three (instruction-cache footprint)
      for (int c = 0; c < 3; ++c) {
        float pe[16];
        oneblob16(c == 0 ? x[0] : (c == 1 ? x[1] : x[2]), pe);
        put_chunk_f16_img(X_hi, X_lo, 2 * c, 2048, row, pe, XIMG(2 * c));
        put_chunk_f16_img(X_hi, X_lo, 2 * c + 1, 2048, row, pe + 8, XIMG(2 * c + 1));
      }
    }
    hashgrid_fwd_to_tile(a.G, a.table, x, X_hi, X_lo, row, 8 * grp, jimg);
    if (ximg) {   // the two grid chunks this thread has just written (its own row): fp16 tile -> bf16 global image
#pragma unroll
      for (int c = 6 + 2 * grp; c < 8 + 2 * grp; ++c) {
        float v[8];
        f16_chunk_to_floats(X_hi, X_lo, c, 2048, row, v);
        store_chunk_img(v, ximg + c * kTile, ximg + (10 + c) * kTile);
      }
    }
  } else {
    const uint4 z4 = make_uint4(0, 0, 0, 0);
    const int c0 = grp ? 8 : 0, c1 = grp ? 10 : 8;
    for (int c = c0; c < c1; ++c) {
      *reinterpret_cast<uint4*>(X_hi + c * 2048 + row * 16) = z4;
      *reinterpret_cast<uint4*>(X_lo + c * 2048 + row * 16) = z4;
      if (ximg) ximg[c * kTile] = ximg[(10 + c) * kTile] = z4;
    }
  }
#undef XIMG
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  constexpr int NH = MODE == kMap ? 64 : 32;  // hidden units computed per point
  if (tid == 0) {  // H = X . W1^T  (fp16 hi / lo halves)
    const uint32_t idesc = umma_idesc_f16(128, NH, 0, 0, 0, 0);
#pragma unroll 1
    for (int ks = 0; ks < 5; ++ks) {
      const uint32_t aoff = ks * 4096, boff = ks * 2048;
      const uint64_t a_hi = umma_desc(smem_u32(X_hi) + aoff, 2048, 128), a_lo = umma_desc(smem_u32(X_lo) + aoff, 2048, 128);
      const uint64_t b_hi = umma_desc(smem_u32(W1_hi) + boff, 1024, 128), b_lo = umma_desc(smem_u32(W1_lo) + boff, 1024, 128);
      umma_bf16(tmem_d, a_hi, b_hi, idesc, ks > 0 ? 1u : 0u);
      umma_bf16(tmem_d, a_lo, b_hi, idesc, 1u);
      umma_bf16(tmem_d, a_hi, b_lo, idesc, 1u);
    }
    umma_commit(&bar);
  }
  mbar_wait_cta(&bar, 0);
  tc_fence_after();
  if (tid == 0) {   // W1 has been consumed: fetch the layer-2 weights over it (W2 hi | lo are contiguous per net)
    mbar_expect_tx(&wbar, (fine ? 2u : 1u) * 2u * kW2Tile);
    bulk_g2s(W2c_hi, wc_all + 640, 2 * kW2Tile, &wbar);
    if (fine) bulk_g2s(W2f_hi, we_net + 640, 2 * kW2Tile, &wbar);
  }
  const uint32_t lane_addr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
  {
    // hidden activations of this thread's share (NH/2 units): next A operand + global tile image (ReLU mask, dW2)
    constexpr int HS = NH / 2;
    uint4* himg = (DNS_DBG(a) & 2) ? nullptr : a.Himg + (int64_t)tile * (2 * (NH / 8) * kTile) + row;
#pragma unroll
    for (int g4 = 0; g4 < HS / 16; ++g4) {
      float v[16];
      tmem_ld16(lane_addr + grp * HS + 16 * g4, v);
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = fmaxf(v[k], 0.f);
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const int c = grp * (HS / 8) + 2 * g4 + c2;
        put_chunk_img(H_hi, H_lo, c, 2048, row, v + 8 * c2, himg ? himg + c * kTile : nullptr,
                      himg ? himg + (NH / 8 + c) * kTile : nullptr);
      }
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();   // accumulator rows read, H tile written: D columns and the X region are free again
  if (tid == 0) {    // O = H . W2^T per net
    tc_fence_after();
    mbar_wait(&wbar, 0);
    const uint32_t idesc = umma_idesc_bf16(128, 48, 0, 0);
    for (int net = 0; net < (fine ? 2 : 1); ++net) {
      const unsigned char* Wh = net ? W2f_hi : W2c_hi;
#pragma unroll 1
      for (int ks = 0; ks < 2; ++ks) {
        const uint32_t aoff = net * 4 * 2048 + ks * 4096, boff = ks * 2 * 768;
        const uint64_t a_hi = umma_desc(smem_u32(H_hi) + aoff, 2048, 128), a_lo = umma_desc(smem_u32(H_lo) + aoff, 2048, 128);
        const uint64_t b_hi = umma_desc(smem_u32(Wh) + boff, 768, 128), b_lo = umma_desc(smem_u32(Wh + kW2Tile) + boff, 768, 128);
        umma_bf16(tmem_d + net * 48, a_hi, b_hi, idesc, ks > 0 ? 1u : 0u);
        umma_bf16(tmem_d + net * 48, a_lo, b_hi, idesc, 1u);
        umma_bf16(tmem_d + net * 48, a_hi, b_lo, idesc, 1u);
      }
    }
    umma_commit(&bar);
  }
  mbar_wait_cta(&bar, 1);
  tc_fence_after();
  float lt = 0.f, fs = 0.f, op = 0.f;
  if (MODE == kTv) {
    if (grp == 0) {
      float v[16];
      tmem_ld16(lane_addr, v);
      if (valid) a.occ[q] = v[0];
    }
  } else if (MODE == kTrack) {
    // the 36-float row: group 0 channels 0..15, group 1 channels 16..35
    float v[16];
    tmem_ld16(lane_addr + 16 * grp, v);
    float4* d4 = valid ? reinterpret_cast<float4*>(a.fine36 + (a.p0 + i) * kOutP + 16 * grp) : nullptr;
    if (valid) {
#pragma unroll
      for (int k = 0; k < 4; ++k) d4[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    }
    if (grp == 1) {
      tmem_ld16(lane_addr + 32, v);
      if (valid) d4[4] = make_float4(v[0], v[1], v[2], v[3]);
    }
  } else {
    // group 0: coarse row + latent loss;  group 1: fine row + free-space / opacity terms.  With the slot-order hand-over
    // group 0 stores  coarse - fine  in the row of its SLOT (what the backward needs of both) instead of the coarse row
    float* dst = nullptr;
    if (valid) {
      if (grp) dst = a.fine36 + (a.p0 + i) * kOutP;
      else if (a.want_coarse_pt) dst = a.coarse36 + (a.p0 + i) * kOutP;
    }
    float4* dslot = (valid && grp == 0 && a.diff36s) ? reinterpret_cast<float4*>(a.diff36s) : nullptr;   // image, see slot_img
    float fo32 = 0.f;
#pragma unroll
    for (int g4 = 0; g4 < 3; ++g4) {
      float vc[16], vf[16];
      if (grp == 0) tmem_ld16(lane_addr + 16 * g4, vc);       // `grp` is uniform over the warp
      if (fine) {
        tmem_ld16(lane_addr + 48 + 16 * g4, vf);
      } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) vf[k] = 0.f;
      }
      if (valid) {
        const int n4 = g4 < 2 ? 4 : 1;                          // 36 = 16 + 16 + 4 floats
        if (grp == 0) {
          if (dst) {
            float4* d4 = reinterpret_cast<float4*>(dst + 16 * g4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < n4) d4[k] = make_float4(vc[4 * k], vc[4 * k + 1], vc[4 * k + 2], vc[4 * k + 3]);
          }
          float dv[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            dv[k] = vc[k] - vf[k];
            if (16 * g4 + k < DNS_LATENT) lt = fmaf(dv[k], dv[k], lt);
          }
          if (dslot) {
            if (g4 < 2) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                dslot[slot_img(q, 4 * g4 + k)] = make_float4(dv[4 * k], dv[4 * k + 1], dv[4 * k + 2], dv[4 * k + 3]);
            } else {
              reinterpret_cast<float*>(dslot + slot_img(q, 8))[0] = dv[0];   // channel 33 belongs to group 1 (fine channel 32)
            }
          }
        } else {
          float4* d4 = reinterpret_cast<float4*>(dst + 16 * g4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < n4) d4[k] = make_float4(vf[4 * k], vf[4 * k + 1], vf[4 * k + 2], vf[4 * k + 3]);
          if (g4 == 2) fo32 = vf[0];
        }
      }
    }
    if (valid && grp == 1 && a.diff36s) reinterpret_cast<float*>(reinterpret_cast<float4*>(a.diff36s) + slot_img(q, 8))[1] = fo32;
    if (valid && grp == 1) {
      float front, band, vd, d = a.gt_depth[r];
      opacity_masks(zv, d, a.trunc, front, band, vd);
      float o = sigmoidf_(10.f * fo32);
      float t1 = o * front * vd;
      fs = t1 * t1;
      float u = (zv - d) / a.sigma;
      float ps = 0.5f * __expf(-0.5f * u * u);
      float e = o * band - ps * band;
      op = e * e;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 128);
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

template <int MODE>
// Three CTAs per SM (<= 85 registers; 86 % carve-out for 3 x 64 KB): since the Jacobian image took the L1-hungry corner
// re-read out of the mapping backward, occupancy beats L1 here -- 1024 x 96 tracking 0.069 -> 0.060 ms, 4096 x 47 mapping
// 0.220 -> 0.207 ms, 131 072 x 47 mapping 5.35 -> 5.20 ms (round 1, with the re-read: 7.6 -> 9.6 ms).
#ifndef DNS_BWD_CTAS
#define DNS_BWD_CTAS 3
#endif
#ifndef DNS_BWD_CARVE_PCT
#define DNS_BWD_CARVE_PCT 86
#endif
__global__ void __launch_bounds__(kTile2, MODE == kTv ? 4 : DNS_BWD_CTAS) k_point_bwd_tc2(PointArgs a, const uint4* __restrict__ wc_all,
                                                             const uint4* __restrict__ we_all) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  // 64 KB.  Phase 1: dOut tiles coarse hi | lo | fine hi | lo (six chunks each, 48 KB) and the W2 tiles.  Phase 2
  // (after the first GEMM): the dH tile [8 chunks hi | lo, 32 KB] and the combined W1 tile (20 KB, prefetched into
  // registers while the GEMM runs) take the place of the dOut tiles.  Three CTAs fit the 196 KB carve-out (DNS_BWD_CTAS
  // above); tracking, which still re-reads the corners for dL/dx, keeps what L1 is left.
  // Single-net modes (tracking, TV) need 36 KB: dOut hi | lo (24 KB) with W2 hi | lo behind them, then dH hi | lo (16 KB) and
  // W1 (20 KB, stored once the first GEMM has consumed dOut and W2) -- four TV CTAs per SM (64 registers, 128 TMEM columns).
  constexpr bool kTwoNets = MODE == kMap;
  unsigned char* DOc_hi = sm;
  unsigned char* DOc_lo = sm + kDOTile;
  unsigned char* DOf_hi = sm + 2 * kDOTile;
  unsigned char* DOf_lo = sm + 3 * kDOTile;
  unsigned char* DH_hi = sm;                        // dH tile [8 | 4 chunks] aliases the dOut tiles
  unsigned char* DH_lo = sm + (kTwoNets ? 8 : 4) * 2048;
  unsigned char* W1_hi = sm + (kTwoNets ? 16 : 8) * 2048;   // behind the dH tile, still inside the (dead) dOut tiles
  unsigned char* W1_lo = W1_hi + kW1Tile;
  unsigned char* W2c_hi = kTwoNets ? sm + 16 * 2048 + 2 * kW1Tile : sm + 2 * kDOTile;
  unsigned char* W2f_hi = W2c_hi + 2 * kW2Tile;
  float* DXS = reinterpret_cast<float*>(sm);        // [3][128] partial d/dx of group 1 (after the last GEMM)
  const int tile = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, row = tid & (kTile - 1), grp = tid >> 7;
  const int n_tiles = a.perm ? a.counts[cTiles] : a.n_tiles_host;
  if (tile >= n_tiles) return;
  DNS_CLK_DECL
  int expert = -1;
  if (MODE == kMap) expert = a.tile_class[tile];
  const bool fine = MODE == kMap && expert >= 0;
  const uint4* we_net = we_all + (int64_t)(fine ? expert : 0) * kNetTc;
  for (int k = tid; k < 384; k += kTile2) {   // W2 coarse / expert, hi | lo contiguous per net
    reinterpret_cast<uint4*>(W2c_hi)[k] = wc_all[640 + k];
    if (kTwoNets) reinterpret_cast<uint4*>(W2f_hi)[k] = fine ? we_net[640 + k] : make_uint4(0, 0, 0, 0);
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 128);
  if (tid == 0) mbar_init(&bar, 1);

  DNS_CLK(a, 0)   // W2 tile loads issued, TMEM allocation
  const int64_t q = (int64_t)tile * kTile + row;
  int64_t i, r;
  float zv, x[3];
  const bool valid = slot_point<MODE>(a, q, i, r, zv, x);
  DNS_CLK(a, 1)   // perm -> ray -> point
  constexpr int DOCH = MODE == kMap ? 10 : 5;   // chunks per half of the dOut image: 5 (40 >= 33 channels) per net
  constexpr int HCH = MODE == kMap ? 8 : 4;     // chunks per half of the H / dH images
  const bool stash = a.need_dparams && !(DNS_DBG(a) & 2);
  uint4* doimg = stash ? a.dOimg + (int64_t)tile * (2 * DOCH * kTile) + row : nullptr;
  // ---- gradients w.r.t. the MLP outputs (rows of 48: 33 used).  Group 0 builds channel chunks 0..2, group 1 chunks 3..5
  {
    const int ch0 = 24 * grp;
    float dc[24], df[24];
#pragma unroll
    for (int k = 0; k < 24; ++k) dc[k] = df[k] = 0.f;
    if (MODE == kTv) {
      if (valid && grp == 0) dc[0] = a.docc[q];
    } else if (MODE == kTrack) {
      if (valid) {
        const float4* s4 = reinterpret_cast<const float4*>(a.dfine36 + (a.p0 + i) * kOutP + ch0);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          if (ch0 + 4 * k < kOutP) {
            float4 v = s4[k];
            dc[4 * k] = v.x; dc[4 * k + 1] = v.y; dc[4 * k + 2] = v.z; dc[4 * k + 3] = v.w;
          }
        }
      }
    } else if (valid) {
      const float g_lt = 2.f * a.lam_lt / (33.f * (float)a.P_total);
      float fo32 = 0.f;
      if (a.dfine36s) {
        // slot-order rows (written by the ray kernel and the forward kernel): unit stride, no permutation chase
        const float4* sd = reinterpret_cast<const float4*>(a.dfine36s);
        const float4* sx = reinterpret_cast<const float4*>(a.diff36s);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          if (ch0 + 4 * k < kOutP) {
            const float4 d = sd[slot_img(q, 6 * grp + k)], x = sx[slot_img(q, 6 * grp + k)];
            const float dd[4] = {d.x, d.y, d.z, d.w}, xx[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int ch = ch0 + 4 * k + e;
              if (ch < DNS_LATENT) {
                const float gl = g_lt * xx[e];
                dc[4 * k + e] = gl;            // coarse net: only the latent loss reaches it in mapping
                df[4 * k + e] = dd[e] - gl;
              } else if (ch == 33) {
                fo32 = xx[e];
              }
            }
          }
        }
      } else {
      const int64_t p = a.p0 + i;
      const float4* sd = reinterpret_cast<const float4*>(a.dfine36 + p * kOutP + ch0);
      const float4* sc = reinterpret_cast<const float4*>(a.coarse36 + p * kOutP + ch0);
      const float4* sf = reinterpret_cast<const float4*>(a.fine36 + p * kOutP + ch0);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        if (ch0 + 4 * k < kOutP) {
          float4 d = sd[k], c = sc[k], f = sf[k];
          const float dd[4] = {d.x, d.y, d.z, d.w}, cc[4] = {c.x, c.y, c.z, c.w}, ff[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int ch = ch0 + 4 * k + e;
            if (ch < DNS_LATENT) {
              const float gl = g_lt * (cc[e] - ff[e]);
              dc[4 * k + e] = gl;            // coarse net: only the latent loss reaches it in mapping
              df[4 * k + e] = dd[e] - gl;
              if (ch == 32) fo32 = ff[e];
            }
          }
        }
      }
      }
      if (grp == 1 && a.counts[cFront] > 0 && a.counts[cBand] > 0) {
        float front, band, vd, d = a.gt_depth[r];
        opacity_masks(zv, d, a.trunc, front, band, vd);
        float o = sigmoidf_(10.f * fo32);
        float u = (zv - d) / a.sigma;
        float ps = 0.5f * __expf(-0.5f * u * u);
        float inv_p = 1.f / (float)a.P_total;
        float d_o = 2.f * a.lam_fs * inv_p * o * front * vd + 2.f * a.lam_op * inv_p * (o - ps) * band;
        df[32 - 24] += d_o * 10.f * o * (1.f - o);
      }
      if (!fine) {
#pragma unroll
        for (int k = 0; k < 24; ++k) df[k] = 0.f;
      }
    }
#pragma unroll
    for (int c3 = 0; c3 < 3; ++c3) {
      const int c = 3 * grp + c3;
      const bool gi = stash && c < 5;   // channels 40..47 are zeros: not part of the global image
      put_chunk_img(DOc_hi, DOc_lo, c, 2048, row, dc + 8 * c3, gi ? doimg + c * kTile : nullptr,
                    gi ? doimg + (DOCH + c) * kTile : nullptr);
      if (MODE == kMap)
        put_chunk_img(DOf_hi, DOf_lo, c, 2048, row, df + 8 * c3, gi ? doimg + (5 + c) * kTile : nullptr,
                      gi ? doimg + (DOCH + 5 + c) * kTile : nullptr);
    }
  }
  DNS_CLK(a, 2)   // dOut rows built
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  DNS_CLK(a, 3)   // first CTA barrier
  const uint32_t tmem_d = tmem_base_s;
  if (tid == 0) {  // dH = dOut . W2   (B = W2 tile MN-major: hidden contiguous; LBO 128 over out rows, SBO 768 over hidden chunks)
    const uint32_t idesc = umma_idesc_bf16(128, 32, 0, 1);
    for (int net = 0; net < (fine ? 2 : 1); ++net) {
      const unsigned char* Ah = net ? DOf_hi : DOc_hi;
      const unsigned char* Al = net ? DOf_lo : DOc_lo;
      const unsigned char* Wh = net ? W2f_hi : W2c_hi;
#pragma unroll 1
      for (int ks = 0; ks < 3; ++ks) {
        const uint32_t aoff = ks * 4096, boff = ks * 256;
        const uint64_t a_hi = umma_desc(smem_u32(Ah) + aoff, 2048, 128), a_lo = umma_desc(smem_u32(Al) + aoff, 2048, 128);
        const uint64_t b_hi = umma_desc(smem_u32(Wh) + boff, 128, 768), b_lo = umma_desc(smem_u32(Wh + kW2Tile) + boff, 128, 768);
        umma_bf16(tmem_d + net * 32, a_hi, b_hi, idesc, ks > 0 ? 1u : 0u);
        umma_bf16(tmem_d + net * 32, a_lo, b_hi, idesc, 1u);
        umma_bf16(tmem_d + net * 32, a_hi, b_lo, idesc, 1u);
      }
    }
    umma_commit(&bar);
  }
  // combined W1 tile (rows 0..31 coarse, 32..63 expert per feature chunk; hi then lo): 1280 elements, five per
  // thread, fetched while the GEMM runs and stored once the dOut tiles are dead
  uint4 w1r[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int e = tid + k * kTile2, half = e >= 640 ? 1 : 0, idx = e - 640 * half, c = idx >> 6, j = idx & 63;
    w1r[k] = j < 32 ? wc_all[320 * half + c * 32 + j]
                    : (fine ? we_net[320 * half + c * 32 + j - 32] : make_uint4(0, 0, 0, 0));
  }
  mbar_wait_cta(&bar, 0);
  tc_fence_after();
  DNS_CLK(a, 4)   // GEMM dH issued + W1 prefetch + wait
#pragma unroll
  for (int k = 0; k < 5; ++k) reinterpret_cast<uint4*>(W1_hi)[tid + k * kTile2] = w1r[k];   // W1_lo follows W1_hi
  const uint32_t lane_addr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
  {
    // dH of this thread's share, masked by the ReLU: the bf16 hi half of the stashed activation is non-zero exactly
    // where the activation was positive.  MAP: group = net (32 units); otherwise half of the single net (16 units).
    constexpr int HS = MODE == kMap ? 32 : 16;
    const bool on = MODE != kMap || grp == 0 || fine;
    const uint4* himg = a.Himg + (int64_t)tile * (2 * HCH * kTile) + row;
    uint4* dhimg = stash ? a.dHimg + (int64_t)tile * (2 * HCH * kTile) + row : nullptr;
#pragma unroll
    for (int g2 = 0; g2 < HS / 16; ++g2) {
      float v[16];
      if (on) tmem_ld16(lane_addr + grp * HS + 16 * g2, v);   // `on` is uniform over the warp
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const int c = grp * (HS / 8) + 2 * g2 + c2;
        const uint4 hv = on ? himg[c * kTile] : make_uint4(0, 0, 0, 0);
        const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
        float dh[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) dh[e] = (on && ((hw[e >> 1] >> (16 * (e & 1))) & 0x7fffu)) ? v[8 * c2 + e] : 0.f;
        put_chunk_img(DH_hi, DH_lo, c, 2048, row, dh, dhimg ? dhimg + c * kTile : nullptr,
                      dhimg ? dhimg + (HCH + c) * kTile : nullptr);
      }
    }
  }
  DNS_CLK(a, 5)   // dH epilogue (H image read, ReLU mask, tile + image stores)
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  DNS_CLK(a, 6)   // second CTA barrier
  if (tid == 0) {  // dX = [dHc | dHf] . [W1c ; W1f]   (B = combined W1 tile MN-major: LBO 128 over hidden rows, SBO 1024)
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, 80, 0, 1);
#pragma unroll 1
    for (int ks = 0; ks < (MODE == kMap ? 4 : 2); ++ks) {
      const uint32_t aoff = ks * 4096, boff = ks * 256;
      const uint64_t a_hi = umma_desc(smem_u32(DH_hi) + aoff, 2048, 128), a_lo = umma_desc(smem_u32(DH_lo) + aoff, 2048, 128);
      const uint64_t b_hi = umma_desc(smem_u32(W1_hi) + boff, 128, 1024), b_lo = umma_desc(smem_u32(W1_lo) + boff, 128, 1024);
      umma_bf16(tmem_d, a_hi, b_hi, idesc, ks > 0 ? 1u : 0u);
      umma_bf16(tmem_d, a_lo, b_hi, idesc, 1u);
      umma_bf16(tmem_d, a_hi, b_lo, idesc, 1u);
    }
    umma_commit(&bar);
  }
  mbar_wait_cta(&bar, 1);
  tc_fence_after();
  DNS_CLK(a, 7)   // GEMM dX + wait
  // dX columns: 0..47 OneBlob (group 0), 48..63 levels 0..7 (group 0), 64..79 levels 8..15 (group 1)
  float dx[3] = {0.f, 0.f, 0.f}, dg[16];
  if (grp == 0) {
    if (a.need_drays) {
#pragma unroll 1   // one copy of the OneBlob backward instead of three (instruction-cache footprint)
      for (int g5 = 0; g5 < 3; ++g5) {
        float v[16];
        tmem_ld16(lane_addr + 16 * g5, v);
        const float d = valid ? oneblob16_bwd(g5 == 0 ? x[0] : (g5 == 1 ? x[1] : x[2]), v) : 0.f;
        dx[0] = g5 == 0 ? d : dx[0];
        dx[1] = g5 == 1 ? d : dx[1];
        dx[2] = g5 == 2 ? d : dx[2];
      }
    }
    tmem_ld16(lane_addr + 48, dg);
  } else {
    tmem_ld16(lane_addr + 64, dg);
  }
  tc_fence_before();
  __syncthreads();   // every accumulator row has been read; the operand tiles are free (DXS aliases them)
  if (warp == 0) tmem_dealloc(tmem_d, 128);
  DNS_CLK(a, 8)   // dX read back, OneBlob backward, third CTA barrier, TMEM release
  float dxg[3] = {0.f, 0.f, 0.f};
  float2* dtab = (a.need_dparams && !(DNS_DBG(a) & 4) && !(DNS_DBG(a) & (grp ? 32 : 16))) ? a.d_table : nullptr;
  const bool want_dx = a.need_drays != 0 && !(DNS_DBG(a) & 8);
  // dL/dx through the grid: from the forward pass's Jacobian image when there is one (12 coalesced loads per thread),
  // else by re-reading the corners
  const bool from_j = MODE != kTv && a.Jst != nullptr;
  // dg leaves the registers for the thread's column of a shared-memory staging area [16][256] behind DXS (the operand tiles
  // are dead): the level loops below are rolled (code size, see hashgrid_bwd_levels)
  float* dgs = reinterpret_cast<float*>(sm) + 3 * kTile + tid;
#pragma unroll
  for (int k = 0; k < 16; ++k) dgs[k * kTile2] = valid ? dg[k] : 0.f;
  float2* dpriv = a.d_priv ? a.d_priv + (size_t)(blockIdx.x % a.priv_copies) * a.priv_end : nullptr;
  const int pl = a.d_priv ? a.priv_levels : 0;
  if (MODE == kTv && a.tv_agg_levels > 0) {
    // coherent lattice: per-cell pre-reduction inside the warp (every lane takes part in the shuffles)
    hashgrid_bwd_rows(a.G, dtab, x, dgs, kTile2, 8 * grp, 8 * grp + 8, valid, q / a.n, a.tv_agg_levels, dpriv, pl);
  } else if (valid) {
    float dxj[3] = {0.f, 0.f, 0.f};
    if (want_dx && from_j)    // issued before the reductions: the 12 loads fly while those drain
      hashgrid_dx_from_jimg<8>(reinterpret_cast<const float4*>(a.Jst) + ((int64_t)tile * 24 + 12 * grp) * kTile + row, dg, dxj);
    hashgrid_bwd_levels(a.G, a.table, dtab, x, dgs, kTile2, 8 * grp, 8 * grp + 8, want_dx && !from_j, dxg, dpriv, pl);
    if (want_dx && from_j) {
#pragma unroll
      for (int c = 0; c < 3; ++c) dxg[c] = dxj[c];
    }
  }
  DNS_CLK(a, 9)   // hash-grid backward of thread 0 (levels 0..7)
  if (a.need_drays && MODE != kTv) {
    if (grp == 1) {
#pragma unroll
      for (int c = 0; c < 3; ++c) DXS[c * kTile + row] = dxg[c];
    }
    __syncthreads();
    if (grp == 0 && valid) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float g = (dx[c] + dxg[c] + DXS[c * kTile + row]) / (float)a.B.ext[c];
        atomicAdd(a.d_rays_o + 3 * r + c, g);
        atomicAdd(a.d_rays_d + 3 * r + c, g * zv);
      }
    }
  }
  DNS_CLK(a, 10)   // ray gradients
}

size_t point_bwd_tc2_smem(int mode) { return mode == kMap ? 16 * 2048 + 2 * kW1Tile + 4 * kW2Tile : 8 * 2048 + 2 * kW1Tile; }
size_t point_fwd_tc2_smem() { return 2 * kXTile + 2 * kW1Tile; }

static void set_attrs2() {
  static unsigned long long seen = 0;
  if (!first_call_on_device(seen)) return;
  const int f = (int)point_fwd_tc2_smem(), b = (int)point_bwd_tc2_smem(kMap);
  cudaFuncSetAttribute(k_point_fwd_tc2<kTrack>, cudaFuncAttributeMaxDynamicSharedMemorySize, f);
  cudaFuncSetAttribute(k_point_fwd_tc2<kMap>, cudaFuncAttributeMaxDynamicSharedMemorySize, f);
  cudaFuncSetAttribute(k_point_fwd_tc2<kTv>, cudaFuncAttributeMaxDynamicSharedMemorySize, f);
  cudaFuncSetAttribute(k_point_bwd_tc2<kTrack>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
  cudaFuncSetAttribute(k_point_bwd_tc2<kMap>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
  cudaFuncSetAttribute(k_point_bwd_tc2<kTv>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
}

// L1 beats occupancy for the gathers: the coarse hash-grid levels live in L1, and a carve-out that fits three CTAs
// (228 KB) leaves only 28 KB of it.  Measured at 131 072 rays x 47: point_fwd (60 KB per CTA) 86 % -> 4.89 ms,
// 72 % -> 3.80, 58 % -> 4.11; point_bwd (64 KB per CTA) 58 % -> 6.79, 72 % -> 7.0, 86 % -> 9.9.  Changing the
// carve-out between consecutive kernels costs a reconfiguration, which shows at SLAM-iteration sizes (mapping
// iteration 2.37 -> 2.60 ms), so the preference is only set for large launches (-DDNS_ABLATE builds read
// DNS_FWD_CARVE / DNS_BWD_CARVE).
struct CarveState {
  int last[16];   // per device
  CarveState() { for (int i = 0; i < 16; ++i) last[i] = -2; }
};
template <typename K>
static void prefer_carveout(K kernel, CarveState& cs, int tiles, const char* env, int tuned) {
  static const int kLargeTiles = 8192;
#ifdef DNS_ABLATE
  if (const char* e = getenv(env)) tuned = atoi(e);
#else
  (void)env;
#endif
  const int want = tiles >= kLargeTiles ? tuned : -1;   // -1: cudaSharedmemCarveoutDefault
  int& last = cs.last[current_device_slot()];
  if (want != last) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, want);
    last = want;
  }
}

int launch_point_fwd_tc(int mode, const PointArgs& pa, int tiles, const uint4* wc, const uint4* we, cudaStream_t st) {
  set_attrs2();
  const size_t smem = point_fwd_tc2_smem();
  static CarveState last[3];
  if (mode == kMap) prefer_carveout(k_point_fwd_tc2<kMap>, last[0], tiles, "DNS_FWD_CARVE", 72);
  else if (mode == kTrack) prefer_carveout(k_point_fwd_tc2<kTrack>, last[1], tiles, "DNS_FWD_CARVE", 72);
  else prefer_carveout(k_point_fwd_tc2<kTv>, last[2], tiles, "DNS_FWD_CARVE", 72);
  if (mode == kMap) k_point_fwd_tc2<kMap><<<tiles, kTile2, smem, st>>>(pa, wc, we);
  else if (mode == kTrack) k_point_fwd_tc2<kTrack><<<tiles, kTile2, smem, st>>>(pa, wc, we);
  else k_point_fwd_tc2<kTv><<<tiles, kTile2, smem, st>>>(pa, wc, we);
  return check_launch("point_fwd_tc2");
}
int launch_point_bwd_tc(int mode, const PointArgs& pa, int tiles, const uint4* wc, const uint4* we, cudaStream_t st) {
  set_attrs2();
  const size_t smem = point_bwd_tc2_smem(mode);
  static CarveState last[3];
  if (mode == kMap) prefer_carveout(k_point_bwd_tc2<kMap>, last[0], tiles, "DNS_BWD_CARVE", DNS_BWD_CARVE_PCT);
  else if (mode == kTrack) prefer_carveout(k_point_bwd_tc2<kTrack>, last[1], tiles, "DNS_BWD_CARVE", DNS_BWD_CARVE_PCT);
  else prefer_carveout(k_point_bwd_tc2<kTv>, last[2], tiles, "DNS_TV_BWD_CARVE", 72);   // no gathers: room for four CTAs
  if (mode == kMap) k_point_bwd_tc2<kMap><<<tiles, kTile2, smem, st>>>(pa, wc, we);
  else if (mode == kTrack) k_point_bwd_tc2<kTrack><<<tiles, kTile2, smem, st>>>(pa, wc, we);
  else k_point_bwd_tc2<kTv><<<tiles, kTile2, smem, st>>>(pa, wc, we);
  return check_launch("point_bwd_tc2");
}

}  // namespace dns
