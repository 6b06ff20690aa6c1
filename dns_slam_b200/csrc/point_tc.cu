// Point kernels with the MLP GEMMs on tcgen05 (sm_100a).
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

template <int MODE>
__global__ void __launch_bounds__(kTile) k_point_fwd_tc(PointArgs a, const uint4* __restrict__ wc_all,
                                                         const uint4* __restrict__ we_all) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float red[32];
  unsigned char* X_hi = sm;
  unsigned char* X_lo = sm + kXTile;
  unsigned char* H_hi = sm;                 // aliases the X tile once the first GEMM has completed
  unsigned char* H_lo = sm + 8 * 2048;
  unsigned char* W1_hi = sm + 2 * kXTile;
  unsigned char* W1_lo = W1_hi + kW1Tile;
  unsigned char* W2c_hi = W1_lo + kW1Tile;  // hi | lo
  unsigned char* W2f_hi = W2c_hi + 2 * kW2Tile;
  const int tile = blockIdx.x, tid = threadIdx.x, warp = tid >> 5;
  const int n_tiles = a.perm ? a.counts[cTiles] : a.n_tiles_host;
  if (tile >= n_tiles) return;
  int expert = -1;
  if (MODE == kMap) expert = a.tile_class[tile];
  const bool fine = MODE == kMap && expert >= 0;
  load_weights_tc(W1_hi, W1_lo, W2c_hi, W2f_hi, wc_all, we_all + (int64_t)(fine ? expert : 0) * kNetTc, fine);
  if (warp == 0) tmem_alloc(&tmem_base_s, 128);
  if (tid == 0) mbar_init(&bar, 1);

  const int64_t q = (int64_t)tile * kTile + tid;
  int64_t i, r;
  float zv, x[3];
  const bool valid = slot_point<MODE>(a, q, i, r, zv, x);
  // encoded inputs of the tile: MMA operand in shared memory + (for the weight gradients) a global tile image
  uint4* ximg = (a.need_dparams && !(a.dbg & 2)) ? a.Ximg + (int64_t)tile * (20 * kTile) + tid : nullptr;
#define XIMG(c) (ximg ? ximg + (c) * kTile : nullptr), (ximg ? ximg + (10 + (c)) * kTile : nullptr)
  if (valid) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float pe[16];
      oneblob16(x[c], pe);
      put_chunk_img(X_hi, X_lo, 2 * c, 2048, tid, pe, XIMG(2 * c));
      put_chunk_img(X_hi, X_lo, 2 * c + 1, 2048, tid, pe + 8, XIMG(2 * c + 1));
    }
    float g[32];
    if (a.dbg & 1) {
#pragma unroll
      for (int k = 0; k < 32; ++k) g[k] = x[k % 3] * 0.01f * k;
    } else
    hashgrid_fwd_regs(a.G, a.table, x, g);
#pragma unroll
    for (int c = 0; c < 4; ++c) put_chunk_img(X_hi, X_lo, 6 + c, 2048, tid, g + 8 * c, XIMG(6 + c));
  } else {
    const uint4 z4 = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int c = 0; c < 10; ++c) {
      *reinterpret_cast<uint4*>(X_hi + c * 2048 + tid * 16) = z4;
      *reinterpret_cast<uint4*>(X_lo + c * 2048 + tid * 16) = z4;
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
  if (tid == 0) {  // H = X . W1^T
    const uint32_t idesc = umma_idesc_bf16(128, NH, 0, 0);
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
  mbar_wait(&bar, 0);
  tc_fence_after();
  const uint32_t lane_addr = tmem_d + ((uint32_t)(warp * 32) << 16);
  {
    float h[NH];
#pragma unroll
    for (int g4 = 0; g4 < NH / 16; ++g4) {
      float v[16];
      tmem_ld16(lane_addr + 16 * g4, v);
#pragma unroll
      for (int k = 0; k < 16; ++k) h[16 * g4 + k] = fmaxf(v[k], 0.f);
    }
    // hidden activations: the next A operand, and a global tile image (ReLU mask + dW2 of the backward)
    uint4* himg = (a.dbg & 2) ? nullptr : a.Himg + (int64_t)tile * (2 * (NH / 8) * kTile) + tid;
#pragma unroll
    for (int c = 0; c < NH / 8; ++c)
      put_chunk_img(H_hi, H_lo, c, 2048, tid, h + 8 * c, himg ? himg + c * kTile : nullptr,
                    himg ? himg + (NH / 8 + c) * kTile : nullptr);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();   // accumulator rows read, H tile written: D columns and the X region are free again
  if (tid == 0) {    // O = H . W2^T per net
    tc_fence_after();
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
  mbar_wait(&bar, 1);
  tc_fence_after();
  float out[48];
#pragma unroll
  for (int g4 = 0; g4 < 3; ++g4) {
    float v[16];
    tmem_ld16(lane_addr + 16 * g4, v);
#pragma unroll
    for (int k = 0; k < 16; ++k) out[16 * g4 + k] = v[k];
  }
  float lt = 0.f, fs = 0.f, op = 0.f;
  if (MODE == kTv) {
    if (valid) a.occ[q] = out[0];
  } else if (MODE == kTrack) {
    if (valid) {
      float4* d4 = reinterpret_cast<float4*>(a.fine36 + (a.p0 + i) * kOutP);
#pragma unroll
      for (int k = 0; k < kOutP / 4; ++k) d4[k] = make_float4(out[4 * k], out[4 * k + 1], out[4 * k + 2], out[4 * k + 3]);
    }
  } else {
    float fo[48];
    if (fine) {
#pragma unroll
      for (int g4 = 0; g4 < 3; ++g4) {
        float v[16];
        tmem_ld16(lane_addr + 48 + 16 * g4, v);
#pragma unroll
        for (int k = 0; k < 16; ++k) fo[16 * g4 + k] = v[k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < 48; ++k) fo[k] = 0.f;
    }
    if (valid) {
      float4* dc = reinterpret_cast<float4*>(a.coarse36 + (a.p0 + i) * kOutP);
      float4* df = reinterpret_cast<float4*>(a.fine36 + (a.p0 + i) * kOutP);
#pragma unroll
      for (int k = 0; k < kOutP / 4; ++k) {
        dc[k] = make_float4(out[4 * k], out[4 * k + 1], out[4 * k + 2], out[4 * k + 3]);
        df[k] = make_float4(fo[4 * k], fo[4 * k + 1], fo[4 * k + 2], fo[4 * k + 3]);
      }
#pragma unroll
      for (int c = 0; c < DNS_LATENT; ++c) {
        float d = out[c] - fo[c];
        lt = fmaf(d, d, lt);
      }
      float front, band, vd, d = a.gt_depth[r];
      opacity_masks(zv, d, a.trunc, front, band, vd);
      float o = sigmoidf_(10.f * fo[32]);
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
__global__ void __launch_bounds__(kTile) k_point_bwd_tc(PointArgs a, const uint4* __restrict__ wc_all,
                                                         const uint4* __restrict__ we_all) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  unsigned char* DOc_hi = sm;                       // dOut tiles: coarse hi | lo | fine hi | lo
  unsigned char* DOc_lo = sm + kDOTile;
  unsigned char* DOf_hi = sm + 2 * kDOTile;
  unsigned char* DOf_lo = sm + 3 * kDOTile;
  unsigned char* DH_hi = sm;                        // dH tile [8 chunks] aliases the dOut tiles
  unsigned char* DH_lo = sm + 8 * 2048;
  unsigned char* W1_hi = sm + 4 * kDOTile;
  unsigned char* W1_lo = W1_hi + kW1Tile;
  unsigned char* W2c_hi = W1_lo + kW1Tile;
  unsigned char* W2f_hi = W2c_hi + 2 * kW2Tile;
  const int tile = blockIdx.x, tid = threadIdx.x, warp = tid >> 5;
  const int n_tiles = a.perm ? a.counts[cTiles] : a.n_tiles_host;
  if (tile >= n_tiles) return;
  int expert = -1;
  if (MODE == kMap) expert = a.tile_class[tile];
  const bool fine = MODE == kMap && expert >= 0;
  load_weights_tc(W1_hi, W1_lo, W2c_hi, W2f_hi, wc_all, we_all + (int64_t)(fine ? expert : 0) * kNetTc, fine);
  if (warp == 0) tmem_alloc(&tmem_base_s, 128);
  if (tid == 0) mbar_init(&bar, 1);

  const int64_t q = (int64_t)tile * kTile + tid;
  int64_t i, r;
  float zv, x[3];
  const bool valid = slot_point<MODE>(a, q, i, r, zv, x);
  // ---- gradients w.r.t. the MLP outputs (rows of 48: 33 used, 3 + 12 zero padding)
  float dc[48], df[48];
#pragma unroll
  for (int k = 0; k < 48; ++k) dc[k] = df[k] = 0.f;
  if (MODE == kTv) {
    if (valid) dc[0] = a.docc[q];
  } else if (MODE == kTrack) {
    if (valid) {
      const float4* s4 = reinterpret_cast<const float4*>(a.dfine36 + (a.p0 + i) * kOutP);
#pragma unroll
      for (int k = 0; k < kOutP / 4; ++k) {
        float4 v = s4[k];
        dc[4 * k] = v.x; dc[4 * k + 1] = v.y; dc[4 * k + 2] = v.z; dc[4 * k + 3] = v.w;
      }
    }
  } else if (valid) {
    const float g_lt = 2.f * a.lam_lt / (33.f * (float)a.P_total);
    const int64_t p = a.p0 + i;
    const float4* sd = reinterpret_cast<const float4*>(a.dfine36 + p * kOutP);
    const float4* sc = reinterpret_cast<const float4*>(a.coarse36 + p * kOutP);
    const float4* sf = reinterpret_cast<const float4*>(a.fine36 + p * kOutP);
    float fo32 = 0.f;
#pragma unroll
    for (int k = 0; k < kOutP / 4; ++k) {
      float4 d = sd[k], c = sc[k], f = sf[k];
      const float dd[4] = {d.x, d.y, d.z, d.w}, cc[4] = {c.x, c.y, c.z, c.w}, ff[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int ch = 4 * k + e;
        if (ch < DNS_LATENT) {
          const float gl = g_lt * (cc[e] - ff[e]);
          dc[ch] = gl;            // coarse net: only the latent loss reaches it in mapping
          df[ch] = dd[e] - gl;
          if (ch == 32) fo32 = ff[e];
        }
      }
    }
    if (a.counts[cFront] > 0 && a.counts[cBand] > 0) {
      float front, band, vd, d = a.gt_depth[r];
      opacity_masks(zv, d, a.trunc, front, band, vd);
      float o = sigmoidf_(10.f * fo32);
      float u = (zv - d) / a.sigma;
      float ps = 0.5f * __expf(-0.5f * u * u);
      float inv_p = 1.f / (float)a.P_total;
      float d_o = 2.f * a.lam_fs * inv_p * o * front * vd + 2.f * a.lam_op * inv_p * (o - ps) * band;
      df[32] += d_o * 10.f * o * (1.f - o);
    }
    if (!fine) {
#pragma unroll
      for (int k = 0; k < 48; ++k) df[k] = 0.f;
    }
  }
  constexpr int DOCH = MODE == kMap ? 10 : 5;   // chunks per half of the dOut image: 5 (40 >= 33 channels) per net
  constexpr int HCH = MODE == kMap ? 8 : 4;     // chunks per half of the H / dH images
  const bool stash = a.need_dparams && !(a.dbg & 2);
  uint4* doimg = stash ? a.dOimg + (int64_t)tile * (2 * DOCH * kTile) + tid : nullptr;
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const bool gi = stash && c < 5;
    put_chunk_img(DOc_hi, DOc_lo, c, 2048, tid, dc + 8 * c, gi ? doimg + c * kTile : nullptr,
                  gi ? doimg + (DOCH + c) * kTile : nullptr);
    if (MODE == kMap)
      put_chunk_img(DOf_hi, DOf_lo, c, 2048, tid, df + 8 * c, gi ? doimg + (5 + c) * kTile : nullptr,
                    gi ? doimg + (DOCH + 5 + c) * kTile : nullptr);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
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
  mbar_wait(&bar, 0);
  tc_fence_after();
  const uint32_t lane_addr = tmem_d + ((uint32_t)(warp * 32) << 16);
  {
    float dh[64];
    // ReLU mask: the bf16 hi half of the stashed activations is non-zero exactly where the activation was positive
    const uint4* himg = a.Himg + (int64_t)tile * (2 * HCH * kTile) + tid;
#pragma unroll
    for (int net = 0; net < 2; ++net) {
      const bool on = net == 0 || fine;
#pragma unroll
      for (int g2 = 0; g2 < 2; ++g2) {
        float v[16];
        if (on) tmem_ld16(lane_addr + net * 32 + 16 * g2, v);   // `on` is uniform over the CTA
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          const uint4 hv = on ? himg[(net * 4 + 2 * g2 + c2) * kTile] : make_uint4(0, 0, 0, 0);
          const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
          for (int e = 0; e < 8; ++e)
            dh[32 * net + 16 * g2 + 8 * c2 + e] = (on && ((hw[e >> 1] >> (16 * (e & 1))) & 0x7fffu)) ? v[8 * c2 + e] : 0.f;
        }
      }
    }
    uint4* dhimg = stash ? a.dHimg + (int64_t)tile * (2 * HCH * kTile) + tid : nullptr;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const bool gi = stash && c < HCH;
      put_chunk_img(DH_hi, DH_lo, c, 2048, tid, dh + 8 * c, gi ? dhimg + c * kTile : nullptr,
                    gi ? dhimg + (HCH + c) * kTile : nullptr);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {  // dX = [dHc | dHf] . [W1c ; W1f]   (B = combined W1 tile MN-major: LBO 128 over hidden rows, SBO 1024)
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, 80, 0, 1);
#pragma unroll 1
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t aoff = ks * 4096, boff = ks * 256;
      const uint64_t a_hi = umma_desc(smem_u32(DH_hi) + aoff, 2048, 128), a_lo = umma_desc(smem_u32(DH_lo) + aoff, 2048, 128);
      const uint64_t b_hi = umma_desc(smem_u32(W1_hi) + boff, 128, 1024), b_lo = umma_desc(smem_u32(W1_lo) + boff, 128, 1024);
      umma_bf16(tmem_d, a_hi, b_hi, idesc, ks > 0 ? 1u : 0u);
      umma_bf16(tmem_d, a_lo, b_hi, idesc, 1u);
      umma_bf16(tmem_d, a_hi, b_lo, idesc, 1u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 1);
  tc_fence_after();
  float dx[3] = {0.f, 0.f, 0.f}, dg[32];
#pragma unroll
  for (int g5 = 0; g5 < 5; ++g5) {
    float v[16];
    tmem_ld16(lane_addr + 16 * g5, v);
    if (g5 < 3) {
      if (a.need_drays && valid) dx[g5] = oneblob16_bwd(x[g5], v);
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) dg[16 * (g5 - 3) + k] = v[k];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 128);
  if (!valid) return;
  float dxg[3];
  hashgrid_bwd_regs(a.G, a.table, (a.need_dparams && !(a.dbg & 4)) ? a.d_table : nullptr, x, dg, a.need_drays != 0 && !(a.dbg & 8), dxg);
  if (a.need_drays && MODE != kTv) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float g = (dx[c] + dxg[c]) / (float)a.B.ext[c];
      atomicAdd(a.d_rays_o + 3 * r + c, g);
      atomicAdd(a.d_rays_d + 3 * r + c, g * zv);
    }
  }
}

size_t point_fwd_tc_smem() { return 2 * kXTile + 2 * kW1Tile + 4 * kW2Tile; }
size_t point_bwd_tc_smem() { return 4 * kDOTile + 2 * kW1Tile + 4 * kW2Tile; }

int prep_nets_tc(const float* coarse, const float* experts, int n_experts, uint4* wc, uint4* we, cudaStream_t st) {
  k_prep_net80_tc<<<1, 128, 0, st>>>(coarse, wc);
  if (experts && n_experts > 0) k_prep_net80_tc<<<n_experts, 128, 0, st>>>(experts, we);
  return check_launch("prep_nets_tc");
}

static void set_attrs() {
  static bool done = false;
  if (done) return;
  const int f = (int)point_fwd_tc_smem(), b = (int)point_bwd_tc_smem();
  cudaFuncSetAttribute(k_point_fwd_tc<kTrack>, cudaFuncAttributeMaxDynamicSharedMemorySize, f);
  cudaFuncSetAttribute(k_point_fwd_tc<kMap>, cudaFuncAttributeMaxDynamicSharedMemorySize, f);
  cudaFuncSetAttribute(k_point_fwd_tc<kTv>, cudaFuncAttributeMaxDynamicSharedMemorySize, f);
  cudaFuncSetAttribute(k_point_bwd_tc<kTrack>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
  cudaFuncSetAttribute(k_point_bwd_tc<kMap>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
  cudaFuncSetAttribute(k_point_bwd_tc<kTv>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
  done = true;
}

// DNS_PT1=1 selects the one-thread-per-slot kernels of this file (A/B measurements); default: point_tc2.cu
static bool one_thread_per_slot() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DNS_PT1");
    v = (e && atoi(e)) ? 1 : 0;
  }
  return v == 1;
}

int launch_point_fwd_tc(int mode, const PointArgs& pa, int tiles, const uint4* wc, const uint4* we, cudaStream_t st) {
  if (!one_thread_per_slot()) return launch_point_fwd_tc2(mode, pa, tiles, wc, we, st);
  set_attrs();
  const size_t smem = point_fwd_tc_smem();
  if (mode == kMap) k_point_fwd_tc<kMap><<<tiles, kTile, smem, st>>>(pa, wc, we);
  else if (mode == kTrack) k_point_fwd_tc<kTrack><<<tiles, kTile, smem, st>>>(pa, wc, we);
  else k_point_fwd_tc<kTv><<<tiles, kTile, smem, st>>>(pa, wc, we);
  return check_launch("point_fwd_tc");
}
int launch_point_bwd_tc(int mode, const PointArgs& pa, int tiles, const uint4* wc, const uint4* we, cudaStream_t st) {
  if (!one_thread_per_slot()) return launch_point_bwd_tc2(mode, pa, tiles, wc, we, st);
  set_attrs();
  const size_t smem = point_bwd_tc_smem();
  if (mode == kMap) k_point_bwd_tc<kMap><<<tiles, kTile, smem, st>>>(pa, wc, we);
  else if (mode == kTrack) k_point_bwd_tc<kTrack><<<tiles, kTile, smem, st>>>(pa, wc, we);
  else k_point_bwd_tc<kTv><<<tiles, kTile, smem, st>>>(pa, wc, we);
  return check_launch("point_bwd_tc");
}

}  // namespace dns
