// Ray kernel on tcgen05, TWO threads per sample point (blockDim = 2 T, T a multiple of 128), its weight
// preparation and its launcher.
//
// Same mathematics as k_ray (render.cu; slams/tracking.py:188-214, slams/mapping.py:603-635,
// utils/common.py:506-537): colour + logit layer 1 per point, colour head, logit layer 2 on the composited hidden
// state, occupancy compositing, p/d/l losses and the full backward.  The two 112x64 contractions per point run on
// tcgen05 (bf16 hi + lo halves, fp32 accumulation in TMEM):
//
//   forward   H[p][0..63]   = X[p][0..111] . W1^T     M = 128 points, N = 64, K = 112
//   backward  dX[p][0..111] = dH[p][0..63] . W1       M = 128 points, N = 112, K = 64
//
// Operand rows are 16-byte feature chunks [chunk][point][8 x bf16] (canonical no-swizzle UMMA layout); the SAME
// shared-memory copy of W1 ([feature chunk][hidden row][8 features]) is the K-major B operand of the forward GEMM
// and the MN-major B operand of the backward GEMM.  A first version ran one thread per point and was issue bound
// (~6.5 k instructions per thread on 10 warps per SM, 176 registers).  Here the threads t and t + T share point t
// (same TMEM lane quarter because T % 128 == 0) and split its work:
//
//   group 0 (t <  T)  OneBlob of the point -> X chunks 0..5; colour hidden units (accumulator columns 0..31),
//                     colour head, colour gradients, dH chunks 0..3; OneBlob backward / ray gradients
//   group 1 (t >= T)  latent + pixel-feature rows -> X chunks 6..13; occupancy compositing (transmittance scans),
//                     logit hidden units (columns 32..63), depth / logit gradients, dH chunks 4..7,
//                     d(latent) and d(feature) rows
//
// Per-ray work (column sums, logits, losses, QV) is spread over all 2 T threads.  T = 128 keeps two CTAs per SM
// (16 warps), T = 256 one CTA of 16 warps with fewer idle rows; pick_ray_block_tc2 chooses by row efficiency.
#include "tc_common.cuh"

namespace dns {

constexpr int kW1oBytes2 = 14 * 64 * 16;  // one bf16 half of the [64 x 112] colour|logit layer-1 weights

// colour | logit layer-1 weights -> bf16 hi / lo chunk tiles [14 feature chunks][64 hidden rows][8 features]
// (+ the colour head's layer 2 as W2cT [32][4], what k_transpose_out computes for the SIMT path: one launch fewer)
__global__ void k_prep_w1o_tc(const float* __restrict__ color, const float* __restrict__ logit, uint4* __restrict__ hi,
                              uint4* __restrict__ lo, uint4* __restrict__ hi16, uint4* __restrict__ lo16,
                              float* __restrict__ W2cT) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (W2cT && i < 128) {
    const int j = i >> 2, c = i & 3;
    W2cT[i] = c < 3 ? color[32 * kIn2 + c * 32 + j] : 0.f;
  }
  if (i >= 14 * 64) return;
  int c = i >> 6, j = i & 63;
  const float* src = (j < 32 ? color + j * kIn2 : logit + (j - 32) * kIn2) + 8 * c;
  float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
  uint4 h, l;
  split8(a, b, h, l);        // bf16 halves: backward GEMM (meets gradients)
  hi[i] = h;
  lo[i] = l;
  split8_f16(a, b, h, l);    // fp16 halves: forward GEMM (decides the ReLU)
  hi16[i] = h;
  lo16[i] = l;
}

__global__ void __launch_bounds__(512) k_ray_tc2(RayArgs a, const uint4* __restrict__ w1_hi, const uint4* __restrict__ w1_lo) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  __shared__ uint64_t bar, wbar;
  __shared__ uint32_t tmem_base_s;
  const int T = a.T, S = a.S, RPC = a.RPC, C = a.C, C4 = a.C4, ld = T + 1;
  const int NT = 2 * T;                        // threads
  const int cs = T * 16;                       // bytes between feature chunks of a point tile
  const int MT = T >> 7;                       // 128-point MMA tiles
  unsigned char* R = smraw;                    // aliased region: X tile -> staging -> dH tile -> staging
  unsigned char* X_hi = R;
  unsigned char* X_lo = R + 14 * cs;
  unsigned char* D_hi = R;
  unsigned char* D_lo = R + 8 * cs;
  float* XC = reinterpret_cast<float*>(R);     // [36][T+1] staging for per-ray reductions
  unsigned char* W_hi = R + 28 * cs;
  unsigned char* W_lo = W_hi + kW1oBytes2;
  float* W2c = reinterpret_cast<float*>(W_lo + kW1oBytes2);  // [32][4]
  float* bs = W2c + 128;                // [T]
  float* us = bs + T;                   // [T]
  float* ws = us + T;                   // [T]
  float* wsh = ws + T;                  // [T] compositing weight of the point (group 1 -> group 0)
  float* tmp = wsh + T;                 // [T] colour part of d_w (group 0 -> group 1)
  float* HB = tmp + T;                  // [RPC][32]
  float* QV = HB + RPC * 32;            // [RPC][32]
  float* RO = QV + RPC * 32;            // [RPC][8]
  float* RG = RO + RPC * 8;             // [RPC][8]
  float* LG = RG + RPC * 8;             // [RPC][C4]
  float* LS = LG + RPC * C4;            // [4]
  float* RT = LS + 4;                   // [RPC][8] ground truth of the CTA's rays, fetched at kernel start: colour 0..2, depth 3,
                                        // label 4 (int bits), mask 5 -- the loss section (one warp per ray) then waits for no load
  float* W2L = RT + RPC * 8;            // [C][32] layer 2 of the logit head (read twice per CTA: forward and backward)
  const int t = threadIdx.x, warp = t >> 5;
  const int grp = t >= T ? 1 : 0, row = t - grp * T;
  if (t == 0) {   // the 28 KB weight tile (fp16 halves for the forward GEMM) arrives by two bulk copies while the
    mbar_init(&bar, 1);   // threads encode their rows; the bf16 halves replace it before the backward GEMM
    mbar_init(&wbar, 1);
    mbar_expect_tx(&wbar, 2 * kW1oBytes2);
    bulk_g2s(W_hi, a.w16_hi, kW1oBytes2, &wbar);
    bulk_g2s(W_lo, a.w16_lo, kW1oBytes2, &wbar);
  }
  for (int i = t; i < 32; i += NT) reinterpret_cast<float4*>(W2c)[i] = reinterpret_cast<const float4*>(a.W2cT)[i];
  if (t < 4) LS[t] = 0.f;
  if (a.w2l_smem)
    for (int i = t; i < C * 8; i += NT) reinterpret_cast<float4*>(W2L)[i] = __ldg(reinterpret_cast<const float4*>(a.logit + 32 * kIn2) + i);
  if (t < RPC) {
    const int64_t rl2 = (int64_t)blockIdx.x * RPC + t, r2 = a.ray0 + rl2;
    if (rl2 < a.Nc) {
      float* rt = RT + t * 8;
      rt[0] = a.gt_color[3 * r2];
      rt[1] = a.gt_color[3 * r2 + 1];
      rt[2] = a.gt_color[3 * r2 + 2];
      rt[3] = a.gt_depth[r2];
      const int64_t lab = a.gt_label[r2];
      rt[4] = __int_as_float((lab < 0 || lab >= C) ? -1 : (int)lab);
      rt[5] = (a.mode == kTrack && a.mask) ? (a.mask[r2] != 0 ? 1.f : 0.f) : 1.f;
    }
  }
  const uint32_t tmem_cols = MT == 1 ? 128 : 256;
  if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
  DNS_CLK_DECL

  const int lr = row / S, s = row - lr * S;
  const int64_t rl = (int64_t)blockIdx.x * RPC + lr;      // chunk-local ray
  const bool valid = lr < RPC && rl < a.Nc;
  const int64_t r = a.ray0 + rl;                          // global ray
  const int64_t p = r * S + s;                            // global point
  const int rb = lr * S;                                  // first row of my ray
  float* xc = XC + row;
  float x[3] = {0.f, 0.f, 0.f}, zv = 0.f, occ = 0.f;
  bool band = true;   // features_band_only: the pixel-feature row exists only for samples inside the truncation band
  // weight-gradient operands leave the kernel as bf16 hi/lo tile images: this CTA owns T rows = T/RS sub-tiles,
  // each laid out [half][chunk][RS rows] (tc.cu: k_dw_img); rows of absent points are written as zeros
  const bool stash = a.need_dparams != 0;
  const int RS = a.RS, sub = row / RS, rr = row - sub * RS;
  const int64_t img_row0 = ((int64_t)blockIdx.x * T + (int64_t)sub * RS) * 2;
  uint4* const x2p = stash ? a.X2img + img_row0 * 14 + rr : nullptr;     // + chunk * RS (hi), + (K + chunk) * RS (lo)
  uint4* const dh2p = stash ? a.dH2img + img_row0 * 8 + rr : nullptr;
  uint4* const hcp = stash ? a.Hcolimg + img_row0 * 4 + rr : nullptr;
  uint4* const dpp = stash ? a.dpreimg + img_row0 + rr : nullptr;
#define IMG2(p, K, c) ((p) ? (p) + (c) * RS : nullptr), ((p) ? (p) + ((K) + (c)) * RS : nullptr)

  // ---- stage this point's row of X = [OneBlob(x) 48 | latent 32 | pixel feature 32] as bf16 hi/lo chunks
  if (valid) {
    zv = a.z[p];
    if (a.feat_band && grp == 1) band = in_band(zv, a.gt_depth[r]);
    if (grp == 0) {
      point_from_ray(a.rays_o + 3 * r, a.rays_d + 3 * r, zv, a.B, x);
#pragma unroll 1   // one copy of the OneBlob + operand-split code instead of three (instruction-cache footprint)
      for (int c = 0; c < 3; ++c) {
        float pe[16];
        oneblob16(c == 0 ? x[0] : (c == 1 ? x[1] : x[2]), pe);
        put_chunk_f16_img(X_hi, X_lo, 2 * c, cs, row, pe, IMG2(x2p, 14, 2 * c));
        put_chunk_f16_img(X_hi, X_lo, 2 * c + 1, cs, row, pe + 8, IMG2(x2p, 14, 2 * c + 1));
      }
    } else {
      {
        float lat[kOutP];
        const float4* s4 = reinterpret_cast<const float4*>(a.fine36 + p * kOutP);
#pragma unroll
        for (int q = 0; q < kOutP / 4; ++q) {
          float4 v = s4[q];
          lat[4 * q] = v.x; lat[4 * q + 1] = v.y; lat[4 * q + 2] = v.z; lat[4 * q + 3] = v.w;
        }
        occ = lat[0];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          put_chunk_f16_img(X_hi, X_lo, 6 + c, cs, row, lat + 1 + 8 * c, IMG2(x2p, 14, 6 + c));
      }
      {
        float ft[32];
        if (a.features && band) {
          const float4* s4 = reinterpret_cast<const float4*>(a.features + p * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 v = s4[q];
            ft[4 * q] = v.x; ft[4 * q + 1] = v.y; ft[4 * q + 2] = v.z; ft[4 * q + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) ft[k] = 0.f;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
          put_chunk_f16_img(X_hi, X_lo, 10 + c, cs, row, ft + 8 * c, IMG2(x2p, 14, 10 + c));
      }
    }
  } else {
    const uint4 z4 = make_uint4(0, 0, 0, 0);
    const int c0 = grp ? 6 : 0, c1 = grp ? 14 : 6;
    for (int c = c0; c < c1; ++c) {
      *reinterpret_cast<uint4*>(X_hi + c * cs + row * 16) = z4;
      *reinterpret_cast<uint4*>(X_lo + c * cs + row * 16) = z4;
      if (stash) x2p[c * RS] = x2p[(14 + c) * RS] = z4;
    }
  }
  DNS_CLK(a, 1)
  // occupancy compositing (common.py:524-532) depends on the latent row only: group 1 runs its scans while the
  // forward GEMM is in flight (its own named barrier; group 0 issues / waits for the MMAs)
  float alpha = 0.f, b = 1.f;
  if (grp == 1) {
    alpha = valid ? sigmoidf_(10.f * occ) : 0.f;
    b = __fadd_rn(1.f - alpha, 1e-10f);
    bs[row] = b;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  DNS_CLK(a, 2)
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  // ---- forward GEMM: H = X . W1^T  (A K-major: LBO = chunk stride, SBO = 128; B K-major: LBO = 1024, SBO = 128)
  if (t == 0) {
    mbar_wait(&wbar, 0);   // weights landed
    const uint32_t idesc = umma_idesc_f16(128, 64, 0, 0, 0, 0);   // fp16 hi / lo halves
    for (int mt = 0; mt < MT; ++mt) {
      const uint32_t d = tmem_d + mt * 112;
#pragma unroll 1
      for (int ks = 0; ks < 7; ++ks) {
        const uint32_t aoff = mt * 2048 + ks * 2 * cs, boff = ks * 2 * 1024;
        const uint64_t a_hi = umma_desc(smem_u32(X_hi) + aoff, cs, 128), a_lo = umma_desc(smem_u32(X_lo) + aoff, cs, 128);
        const uint64_t b_hi = umma_desc(smem_u32(W_hi) + boff, 1024, 128), b_lo = umma_desc(smem_u32(W_lo) + boff, 1024, 128);
        umma_bf16(d, a_hi, b_hi, idesc, ks > 0 ? 1u : 0u);
        umma_bf16(d, a_lo, b_hi, idesc, 1u);
        umma_bf16(d, a_hi, b_lo, idesc, 1u);
      }
    }
    umma_commit(&bar);
  }
  float Ts = 1.f, u = 0.f, sumu = 0.f, w = 0.f;
  if (grp == 1) {
    if (valid)
      for (int j = 0; j < s; ++j) Ts *= bs[rb + j];
    u = alpha * Ts;
    us[row] = u;
    asm volatile("bar.sync 1, %0;" ::"r"(T) : "memory");   // group 1 only (T threads, whole warps)
    if (valid)
      for (int j = 0; j < S; ++j) sumu += us[rb + j];
    w = valid ? (a.fwd_only == 2 ? 1.f : u / sumu) : 0.f;   // 2: free-point query, no compositing
    wsh[row] = w;
  }
  mbar_wait_cta(&bar, 0);
  DNS_CLK(a, 3)
  tc_fence_after();
  if (t == 0 && !a.fwd_only) {   // the forward GEMM has consumed the fp16 weight tile: fetch the bf16 halves over it
    mbar_expect_tx(&wbar, 2 * kW1oBytes2);
    bulk_g2s(W_hi, w1_hi, kW1oBytes2, &wbar);
    bulk_g2s(W_lo, w1_lo, kW1oBytes2, &wbar);
  }
  // this thread's half of the hidden row: group 0 colour units, group 1 logit units
  const uint32_t taddr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((row >> 7) * 112);
  float h[32];
  {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      float v[16];
      tmem_ld16(taddr + 32 * grp + 16 * g, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) h[16 * g + i] = valid ? fmaxf(v[i], 0.f) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();   // every thread has read its accumulator half: region R may be reused as staging
  DNS_CLK(a, 4)
  float rgb[3] = {0.f, 0.f, 0.f};
  if (grp == 0) {    // colour head: 32 -> 3, sigmoid (decoder.py:123)
    float pre[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float4 wv = *reinterpret_cast<const float4*>(W2c + 4 * j);
      pre[0] = fmaf(h[j], wv.x, pre[0]);
      pre[1] = fmaf(h[j], wv.y, pre[1]);
      pre[2] = fmaf(h[j], wv.z, pre[2]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) rgb[c] = sigmoidf_(pre[c]);
  }
  if (grp == 0) w = wsh[row];
  if (valid) {
    if (grp == 1) {
#pragma unroll
      for (int j = 0; j < 32; ++j) xc[j * ld] = w * h[j];
      xc[35 * ld] = w * zv;
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) xc[(32 + c) * ld] = w * rgb[c];
    }
  }
  __syncthreads();
  DNS_CLK(a, 5)
  for (int e = t; e < RPC * 36; e += NT) {
    int l2 = e / 36, o = e - l2 * 36;
    if ((int64_t)blockIdx.x * RPC + l2 < a.Nc) {
      const float* col = XC + o * ld + l2 * S;
      float acc = 0.f;
      for (int j = 0; j < S; ++j) acc += col[j];
      if (o < 32) HB[l2 * 32 + o] = acc;
      else RO[l2 * 8 + (o - 32)] = acc;
    }
  }
  __syncthreads();
  DNS_CLK(a, 6)
  float dz = 0.f;
  if (grp == 1) {
    dz = valid ? zv - RO[lr * 8 + 3] : 0.f;
    bs[row] = w * dz * dz;
    us[row] = w * dz;
  }
  const float* W2l = a.w2l_smem ? W2L : a.logit + 32 * kIn2;   // (many classes x many rays per CTA: the copy does not fit)
  for (int e = t; e < RPC * C; e += NT) {
    int l2 = e / C, c = e - l2 * C;
    if ((int64_t)blockIdx.x * RPC + l2 < a.Nc) {
      const float4* wr = reinterpret_cast<const float4*>(W2l + c * 32);
      const float* hb = HB + l2 * 32;
      float acc = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 v = wr[q];
        acc = fmaf(hb[4 * q], v.x, acc);
        acc = fmaf(hb[4 * q + 1], v.y, acc);
        acc = fmaf(hb[4 * q + 2], v.z, acc);
        acc = fmaf(hb[4 * q + 3], v.w, acc);
      }
      LG[l2 * C4 + c] = acc;
    }
  }
  __syncthreads();
  DNS_CLK(a, 7)
  // ---- per-ray losses and their gradients: one warp per ray, lanes over samples / classes
  for (int l2 = warp; l2 < RPC; l2 += (NT >> 5)) {
    const int64_t rl2 = (int64_t)blockIdx.x * RPC + l2, r2 = a.ray0 + rl2;
    if (rl2 >= a.Nc) break;   // uniform over the warp
    const int lane = t & 31;
    float var = 0.f, swdz = 0.f;
    for (int j = lane; j < S; j += 32) {
      var += bs[l2 * S + j];
      swdz += us[l2 * S + j];
    }
    float* ro = RO + l2 * 8;
    float* rg = RG + l2 * 8;
    float* lg = LG + l2 * C4;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) {
      const float v = lg[c];
      a.pred_logits[r2 * C + c] = v;
      mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      var += __shfl_xor_sync(0xffffffffu, var, o);
      swdz += __shfl_xor_sync(0xffffffffu, swdz, o);
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += __expf(lg[c] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    // scalars of the ray: evaluated by every lane (same values), written by lane 0
    const float dhat = ro[3];
    const float* rt = RT + l2 * 8;
    const float gd = rt[3];
    int lab = __float_as_int(rt[4]);
    if (lab < 0) {   // label outside [0, C): torch's cross_entropy raises here; the flag surfaces as losses[7] = -3
      if (lane == 0) *a.err = 3;
      lab = 0;
    }
    const bool track = a.mode == kTrack;
    const bool m = rt[5] != 0.f;
    const float n_ray = track ? (float)a.counts[cMask] : (float)a.N_total;
    float lp = 0.f, ldp = 0.f, ll = 0.f;
    float g_rgb[3] = {0.f, 0.f, 0.f}, g_d = 0.f, g_var = 0.f, g_ce = 0.f, lse = 0.f;
    if (m) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float e = ro[c] - rt[c];
        lp = fmaf(e, e, lp);
        g_rgb[c] = a.lam_p * 2.f * e / (3.f * n_ray);
      }
      float diff = dhat - gd;
      float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
      if (track) {
        float vv = var + 1e-10f, inv = rsqrtf(vv);
        ldp = fabsf(diff) * inv;
        g_d = a.lam_d * sgn * inv / n_ray;
        g_var = a.lam_d * fabsf(diff) * (-0.5f) * inv / vv / n_ray;
        g_d += g_var * (-2.f) * swdz;
      } else if (gd > 0.f) {
        ldp = fabsf(diff);
        g_d = a.lam_d * sgn / (float)a.counts[cDpos];
      }
      lse = logf(se) + mx;
      ll = lse - lg[lab];
      g_ce = a.lam_l / n_ray;
    }
    __syncwarp();   // lg[lab] has been read by every lane before any entry is overwritten
    for (int c = lane; c < C4; c += 32) {
      const float g = (m && c < C) ? g_ce * (__expf(lg[c] - lse) - (c == lab ? 1.f : 0.f)) : 0.f;
      lg[c] = g;
      if (a.need_dparams) a.dlogit[rl2 * C4 + c] = g;
    }
    if (a.need_dparams) a.Hbar[rl2 * 32 + lane] = HB[l2 * 32 + lane];
    if (lane == 0) {
      a.pred_color[3 * r2] = ro[0];
      a.pred_color[3 * r2 + 1] = ro[1];
      a.pred_color[3 * r2 + 2] = ro[2];
      a.pred_depth[r2] = dhat;
      a.pred_var[r2] = var;
      rg[0] = g_rgb[0];
      rg[1] = g_rgb[1];
      rg[2] = g_rgb[2];
      rg[3] = g_d;
      rg[4] = g_var;
      atomicAdd(LS + 0, lp);
      atomicAdd(LS + 1, ldp);
      atomicAdd(LS + 2, ll);
    }
  }
  __syncthreads();
  DNS_CLK(a, 8)
  if (t == 0) {
    atomicAdd(a.raw + rP, LS[0]);
    atomicAdd(a.raw + rD, LS[1]);
    atomicAdd(a.raw + rL, LS[2]);
  }
  if (a.fwd_only) {   // inference: predictions are out, nothing to differentiate (uniform over the CTA)
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, tmem_cols);
    return;
  }
  for (int e = t; e < RPC * 32; e += NT) {
    int l2 = e >> 5, j = e & 31;
    if ((int64_t)blockIdx.x * RPC + l2 < a.Nc) {
      const float* lg = LG + l2 * C4;
      float acc = 0.f;
      for (int c = 0; c < C; ++c) acc = fmaf(lg[c], W2l[c * 32 + j], acc);
      QV[e] = acc;
    }
  }
  // ---- per-point backward: d_w = dL/dw of the point, assembled from both groups
  if (grp == 0) {
    float part = 0.f;
    if (valid) {
      const float* rg = RG + lr * 8;
      part = rg[0] * rgb[0] + rg[1] * rgb[1] + rg[2] * rgb[2];
    }
    tmp[row] = part;
  }
  __syncthreads();
  DNS_CLK(a, 9)
  float d_w = 0.f;
  if (grp == 1) {
    if (valid) {
      const float* rg = RG + lr * 8;
      const float* qv = QV + lr * 32;
      d_w = tmp[row] + rg[3] * zv + rg[4] * dz * dz;
#pragma unroll
      for (int j = 0; j < 32; ++j) d_w = fmaf(h[j], qv[j], d_w);
    }
    bs[row] = w * d_w;
  }
  __syncthreads();
  DNS_CLK(a, 10)
  float d_u = 0.f, d_occ = 0.f;
  if (grp == 1) {
    if (valid) {
      // G = sum_j w_j d_w_j;  d_u_j = (d_w_j - G) / sumu;  suf = sum_{j>s} d_u_j u_j = sum_{j>s} w_j d_w_j - G sum_{j>s} w_j
      float G = 0.f, sufB = 0.f, sufW = 0.f;
      for (int j = 0; j < S; ++j) {
        const float bj = bs[rb + j];
        G += bj;
        if (j > s) {
          sufB += bj;
          sufW += wsh[rb + j];
        }
      }
      d_u = (d_w - G) / sumu;
      const float suf = sufB - G * sufW;
      float d_alpha = d_u * Ts - suf / b;
      d_occ = d_alpha * 10.f * alpha * (1.f - alpha);
      const float* qv = QV + lr * 32;
#pragma unroll
      for (int j = 0; j < 32; ++j) h[j] = h[j] > 0.f ? w * qv[j] : 0.f;
    }
  } else {
    float dp[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (valid) {
      const float* rg = RG + lr * 8;
#pragma unroll
      for (int c = 0; c < 3; ++c) dp[c] = w * rg[c] * rgb[c] * (1.f - rgb[c]);
    }
    if (stash) {   // colour hidden activations (zeros for absent points) and the pre-sigmoid colour gradient
#pragma unroll
      for (int c = 0; c < 4; ++c) store_chunk_img(h + 8 * c, hcp + c * RS, hcp + (4 + c) * RS);
      store_chunk_img(dp, dpp, dpp + RS);
    }
    if (valid) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float4 wv = *reinterpret_cast<const float4*>(W2c + 4 * j);
        h[j] = h[j] > 0.f ? dp[0] * wv.x + dp[1] * wv.y + dp[2] * wv.z : 0.f;
      }
    }
  }
  // ---- backward GEMM: dX = dH . W1  (A = dH K-major over hidden; B = W1 MN-major: features contiguous)
#pragma unroll
  for (int c = 0; c < 4; ++c)   // invalid threads hold zeros
    put_chunk_img(D_hi, D_lo, 4 * grp + c, cs, row, h + 8 * c, IMG2(dh2p, 8, 4 * grp + c));
#undef IMG2
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  DNS_CLK(a, 11)
  if (t == 0) {
    tc_fence_after();
    mbar_wait(&wbar, 1);   // bf16 weight halves landed
    const uint32_t idesc = umma_idesc_bf16(128, 112, 0, 1);
    for (int mt = 0; mt < MT; ++mt) {
      const uint32_t d = tmem_d + mt * 112;
#pragma unroll 1
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t aoff = mt * 2048 + ks * 2 * cs, boff = ks * 256;
        const uint64_t a_hi = umma_desc(smem_u32(D_hi) + aoff, cs, 128), a_lo = umma_desc(smem_u32(D_lo) + aoff, cs, 128);
        const uint64_t b_hi = umma_desc(smem_u32(W_hi) + boff, 128, 1024), b_lo = umma_desc(smem_u32(W_lo) + boff, 128, 1024);
        umma_bf16(d, a_hi, b_hi, idesc, ks > 0 ? 1u : 0u);
        umma_bf16(d, a_lo, b_hi, idesc, 1u);
        umma_bf16(d, a_hi, b_lo, idesc, 1u);
      }
    }
    umma_commit(&bar);
  }
  mbar_wait_cta(&bar, 1);
  DNS_CLK(a, 12)
  tc_fence_after();
  float g3[3] = {0.f, 0.f, 0.f};
  if (grp == 0) {   // dX columns 0..47: OneBlob backward -> d(ray)
    if (a.need_drays) {
#pragma unroll 1   // one copy of the OneBlob backward instead of three
      for (int g = 0; g < 3; ++g) {
        float v[16];
        tmem_ld16(taddr + 16 * g, v);
        const float gx = valid ? oneblob16_bwd(g == 0 ? x[0] : (g == 1 ? x[1] : x[2]), v) / (float)a.B.ext[g] : 0.f;
        g3[0] = g == 0 ? gx : g3[0];
        g3[1] = g == 1 ? gx : g3[1];
        g3[2] = g == 2 ? gx : g3[2];
      }
    }
  } else {          // columns 48..79: d(latent) (with d_occ in channel 0); columns 80..111: d(pixel feature)
    float lat[kOutP];
    lat[0] = d_occ;
    lat[33] = lat[34] = lat[35] = 0.f;
#pragma unroll
    for (int g = 3; g < 7; ++g) {
      float v[16];
      if (g < 5 || a.need_dfeat) tmem_ld16(taddr + 16 * g, v);   // (warp-uniform condition: tcgen05.ld is warp-collective)
      if (g < 5) {
#pragma unroll
        for (int i = 0; i < 16; ++i) lat[1 + 16 * (g - 3) + i] = v[i];
      } else if (a.need_dfeat && valid && band) {
        float4* d4 = reinterpret_cast<float4*>(a.d_features + p * 32 + 16 * (g - 5));
#pragma unroll
        for (int q = 0; q < 4; ++q) d4[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    }
    if (valid) {
      // point order, or (tcgen05 MAP path) the row of the point's SLOT: fire-and-forget here, unit stride for the reader
      if (a.inv) {       // slot-order latent image (render.cuh slot_img): chunk c of slot sq
        const int64_t sq = a.inv[p - a.ray0 * S];
        float4* img = reinterpret_cast<float4*>(a.dfine36s);
#pragma unroll
        for (int q = 0; q < kOutP / 4; ++q)
          img[slot_img(sq, q)] = make_float4(lat[4 * q], lat[4 * q + 1], lat[4 * q + 2], lat[4 * q + 3]);
      } else {
        float4* d4 = reinterpret_cast<float4*>(a.dfine36 + p * kOutP);
#pragma unroll
        for (int q = 0; q < kOutP / 4; ++q) d4[q] = make_float4(lat[4 * q], lat[4 * q + 1], lat[4 * q + 2], lat[4 * q + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  DNS_CLK(a, 13)
  if (warp == 0) tmem_dealloc(tmem_d, tmem_cols);
  if (a.need_drays) {
    if (valid && grp == 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        xc[c * ld] = g3[c];
        xc[(3 + c) * ld] = g3[c] * zv;
      }
    }
    __syncthreads();
  DNS_CLK(a, 14)
    for (int e = t; e < RPC * 6; e += NT) {
      int l2 = e / 6, o = e - l2 * 6;
      int64_t rr2 = (int64_t)blockIdx.x * RPC + l2;
      if (rr2 < a.Nc) {
        const float* col = XC + o * ld + l2 * S;
        float acc = 0.f;
        for (int j = 0; j < S; ++j) acc += col[j];
        if (o < 3) a.d_rays_o[3 * (a.ray0 + rr2) + o] = acc;
        else a.d_rays_d[3 * (a.ray0 + rr2) + o - 3] = acc;
      }
    }
  }
  DNS_CLK(a, 15)
}

// T in {128, 256}: the larger row efficiency wins (ties -> 128, two CTAs per SM); -DDNS_ABLATE builds read DNS_RAY_T
void pick_ray_block_tc2(int S, int& T, int& RPC) {
  int forced = 0;
#ifdef DNS_ABLATE
  if (const char* e = getenv("DNS_RAY_T")) forced = atoi(e);
#endif
  int best_t = 0;
  double best = -1.0;
  for (int tt = 128; tt <= 256; tt += 128) {
    int rr = tt / S;
    if (rr > 64) rr = 64;
    if (rr < 1) continue;
    double eff = (double)(rr * S) / tt;
    if (forced == tt) eff += 10.0;
    if (eff > best + 1e-9) {
      best = eff;
      best_t = tt;
    }
  }
  if (best_t == 0) {   // S > 256: rejected by dns_render_fwd_bwd; keep the workspace query well defined
    T = 256;
    RPC = 1;
    return;
  }
  T = best_t;
  RPC = T / S > 64 ? 64 : T / S;
}

constexpr size_t kRaySmemMax = 226 * 1024;   // dynamic; the kernel's static shared memory needs the rest of the 227 KB
size_t ray_tc2_smem_bytes(int T, int RPC, int C4, bool w2l) {
  return (size_t)28 * T * 16 + 2 * kW1oBytes2 + sizeof(float) * (128 + 5 * T + RPC * (32 + 32 + 8 + 8 + 8 + C4) + 8 + (w2l ? C4 * 32 : 0));
}
size_t ray_tc2_smem_bytes(int T, int RPC, int C4) { return ray_tc2_smem_bytes(T, RPC, C4, false); }

int launch_ray_tc2(const RayArgs& ra, uint4* w1_hi, uint4* w1_lo, int64_t n_rays_chunk, cudaStream_t st) {
  static unsigned long long seen = 0;
  if (first_call_on_device(seen)) cudaFuncSetAttribute(k_ray_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRaySmemMax);
  RayArgs a2 = ra;
  a2.w2l_smem = ray_tc2_smem_bytes(ra.T, ra.RPC, ra.C4, true) <= kRaySmemMax;
  const size_t smem = ray_tc2_smem_bytes(ra.T, ra.RPC, ra.C4, a2.w2l_smem != 0);
  if (smem > kRaySmemMax) {
    set_error("ray_tc2: %lld bytes of shared memory needed (n_samples x n_class too large for the tcgen05 ray kernel)", (long long)smem);
    return DNS_ERR_UNSUPPORTED;
  }
  k_ray_tc2<<<(int)((n_rays_chunk + ra.RPC - 1) / ra.RPC), 2 * ra.T, smem, st>>>(a2, w1_hi, w1_lo);
  return check_launch("ray_tc2");
}

void pick_ray_block_any(int S, int& T, int& RPC) { pick_ray_block_tc2(S, T, RPC); }

int launch_ray_tc(const RayArgs& ra, const float* color, const float* logit, uint4* w1_hi, uint4* w1_lo, bool prep,
                  int64_t n_rays_chunk, cudaStream_t st) {
  if (prep)
    k_prep_w1o_tc<<<(14 * 64 + 127) / 128, 128, 0, st>>>(color, logit, w1_hi, w1_lo, const_cast<uint4*>(ra.w16_hi),
                                                         const_cast<uint4*>(ra.w16_lo), const_cast<float*>(ra.W2cT));
  return launch_ray_tc2(ra, w1_hi, w1_lo, n_rays_chunk, st);
}

}  // namespace dns
