// Pixel-feature branch in ONE kernel per direction: feature matching + Merge + truncation mask.
//
//   reference   utils/common.py:645-679 (feature_matching: up-sample, project, round, mask, gather),
//               models/decoder.py:67-77 (Merge: OneBlob(refer_p) || feature -> MLP 112 -> 32 -> 32 -> mean over views),
//               slams/tracking.py:163-171 / slams/mapping.py:549-557 (x truncation mask of the sample)
//
// Only the samples inside the truncation band (z within +-5 % of a positive gt depth; about a third of them) pass
// through the branch at all -- the reference evaluates every sample and multiplies two thirds by zero.  A row list of
// the band samples is built first (k_band_rows); the kernels are persistent (one CTA walks tiles) and work on rows
// (band sample, view): tile = 128 / R samples x R views, row t = sample t / R, view t % R.
//
//   forward   per row: point = o + d z (fp32), project into the view (w2c, flip y/z, K, /(z + 1e-5), rint), mask,
//             4-tap bilinear fetch of the half-resolution channels-last map at the rounded pixel (align_corners),
//             X = [OneBlob((point - cam_o - lo) / ext) 48 | feature 64] -> H = relu(X W1^T) -> O = H W2^T;
//             features[sample] = mean over the R rows.  Nothing but the [N,S,32] result touches HBM.
//   backward  X and H are RECOMPUTED (the gathered features have no gradient and no stash exists), then
//             dO = d_features[sample] / R, dH = (dO W2) [H > 0], dX[:, :48] = dH W1[:, :48] -> OneBlob backward ->
//             d(point) -> d_rays_o / d_rays_d (+=), and the weight gradients dW1 = X^T dH, dW2 = H^T dO accumulate
//             in TMEM across ALL tiles of the CTA (every row shares the Merge weights) and are flushed once.
//
// GEMMs on tcgen05 as in merge_tc.cu: bf16 hi + lo operand halves, three products, fp32 accumulation in TMEM, canonical
// no-swizzle UMMA tiles [chunk of 8 features][row][16 B]; two threads per row (tid, tid + 128: same TMEM lane quarter).
#include <string.h>

#include "tc_common.cuh"

namespace dns {

constexpr int kFW1 = 14 * 32;                  // uint4 per half of the W1 tile [14 feature chunks][32 hidden rows]
constexpr int kFW2 = 4 * 32;                   // uint4 per half of the W2 tile [4 hidden chunks][32 output rows]
constexpr int kFW = 2 * kFW1 + 2 * kFW2;       // W1 hi | W1 lo | W2 hi | W2 lo
constexpr int kFT = 256;                       // threads: two per row
constexpr int kMaxViews = 32;

struct FmArgs {
  int S, R, F, PPT;           // samples per ray, views per frame, frames, band samples per tile (128 / R)
  int ray_start[DNS_MAX_FRAMES + 1];
  int H, W, h, w;
  int64_t P;                  // N * S
  Bound B;
  const float* K;             // [3][3]
  const float* w2c;           // [F*R][16]
  const float* cam_o;         // [F*R][3]
  const float* feats[DNS_MAX_FRAMES];
  const float* rays_o;
  const float* rays_d;
  const float* z;
  const int* rows;            // band row list (sample ids), NULL = every sample
  const int* n_rows_dev;      // number of band samples (device), NULL => n_rows_host
  int64_t n_rows_host;
  const uint4* wts;           // prepared weights (kFW)
  float* out;                 // [P][32]
  const float* d_out;         // [P][32]
  float* d_params;            // tcnn layout W1[32][112] | W2[32][32], +=
  float* d_rays_o;            // [N][3] +=
  float* d_rays_d;
  int need_dparams, need_drays;
  unsigned long long* phase_clk;   // -DDNS_ABLATE builds only (DNS_PHASE_CLK_FM): [32] clock64 deltas of thread 0 (fwd 0.., bwd 16..)
  uint4* Ximg;                // optional stash: the X operand tiles of the forward pass, [tile][hi | lo][14 chunks][128 rows]
  int64_t stash_bytes;
};
constexpr int kXImgBytes = 28 * 2048;   // one tile image = the shared-memory X tile, byte for byte

// params = W1[32][112] | W2[32][32] (tinycudann layout) -> bf16 hi/lo chunk tiles
__global__ void k_prep_featmerge(const float* __restrict__ params, uint4* __restrict__ out) {
  for (int i = threadIdx.x; i < kFW1 + kFW2; i += blockDim.x) {
    const float* src;
    int dst_hi, dst_lo;
    if (i < kFW1) {
      int c = i >> 5, j = i & 31;
      src = params + j * kIn2 + 8 * c;
      dst_hi = i;
      dst_lo = kFW1 + i;
    } else {
      int k = i - kFW1, c = k >> 5, j = k & 31;
      src = params + 32 * kIn2 + j * 32 + 8 * c;
      dst_hi = 2 * kFW1 + k;
      dst_lo = 2 * kFW1 + kFW2 + k;
    }
    uint4 h, l;
    split8(*reinterpret_cast<const float4*>(src), *reinterpret_cast<const float4*>(src + 4), h, l);
    out[dst_hi] = h;
    out[dst_lo] = l;
  }
}


// Row list of the band samples.  A block owns 4096 consecutive samples and appends its (ordered) hits with ONE
// atomic on the counter, so the samples of a ray stay adjacent; block order is arbitrary (results do not depend on it).
constexpr int kBandPer = 16;
__global__ void __launch_bounds__(256) k_band_rows(const float* __restrict__ z, const float* __restrict__ gt_depth, int64_t P,
                                                   int S, int* __restrict__ rows, int* __restrict__ counter) {
  __shared__ int warp_tot[8];
  __shared__ int base_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t p0 = ((int64_t)blockIdx.x * 256 + tid) * kBandPer;
  unsigned hits = 0;
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < kBandPer; ++k) {
    const int64_t p = p0 + k;
    if (p < P && in_band(z[p], gt_depth[p / S])) {
      hits |= 1u << k;
      ++cnt;
    }
  }
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (tid == 0) {
    int tot = 0;
    for (int w = 0; w < 8; ++w) {
      const int t = warp_tot[w];
      warp_tot[w] = tot;
      tot += t;
    }
    base_s = tot > 0 ? atomicAdd(counter, tot) : 0;
  }
  __syncthreads();
  int pos = base_s + warp_tot[warp] + incl - cnt;
#pragma unroll
  for (int k = 0; k < kBandPer; ++k)
    if (hits & (1u << k)) rows[pos++] = (int)(p0 + k);
}

// Everything a row needs to build its part of the operand tile.
struct RowGeom {
  bool valid;       // row exists (sample inside the list, view < R)
  bool vis;         // projection inside the image and in front of the camera (common.py:658-660)
  int64_t p, r;     // sample, ray
  int view;         // f * R + v
  float zv;
  float x[3];       // normalised refer_p (OneBlob input)
  const float *f00, *f01, *f10, *f11;
  float wy0, wy1, wx0, wx1;
};

// sample id of row `row` of tile `tile` (-1: the row does not exist); the kernels fetch it ONE TILE AHEAD so that the
// dependent chain rows[] -> z / rays / d_out of the next tile starts with the id already in a register
__device__ __forceinline__ int64_t row_sample(const FmArgs& a, int64_t n_rows, int64_t n_tiles, int64_t tile, int row) {
  if (tile >= n_tiles) return -1;
  const int pslot = row / a.R;
  const int64_t bi = tile * a.PPT + pslot;
  if (pslot >= a.PPT || bi >= n_rows) return -1;
  return a.rows ? (int64_t)a.rows[bi] : bi;
}
template <bool PROJECT = true>
__device__ __forceinline__ void row_geometry(const FmArgs& a, int64_t p_row, int row, const float* sK,
                                             const float* sW2c, const float* sCamO, RowGeom& g) {
  const int pslot = row / a.R, v = row - pslot * a.R;
  g.valid = p_row >= 0;
  g.vis = false;
  g.p = g.r = 0;
  g.view = 0;
  g.zv = 0.f;
  g.x[0] = g.x[1] = g.x[2] = 0.f;
  if (!g.valid) return;
  g.p = p_row;
  g.r = g.p < 0x7fffffffLL ? (int64_t)((uint32_t)g.p / (uint32_t)a.S) : g.p / a.S;   // 32-bit division whenever the index allows
  g.zv = a.z[g.p];
  int f = 0;
  while (f + 1 < a.F && g.r >= a.ray_start[f + 1]) ++f;
  g.view = f * a.R + v;
  float pt[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) pt[c] = __fadd_rn(a.rays_o[3 * g.r + c], __fmul_rn(a.rays_d[3 * g.r + c], g.zv));
  const float* co = sCamO + 3 * g.view;
#pragma unroll
  for (int c = 0; c < 3; ++c) g.x[c] = (float)(((double)__fsub_rn(pt[c], co[c]) - a.B.lo[c]) / a.B.ext[c]);
  if (!PROJECT) return;   // the backward with a stashed X tile needs the OneBlob argument only
  // projection (utils/common.py:648-660), the arithmetic of k_feature_gather (sample.cu)
  const float* M = sW2c + 16 * g.view;
  float cam[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) cam[c] = fmaf(M[4 * c + 2], pt[2], fmaf(M[4 * c + 1], pt[1], fmaf(M[4 * c], pt[0], M[4 * c + 3])));
  cam[1] = -cam[1];
  cam[2] = -cam[2];
  float img[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) img[c] = fmaf(sK[3 * c + 2], cam[2], fmaf(sK[3 * c + 1], cam[1], sK[3 * c] * cam[0]));
  const float den = img[2] + 1e-5f;
  const float u = rintf(img[0] / den), vv = rintf(img[1] / den);
  g.vis = (u > 0.f) && (u < (float)(a.W - 1)) && (vv > 0.f) && (vv < (float)(a.H - 1)) && (cam[2] > 0.f);
  if (g.vis) {
    // F.interpolate(..., mode='bilinear', align_corners=True) sampled at the integer pixel (vi, ui)
    const int ui = (int)u, vi = (int)vv;
    const float sy = a.H > 1 ? (float)(a.h - 1) / (float)(a.H - 1) : 0.f, sx = a.W > 1 ? (float)(a.w - 1) / (float)(a.W - 1) : 0.f;
    const float fy = sy * (float)vi, fx = sx * (float)ui;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < a.h - 1 ? 1 : 0), x1 = x0 + (x0 < a.w - 1 ? 1 : 0);
    g.wy1 = fy - (float)y0;
    g.wy0 = 1.f - g.wy1;
    g.wx1 = fx - (float)x0;
    g.wx0 = 1.f - g.wx1;
    const float* fm = a.feats[f] + (int64_t)v * a.h * a.w * 64;
    g.f00 = fm + ((int64_t)y0 * a.w + x0) * 64;
    g.f01 = fm + ((int64_t)y0 * a.w + x1) * 64;
    g.f10 = fm + ((int64_t)y1 * a.w + x0) * 64;
    g.f11 = fm + ((int64_t)y1 * a.w + x1) * 64;
  }
}

// This thread's share of the row's X = [OneBlob 48 | feature 64]: group 0 the OneBlob chunks 0..5 and feature
// chunks 6..8 (24 channels), group 1 feature chunks 9..13 (40 channels).
// ``img``: this row's slot in the tile image (NULL: no stash), chunk c hi at img[c * 128], lo at img[(14 + c) * 128].
template <bool FEATURES = true>
__device__ __forceinline__ void build_x(const RowGeom& g, int grp, int row, unsigned char* X_hi, unsigned char* X_lo, uint4* img) {
  const uint4 z4 = make_uint4(0, 0, 0, 0);
  const int c0 = grp ? 9 : 6, c1 = grp ? 14 : 9;
  if (!FEATURES) {   // OneBlob chunks only (group 0); the feature chunks come from gather_features_coop
    if (grp == 0) {
#pragma unroll 1   // one copy of the OneBlob + operand-split code (instruction-cache footprint)
      for (int c = 0; c < 3; ++c) {
        float pe[16];
        if (g.valid) oneblob16(c == 0 ? g.x[0] : (c == 1 ? g.x[1] : g.x[2]), pe);
        else {
#pragma unroll
          for (int e = 0; e < 16; ++e) pe[e] = 0.f;
        }
        put_chunk(X_hi, X_lo, 2 * c, 2048, row, pe);
        put_chunk(X_hi, X_lo, 2 * c + 1, 2048, row, pe + 8);
      }
    }
    return;
  }
  if (!g.valid) {
    for (int c = grp ? 9 : 0; c < c1; ++c) {
      *reinterpret_cast<uint4*>(X_hi + c * 2048 + row * 16) = z4;
      *reinterpret_cast<uint4*>(X_lo + c * 2048 + row * 16) = z4;
      if (img) img[c * kTile] = img[(14 + c) * kTile] = z4;
    }
    return;
  }
  if (grp == 0) {
#pragma unroll 1   // one copy of the OneBlob + operand-split code (instruction-cache footprint)
    for (int c = 0; c < 3; ++c) {
      float pe[16];
      oneblob16(c == 0 ? g.x[0] : (c == 1 ? g.x[1] : g.x[2]), pe);
      put_chunk_img(X_hi, X_lo, 2 * c, 2048, row, pe, img ? img + (2 * c) * kTile : nullptr, img ? img + (14 + 2 * c) * kTile : nullptr);
      put_chunk_img(X_hi, X_lo, 2 * c + 1, 2048, row, pe + 8, img ? img + (2 * c + 1) * kTile : nullptr,
                    img ? img + (15 + 2 * c) * kTile : nullptr);
    }
  }
  if (!g.vis) {   // code * mask (common.py:676): a hidden view contributes its OneBlob part only
    for (int c = c0; c < c1; ++c) {
      *reinterpret_cast<uint4*>(X_hi + c * 2048 + row * 16) = z4;
      *reinterpret_cast<uint4*>(X_lo + c * 2048 + row * 16) = z4;
      if (img) img[c * kTile] = img[(14 + c) * kTile] = z4;
    }
    return;
  }
#pragma unroll 1
  for (int c = c0; c < c1; ++c) {
    const int ch = 8 * (c - 6);
    float f[8];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float4 a00 = __ldg(reinterpret_cast<const float4*>(g.f00 + ch) + q);
      const float4 a01 = __ldg(reinterpret_cast<const float4*>(g.f01 + ch) + q);
      const float4 a10 = __ldg(reinterpret_cast<const float4*>(g.f10 + ch) + q);
      const float4 a11 = __ldg(reinterpret_cast<const float4*>(g.f11 + ch) + q);
      f[4 * q + 0] = g.wy0 * (g.wx0 * a00.x + g.wx1 * a01.x) + g.wy1 * (g.wx0 * a10.x + g.wx1 * a11.x);
      f[4 * q + 1] = g.wy0 * (g.wx0 * a00.y + g.wx1 * a01.y) + g.wy1 * (g.wx0 * a10.y + g.wx1 * a11.y);
      f[4 * q + 2] = g.wy0 * (g.wx0 * a00.z + g.wx1 * a01.z) + g.wy1 * (g.wx0 * a10.z + g.wx1 * a11.z);
      f[4 * q + 3] = g.wy0 * (g.wx0 * a00.w + g.wx1 * a01.w) + g.wy1 * (g.wx0 * a10.w + g.wx1 * a11.w);
    }
    put_chunk_img(X_hi, X_lo, c, 2048, row, f, img ? img + c * kTile : nullptr, img ? img + (14 + c) * kTile : nullptr);
  }
}

// Cooperative form of the feature part of build_x for the forward kernel.  build_x lets every thread fetch ITS row's taps:
// the 32 lanes of a load instruction then touch 32 different cache lines (32 L1 wavefronts for 512 useful bytes), and the
// kernel sat on that rate.  Here the row owners publish their tap descriptors in shared memory and each warp walks 16
// rows, four at a time: the 8 lanes of a row read the tap's 64 channels as two contiguous 128-byte lines (lane k: channels
// 4k..4k+3 and 32+4k..32+4k+3), so an instruction touches 4 lines instead of 32.  Neighbouring lanes then swap halves so
// that each owns one 8-channel operand chunk (even k: chunk 6 + k/2, odd k: chunk 10 + k/2).  Same blend expression as
// build_x: the tile is bit identical.
struct TapDesc {
  const float* f00;   // NULL: row absent or view hidden (zeros)
  int dx, dy;         // element offsets of the x1 / y1 taps
  float wx1, wy1;
};
__device__ __forceinline__ void publish_taps(const RowGeom& g, TapDesc* d) {
  d->f00 = (g.valid && g.vis) ? g.f00 : nullptr;
  if (g.valid && g.vis) {
    d->dx = (int)(g.f01 - g.f00);
    d->dy = (int)(g.f10 - g.f00);
    d->wx1 = g.wx1;
    d->wy1 = g.wy1;
  }
}
__device__ __forceinline__ void gather_features_coop(const TapDesc* taps, int row0, int n_it, int lane, unsigned char* X_hi,
                                                     unsigned char* X_lo) {
  const int k = lane & 7;
#pragma unroll 2
  for (int it = 0; it < n_it; ++it) {
    const int row = row0 + 4 * it + (lane >> 3);
    const TapDesc td = taps[row];
    float lo4[4] = {0.f, 0.f, 0.f, 0.f}, hi4[4] = {0.f, 0.f, 0.f, 0.f};   // channels 4k.., 32+4k..
    if (td.f00) {
      const float wx1 = td.wx1, wx0 = 1.f - wx1, wy1 = td.wy1, wy0 = 1.f - wy1;
      const float4* p00 = reinterpret_cast<const float4*>(td.f00) + k;
      const float4* p01 = reinterpret_cast<const float4*>(td.f00 + td.dx) + k;
      const float4* p10 = reinterpret_cast<const float4*>(td.f00 + td.dy) + k;
      const float4* p11 = reinterpret_cast<const float4*>(td.f00 + td.dy + td.dx) + k;
      const float4 a00 = __ldg(p00), a01 = __ldg(p01), a10 = __ldg(p10), a11 = __ldg(p11);
      const float4 b00 = __ldg(p00 + 8), b01 = __ldg(p01 + 8), b10 = __ldg(p10 + 8), b11 = __ldg(p11 + 8);
      lo4[0] = wy0 * (wx0 * a00.x + wx1 * a01.x) + wy1 * (wx0 * a10.x + wx1 * a11.x);
      lo4[1] = wy0 * (wx0 * a00.y + wx1 * a01.y) + wy1 * (wx0 * a10.y + wx1 * a11.y);
      lo4[2] = wy0 * (wx0 * a00.z + wx1 * a01.z) + wy1 * (wx0 * a10.z + wx1 * a11.z);
      lo4[3] = wy0 * (wx0 * a00.w + wx1 * a01.w) + wy1 * (wx0 * a10.w + wx1 * a11.w);
      hi4[0] = wy0 * (wx0 * b00.x + wx1 * b01.x) + wy1 * (wx0 * b10.x + wx1 * b11.x);
      hi4[1] = wy0 * (wx0 * b00.y + wx1 * b01.y) + wy1 * (wx0 * b10.y + wx1 * b11.y);
      hi4[2] = wy0 * (wx0 * b00.z + wx1 * b01.z) + wy1 * (wx0 * b10.z + wx1 * b11.z);
      hi4[3] = wy0 * (wx0 * b00.w + wx1 * b01.w) + wy1 * (wx0 * b10.w + wx1 * b11.w);
    }
    // even k keeps its first set and takes the partner's first set (channels 4k..4k+7 = chunk 6 + k/2); odd k keeps its
    // second set behind the partner's (channels 32+4(k-1)..32+4k+3 = chunk 10 + k/2)
    const bool odd = k & 1;
    float f[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float send = odd ? lo4[e] : hi4[e];
      const float got = __shfl_xor_sync(0xffffffffu, send, 1);
      f[e] = odd ? got : lo4[e];
      f[4 + e] = odd ? hi4[e] : got;
    }
    put_chunk(X_hi, X_lo, (odd ? 10 : 6) + (k >> 1), 2048, row, f);
  }
}

// H = X . W1^T  (M = 128 rows, N = 32, K = 112), issued by one thread
__device__ __forceinline__ void mma_hidden(uint32_t tmem_h, const unsigned char* X_hi, const unsigned char* X_lo,
                                           const unsigned char* W1_hi, const unsigned char* W1_lo) {
  const uint32_t idesc = umma_idesc_bf16(128, 32, 0, 0);
#pragma unroll 1
  for (int ks = 0; ks < 7; ++ks) {
    const uint32_t aoff = ks * 4096, boff = ks * 1024;
    const uint64_t a_hi = umma_desc(smem_u32(X_hi) + aoff, 2048, 128), a_lo = umma_desc(smem_u32(X_lo) + aoff, 2048, 128);
    const uint64_t b_hi = umma_desc(smem_u32(W1_hi) + boff, 512, 128), b_lo = umma_desc(smem_u32(W1_lo) + boff, 512, 128);
    umma_bf16(tmem_h, a_hi, b_hi, idesc, ks > 0 ? 1u : 0u);
    umma_bf16(tmem_h, a_lo, b_hi, idesc, 1u);
    umma_bf16(tmem_h, a_hi, b_lo, idesc, 1u);
  }
}

// views / intrinsics of the call into shared memory (<= 64 views)
__device__ __forceinline__ void load_views(const FmArgs& a, float* sK, float* sW2c, float* sCamO) {
  const int nv = a.F * a.R;
  for (int i = threadIdx.x; i < 9; i += blockDim.x) sK[i] = a.K[i];
  for (int i = threadIdx.x; i < 16 * nv; i += blockDim.x) sW2c[i] = a.w2c[i];
  for (int i = threadIdx.x; i < 3 * nv; i += blockDim.x) sCamO[i] = a.cam_o[i];
}

constexpr int kFmHB = 17408;                                           // H tile (16 KB) / output staging [128][33] floats
constexpr int kFmFwdSmem = 28 * 2048 + kFmHB + kFW * 16;               // X tile | H / staging | weights
constexpr int kFmBwdSmem = 28 * 2048 + 8 * 2048 + 8 * 2048 + kFW * 16;  // X | H | dO / dH | weights

__global__ void __launch_bounds__(kFT, 2) k_featmerge_fwd(FmArgs a) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float sK[9], sW2c[16 * kMaxViews], sCamO[3 * kMaxViews];
  __shared__ TapDesc sTaps[kTile];
  unsigned char* X_hi = sm;
  unsigned char* X_lo = sm + 14 * 2048;
  // H (and, after the second GEMM, the output staging) has a region of its own: the bulk store of the X tile (the image
  // for the backward) then only has to finish before the NEXT tile is built, not before the H epilogue
  unsigned char* H_hi = sm + 28 * 2048;
  unsigned char* H_lo = H_hi + 4 * 2048;
  float* OB = reinterpret_cast<float*>(H_hi);            // [128][33] output rows (H is dead by then)
  unsigned char* Wt = sm + 28 * 2048 + kFmHB;
  unsigned char* W1_hi = Wt;
  unsigned char* W1_lo = Wt + kFW1 * 16;
  unsigned char* W2_hi = Wt + 2 * kFW1 * 16;
  unsigned char* W2_lo = W2_hi + kFW2 * 16;
  const int tid = threadIdx.x, warp = tid >> 5, row = tid & (kTile - 1), grp = tid >> 7;
  const int64_t n_rows = a.n_rows_dev ? (int64_t)*a.n_rows_dev : a.n_rows_host;
  const int64_t n_tiles = (n_rows + a.PPT - 1) / a.PPT;
  if ((int64_t)blockIdx.x >= n_tiles) return;
  for (int i = tid; i < kFW; i += kFT) reinterpret_cast<uint4*>(Wt)[i] = a.wts[i];
  load_views(a, sK, sW2c, sCamO);
  if (warp == 0) tmem_alloc(&tmem_base_s, 64);
  if (tid == 0) mbar_init(&bar, 1);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  const uint32_t lane_addr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
  const float inv_r = 1.f / (float)a.R;
  // the X tiles are kept for the backward when the caller's stash holds all of them (decided on the device: no host sync)
  const bool use_img = a.Ximg != nullptr && n_tiles * (int64_t)kXImgBytes <= a.stash_bytes;
  uint32_t phase = 0;
  DNS_CLK_DECL
  int64_t p_next = row_sample(a, n_rows, n_tiles, blockIdx.x, row);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    RowGeom g;
    const int64_t p_row = p_next;
    p_next = row_sample(a, n_rows, n_tiles, tile + gridDim.x, row);      // one tile ahead
    row_geometry(a, p_row, row, sK, sW2c, sCamO, g);
    if (grp == 1) publish_taps(g, sTaps + row);
    if (tid == 0 && use_img) bulk_wait_read();   // the previous tile's image store has read the X region
    __syncthreads();
    DNS_CLK(a, 1)
    build_x<false>(g, grp, row, X_hi, X_lo, nullptr);            // OneBlob chunks (group 0)
    // feature chunks, coalesced: the warps of group 0 (busy with the OneBlob) take 8 rows each, those of group 1 24
    if (grp == 0) gather_features_coop(sTaps, 8 * warp, 2, tid & 31, X_hi, X_lo);
    else gather_features_coop(sTaps, 32 + 24 * (warp - 4), 6, tid & 31, X_hi, X_lo);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    DNS_CLK(a, 2)
    if (tid == 0) {
      tc_fence_after();
      mma_hidden(tmem_d, X_hi, X_lo, W1_hi, W1_lo);
      // the tile image for the backward = the X tile byte for byte: ONE bulk store, waited for at the top of the next tile
      if (use_img) bulk_s2g(a.Ximg + tile * (kXImgBytes / 16), X_hi, kXImgBytes);
      umma_commit(&bar);
    }
    mbar_wait_cta(&bar, phase);
    DNS_CLK(a, 3)
    phase ^= 1;
    tc_fence_after();
    {   // hidden activations of this thread's 16 units -> H tile
      float v[16];
      tmem_ld16(lane_addr + 16 * grp, v);
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = fmaxf(v[k], 0.f);
      put_chunk(H_hi, H_lo, 2 * grp, 2048, row, v);
      put_chunk(H_hi, H_lo, 2 * grp + 1, 2048, row, v + 8);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    DNS_CLK(a, 4)
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
    mbar_wait_cta(&bar, phase);
    DNS_CLK(a, 5)
    phase ^= 1;
    tc_fence_after();
    {
      float v[16];
      tmem_ld16(lane_addr + 32 + 16 * grp, v);
#pragma unroll
      for (int k = 0; k < 16; ++k) OB[row * 33 + 16 * grp + k] = g.valid ? v[k] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    DNS_CLK(a, 6)
    // mean over the R view rows of a sample (decoder.py:77), one float4 per thread and step
    for (int e = tid; e < a.PPT * 8; e += kFT) {
      const int pslot = e >> 3, q = e & 7;
      const int64_t bi = tile * a.PPT + pslot;
      if (bi < n_rows) {
        const int64_t p = a.rows ? (int64_t)a.rows[bi] : bi;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int v = 0; v < a.R; ++v) {
          const float* o = OB + (pslot * a.R + v) * 33 + 4 * q;
          acc.x += o[0]; acc.y += o[1]; acc.z += o[2]; acc.w += o[3];
        }
        *reinterpret_cast<float4*>(a.out + p * 32 + 4 * q) = make_float4(acc.x * inv_r, acc.y * inv_r, acc.z * inv_r, acc.w * inv_r);
      }
    }
    __syncthreads();   // OB / H are overwritten by the next tile's X
    DNS_CLK(a, 7)
  }
  if (tid == 0 && use_img) bulk_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 64);
}

__global__ void __launch_bounds__(kFT, 2) k_featmerge_bwd(FmArgs a) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar, xbar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float sK[9], sW2c[16 * kMaxViews], sCamO[3 * kMaxViews];
  __shared__ float DXS[kTile * 3];
  // X (56 KB) | H (16 KB) | dO, later dH (16 KB) | weights (18 KB).  The 128-lane MMA footprint of the weight-gradient
  // A operands (X^T: 14 of 16 chunks; H^T: 4 of 16) runs into the regions behind them; those lanes are never read.
  unsigned char* X_hi = sm;
  unsigned char* X_lo = sm + 14 * 2048;
  unsigned char* H_hi = sm + 28 * 2048;
  unsigned char* H_lo = H_hi + 4 * 2048;
  unsigned char* D_hi = H_hi + 8 * 2048;     // dO tile, overwritten by the dH tile
  unsigned char* D_lo = D_hi + 4 * 2048;
  unsigned char* Wt = D_hi + 8 * 2048;
  unsigned char* W1_hi = Wt;
  unsigned char* W1_lo = Wt + kFW1 * 16;
  unsigned char* W2_hi = Wt + 2 * kFW1 * 16;
  unsigned char* W2_lo = W2_hi + kFW2 * 16;
  const int tid = threadIdx.x, warp = tid >> 5, row = tid & (kTile - 1), grp = tid >> 7;
  const int64_t n_rows = a.n_rows_dev ? (int64_t)*a.n_rows_dev : a.n_rows_host;
  const int64_t n_tiles = (n_rows + a.PPT - 1) / a.PPT;
  if ((int64_t)blockIdx.x >= n_tiles) return;
  for (int i = tid; i < kFW; i += kFT) reinterpret_cast<uint4*>(Wt)[i] = a.wts[i];
  load_views(a, sK, sW2c, sCamO);
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init(&xbar, 1);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // TMEM columns: H 0..31 | dH (pre-mask) 32..63 | dX 64..111 | dW2 accumulator 112..143 | dW1 accumulator 144..175
  const uint32_t tmem_d = tmem_base_s;
  const uint32_t lane_addr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
  const float inv_r = 1.f / (float)a.R;
  uint32_t phase = 0, xphase = 0;
  bool have_acc = false;
  // X tiles stashed by the forward pass (same device-side rule): ONE bulk copy per tile instead of the gather + encode
  const bool use_img = a.Ximg != nullptr && n_tiles * (int64_t)kXImgBytes <= a.stash_bytes;
  DNS_CLK_DECL
  int64_t p_next = row_sample(a, n_rows, n_tiles, blockIdx.x, row);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    RowGeom g;
    const int64_t p_row = p_next;
    p_next = row_sample(a, n_rows, n_tiles, tile + gridDim.x, row);      // one tile ahead
    if (use_img) {
      if (tid == 0) {   // the previous tile's MMAs have completed (end-of-loop barrier): the X region is free
        mbar_expect_tx(&xbar, kXImgBytes);
        bulk_g2s(X_hi, a.Ximg + tile * (kXImgBytes / 16), kXImgBytes, &xbar);
      }
      row_geometry<false>(a, p_row, row, sK, sW2c, sCamO, g);
    } else {
      row_geometry(a, p_row, row, sK, sW2c, sCamO, g);
      build_x(g, grp, row, X_hi, X_lo, nullptr);
    }
    {   // dO = d_features[sample] / R (the same for the R views of the sample): 16 channels per thread
      float f[16];
      if (g.valid) {
        const float4* s4 = reinterpret_cast<const float4*>(a.d_out + g.p * 32 + 16 * grp);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 u = s4[q];
          f[4 * q] = u.x * inv_r; f[4 * q + 1] = u.y * inv_r; f[4 * q + 2] = u.z * inv_r; f[4 * q + 3] = u.w * inv_r;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) f[k] = 0.f;
      }
      put_chunk(D_hi, D_lo, 2 * grp, 2048, row, f);
      put_chunk(D_hi, D_lo, 2 * grp + 1, 2048, row, f + 8);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    DNS_CLK(a, 17)
    if (tid == 0) {
      tc_fence_after();
      if (use_img) {
        mbar_wait(&xbar, xphase);   // X tile landed
        xphase ^= 1;
      }
      mma_hidden(tmem_d, X_hi, X_lo, W1_hi, W1_lo);
      umma_commit(&bar);
    }
    mbar_wait_cta(&bar, phase);
    DNS_CLK(a, 18)
    phase ^= 1;
    tc_fence_after();
    float hv[16];   // this thread's 16 hidden activations (kept for the ReLU mask)
    tmem_ld16(lane_addr + 16 * grp, hv);
#pragma unroll
    for (int k = 0; k < 16; ++k) hv[k] = g.valid ? fmaxf(hv[k], 0.f) : 0.f;
    put_chunk(H_hi, H_lo, 2 * grp, 2048, row, hv);
    put_chunk(H_hi, H_lo, 2 * grp + 1, 2048, row, hv + 8);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    DNS_CLK(a, 19)
    if (tid == 0) {
      tc_fence_after();
      {   // dH = dO . W2   (B = W2 tile MN-major: hidden contiguous; LBO 128 over output rows, SBO 512 over hidden chunks)
        const uint32_t idesc = umma_idesc_bf16(128, 32, 0, 1);
#pragma unroll 1
        for (int ks = 0; ks < 2; ++ks) {
          const uint32_t aoff = ks * 4096, boff = ks * 256;
          const uint64_t a_hi = umma_desc(smem_u32(D_hi) + aoff, 2048, 128), a_lo = umma_desc(smem_u32(D_lo) + aoff, 2048, 128);
          const uint64_t b_hi = umma_desc(smem_u32(W2_hi) + boff, 128, 512), b_lo = umma_desc(smem_u32(W2_lo) + boff, 128, 512);
          umma_bf16(tmem_d + 32, a_hi, b_hi, idesc, ks > 0 ? 1u : 0u);
          umma_bf16(tmem_d + 32, a_lo, b_hi, idesc, 1u);
          umma_bf16(tmem_d + 32, a_hi, b_lo, idesc, 1u);
        }
      }
      if (a.need_dparams) {   // dW2[hidden lane][out column] += H^T dO  (both MN-major, K = 16 rows per MMA; tc.cu)
        const uint32_t idesc = umma_idesc_bf16(128, 32, 1, 1);
#pragma unroll 1
        for (int k = 0; k < kTile / 16; ++k) {
          const uint32_t koff = k * 256;
          const uint64_t a_hi = umma_desc(smem_u32(H_hi) + koff, 128, 2048), a_lo = umma_desc(smem_u32(H_lo) + koff, 128, 2048);
          const uint64_t b_hi = umma_desc(smem_u32(D_hi) + koff, 128, 2048), b_lo = umma_desc(smem_u32(D_lo) + koff, 128, 2048);
          umma_bf16(tmem_d + 112, a_hi, b_hi, idesc, (have_acc || k > 0) ? 1u : 0u);
          umma_bf16(tmem_d + 112, a_lo, b_hi, idesc, 1u);
          umma_bf16(tmem_d + 112, a_hi, b_lo, idesc, 1u);
        }
      }
      umma_commit(&bar);
    }
    mbar_wait_cta(&bar, phase);
    DNS_CLK(a, 20)
    phase ^= 1;
    tc_fence_after();
    {   // dH = (dO W2) [H > 0] -> the dH tile takes the place of the dO tile (its MMAs have completed)
      float v[16];
      tmem_ld16(lane_addr + 32 + 16 * grp, v);
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = hv[k] > 0.f ? v[k] : 0.f;
      put_chunk(D_hi, D_lo, 2 * grp, 2048, row, v);
      put_chunk(D_hi, D_lo, 2 * grp + 1, 2048, row, v + 8);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    DNS_CLK(a, 21)
    if (tid == 0) {
      tc_fence_after();
      if (a.need_drays) {   // dX[:, 0..47] = dH . W1[:, 0..47]   (B = W1 tile MN-major: LBO 128 over hidden rows, SBO 512)
        const uint32_t idesc = umma_idesc_bf16(128, 48, 0, 1);
#pragma unroll 1
        for (int ks = 0; ks < 2; ++ks) {
          const uint32_t aoff = ks * 4096, boff = ks * 256;
          const uint64_t a_hi = umma_desc(smem_u32(D_hi) + aoff, 2048, 128), a_lo = umma_desc(smem_u32(D_lo) + aoff, 2048, 128);
          const uint64_t b_hi = umma_desc(smem_u32(W1_hi) + boff, 128, 512), b_lo = umma_desc(smem_u32(W1_lo) + boff, 128, 512);
          umma_bf16(tmem_d + 64, a_hi, b_hi, idesc, ks > 0 ? 1u : 0u);
          umma_bf16(tmem_d + 64, a_lo, b_hi, idesc, 1u);
          umma_bf16(tmem_d + 64, a_hi, b_lo, idesc, 1u);
        }
      }
      if (a.need_dparams) {   // dW1[feature lane][hidden column] += X^T dH
        const uint32_t idesc = umma_idesc_bf16(128, 32, 1, 1);
#pragma unroll 1
        for (int k = 0; k < kTile / 16; ++k) {
          const uint32_t koff = k * 256;
          const uint64_t a_hi = umma_desc(smem_u32(X_hi) + koff, 128, 2048), a_lo = umma_desc(smem_u32(X_lo) + koff, 128, 2048);
          const uint64_t b_hi = umma_desc(smem_u32(D_hi) + koff, 128, 2048), b_lo = umma_desc(smem_u32(D_lo) + koff, 128, 2048);
          umma_bf16(tmem_d + 144, a_hi, b_hi, idesc, (have_acc || k > 0) ? 1u : 0u);
          umma_bf16(tmem_d + 144, a_lo, b_hi, idesc, 1u);
          umma_bf16(tmem_d + 144, a_hi, b_lo, idesc, 1u);
        }
      }
      umma_commit(&bar);
    }
    have_acc = true;
    mbar_wait_cta(&bar, phase);
    DNS_CLK(a, 22)
    phase ^= 1;
    tc_fence_after();
    if (a.need_drays) {
      // OneBlob backward per coordinate (group 0: x, y; group 1: z), summed over the R view rows of the sample, then
      // d(point) -> d_rays_o, d_rays_d (point = o + d z)
      const int c0 = grp ? 2 : 0, c1 = grp ? 3 : 2;
#pragma unroll 1
      for (int c = c0; c < c1; ++c) {
        float v[16];
        tmem_ld16(lane_addr + 64 + 16 * c, v);
        DXS[row * 3 + c] = g.valid ? oneblob16_bwd(c == 0 ? g.x[0] : (c == 1 ? g.x[1] : g.x[2]), v) / (float)a.B.ext[c] : 0.f;
      }
      tc_fence_before();
      __syncthreads();
      for (int e = tid; e < a.PPT * 3; e += kFT) {
        const int pslot = e / 3, c = e - pslot * 3;
        const int64_t bi = tile * a.PPT + pslot;
        if (bi < n_rows) {
          const int64_t p = a.rows ? (int64_t)a.rows[bi] : bi, r = p / a.S;
          float acc = 0.f;
          for (int v = 0; v < a.R; ++v) acc += DXS[(pslot * a.R + v) * 3 + c];
          atomicAdd(a.d_rays_o + 3 * r + c, acc);
          atomicAdd(a.d_rays_d + 3 * r + c, acc * a.z[p]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // accumulator columns read, operand tiles free for the next tile
    DNS_CLK(a, 23)
    tc_fence_after();
  }
  if (a.need_dparams && have_acc) {
    // flush: dW1 accumulator lane = input feature i (< 112), column = hidden unit j -> W1[j][i];
    //        dW2 accumulator lane = hidden unit k (< 32), column = output j -> W2[j][k]
    if (grp == 0) {
#pragma unroll
      for (int g2 = 0; g2 < 2; ++g2) {
        float v[16];
        tmem_ld16(lane_addr + 144 + 16 * g2, v);
        if (row < kIn2) {
#pragma unroll
          for (int k = 0; k < 16; ++k)
            if (v[k] != 0.f) atomicAdd(a.d_params + (16 * g2 + k) * kIn2 + row, v[k]);
        }
      }
    } else if (warp == 4) {   // lanes 0..31
#pragma unroll
      for (int g2 = 0; g2 < 2; ++g2) {
        float v[16];
        tmem_ld16(lane_addr + 112 + 16 * g2, v);
#pragma unroll
        for (int k = 0; k < 16; ++k)
          if (v[k] != 0.f) atomicAdd(a.d_params + 32 * kIn2 + (16 * g2 + k) * 32 + row, v[k]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 256);
}

static int fm_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

struct FmWs {
  int* counter;
  uint4* wts;
  int* rows;
};
static int64_t fm_carve(FmWs& w, char* base, int64_t P) {
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    char* q = base ? base + off : nullptr;
    off += (bytes + 255) & ~(int64_t)255;
    return q;
  };
  w.counter = (int*)take(256);
  w.wts = (uint4*)take(kFW * 16);
  w.rows = (int*)take(P * 4);
  return off + 256;
}

static int fm_fill(FmArgs& m, FmWs& w, const dns_featmerge_args* a, const char* what) {
  const int64_t N = a->n_rays, S = a->n_samples;
  if (N <= 0 || S <= 0 || N * S >= 2147483647LL) {
    set_error("%s: need 0 < n_rays * n_samples < 2^31", what);
    return DNS_ERR_ARG;
  }
  if (a->n_frames < 1 || a->n_frames > DNS_MAX_FRAMES || a->n_views < 1 || a->n_views > 8) {
    set_error("%s: 1 <= n_frames <= %d, 1 <= n_views <= 8 (got %d, %d)", what, DNS_MAX_FRAMES, a->n_frames, a->n_views);
    return DNS_ERR_UNSUPPORTED;
  }
  if (a->n_frames * a->n_views > kMaxViews) {
    set_error("%s: at most %d views in one call", what, kMaxViews);
    return DNS_ERR_UNSUPPORTED;
  }
  if (a->ray_start[0] != 0 || a->ray_start[a->n_frames] != N) {
    set_error("%s: ray_start must run from 0 to n_rays", what);
    return DNS_ERR_ARG;
  }
  if (!a->workspace || a->workspace_bytes < dns_featmerge_workspace_bytes(a->n_rays, a->n_samples)) {
    set_error("%s: workspace too small", what);
    return DNS_ERR_ARG;
  }
  if (((uintptr_t)a->params & 15)) {
    set_error("%s: params must be 16-byte aligned", what);
    return DNS_ERR_ARG;
  }
  memset(&m, 0, sizeof(m));
  fm_carve(w, (char*)a->workspace, N * S);
  m.S = (int)S; m.R = a->n_views; m.F = a->n_frames; m.PPT = kTile / a->n_views;
  for (int f = 0; f <= a->n_frames; ++f) m.ray_start[f] = a->ray_start[f];
  m.H = a->H; m.W = a->W; m.h = a->h; m.w = a->w; m.P = N * S;
  for (int c = 0; c < 3; ++c) {
    m.B.lo[c] = a->bound[c][0];
    m.B.ext[c] = a->bound[c][1] - a->bound[c][0];
  }
  m.K = a->K; m.w2c = a->w2c; m.cam_o = a->cam_o;
  for (int f = 0; f < a->n_frames; ++f) {
    if (((uintptr_t)a->feats[f] & 15) || !a->feats[f]) {
      set_error("%s: feature maps must be 16-byte aligned device pointers", what);
      return DNS_ERR_ARG;
    }
    m.feats[f] = a->feats[f];
  }
  m.rays_o = a->rays_o; m.rays_d = a->rays_d; m.z = a->z_vals;
#ifdef DNS_ABLATE
  if (const char* e = getenv("DNS_PHASE_CLK_FM")) m.phase_clk = (unsigned long long*)strtoull(e, nullptr, 0);   // device pointer
#endif
  m.rows = a->apply_trunc ? w.rows : nullptr;
  m.n_rows_dev = a->apply_trunc ? w.counter : nullptr;
  m.n_rows_host = N * S;
  m.wts = w.wts;
  if (a->stash && a->stash_bytes >= kXImgBytes && ((uintptr_t)a->stash & 15) == 0) {
    m.Ximg = (uint4*)a->stash;
    m.stash_bytes = a->stash_bytes;
  }
  return DNS_OK;
}

}  // namespace dns

using namespace dns;

extern "C" {

int64_t dns_featmerge_workspace_bytes(int n_rays, int n_samples) {
  FmWs w;
  return fm_carve(w, nullptr, (int64_t)(n_rays < 1 ? 1 : n_rays) * (n_samples < 1 ? 1 : n_samples));
}

int dns_featmerge_fwd(const dns_featmerge_args* a, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  FmArgs m;
  FmWs w;
  if (int e = fm_fill(m, w, a, "featmerge_fwd")) return e;
  if (!a->features) {
    set_error("featmerge_fwd: features output is NULL");
    return DNS_ERR_ARG;
  }
  m.out = a->features;
  cudaFuncSetAttribute(k_featmerge_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, kFmFwdSmem);
  PhaseScope ph(phFeature, st, a->apply_trunc ? 5 : 2);
  k_prep_featmerge<<<1, 128, 0, st>>>(a->params, w.wts);
  const int64_t P = m.P;
  if (a->apply_trunc) {
    cudaMemsetAsync(w.counter, 0, sizeof(int), st);
    if (!a->no_zero_fill) cudaMemsetAsync(a->features, 0, sizeof(float) * P * 32, st);
    const int64_t blocks = (P + 256 * kBandPer - 1) / (256 * kBandPer);
    k_band_rows<<<(unsigned)blocks, 256, 0, st>>>(a->z_vals, a->gt_depth, P, m.S, w.rows, w.counter);
  }
  const int64_t tiles_max = (P + m.PPT - 1) / m.PPT;
  const int64_t grid = tiles_max < 2 * fm_sm_count() ? tiles_max : 2 * fm_sm_count();
  k_featmerge_fwd<<<(unsigned)grid, kFT, kFmFwdSmem, st>>>(m);
  return check_launch("featmerge_fwd");
}

int dns_featmerge_bwd(const dns_featmerge_args* a, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  FmArgs m;
  FmWs w;
  if (int e = fm_fill(m, w, a, "featmerge_bwd")) return e;
  m.need_dparams = a->need_dparams && a->d_params;
  m.need_drays = a->need_drays && a->d_rays_o && a->d_rays_d;
  if (!a->d_features) {
    set_error("featmerge_bwd: d_features is NULL");
    return DNS_ERR_ARG;
  }
  if (!m.need_dparams && !m.need_drays) return DNS_OK;
  m.d_out = a->d_features; m.d_params = a->d_params; m.d_rays_o = a->d_rays_o; m.d_rays_d = a->d_rays_d;
  cudaFuncSetAttribute(k_featmerge_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, kFmBwdSmem);
  PhaseScope ph(phFeature, st, 1);
  // the row list and the prepared weights in the workspace are those of the forward call
  const int64_t tiles_max = (m.P + m.PPT - 1) / m.PPT;
  const int64_t grid = tiles_max < 2 * fm_sm_count() ? tiles_max : 2 * fm_sm_count();
  k_featmerge_bwd<<<(unsigned)grid, kFT, kFmBwdSmem, st>>>(m);
  return check_launch("featmerge_bwd");
}

}  // extern "C"
