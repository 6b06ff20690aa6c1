// Shared device helpers of the dns_slam_b200 kernels (sm_100a).
//
// Encodings follow the tiny-cuda-nn algorithms the reference reaches through
// models/pos_encoding.py:31-46 (HashGrid) and :61-71 (OneBlob); MLPs follow
// models/decoder.py:58-65 (bias-free 1-hidden-layer ReLU networks of width 32).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dns_slam_b200.h"

namespace dns {

constexpr int kTile = 128;   // slots per point tile / threads per point CTA
constexpr int kXld = 81;     // odd leading dimension of the [slot][80] smem tile: conflict free
constexpr int kIn1 = 80;     // OneBlob 48 + grid 32
constexpr int kIn2 = 112;    // OneBlob 48 + latent 32 + pixel feature 32
constexpr int kOutP = 36;    // 33 latent channels padded to a multiple of 4
constexpr int kNetT = kIn1 * 32 + 32 * kOutP;  // transposed coarse/expert block in the workspace

void set_error(const char* fmt, ...);
int check_launch(const char* what);

// Phase accounting: kernel-launch counters (always) and CUDA-event timing on the launching stream
// (when enabled through dns_profile_enable) -- this is what bench.py's roofline numbers come from.
enum Phase { phPrep = 0, phClassPrep, phPointFwd, phRay, phPointBwd, phDwGemm, phFinalize, phAdam, phTvFwd,
             phTvBwd, phSample, phFeature, phOps, phCount };
struct PhaseScope {
  int phase;
  cudaStream_t st;
  int slot;
  PhaseScope(int phase, cudaStream_t st, int n_launches);
  ~PhaseScope();
};

// tcgen05 weight-gradient GEMM (tc.cu):  out[l][c] += sum_p L[p][l] * Cc[p][c]
struct DwArgs {
  const float* L;   // lane-side operand [rows][ldl] (nL <= 128)
  int ldl, nL;
  const float* Cc;  // column-side operand [rows][ldcc] (nC <= 128)
  int ldcc, nC;
  int64_t n_rows;
  const int* n_tiles_dev;
  int n_tiles_host;
  const int* tile_class;
  // output 0 takes columns [0, split), output 1 columns [split, nC); element (l, c) -> o[l*sl + c*sc]
  float* out0;
  float* out1;
  int split;
  int64_t sl0, sc0, cls0, sl1, sc1, cls1;
  int l_f16, c_f16;   // operand staged as fp16 hi/lo instead of bf16 hi/lo (test entry dns_debug_gemm_fmt)
};
int launch_dw_gemm_tc2(DwArgs a, cudaStream_t st);

// Operand already stored as a bf16 hi/lo UMMA tile IMAGE: [sub-tile][half: hi, lo][chunk][RS rows] of 16-byte
// (8 x bf16) elements; a sub-tile is RS consecutive rows (points), RS % 16 == 0.
struct DwImg {
  const uint4* ptr;
  int chunks_total;  // chunks per half in the image
  int chunk0;        // first chunk used by this GEMM
  int chunks_used;
  int n_valid;       // valid features (lanes / columns) among chunks_used * 8
};
struct DwImgArgs {
  DwImg L, Cc;       // lane-side / column-side operand
  int RS;            // rows per sub-tile
  int subs_per_tile; // sub-tiles per class tile (tile_class is indexed by tile)
  const int* n_tiles_dev;
  int n_tiles_host;  // class tiles (NOT sub-tiles)
  const int* tile_class;
  float* out0;
  float* out1;
  int split;
  int64_t sl0, sc0, cls0, sl1, sc1, cls1;
  // Block-diagonal mode (diag_l > 0): the lanes are two blocks of diag_l features, the columns two blocks of diag_c; only
  // the diagonal blocks are kept -- (lane block 0) x (column block 0) -> out0, block 1 x block 1 -> out1, element
  // (l mod diag_l, c mod diag_c); L.n_valid / Cc.n_valid count the valid features PER BLOCK.  Two small GEMMs that share
  // nothing but the tile loop (dOut^T H of the coarse net and of the class expert) then cost ONE pass: the kernel's time
  // per tile is its fixed MMA / barrier chain, not the operand bytes.
  int diag_l, diag_c;
  // Second operand pair (L2.ptr != NULL): its lanes / columns follow those of L / Cc in the stage, and the product keeps
  // two blocks -- L x Cc -> out0 / out1 as usual, L2 x C2 -> out2 (element (l, c) -> out2[l*sl2 + c*sc2]); the cross
  // blocks are dropped.  Same sub-tiling (RS) for all four images.
  DwImg L2, C2;
  float* out2;
  int64_t sl2, sc2;
};
int launch_dw_img(const DwImgArgs& a, cudaStream_t st);

// Function attributes (dynamic shared-memory limit, carve-out) are per DEVICE: a guard that is true the first time it
// is asked on each device of the process (up to 64), so that a second GPU in the same process gets them too.
inline bool first_call_on_device(unsigned long long& seen) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (seen & bit) return false;
  seen |= bit;
  return true;
}
inline int current_device_slot() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev & 15;
}

// Ablation switches exist only in -DDNS_ABLATE builds (scratch measurements); release kernels carry none.
#ifdef DNS_ABLATE
#define DNS_DBG(a) ((a).dbg)
#define DNS_CLK_DECL long long _clk_prev = clock64();
#define DNS_CLK(a, i)                                                                          \
  if ((a).phase_clk && threadIdx.x == 0) {                                                     \
    const long long _now = clock64();                                                          \
    atomicAdd((a).phase_clk + (i), (unsigned long long)(_now - _clk_prev));                    \
    _clk_prev = _now;                                                                          \
  }
#else
#define DNS_DBG(a) 0
#define DNS_CLK_DECL
#define DNS_CLK(a, i)
#endif

// ---------------------------------------------------------------------------------------
// OneBlob (16-bin periodic quartic kernel)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float q_cdf(float u) {
  float u2 = u * u, u4 = u2 * u2;
  float v = (15.0f / 16.0f) * u * (1.0f - (2.0f / 3.0f) * u2 + (1.0f / 5.0f) * u4) + 0.5f;
  return fminf(fmaxf(v, 0.0f), 1.0f);
}
__device__ __forceinline__ float q_pdf(float u) {  // d q_cdf / du (0 where the clamp is active)
  float t = 1.0f - u * u;
  return fabsf(u) < 1.0f ? (15.0f / 16.0f) * t * t : 0.0f;
}
__device__ __forceinline__ float cdf3(float d, float nb) {
  return q_cdf(d * nb) + q_cdf((d - 1.0f) * nb) + q_cdf((d + 1.0f) * nb);
}
__device__ __forceinline__ float pdf3(float d, float nb) {
  return q_pdf(d * nb) + q_pdf((d - 1.0f) * nb) + q_pdf((d + 1.0f) * nb);
}
// out[b] = cdf3((b+1)/nb - x) - cdf3(b/nb - x), b = 0..nb-1, written with stride `st`
__device__ __forceinline__ void oneblob_fwd(float x, int nb, float* out, int st) {
  float fnb = (float)nb;
  float prev = cdf3(0.0f - x, fnb);
  for (int b = 0; b < nb; ++b) {
    float cur = cdf3((float)(b + 1) / fnb - x, fnb);
    out[b * st] = cur - prev;
    prev = cur;
  }
}
// dL/dx = sum_b d_out[b] * -(nb) * (pdf3(right - x) - pdf3(left - x))
__device__ __forceinline__ float oneblob_bwd(float x, int nb, const float* d_out, int st) {
  float fnb = (float)nb;
  float prev = pdf3(0.0f - x, fnb);
  float acc = 0.0f;
  for (int b = 0; b < nb; ++b) {
    float cur = pdf3((float)(b + 1) / fnb - x, fnb);
    acc += d_out[b * st] * (cur - prev);
    prev = cur;
  }
  return -fnb * acc;
}

// ---------------------------------------------------------------------------------------
// Hash grid
// ---------------------------------------------------------------------------------------
// pos = fmaf(scale, x, 0.5) emulated through double so that CPU oracle and GPU agree bit for
// bit (the double product of two floats is exact).
__device__ __forceinline__ void grid_pos(float x, float scale, uint32_t& g, float& w) {
  double pd = (double)x * (double)scale + 0.5;
  float pos = (float)pd;
  float fl = floorf(pos);
  g = (uint32_t)(int)fl;
  w = pos - fl;
}
__device__ __forceinline__ uint32_t corner_index(const dns_grid& G, int l, uint32_t cx, uint32_t cy,
                                                 uint32_t cz) {
  uint32_t size = G.size[l], idx;
  if (G.hashed[l]) {
    idx = (cx ^ (cy * 2654435761u) ^ (cz * 805459861u)) & (size - 1u);  // size == 2^log2_T
  } else {
    uint32_t res = G.res[l];
    idx = cx + cy * res + cz * (res * res);
    if (idx >= size) idx %= size;
  }
  return idx + G.offset[l];
}

// The eight corner indices of cell g at level l (corner c = (c & 1, (c >> 1) & 1, c >> 2)), bit for bit what corner_index
// gives corner by corner.  Dense levels: corner = base + delta (mod 2^32) with base = gx + gy res + gz res^2, so a cell
// whose far corner stays below the level size needs no modulo at all, and a cell outside the bound (the TV lattice of
// mapping.py:129-159 reaches far beyond it: ncu showed 16 % of the TV forward's instructions in `idx %= size`) needs ONE
// modulo for the base and a conditional subtraction per corner (delta <= 1 + res + res^2 < size).  Only a base within delta
// of 2^32 (slightly negative cells) takes the per-corner form, where the 32-bit wrap-around matters.
__device__ __forceinline__ void corner_indices8(const dns_grid& G, int l, const uint32_t g[3], uint32_t (&idx)[8]) {
  const uint32_t size = G.size[l], off = G.offset[l];
  if (G.hashed[l]) {
    const uint32_t m = size - 1u;   // size == 2^log2_T
    const uint32_t y0 = g[1] * 2654435761u, y1 = y0 + 2654435761u, z0 = g[2] * 805459861u, z1 = z0 + 805459861u;
#pragma unroll
    for (int c = 0; c < 8; ++c) idx[c] = (((g[0] + (c & 1)) ^ ((c & 2) ? y1 : y0) ^ ((c & 4) ? z1 : z0)) & m) + off;
  } else {
    const uint32_t res = G.res[l], r2 = res * res;
    const uint32_t base = g[0] + g[1] * res + g[2] * r2, dmax = 1u + res + r2;
    if (base <= 0xffffffffu - dmax && dmax < size) {
      uint32_t r0 = base;
      if (base + dmax >= size) r0 = base % size;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t v = r0 + (c & 1) + ((c & 2) ? res : 0u) + ((c & 4) ? r2 : 0u);
        if (v >= size) v -= size;
        idx[c] = v + off;
      }
    } else {
#pragma unroll 1   // rare (cells within delta of the 2^32 wrap): one copy of the modulo, idx[] written by an unrolled select
      for (int c = 0; c < 8; ++c) {
        uint32_t v = base + (c & 1) + ((c & 2) ? res : 0u) + ((c & 4) ? r2 : 0u);
        if (v >= size) v %= size;
        v += off;
#pragma unroll
        for (int j = 0; j < 8; ++j) idx[j] = j == c ? v : idx[j];
      }
    }
  }
}

// forward of all levels for one point; out[2l+f] written with stride st
__device__ __forceinline__ void hashgrid_fwd(const dns_grid& G, const float2* __restrict__ table,
                                             const float x[3], float* out, int st) {
#pragma unroll 8
  for (int l = 0; l < G.n_levels; ++l) {
    uint32_t g[3];
    float w[3];
    float sc = G.scale[l];
    grid_pos(x[0], sc, g[0], w[0]);
    grid_pos(x[1], sc, g[1], w[1]);
    grid_pos(x[2], sc, g[2], w[2]);
    float2 v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
      v[c] = __ldg(table + corner_index(G, l, g[0] + (c & 1), g[1] + ((c >> 1) & 1), g[2] + (c >> 2)));
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float wt = ((c & 1) ? w[0] : 1.f - w[0]) * ((c & 2) ? w[1] : 1.f - w[1]) * ((c & 4) ? w[2] : 1.f - w[2]);
      a0 += wt * v[c].x;
      a1 += wt * v[c].y;
    }
    out[(2 * l) * st] = a0;
    out[(2 * l + 1) * st] = a1;
  }
}

// backward of all levels for one point: scatter into d_table (if non-null) and return dL/dx
// (if want_dx).  d_out[2l+f] read with stride st.
__device__ __forceinline__ void hashgrid_bwd(const dns_grid& G, const float2* __restrict__ table,
                                             float2* d_table, const float x[3], const float* d_out,
                                             int st, bool want_dx, float dx[3]) {
  dx[0] = dx[1] = dx[2] = 0.f;
#pragma unroll 1
  for (int l = 0; l < G.n_levels; ++l) {
    float g0 = d_out[(2 * l) * st], g1 = d_out[(2 * l + 1) * st];
    if (g0 == 0.f && g1 == 0.f) continue;
    uint32_t g[3];
    float w[3];
    float sc = G.scale[l];
    grid_pos(x[0], sc, g[0], w[0]);
    grid_pos(x[1], sc, g[1], w[1]);
    grid_pos(x[2], sc, g[2], w[2]);
    uint32_t idx[8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
      idx[c] = corner_index(G, l, g[0] + (c & 1), g[1] + ((c >> 1) & 1), g[2] + (c >> 2));
    if (d_table) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float wt = ((c & 1) ? w[0] : 1.f - w[0]) * ((c & 2) ? w[1] : 1.f - w[1]) * ((c & 4) ? w[2] : 1.f - w[2]);
        atomicAdd(d_table + idx[c], make_float2(wt * g0, wt * g1));
      }
    }
    if (want_dx) {
      float s[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float2 v = __ldg(table + idx[c]);
        s[c] = v.x * g0 + v.y * g1;
      }
      float wx0 = 1.f - w[0], wy0 = 1.f - w[1], wz0 = 1.f - w[2];
      // d/dw_x: corners differing in bit 0
      dx[0] += sc * (wy0 * wz0 * (s[1] - s[0]) + w[1] * wz0 * (s[3] - s[2]) + wy0 * w[2] * (s[5] - s[4]) +
                     w[1] * w[2] * (s[7] - s[6]));
      dx[1] += sc * (wx0 * wz0 * (s[2] - s[0]) + w[0] * wz0 * (s[3] - s[1]) + wx0 * w[2] * (s[6] - s[4]) +
                     w[0] * w[2] * (s[7] - s[5]));
      dx[2] += sc * (wx0 * wy0 * (s[4] - s[0]) + w[0] * wy0 * (s[5] - s[1]) + wx0 * w[1] * (s[6] - s[2]) +
                     w[0] * w[1] * (s[7] - s[3]));
    }
  }
}

// ---------------------------------------------------------------------------------------
// point geometry (slams/tracking.py:160,189-190; slams/mapping.py:531,604-608)
// ---------------------------------------------------------------------------------------
struct Bound {
  double lo[3];
  double ext[3];  // hi - lo, float64
};
// pts = o + d * z in fp32 (separate multiply and add, as torch evaluates it), then
// x = float((double(pts) - lo) / (hi - lo))
__device__ __forceinline__ void point_from_ray(const float* __restrict__ o, const float* __restrict__ d,
                                               float z, const Bound& B, float x[3]) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float pt = __fadd_rn(o[a], __fmul_rn(d[a], z));
    x[a] = (float)(((double)pt - B.lo[a]) / B.ext[a]);
  }
}

// Truncation band of tracking.py:167-170: front = z < 0.95 d, back = z > 1.05 d, keep = !front & !back & d > 0.
__device__ __forceinline__ bool in_band(float zv, float d) {
  return !(zv < __fmul_rn(d, 0.95f)) && !(zv > __fmul_rn(d, 1.05f)) && d > 0.f;
}

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + __expf(-v)); }

// ---------------------------------------------------------------------------------------
// thread-per-point MLP pieces; weights in shared memory, k-major (W1T [K][32]) so that
// every lane reads the same 16 bytes (broadcast), activations in registers.
// ---------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void zero(float (&v)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = 0.f;
}
// h[0..31] += sum_k x[k*xs] * W1T[(k0+k)*ld + col0 + j]
template <int NOUT>
__device__ __forceinline__ void accum_layer(const float* __restrict__ x, int xs, int K,
                                            const float* __restrict__ WT, int ld, float (&h)[NOUT]) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    float xv = x[k * xs];
    const float4* w = reinterpret_cast<const float4*>(WT + k * ld);
#pragma unroll
    for (int q = 0; q < NOUT / 4; ++q) {
      float4 v = w[q];
      h[4 * q + 0] = fmaf(xv, v.x, h[4 * q + 0]);
      h[4 * q + 1] = fmaf(xv, v.y, h[4 * q + 1]);
      h[4 * q + 2] = fmaf(xv, v.z, h[4 * q + 2]);
      h[4 * q + 3] = fmaf(xv, v.w, h[4 * q + 3]);
    }
  }
}
// dot of a register vector with one row of a k-major weight matrix: sum_j v[j] * W[row*ld + j]
template <int N>
__device__ __forceinline__ float dot_row(const float (&v)[N], const float* __restrict__ Wrow) {
  const float4* w = reinterpret_cast<const float4*>(Wrow);
  float acc = 0.f;
#pragma unroll
  for (int q = 0; q < N / 4; ++q) {
    float4 t = w[q];
    acc = fmaf(v[4 * q + 0], t.x, acc);
    acc = fmaf(v[4 * q + 1], t.y, acc);
    acc = fmaf(v[4 * q + 2], t.z, acc);
    acc = fmaf(v[4 * q + 3], t.w, acc);
  }
  return acc;
}

__device__ __forceinline__ float block_reduce_sum(float v, float* red /*[32]*/) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  v = (l < nw) ? red[l] : 0.f;
  if (w == 0)
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;  // valid in warp 0
}

}  // namespace dns
