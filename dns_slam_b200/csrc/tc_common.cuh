// Helpers shared by the tcgen05 point / ray kernels: unrolled 16-bin OneBlob in registers and the
// bf16 hi/lo chunk writer of the canonical UMMA operand layout  [chunk][point][8 x bf16].
#pragma once
#include "render.cuh"
#include "umma.cuh"

namespace dns {

// The quartic kernel has support +-1/16, so only the bins floor(16x)-1 .. floor(16x)+1 (periodic) are
// non-zero: those three are evaluated with the very formula of oneblob_fwd (identical values), the other
// thirteen are exact zeros there too (saturated CDFs cancel).  x far outside [0,1] takes the dense path.
__device__ __forceinline__ void oneblob16(float x, float (&o)[16]) {
  if (x < -0.5f || x > 1.5f) {
    // cold path, ROLLED (same operations in the same order): unrolled, its 51 quartic CDFs were ~800 instructions per
    // coordinate in every kernel that encodes a point -- a third of k_ray_tc2's 165 KB of SASS for a path that only the
    // out-of-bound part of the TV lattice takes.  The bin is written by an unrolled select so that o[] stays in registers.
    float prev = cdf3(0.0f - x, 16.0f);
#pragma unroll 1
    for (int b = 0; b < 16; ++b) {
      const float cur = cdf3((float)(b + 1) / 16.0f - x, 16.0f);
      const float v = cur - prev;
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = j == b ? v : o[j];
      prev = cur;
    }
    return;
  }
  const int bx = (int)floorf(x * 16.0f);
  int bb[3];
  float vv[3];
  // the upper edge of a bin is the lower edge of the next one unless the bins wrap around (15 -> 0): the same expression on
  // the same operands, so the value is reused (4 CDF triples per coordinate instead of 6)
  float up_prev = 0.f;
  int b_prev = -2;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int b = (bx + k - 1) & 15;
    bb[k] = b;
    float lo;
    if (b == b_prev + 1) lo = up_prev;
    else lo = cdf3((float)b / 16.0f - x, 16.0f);
    const float up = cdf3((float)(b + 1) / 16.0f - x, 16.0f);
    vv[k] = up - lo;
    up_prev = up;
    b_prev = b;
  }
#pragma unroll
  for (int b = 0; b < 16; ++b) o[b] = b == bb[0] ? vv[0] : (b == bb[1] ? vv[1] : (b == bb[2] ? vv[2] : 0.0f));
}
__device__ __forceinline__ float oneblob16_bwd(float x, const float (&d)[16]) {
  if (x < -0.5f || x > 1.5f) {   // cold path, rolled (see oneblob16)
    float prev = pdf3(0.0f - x, 16.0f), acc = 0.f;
#pragma unroll 1
    for (int b = 0; b < 16; ++b) {
      const float cur = pdf3((float)(b + 1) / 16.0f - x, 16.0f);
      float db = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) db = j == b ? d[j] : db;
      acc += db * (cur - prev);
      prev = cur;
    }
    return -16.0f * acc;
  }
  const int bx = (int)floorf(x * 16.0f);
  float acc = 0.f, up_prev = 0.f;
  int b_prev = -2;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int b = (bx + k - 1) & 15;
    float db = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) db = j == b ? d[j] : db;
    float lo;   // shared bin edge, see oneblob16
    if (b == b_prev + 1) lo = up_prev;
    else lo = pdf3((float)b / 16.0f - x, 16.0f);
    const float up = pdf3((float)(b + 1) / 16.0f - x, 16.0f);
    acc += db * (up - lo);
    up_prev = up;
    b_prev = b;
  }
  return -16.0f * acc;
}
// Table entries of the x and x+1 corners of one cell edge.  They are neighbours in memory (dense levels: consecutive
// indices; hashed levels: the x prime is 1, so an even x only flips bit 0); when they share an aligned 16-byte pair,
// one 16-byte load fetches both (a quarter fewer L1 / L2 requests over the 8 corners).
__device__ __forceinline__ void load_corner_pair(const float2* __restrict__ table, uint32_t i0, uint32_t i1, float2& v0,
                                                 float2& v1) {
  if ((i0 ^ i1) == 1u) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(table + (i0 & ~1u)));
    const bool odd = i0 & 1u;
    v0 = odd ? make_float2(q.z, q.w) : make_float2(q.x, q.y);
    v1 = odd ? make_float2(q.x, q.y) : make_float2(q.z, q.w);
  } else {
    v0 = __ldg(table + i0);
    v1 = __ldg(table + i1);
  }
}
// Rolled variant for the forward tile: each level's feature pair goes straight into the bf16 hi / lo operand tile
// (4 bytes at  chunk 6 + l/4, row, element pair l%4), so no per-thread feature array exists and the loop body is
// emitted twice instead of eight times (the fully unrolled form showed instruction-fetch stalls in ncu).
// `jimg` != nullptr: the level's Jacobian d(feature pair) / d(x) -- six floats, scale included -- goes into the thread's
// slot-order Jacobian image (float4 chunk k of row `row` at jimg[k * 128]; two levels = three chunks), from which the
// backward kernel forms dL/dx = sum_l g_l . J_l with 12 coalesced 16-byte loads per thread instead of re-reading the
// 64 corners (one L1 wavefront per LANE: the unrelated points of a warp share no cache line).
// The level range is a RUN-TIME argument (l0 even, eight levels): the two thread groups of a CTA run the same copy of the
// loop body instead of one template instance each (half the instruction-cache footprint of the kernel's hottest loop).
__device__ __forceinline__ void hashgrid_fwd_to_tile(const dns_grid& G, const float2* __restrict__ table, const float x[3],
                                                     unsigned char* X_hi, unsigned char* X_lo, int row, int l0,
                                                     float4* jimg = nullptr) {
  float jj[12];
#pragma unroll 1
  for (int lb = l0; lb < l0 + 8; lb += 2)
#pragma unroll
  for (int l = lb; l < lb + 2; ++l) {
    uint32_t g[3];
    float w[3];
    const float sc = G.scale[l];
    grid_pos(x[0], sc, g[0], w[0]);
    grid_pos(x[1], sc, g[1], w[1]);
    grid_pos(x[2], sc, g[2], w[2]);
    float2 v[8];
    uint32_t ci[8];
    corner_indices8(G, l, g, ci);
#pragma unroll
    for (int c = 0; c < 8; c += 2) load_corner_pair(table, ci[c], ci[c + 1], v[c], v[c + 1]);
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float wt = ((c & 1) ? w[0] : 1.f - w[0]) * ((c & 2) ? w[1] : 1.f - w[1]) * ((c & 4) ? w[2] : 1.f - w[2]);
      a0 += wt * v[c].x;
      a1 += wt * v[c].y;
    }
    // fp16 hi / lo halves: the tile feeds the GEMM whose result decides the ReLU (22 mantissa bits, see put_chunk_f16_img)
    const __half2 hh = __floats2half2_rn(a0, a1);
    const float2 back = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(a0 - back.x, a1 - back.y);
    const int off = (6 + (l >> 2)) * 2048 + row * 16 + (l & 3) * 4;
    *reinterpret_cast<__half2*>(X_hi + off) = hh;
    *reinterpret_cast<__half2*>(X_lo + off) = ll;
    if (jimg) {
      const float wx0 = 1.f - w[0], wy0 = 1.f - w[1], wz0 = 1.f - w[2];
      float* j6 = jj + 6 * (l - lb);     // (the loop is unrolled by two: the parity is a compile-time constant)
      // the same corner differences and weights as the backward's dL/dx (hashgrid_bwd_levels), per feature
      j6[0] = sc * (wy0 * wz0 * (v[1].x - v[0].x) + w[1] * wz0 * (v[3].x - v[2].x) + wy0 * w[2] * (v[5].x - v[4].x) + w[1] * w[2] * (v[7].x - v[6].x));
      j6[1] = sc * (wx0 * wz0 * (v[2].x - v[0].x) + w[0] * wz0 * (v[3].x - v[1].x) + wx0 * w[2] * (v[6].x - v[4].x) + w[0] * w[2] * (v[7].x - v[5].x));
      j6[2] = sc * (wx0 * wy0 * (v[4].x - v[0].x) + w[0] * wy0 * (v[5].x - v[1].x) + wx0 * w[1] * (v[6].x - v[2].x) + w[0] * w[1] * (v[7].x - v[3].x));
      j6[3] = sc * (wy0 * wz0 * (v[1].y - v[0].y) + w[1] * wz0 * (v[3].y - v[2].y) + wy0 * w[2] * (v[5].y - v[4].y) + w[1] * w[2] * (v[7].y - v[6].y));
      j6[4] = sc * (wx0 * wz0 * (v[2].y - v[0].y) + w[0] * wz0 * (v[3].y - v[1].y) + wx0 * w[2] * (v[6].y - v[4].y) + w[0] * w[2] * (v[7].y - v[5].y));
      j6[5] = sc * (wx0 * wy0 * (v[4].y - v[0].y) + w[0] * wy0 * (v[5].y - v[1].y) + wx0 * w[1] * (v[6].y - v[2].y) + w[0] * w[1] * (v[7].y - v[3].y));
      if (l - lb) {
        float4* dst = jimg + (3 * ((lb - l0) >> 1)) * 128;
        dst[0] = make_float4(jj[0], jj[1], jj[2], jj[3]);
        dst[128] = make_float4(jj[4], jj[5], jj[6], jj[7]);
        dst[256] = make_float4(jj[8], jj[9], jj[10], jj[11]);
      }
    }
  }
}
// dL/dx of levels L0..L1-1 from the Jacobian image of the forward pass (see hashgrid_fwd_to_tile): dg = dL/d(feature pairs)
template <int NL>
__device__ __forceinline__ void hashgrid_dx_from_jimg(const float4* __restrict__ jimg, const float (&dg)[2 * NL], float dx[3]) {
  static_assert(NL % 2 == 0, "two levels per three chunks");
  dx[0] = dx[1] = dx[2] = 0.f;
#pragma unroll
  for (int p = 0; p < NL / 2; ++p) {
    const float4 a = jimg[(3 * p) * 128], b = jimg[(3 * p + 1) * 128], c = jimg[(3 * p + 2) * 128];
    const float j[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float g0 = dg[2 * (2 * p + h)], g1 = dg[2 * (2 * p + h) + 1];
#pragma unroll
      for (int d = 0; d < 3; ++d) dx[d] += g0 * j[6 * h + d] + g1 * j[6 * h + 3 + d];
    }
  }
}
// Hash-table scatter (and, for tracking, the corner re-read of dL/dx) of levels [l0, l1) of one point.  The level loop is
// ROLLED: dg comes from the thread's column of a shared-memory staging area (element k at dgs[k * dgs_stride]) instead of a
// register array.  Unrolled over eight levels (round 1 / start of round 2) the scatter alone was ~60 KB of SASS per thread
// group and k_point_bwd_tc2<MAP> 200 KB -- beyond the instruction cache, with every warp of the SM in a different phase
// (ncu: no_instructions 13 % of the stall samples).
// The x and x+1 corners of a cell edge are neighbours in memory (dense levels: consecutive indices; hashed levels: the x
// prime is 1, so an even x only flips bit 0).  When they share an aligned 16-byte pair, ONE vector reduction carries both:
// a quarter fewer operations for the L2 atomic units that bound this kernel.  (Pairing the re-reads of the tracking path
// like the forward gathers was measured slower: 6.9 -> 7.7 ms.)
__device__ __forceinline__ void hashgrid_bwd_levels(const dns_grid& G, const float2* __restrict__ table, float2* d_table_all,
                                                    const float x[3], const float* dgs, int dgs_stride, int l0, int l1,
                                                    bool want_dx, float dx[3], float2* d_priv, int priv_levels) {
  dx[0] = dx[1] = dx[2] = 0.f;
#pragma unroll 1
  for (int l = l0; l < l1; ++l) {
    float2* d_table = (d_table_all && l < priv_levels) ? d_priv : d_table_all;
    const float g0 = dgs[(2 * (l - l0)) * dgs_stride], g1 = dgs[(2 * (l - l0) + 1) * dgs_stride];
    if (g0 == 0.f && g1 == 0.f) continue;
    uint32_t g[3];
    float w[3];
    const float sc = G.scale[l];
    grid_pos(x[0], sc, g[0], w[0]);
    grid_pos(x[1], sc, g[1], w[1]);
    grid_pos(x[2], sc, g[2], w[2]);
    uint32_t idx[8];
    corner_indices8(G, l, g, idx);
    if (d_table) {
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        const float wyz = ((c & 2) ? w[1] : 1.f - w[1]) * ((c & 4) ? w[2] : 1.f - w[2]);
        const float w0 = (1.f - w[0]) * wyz, w1 = w[0] * wyz;
        const uint32_t i0 = idx[c], i1 = idx[c + 1];
        if ((i0 ^ i1) == 1u) {   // one vector reduction for an aligned x / x+1 pair
          const bool odd = i0 & 1u;
          const float wa = odd ? w1 : w0, wb = odd ? w0 : w1;
          atomicAdd(reinterpret_cast<float4*>(d_table + (i0 & ~1u)), make_float4(wa * g0, wa * g1, wb * g0, wb * g1));
        } else {
          atomicAdd(d_table + i0, make_float2(w0 * g0, w0 * g1));
          atomicAdd(d_table + i1, make_float2(w1 * g0, w1 * g1));
        }
      }
    }
    if (want_dx) {
      float s[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float2 v = __ldg(table + idx[c]);
        s[c] = v.x * g0 + v.y * g1;
      }
      const float wx0 = 1.f - w[0], wy0 = 1.f - w[1], wz0 = 1.f - w[2];
      dx[0] += sc * (wy0 * wz0 * (s[1] - s[0]) + w[1] * wz0 * (s[3] - s[2]) + wy0 * w[2] * (s[5] - s[4]) + w[1] * w[2] * (s[7] - s[6]));
      dx[1] += sc * (wx0 * wz0 * (s[2] - s[0]) + w[0] * wz0 * (s[3] - s[1]) + wx0 * w[2] * (s[6] - s[4]) + w[0] * w[2] * (s[7] - s[5]));
      dx[2] += sc * (wx0 * wy0 * (s[4] - s[0]) + w[0] * wy0 * (s[5] - s[1]) + wx0 * w[1] * (s[6] - s[2]) + w[0] * w[1] * (s[7] - s[3]));
    }
  }
}
// Hash-table scatter of the TV lattice (mapping.py:129-159).  Slots run x fastest (slot_point<kTv>), so the lanes of a warp
// are consecutive lattice points of one x row: same y, z (bit-identical weights) and non-decreasing cells along x.  At the
// levels whose cells are wider than the lattice spacing (l < agg_levels) several lanes in a row fall into the SAME cell and
// the next cell's x plane is this cell's x+1 plane, so the warp first sums (1-wx).g and wx.g per cell (segmented scan over
// the runs of equal cells: five shuffle rounds on four values), hands a run's x+1 sums to the run of the neighbouring cell,
// and only the LAST lane of a run issues reductions: four per x plane instead of eight (six with pairing) per lane.  The
// reduction count per lane, not bytes, is what bounds the scatter.  Must be called by all 32 lanes (dg of an invalid lane = 0);
// `rowid` identifies the x row of the lane's lattice point.  Levels >= agg_levels take the per-lane path.
// The level loop is rolled (dg comes from the thread's column of a shared-memory staging area, element k at dgs[k * dgs_stride]):
// unrolled, the eight scan bodies per thread missed the instruction cache (ncu: no_instructions 31 % of the stalls).
__device__ __forceinline__ void hashgrid_bwd_rows(const dns_grid& G, float2* d_table_all, const float x[3], const float* dgs,
                                                  int dgs_stride, int l0, int l1, bool valid, int64_t rowid, int agg_levels,
                                                  float2* d_priv, int priv_levels) {
  constexpr unsigned kAll = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int64_t row_up = __shfl_up_sync(kAll, rowid, 1);
  const int valid_up = __shfl_up_sync(kAll, (int)valid, 1);
  const bool same_row = lane > 0 && valid && valid_up && row_up == rowid;
#pragma unroll 1
  for (int l = l0; l < l1; ++l) {
    float2* d_table = l < priv_levels ? d_priv : d_table_all;
    const float g0 = dgs[(2 * (l - l0)) * dgs_stride], g1 = dgs[(2 * (l - l0) + 1) * dgs_stride];
    uint32_t g[3];
    float w[3];
    const float sc = G.scale[l];
    grid_pos(x[0], sc, g[0], w[0]);
    grid_pos(x[1], sc, g[1], w[1]);
    grid_pos(x[2], sc, g[2], w[2]);
    if (l < agg_levels) {   // uniform over the grid
      const uint32_t gx_up = __shfl_up_sync(kAll, g[0], 1);
      const bool head = !(same_row && gx_up == g[0]);
      const bool adj = same_row && gx_up + 1u == g[0];   // the previous run is the x-neighbour cell: its x+1 plane is my x plane
      const unsigned heads = __ballot_sync(kAll, head);
      const int start = 31 - __clz(heads & (kAll >> (31 - lane)));   // first lane of my run (lane 0 is always a head)
      const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
      float a0 = (1.f - w[0]) * g0, a1 = (1.f - w[0]) * g1, b0 = w[0] * g0, b1 = w[0] * g1;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const float ta0 = __shfl_up_sync(kAll, a0, d), ta1 = __shfl_up_sync(kAll, a1, d);
        const float tb0 = __shfl_up_sync(kAll, b0, d), tb1 = __shfl_up_sync(kAll, b1, d);
        if (lane - d >= start) { a0 += ta0; a1 += ta1; b0 += tb0; b1 += tb1; }
      }
      // the last lane of a run holds the run's sums; the previous run ends at lane start - 1
      const float pb0 = __shfl_sync(kAll, b0, (start - 1) & 31), pb1 = __shfl_sync(kAll, b1, (start - 1) & 31);
      const int adj_head = __shfl_sync(kAll, (int)adj, start);
      const int next_adj = __shfl_down_sync(kAll, (int)adj, 1);    // lane + 1 heads the next run when this lane is a tail
      if (tail && valid && d_table_all) {
        if (adj_head) { a0 += pb0; a1 += pb1; }
        const bool plus = lane == 31 || !next_adj;                 // nobody takes over the x+1 plane
        uint32_t ci[8];
        corner_indices8(G, l, g, ci);
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
          const float wyz = ((c & 2) ? w[1] : 1.f - w[1]) * ((c & 4) ? w[2] : 1.f - w[2]);
          if (a0 != 0.f || a1 != 0.f) atomicAdd(d_table + ci[c], make_float2(wyz * a0, wyz * a1));
          if (plus && (b0 != 0.f || b1 != 0.f)) atomicAdd(d_table + ci[c + 1], make_float2(wyz * b0, wyz * b1));
        }
      }
    } else if (valid && d_table_all && (g0 != 0.f || g1 != 0.f)) {
      uint32_t ci[8];
      corner_indices8(G, l, g, ci);
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        const float wyz = ((c & 2) ? w[1] : 1.f - w[1]) * ((c & 4) ? w[2] : 1.f - w[2]);
        const float w0 = (1.f - w[0]) * wyz, w1 = w[0] * wyz;
        const uint32_t i0 = ci[c], i1 = ci[c + 1];
        if ((i0 ^ i1) == 1u) {
          const bool odd = i0 & 1u;
          const float wa = odd ? w1 : w0, wb = odd ? w0 : w1;
          atomicAdd(reinterpret_cast<float4*>(d_table + (i0 & ~1u)), make_float4(wa * g0, wa * g1, wb * g0, wb * g1));
        } else {
          atomicAdd(d_table + i0, make_float2(w0 * g0, w0 * g1));
          atomicAdd(d_table + i1, make_float2(w1 * g0, w1 * g1));
        }
      }
    }
  }
}
__device__ __forceinline__ void put_chunk(unsigned char* hi_tile, unsigned char* lo_tile, int chunk, int cs, int point,
                                          const float* v8) {
  uint4 h, l;
  split8(make_float4(v8[0], v8[1], v8[2], v8[3]), make_float4(v8[4], v8[5], v8[6], v8[7]), h, l);
  *reinterpret_cast<uint4*>(hi_tile + chunk * cs + point * 16) = h;
  *reinterpret_cast<uint4*>(lo_tile + chunk * cs + point * 16) = l;
}
// 8 values -> bf16 hi / lo chunks of a global tile image only
__device__ __forceinline__ void store_chunk_img(const float* v8, uint4* g_hi, uint4* g_lo) {
  uint4 h, l;
  split8(make_float4(v8[0], v8[1], v8[2], v8[3]), make_float4(v8[4], v8[5], v8[6], v8[7]), h, l);
  *g_hi = h;
  *g_lo = l;
}
// Layer-1 INPUT tiles (X of the point kernels, X of the ray kernel): the shared-memory operand is split into fp16 hi + lo
// halves (11 + 11 mantissa bits) because the GEMM it feeds decides on which side of the ReLU a hidden unit lands -- with
// bf16 halves (8 + 8 bits) about one pre-activation in 10^5 flipped against an fp32 evaluation, which is invisible in
// the forward pass but switches that unit's sub-gradient (round-1 parity outliers).  The inputs are bounded (OneBlob in
// [0,1], grid features, latents): fp16's range is enough, as it is for the reference's fp16 tinycudann networks.  The
// global tile IMAGE of the same values stays bf16 hi + lo: it meets gradients (full fp32 range) in the weight-gradient
// GEMM, and both operands of one tcgen05.mma must share the element format (a mixed descriptor traps on B200).
__device__ __forceinline__ void put_chunk_f16_img(unsigned char* hi_tile, unsigned char* lo_tile, int chunk, int cs, int point,
                                                  const float* v8, uint4* g_hi, uint4* g_lo) {
  const float4 a = make_float4(v8[0], v8[1], v8[2], v8[3]), b = make_float4(v8[4], v8[5], v8[6], v8[7]);
  uint4 h, l;
  split8_f16(a, b, h, l);
  *reinterpret_cast<uint4*>(hi_tile + chunk * cs + point * 16) = h;
  *reinterpret_cast<uint4*>(lo_tile + chunk * cs + point * 16) = l;
  if (g_hi) {
    split8(a, b, h, l);
    *g_hi = h;
    *g_lo = l;
  }
}
// one fp16 hi / lo chunk of a tile back to 8 floats (hi + lo is exact to 22 bits)
__device__ __forceinline__ void f16_chunk_to_floats(const unsigned char* hi_tile, const unsigned char* lo_tile, int chunk, int cs,
                                                    int point, float (&v)[8]) {
  const uint4 h = *reinterpret_cast<const uint4*>(hi_tile + chunk * cs + point * 16);
  const uint4 l = *reinterpret_cast<const uint4*>(lo_tile + chunk * cs + point * 16);
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hw[i]));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&lw[i]));
    v[2 * i] = a.x + b.x;
    v[2 * i + 1] = a.y + b.y;
  }
}
// same, plus a copy of both halves into a global tile image for the weight-gradient GEMMs (tc.cu: k_dw_img).
// Within one chunk neighbouring rows are 16 B apart, so a warp writes 512 contiguous bytes per store.
__device__ __forceinline__ void put_chunk_img(unsigned char* hi_tile, unsigned char* lo_tile, int chunk, int cs, int point,
                                              const float* v8, uint4* g_hi, uint4* g_lo) {
  uint4 h, l;
  split8(make_float4(v8[0], v8[1], v8[2], v8[3]), make_float4(v8[4], v8[5], v8[6], v8[7]), h, l);
  *reinterpret_cast<uint4*>(hi_tile + chunk * cs + point * 16) = h;
  *reinterpret_cast<uint4*>(lo_tile + chunk * cs + point * 16) = l;
  if (g_hi) {
    *g_hi = h;
    *g_lo = l;
  }
}

}  // namespace dns
