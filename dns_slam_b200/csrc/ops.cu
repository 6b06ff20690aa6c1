// Operator-surface kernels: the tinycudann module API the reference calls
// (models/pos_encoding.py:31-46,61-71; models/decoder.py:58-65), plus the shared weight-gradient
// GEMM and Adam (slams/mapping.py:464-466,910).
#include <math.h>
#include <stdarg.h>

#include "common.cuh"

namespace dns {

static thread_local char g_err[512] = "ok";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return DNS_ERR_CUDA;
  }
  return DNS_OK;
}

// ---------------------------------------------------------------------------------------
// phase accounting
// ---------------------------------------------------------------------------------------
static bool g_profile = false;
static long long g_launches[phCount] = {0};
struct EvPair {
  int phase;
  cudaEvent_t a, b;
};
constexpr int kMaxEvents = 4096;
static EvPair g_events[kMaxEvents];
static int g_n_events = 0;      // intervals recorded since the last reset
static int g_n_created = 0;     // events are created once and re-used: no driver calls besides the records
PhaseScope::PhaseScope(int phase_, cudaStream_t st_, int n_launches) : phase(phase_), st(st_), slot(-1) {
  g_launches[phase] += n_launches;
  if (g_profile && g_n_events < kMaxEvents) {
    slot = g_n_events++;
    if (slot >= g_n_created) {
      cudaEventCreate(&g_events[slot].a);
      cudaEventCreate(&g_events[slot].b);
      g_n_created = slot + 1;
    }
    g_events[slot].phase = phase;
    cudaEventRecord(g_events[slot].a, st);
  }
}
PhaseScope::~PhaseScope() {
  if (slot >= 0) cudaEventRecord(g_events[slot].b, st);
}

// ---------------------------------------------------------------------------------------
// OneBlob
// ---------------------------------------------------------------------------------------
__global__ void k_oneblob_fwd(const float* __restrict__ x, int64_t n /*P*D*/, int nb, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  oneblob_fwd(x[i], nb, out + i * nb, 1);
}
__global__ void k_oneblob_bwd(const float* __restrict__ x, const float* __restrict__ d_out, int64_t n, int nb,
                              float* __restrict__ d_x) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  d_x[i] = oneblob_bwd(x[i], nb, d_out + i * nb, 1);
}

// ---------------------------------------------------------------------------------------
// HashGrid
// ---------------------------------------------------------------------------------------
__global__ void k_hashgrid_fwd(dns_grid G, const float* __restrict__ x, const float2* __restrict__ table,
                               int64_t P, float* __restrict__ out) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float xv[3] = {x[3 * p], x[3 * p + 1], x[3 * p + 2]};
  float o[2 * DNS_MAX_LEVELS];
  hashgrid_fwd(G, table, xv, o, 1);
  int w = 2 * G.n_levels;
  for (int i = 0; i < w; ++i) out[p * w + i] = o[i];
}
__global__ void k_hashgrid_bwd(dns_grid G, const float* __restrict__ x, const float2* __restrict__ table,
                               const float* __restrict__ d_out, int64_t P, float2* d_table, float* d_x) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float xv[3] = {x[3 * p], x[3 * p + 1], x[3 * p + 2]};
  float dx[3];
  hashgrid_bwd(G, table, d_table, xv, d_out + p * 2 * G.n_levels, 1, d_x != nullptr, dx);
  if (d_x) {
    d_x[3 * p] = dx[0];
    d_x[3 * p + 1] = dx[1];
    d_x[3 * p + 2] = dx[2];
  }
}
__global__ void k_hashgrid_indices(dns_grid G, const float* __restrict__ x, int64_t P, uint32_t* __restrict__ idx) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  for (int l = 0; l < G.n_levels; ++l) {
    uint32_t g[3];
    float w[3];
    for (int a = 0; a < 3; ++a) grid_pos(x[3 * p + a], G.scale[l], g[a], w[a]);
    uint32_t i8[8];
    corner_indices8(G, l, g, i8);   // the per-cell form the fused kernels use; must equal the per-corner form
    for (int c = 0; c < 8; ++c) {
      const uint32_t ic = corner_index(G, l, g[0] + (c & 1), g[1] + ((c >> 1) & 1), g[2] + (c >> 2));
      idx[(p * G.n_levels + l) * 8 + c] = ic == i8[c] ? ic : 0xffffffffu;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Weight-gradient GEMM:  C[m][n] += sum_p A[p][m] * B[p][n]
// Rows p come in tiles of 128; with tile_class != NULL tile t adds into C + tile_class[t] *
// c_stride (class-wise experts, slams/mapping.py:590-601), negative classes are skipped.
// Each CTA owns a contiguous range of tiles, accumulates in registers and flushes with atomics.
// ---------------------------------------------------------------------------------------
constexpr int kGemmRows = 64;
__global__ void __launch_bounds__(256)
k_dw_gemm(const float* __restrict__ A, int lda, int M, const float* __restrict__ B, int ldb, int N,
          int64_t n_rows, const int* __restrict__ n_tiles_dev, int n_tiles_host,
          const int* __restrict__ tile_class, float* C, int ldc, int64_t c_stride) {
  extern __shared__ float sm[];
  const int Ms = (M + 3) & ~3, Ns = (N + 3) & ~3;
  float* As = sm;                      // [64][Ms]
  float* Bs = sm + kGemmRows * Ms;     // [64][Ns]
  const int n_tiles = n_tiles_dev ? *n_tiles_dev : n_tiles_host;
  const int per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int t0 = blockIdx.x * per, t1 = min(n_tiles, t0 + per);
  if (t0 >= t1) return;
  const int nt = Ns >> 2, mt = Ms >> 2;
  const int tid = threadIdx.x;
  const bool active = tid < mt * nt;
  const int tm = tid / nt, tn = tid % nt;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  int cur_class = -1;
  auto flush = [&]() {
    if (cur_class >= 0 && active) {
      float* Cc = C + (int64_t)cur_class * c_stride;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int m = 4 * tm + i, n = 4 * tn + j;
          if (m < M && n < N && acc[i][j] != 0.f) atomicAdd(Cc + (int64_t)m * ldc + n, acc[i][j]);
          acc[i][j] = 0.f;
        }
    }
  };
  for (int t = t0; t < t1; ++t) {
    int cls = tile_class ? tile_class[t] : 0;
    if (cls != cur_class) {
      flush();
      cur_class = cls;
    }
    if (cls < 0) continue;
    for (int half = 0; half < kTile / kGemmRows; ++half) {
      int64_t r0 = (int64_t)t * kTile + half * kGemmRows;
      if (r0 >= n_rows) break;
      int rows = (int)min((int64_t)kGemmRows, n_rows - r0);
      __syncthreads();
      for (int i = tid; i < kGemmRows * Ms; i += blockDim.x) {
        int r = i / Ms, m = i - r * Ms;
        As[i] = (r < rows && m < M) ? A[(r0 + r) * lda + m] : 0.f;
      }
      for (int i = tid; i < kGemmRows * Ns; i += blockDim.x) {
        int r = i / Ns, n = i - r * Ns;
        Bs[i] = (r < rows && n < N) ? B[(r0 + r) * ldb + n] : 0.f;
      }
      __syncthreads();
      if (active) {
        const float4* ap = reinterpret_cast<const float4*>(As + 4 * tm);
        const float4* bp = reinterpret_cast<const float4*>(Bs + 4 * tn);
#pragma unroll 4
        for (int r = 0; r < kGemmRows; ++r) {
          float4 a = ap[r * mt], b = bp[r * nt];
          acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
          acc[0][2] = fmaf(a.x, b.z, acc[0][2]); acc[0][3] = fmaf(a.x, b.w, acc[0][3]);
          acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
          acc[1][2] = fmaf(a.y, b.z, acc[1][2]); acc[1][3] = fmaf(a.y, b.w, acc[1][3]);
          acc[2][0] = fmaf(a.z, b.x, acc[2][0]); acc[2][1] = fmaf(a.z, b.y, acc[2][1]);
          acc[2][2] = fmaf(a.z, b.z, acc[2][2]); acc[2][3] = fmaf(a.z, b.w, acc[2][3]);
          acc[3][0] = fmaf(a.w, b.x, acc[3][0]); acc[3][1] = fmaf(a.w, b.y, acc[3][1]);
          acc[3][2] = fmaf(a.w, b.z, acc[3][2]); acc[3][3] = fmaf(a.w, b.w, acc[3][3]);
        }
      }
    }
  }
  flush();
}

int launch_dw_gemm_tc(const float* A, int lda, int M, const float* B, int ldb, int N, int64_t n_rows,
                      const int* n_tiles_dev, int n_tiles_host, const int* tile_class, float* C, int ldc,
                      int64_t c_stride, cudaStream_t st);

int launch_dw_gemm(const float* A, int lda, int M, const float* B, int ldb, int N, int64_t n_rows,
                   const int* n_tiles_dev, int n_tiles_host, const int* tile_class, float* C, int ldc,
                   int64_t c_stride, cudaStream_t st, bool tensor_cores) {
  // the tcgen05 kernel stages rows with 16-byte loads: rows must be float4-aligned and padded
  if (tensor_cores && M <= 128 && N <= 128 && (lda & 3) == 0 && (ldb & 3) == 0 && ((M + 3) & ~3) <= lda &&
      ((N + 3) & ~3) <= ldb && (((uintptr_t)A | (uintptr_t)B) & 15) == 0)
    return launch_dw_gemm_tc(A, lda, M, B, ldb, N, n_rows, n_tiles_dev, n_tiles_host, tile_class, C, ldc, c_stride, st);
  int Ms = (M + 3) & ~3, Ns = (N + 3) & ~3;
  if ((Ms >> 2) * (Ns >> 2) > 256) {
    set_error("dw_gemm: M*N too large (%d x %d)", M, N);
    return DNS_ERR_UNSUPPORTED;
  }
  if (n_tiles_host <= 0) return DNS_OK;
  size_t smem = (size_t)kGemmRows * (Ms + Ns) * sizeof(float);
  int grid = n_tiles_host < 592 ? n_tiles_host : 592;
  static unsigned long long seen = 0;
  if (first_call_on_device(seen)) cudaFuncSetAttribute(k_dw_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  k_dw_gemm<<<grid, 256, smem, st>>>(A, lda, M, B, ldb, N, n_rows, n_tiles_dev, n_tiles_host, tile_class, C, ldc,
                                     c_stride);
  return check_launch("dw_gemm");
}

// ---------------------------------------------------------------------------------------
// tcnn.Network forward / backward (generic n_in <= 128, n_out <= 128)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTile)
k_mlp_fwd(const float* __restrict__ x, const float* __restrict__ params, int64_t P, int n_in, int n_out,
          float* __restrict__ out, float* __restrict__ hidden) {
  extern __shared__ float sm[];
  const int xld = n_in + 1, n_out4 = (n_out + 3) & ~3;
  float* W1T = sm;                       // [n_in][32]
  float* W2T = W1T + n_in * 32;          // [32][n_out4]
  float* XS = W2T + 32 * n_out4;         // [128][n_in+1]
  const int tid = threadIdx.x;
  const float* W1 = params;              // [32][n_in]
  const float* W2 = params + 32 * n_in;  // [out_pad][32]
  for (int i = tid; i < 32 * n_in; i += kTile) {
    int j = i / n_in, k = i - j * n_in;
    W1T[k * 32 + j] = W1[i];
  }
  for (int i = tid; i < 32 * n_out4; i += kTile) {
    int j = i / n_out4, c = i - j * n_out4;
    W2T[i] = c < n_out ? W2[c * 32 + j] : 0.f;
  }
  const int64_t n_tiles = (P + kTile - 1) / kTile;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t p0 = t * kTile;
    const int rows = (int)min((int64_t)kTile, P - p0);
    __syncthreads();
    for (int i = tid; i < kTile * n_in; i += kTile) {
      int r = i / n_in, k = i - r * n_in;
      XS[r * xld + k] = r < rows ? x[p0 * n_in + i] : 0.f;
    }
    __syncthreads();
    float h[32];
    zero(h);
    accum_layer<32>(XS + tid * xld, 1, n_in, W1T, 32, h);
#pragma unroll
    for (int j = 0; j < 32; ++j) h[j] = fmaxf(h[j], 0.f);
    const int64_t p = p0 + tid;
    if (tid < rows) {
      if (hidden) {
        float4* hp = reinterpret_cast<float4*>(hidden + p * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) hp[q] = make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
      }
      for (int c = 0; c < n_out; c += 4) {
        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float4 w = *reinterpret_cast<const float4*>(W2T + j * n_out4 + c);
          o[0] = fmaf(h[j], w.x, o[0]);
          o[1] = fmaf(h[j], w.y, o[1]);
          o[2] = fmaf(h[j], w.z, o[2]);
          o[3] = fmaf(h[j], w.w, o[3]);
        }
        for (int i = 0; i < 4 && c + i < n_out; ++i) out[p * n_out + c + i] = o[i];
      }
    }
  }
}

// d_hidden = relu'(hidden) * (d_out @ W2);  d_x = d_hidden @ W1
__global__ void __launch_bounds__(kTile)
k_mlp_bwd(const float* __restrict__ params, const float* __restrict__ hidden, const float* __restrict__ d_out,
          int64_t P, int n_in, int n_out, float* __restrict__ d_hidden, float* __restrict__ d_x) {
  extern __shared__ float sm[];
  float* W1s = sm;                 // [32][n_in]   (tcnn layout: row j contiguous over k)
  float* W2s = W1s + 32 * n_in;    // [n_out][32]
  const int tid = threadIdx.x;
  for (int i = tid; i < 32 * n_in; i += kTile) W1s[i] = params[i];
  for (int i = tid; i < n_out * 32; i += kTile) W2s[i] = params[32 * n_in + i];
  __syncthreads();
  const int64_t n_tiles = (P + kTile - 1) / kTile;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t p = t * kTile + tid;
    if (p >= P) continue;
    float dh[32];
    zero(dh);
    for (int c = 0; c < n_out; ++c) {
      float g = d_out[p * n_out + c];
      const float4* w = reinterpret_cast<const float4*>(W2s + c * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 v = w[q];
        dh[4 * q] = fmaf(g, v.x, dh[4 * q]);
        dh[4 * q + 1] = fmaf(g, v.y, dh[4 * q + 1]);
        dh[4 * q + 2] = fmaf(g, v.z, dh[4 * q + 2]);
        dh[4 * q + 3] = fmaf(g, v.w, dh[4 * q + 3]);
      }
    }
    const float4* hp = reinterpret_cast<const float4*>(hidden + p * 32);
    float4* dp = reinterpret_cast<float4*>(d_hidden + p * 32);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4 hv = hp[q];
      dh[4 * q] = hv.x > 0.f ? dh[4 * q] : 0.f;
      dh[4 * q + 1] = hv.y > 0.f ? dh[4 * q + 1] : 0.f;
      dh[4 * q + 2] = hv.z > 0.f ? dh[4 * q + 2] : 0.f;
      dh[4 * q + 3] = hv.w > 0.f ? dh[4 * q + 3] : 0.f;
      dp[q] = make_float4(dh[4 * q], dh[4 * q + 1], dh[4 * q + 2], dh[4 * q + 3]);
    }
    if (d_x) {
      for (int k = 0; k < n_in; k += 4) {
        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float4 w = *reinterpret_cast<const float4*>(W1s + j * n_in + k);
          o[0] = fmaf(dh[j], w.x, o[0]);
          o[1] = fmaf(dh[j], w.y, o[1]);
          o[2] = fmaf(dh[j], w.z, o[2]);
          o[3] = fmaf(dh[j], w.w, o[3]);
        }
        *reinterpret_cast<float4*>(d_x + p * n_in + k) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Adam (torch.optim.Adam defaults: no weight decay, no amsgrad)
// ---------------------------------------------------------------------------------------
__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                       float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float bc1,
                       float bc2_sqrt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float gi = g[i];
    float mi = m[i] + (1.f - b1) * (gi - m[i]);        // lerp, as torch's foreach path
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - (lr / bc1) * (mi / denom);
  }
}

// Several parameter segments (own learning rate each) in ONE launch, step counter on the device so that a captured
// CUDA graph advances it on every replay (slams/tracking.py:119-124: translation / quaternion groups;
// slams/mapping.py:464-466: decoder / quaternions / translations).
__global__ void k_adam_tick(int* step) { *step += 1; }
// beta^t for the integer step count by repeated squaring in double (<= 2 log2 t multiplications, relative error ~1e-15:
// the float bias corrections are those of pow(); every thread of the launch evaluating pow() twice cost 10 us of FP64).
__device__ __forceinline__ double ipow(double b, int t) {
  double r = 1.0;
  for (; t > 0; t >>= 1, b *= b)
    if (t & 1) r *= b;
  return r;
}
// One Adam update of up to four consecutive elements held in registers.
__device__ __forceinline__ void adam4(float4& p, const float4& g, float4& m, float4& v, float b1, float b2, float eps,
                                      float step_size, float bc2_sqrt) {
  float* pp = &p.x; const float* gg = &g.x; float* mm = &m.x; float* vv = &v.x;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float mi = mm[e] + (1.f - b1) * (gg[e] - mm[e]);       // lerp, as torch's foreach path
    const float vi = b2 * vv[e] + (1.f - b2) * gg[e] * gg[e];
    mm[e] = mi;
    vv[e] = vi;
    pp[e] = pp[e] - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  }
}
// A block's pass over [base, base + len) of one segment, 16-byte accesses and kAdamBatch of them in flight per thread (the
// scalar grid-stride form waited one DRAM latency per element: 48 us for the 1.9 M parameters of a Replica decoder, most
// of it in the 40 blocks that own a class-expert row each).  The pointers of a segment are 16-byte aligned whenever its
// offset in the flat buffer is a multiple of four floats; anything else takes the scalar tail loop.
constexpr int kAdamBatch = 2;
__device__ __forceinline__ void adam_span(const dns_adam_seg& sg, int64_t base, int64_t len, int tid, int nthreads, float b1,
                                          float b2, float eps, float step_size, float bc2_sqrt) {
  const bool vec = ((((uintptr_t)(sg.p + base)) | ((uintptr_t)(sg.g + base)) | ((uintptr_t)(sg.m + base)) |
                     ((uintptr_t)(sg.v + base))) & 15) == 0;
  const int64_t n4 = vec ? len >> 2 : 0;
  float4* p4 = reinterpret_cast<float4*>(sg.p + base);
  const float4* g4 = reinterpret_cast<const float4*>(sg.g + base);
  float4* m4 = reinterpret_cast<float4*>(sg.m + base);
  float4* v4 = reinterpret_cast<float4*>(sg.v + base);
  for (int64_t i0 = tid; i0 < n4; i0 += (int64_t)kAdamBatch * nthreads) {
    float4 p[kAdamBatch], g[kAdamBatch], m[kAdamBatch], v[kAdamBatch];
#pragma unroll
    for (int k = 0; k < kAdamBatch; ++k) {
      const int64_t i = i0 + (int64_t)k * nthreads;
      if (i < n4) { g[k] = g4[i]; m[k] = m4[i]; v[k] = v4[i]; p[k] = p4[i]; }
    }
#pragma unroll
    for (int k = 0; k < kAdamBatch; ++k) {
      const int64_t i = i0 + (int64_t)k * nthreads;
      if (i < n4) {
        adam4(p[k], g[k], m[k], v[k], b1, b2, eps, step_size, bc2_sqrt);
        m4[i] = m[k]; v4[i] = v[k]; p4[i] = p[k];
      }
    }
  }
  for (int64_t i = base + 4 * n4 + tid; i < base + len; i += nthreads) {
    const float gi = sg.g[i];
    const float mi = sg.m[i] + (1.f - b1) * (gi - sg.m[i]);
    const float vi = b2 * sg.v[i] + (1.f - b2) * gi * gi;
    sg.m[i] = mi;
    sg.v[i] = vi;
    sg.p[i] = sg.p[i] - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  }
}
// 148 x 4 blocks of 256 threads at <= 64 registers: the whole grid is resident at once (at 127 registers and 1184 blocks it
// ran four waves of latency-bound blocks: 28 us for 30 MB)
__global__ void __launch_bounds__(256, 4) k_adam_multi(const dns_adam_seg* __restrict__ segs, const int* __restrict__ step, float b1, float b2,
                             float eps) {
  const dns_adam_seg sg = segs[blockIdx.y];
  if (sg.row_len > 0) {
    // independent tensors per row: a row without gradient is skipped like a parameter whose .grad is None, the others
    // advance their OWN step count (torch.optim.Adam keeps `step` per parameter tensor)
    __shared__ int s_step;
    const int64_t rows = sg.n / sg.row_len;
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
      const int64_t base = row * sg.row_len;
      int any = 0;
      if ((((uintptr_t)(sg.g + base)) & 15) == 0 && (sg.row_len & 3) == 0) {
        const float4* g4 = reinterpret_cast<const float4*>(sg.g + base);
        for (int i = threadIdx.x; i < (sg.row_len >> 2); i += blockDim.x) {
          const float4 g = g4[i];
          any |= (g.x != 0.f) | (g.y != 0.f) | (g.z != 0.f) | (g.w != 0.f);
        }
      } else {
        for (int i = threadIdx.x; i < sg.row_len; i += blockDim.x) any |= sg.g[base + i] != 0.f;
      }
      any = __syncthreads_or(any);
      if (!any) continue;                    // uniform over the block
      if (threadIdx.x == 0) s_step = ++sg.row_steps[row];
      __syncthreads();
      const float bc1 = (float)(1.0 - ipow((double)b1, s_step)), bc2_sqrt = (float)sqrt(1.0 - ipow((double)b2, s_step));
      adam_span(sg, base, sg.row_len, threadIdx.x, blockDim.x, b1, b2, eps, sg.lr / bc1, bc2_sqrt);
      __syncthreads();                       // s_step is rewritten for the next row
    }
    return;
  }
  const int t = *step;
  const float bc1 = (float)(1.0 - ipow((double)b1, t)), bc2_sqrt = (float)sqrt(1.0 - ipow((double)b2, t));
  // contiguous spans per block (multiples of four floats, so an aligned segment stays aligned in every block)
  const int64_t per = ((sg.n + gridDim.x - 1) / gridDim.x + 3) & ~(int64_t)3;
  const int64_t b0 = (int64_t)blockIdx.x * per;
  if (b0 >= sg.n) return;
  adam_span(sg, b0, (sg.n - b0 < per ? sg.n - b0 : per), threadIdx.x, blockDim.x, b1, b2, eps, sg.lr / bc1, bc2_sqrt);
}

}  // namespace dns

using namespace dns;

extern "C" {

void dns_profile_enable(int on) {
  g_profile = on != 0;
  if (g_profile) {
    for (; g_n_created < 512; ++g_n_created) {
      cudaEventCreate(&g_events[g_n_created].a);
      cudaEventCreate(&g_events[g_n_created].b);
    }
  }
}
// Adds the elapsed milliseconds of every recorded phase interval to ms[phase] and copies the launch
// counters; waits for the recorded events.  reset != 0 clears counters and intervals.
int dns_profile_read(double* ms, long long* launches, int reset) {
  for (int i = 0; i < g_n_events; ++i) {
    float t = 0.f;
    cudaEventSynchronize(g_events[i].b);
    if (cudaEventElapsedTime(&t, g_events[i].a, g_events[i].b) == cudaSuccess && ms) ms[g_events[i].phase] += t;
  }
  if (launches)
    for (int i = 0; i < phCount; ++i) launches[i] = g_launches[i];
  if (reset) {
    g_n_events = 0;
    for (int i = 0; i < phCount; ++i) g_launches[i] = 0;
  }
  return phCount;
}

const char* dns_last_error(void) { return g_err; }
int dns_version(void) { return 100; }
void dns_struct_sizes(int64_t out[5]) {
  out[0] = sizeof(dns_grid);
  out[1] = sizeof(dns_render_args);
  out[2] = sizeof(dns_tv_args);
  out[3] = sizeof(dns_sample_args);
  out[4] = sizeof(dns_featmerge_args);
}

int dns_oneblob_fwd(const float* x, int64_t P, int D, int n_bins, float* out, void* stream) {
  int64_t n = P * D;
  if (n <= 0) return DNS_OK;
  PhaseScope ph(phOps, (cudaStream_t)stream, 1);
  k_oneblob_fwd<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, n, n_bins, out);
  return check_launch("oneblob_fwd");
}
int dns_oneblob_bwd(const float* x, const float* d_out, int64_t P, int D, int n_bins, float* d_x, void* stream) {
  int64_t n = P * D;
  if (n <= 0) return DNS_OK;
  PhaseScope ph(phOps, (cudaStream_t)stream, 1);
  k_oneblob_bwd<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, d_out, n, n_bins, d_x);
  return check_launch("oneblob_bwd");
}
int dns_hashgrid_fwd(const dns_grid* g, const float* x, const float* table, int64_t P, float* out, void* stream) {
  if (!g || g->n_levels > DNS_MAX_LEVELS || g->n_features != 2) {
    set_error("hashgrid: unsupported grid (levels<=16, features==2)");
    return DNS_ERR_UNSUPPORTED;
  }
  if (P <= 0) return DNS_OK;
  PhaseScope ph(phOps, (cudaStream_t)stream, 1);
  k_hashgrid_fwd<<<(unsigned)((P + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*g, x, (const float2*)table, P, out);
  return check_launch("hashgrid_fwd");
}
int dns_hashgrid_bwd(const dns_grid* g, const float* x, const float* table, const float* d_out, int64_t P,
                     float* d_table, float* d_x, void* stream) {
  if (!g || g->n_levels > DNS_MAX_LEVELS || g->n_features != 2) {
    set_error("hashgrid: unsupported grid (levels<=16, features==2)");
    return DNS_ERR_UNSUPPORTED;
  }
  if (P <= 0) return DNS_OK;
  PhaseScope ph(phOps, (cudaStream_t)stream, 1);
  k_hashgrid_bwd<<<(unsigned)((P + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*g, x, (const float2*)table, d_out, P,
                                                                             (float2*)d_table, d_x);
  return check_launch("hashgrid_bwd");
}
int dns_hashgrid_indices(const dns_grid* g, const float* x, int64_t P, uint32_t* idx, void* stream) {
  if (P <= 0) return DNS_OK;
  k_hashgrid_indices<<<(unsigned)((P + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*g, x, P, idx);
  return check_launch("hashgrid_indices");
}

static int mlp_check(int n_in, int n_out) {
  if (n_in % 16 != 0 || n_in > 128 || n_out < 1 || n_out > 128) {
    set_error("mlp: need n_in %% 16 == 0, n_in <= 128, 1 <= n_out <= 128 (got %d, %d)", n_in, n_out);
    return DNS_ERR_UNSUPPORTED;
  }
  return DNS_OK;
}
int dns_mlp_fwd(const float* x, const float* params, int64_t P, int n_in, int n_out, float* out, float* hidden,
                void* stream) {
  if (int e = mlp_check(n_in, n_out)) return e;
  if (P <= 0) return DNS_OK;
  int n_out4 = (n_out + 3) & ~3;
  size_t smem = sizeof(float) * ((size_t)n_in * 32 + 32 * n_out4 + (size_t)kTile * (n_in + 1));
  cudaFuncSetAttribute(k_mlp_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  int64_t tiles = (P + kTile - 1) / kTile;
  int grid = (int)(tiles < 444 ? tiles : 444);
  PhaseScope ph(phOps, (cudaStream_t)stream, 1);
  k_mlp_fwd<<<grid, kTile, smem, (cudaStream_t)stream>>>(x, params, P, n_in, n_out, out, hidden);
  return check_launch("mlp_fwd");
}
int dns_mlp_bwd(const float* x, const float* params, const float* hidden, const float* d_out, int64_t P, int n_in,
                int n_out, float* d_hidden, float* d_x, float* d_params, void* stream) {
  if (int e = mlp_check(n_in, n_out)) return e;
  if (P <= 0) return DNS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  size_t smem = sizeof(float) * ((size_t)n_in * 32 + (size_t)n_out * 32);
  cudaFuncSetAttribute(k_mlp_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  int64_t tiles = (P + kTile - 1) / kTile;
  int grid = (int)(tiles < 592 ? tiles : 592);
  PhaseScope ph(phOps, st, d_params ? 3 : 1);
  k_mlp_bwd<<<grid, kTile, smem, st>>>(params, hidden, d_out, P, n_in, n_out, d_hidden, d_x);
  if (int e = check_launch("mlp_bwd")) return e;
  if (d_params) {
    // dW1[j][k] = sum_p dH[p][j] X[p][k];  dW2[c][j] = sum_p dOut[p][c] H[p][j]
    if (int e = launch_dw_gemm(d_hidden, 32, 32, x, n_in, n_in, P, nullptr, (int)tiles, nullptr, d_params, n_in, 0, st, true))
      return e;
    if (int e = launch_dw_gemm(d_out, n_out, n_out, hidden, 32, 32, P, nullptr, (int)tiles, nullptr,
                               d_params + 32 * n_in, 32, 0, st, true))
      return e;
  }
  return DNS_OK;
}

int dns_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                  float beta1, float beta2, float eps, int step, void* stream) {
  if (n <= 0) return DNS_OK;
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  int64_t blocks = (n + 255) / 256;
  int grid = (int)(blocks < 148 * 16 ? blocks : 148 * 16);
  PhaseScope ph(phAdam, (cudaStream_t)stream, 1);
  k_adam<<<grid, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                 (float)bc1, (float)sqrt(bc2));
  return check_launch("adam");
}

int dns_adam_multi(const dns_adam_seg* segs_dev, int n_segs, int64_t max_n, int* step_dev, float beta1, float beta2,
                   float eps, void* stream) {
  if (n_segs <= 0 || max_n <= 0) return DNS_OK;
  if (!segs_dev || !step_dev || n_segs > 65535) {
    set_error("dns_adam_multi: bad arguments (%d segments)", n_segs);
    return DNS_ERR_ARG;
  }
  int64_t blocks = (max_n + 255) / 256;
  const int bx = (int)(blocks < 148 * 4 ? blocks : 148 * 4);
  PhaseScope ph(phAdam, (cudaStream_t)stream, 2);
  k_adam_tick<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
  k_adam_multi<<<dim3(bx, n_segs), 256, 0, (cudaStream_t)stream>>>(segs_dev, step_dev, beta1, beta2, eps);
  return check_launch("adam_multi");
}

}  // extern "C"
