// tcgen05 (5th-gen tensor core) weight-gradient GEMM for sm_100a.
//
//   C[m][n] += sum_p A[p][m] * B[p][n]        (dW = X^T dH of every MLP; models/decoder.py:58-117)
//
// The reduction runs over sample points (rows p), 128 per tile.  Both operands are staged in shared
// memory as bf16 hi + lo halves (x = hi + lo, |lo| <= 2^-9 |x|), three products hi*hi + lo*hi + hi*lo
// are accumulated in fp32 in TMEM, so the result carries ~2^-17 relative error per term -- inside the
// 1e-3 parity bar where plain bf16 / tf32 would not be.  One operand ("lane side", <= 128 wide) maps
// to TMEM lanes, the other ("column side", <= 128 wide, padded to 16) to TMEM columns; the wider one
// goes to the lanes.
//
// Shared-memory operand layout (no swizzle, MN-major canonical UMMA layout): a tile is a stack of
// 16-byte "feature chunks" (8 bf16 features), chunk c holding [128 points][8 features]:
//     byte address = c * 2048 + point * 16 + (feature % 8) * 2
// i.e. core matrix = 8 points x 16 B contiguous (128 B), LBO (next 8 points, K direction) = 128 B,
// SBO (next 8 features, MN direction) = 2048 B.  Each thread owns one point row and writes its
// 16-byte chunks with conflict-free st.shared.v4.
//
// One CTA owns a contiguous range of tiles, keeps the accumulator in TMEM across them and flushes it
// with atomics once per range (or when the class of a class-grouped tile stream changes).
#include <cuda_bf16.h>

#include "common.cuh"

namespace dns {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
  return d;                // base_offset 0, layout_type 0 = no swizzle
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both MN-major
__device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// fp32 -> bf16 hi / lo halves, 8 values -> two 16-byte chunks
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * i]), h1 = __float2bfloat16_rn(x[2 * i + 1]);
    __nv_bfloat16 l0 = __float2bfloat16_rn(x[2 * i] - __bfloat162float(h0));
    __nv_bfloat16 l1 = __float2bfloat16_rn(x[2 * i + 1] - __bfloat162float(h1));
    h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    l[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// stage one fp32 row (n values, row stride known to the caller) of this thread's point into a chunk tile
__device__ __forceinline__ void stage_row(const float* __restrict__ row, int n, bool valid, unsigned char* hi_tile,
                                          unsigned char* lo_tile, int point) {
  const int chunks = (n + 7) >> 3;
  for (int c = 0; c < chunks; ++c) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int k = 8 * c + i;
      x[i] = (valid && k < n) ? row[k] : 0.f;
    }
    uint4 hi, lo;
    split8(x, hi, lo);
    *reinterpret_cast<uint4*>(hi_tile + c * 2048 + point * 16) = hi;
    *reinterpret_cast<uint4*>(lo_tile + c * 2048 + point * 16) = lo;
  }
}

constexpr int kLaneTile = 16 * 2048;  // lane-side tile always spans 16 chunks (128 features); unused ones stay zero

// L: lane-side operand [rows][ldl] (nL <= 128 columns used), Cc: column-side operand [rows][ldcc] (nC <= 128).
// out[l * stride_l + c * stride_c] += sum_p L[p][l] * Cc[p][c]
__global__ void __launch_bounds__(kTile)
k_dw_gemm_tc(const float* __restrict__ L, int ldl, int nL, const float* __restrict__ Cc, int ldcc, int nC,
             int64_t n_rows, const int* __restrict__ n_tiles_dev, int n_tiles_host,
             const int* __restrict__ tile_class, float* out, int64_t stride_l, int64_t stride_c, int64_t class_stride) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int nC16 = (nC + 15) & ~15;
  const int c_chunks = nC16 >> 3;
  unsigned char* L_hi = smem;
  unsigned char* L_lo = L_hi + kLaneTile;
  unsigned char* C_hi = L_lo + kLaneTile;
  unsigned char* C_lo = C_hi + c_chunks * 2048;

  const int n_tiles = n_tiles_dev ? *n_tiles_dev : n_tiles_host;
  const int per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int t0 = blockIdx.x * per, t1 = min(n_tiles, t0 + per);
  if (t0 >= t1) return;

  // zero the operand tiles once: chunks beyond the staged width must read as zero
  for (int i = tid; i < (2 * kLaneTile + 2 * c_chunks * 2048) / 16; i += kTile)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  const uint32_t tmem_cols = nC16 <= 32 ? 32 : (nC16 <= 64 ? 64 : 128);
  if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
  if (tid == 0) mbar_init(&bar, 1);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  const uint32_t idesc = umma_idesc_bf16(128, nC16, 1, 1);
  uint32_t phase = 0;
  int cur_class = -1;
  bool have_acc = false;

  auto flush = [&]() {  // all threads: wait for the MMAs, read the accumulator, atomically add
    if (!have_acc) return;
    tc_fence_after();
    if (cur_class >= 0) {
      float* o = out + (int64_t)cur_class * class_stride;
      const int l = tid;  // TMEM lane == lane-side index (warp w owns lanes 32w .. 32w+31)
      for (int c0 = 0; c0 < nC16; c0 += 16) {
        float v[16];
        tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        if (l < nL) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c0 + i < nC && v[i] != 0.f) atomicAdd(o + (int64_t)l * stride_l + (int64_t)(c0 + i) * stride_c, v[i]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    have_acc = false;
  };

  bool pending = false;  // MMAs issued and not yet waited for
  for (int t = t0; t < t1; ++t) {
    const int cls = tile_class ? tile_class[t] : 0;
    if (pending) {  // the previous tile's MMAs still read the operand tiles
      mbar_wait(&bar, phase);
      phase ^= 1;
      pending = false;
    }
    if (cls != cur_class) {
      flush();
      cur_class = cls;
    }
    if (cls < 0) continue;
    const int64_t p = (int64_t)t * kTile + tid;
    const bool valid = p < n_rows;
    stage_row(L + p * ldl, nL, valid, L_hi, L_lo, tid);
    stage_row(Cc + p * ldcc, nC, valid, C_hi, C_lo, tid);
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll 1
      for (int k = 0; k < kTile / 16; ++k) {  // 16 points per MMA: two K core matrices of 128 B
        const uint32_t koff = k * 256;
        const uint64_t a_hi = umma_desc(smem_u32(L_hi) + koff, 128, 2048), a_lo = umma_desc(smem_u32(L_lo) + koff, 128, 2048);
        const uint64_t b_hi = umma_desc(smem_u32(C_hi) + koff, 128, 2048), b_lo = umma_desc(smem_u32(C_lo) + koff, 128, 2048);
        umma_bf16(tmem_d, a_hi, b_hi, idesc, (have_acc || k > 0) ? 1u : 0u);
        umma_bf16(tmem_d, a_lo, b_hi, idesc, 1u);
        umma_bf16(tmem_d, a_hi, b_lo, idesc, 1u);
      }
      umma_commit(&bar);
    }
    have_acc = true;
    pending = true;
  }
  if (pending) {
    mbar_wait(&bar, phase);
    phase ^= 1;
  }
  flush();
  if (warp == 0) tmem_dealloc(tmem_d, tmem_cols);
}

static bool g_use_tc = true;

int launch_dw_gemm_tc(const float* A, int lda, int M, const float* B, int ldb, int N, int64_t n_rows,
                      const int* n_tiles_dev, int n_tiles_host, const int* tile_class, float* C, int ldc,
                      int64_t c_stride, cudaStream_t st) {
  // C[m][n] (ldc) += sum_p A[p][m] B[p][n]; the wider operand goes to the TMEM lanes
  if (M > 128 || N > 128) {
    set_error("dw_gemm_tc: operand wider than 128 (%d, %d)", M, N);
    return DNS_ERR_UNSUPPORTED;
  }
  if (n_tiles_host <= 0) return DNS_OK;
  const bool a_on_lanes = M > N;
  const float* Lp = a_on_lanes ? A : B;
  const float* Cp = a_on_lanes ? B : A;
  const int ldl = a_on_lanes ? lda : ldb, ldcc = a_on_lanes ? ldb : lda;
  const int nL = a_on_lanes ? M : N, nC = a_on_lanes ? N : M;
  const int64_t stride_l = a_on_lanes ? ldc : 1, stride_c = a_on_lanes ? 1 : ldc;
  const int nC16 = (nC + 15) & ~15;
  size_t smem = 2 * (size_t)kLaneTile + 2 * (size_t)(nC16 >> 3) * 2048;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_dw_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    attr = true;
  }
  int grid = n_tiles_host < 296 ? n_tiles_host : 296;
  k_dw_gemm_tc<<<grid, kTile, smem, st>>>(Lp, ldl, nL, Cp, ldcc, nC, n_rows, n_tiles_dev, n_tiles_host, tile_class, C,
                                         stride_l, stride_c, c_stride);
  return check_launch("dw_gemm_tc");
}

bool use_tensor_cores() { return g_use_tc; }

}  // namespace dns

extern "C" {
// 1: weight-gradient GEMMs on tcgen05 (default); 0: fp32 SIMT path (A/B comparisons)
void dns_set_tensor_cores(int on) { dns::g_use_tc = on != 0; }

// debug / test entry: C[m][n] (ldc = N) += sum_p A[p][m] B[p][n] through the tcgen05 kernel
int dns_debug_gemm_tc(const float* A, int lda, int M, const float* B, int ldb, int N, int64_t rows, float* C, void* stream) {
  int tiles = (int)((rows + dns::kTile - 1) / dns::kTile);
  return dns::launch_dw_gemm_tc(A, lda, M, B, ldb, N, rows, nullptr, tiles, nullptr, C, N, 0, (cudaStream_t)stream);
}
}
