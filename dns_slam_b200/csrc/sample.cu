// Ray / point sampling and the pixel-feature gather (sm_100a).
//
//   dns_sample_rays     utils/common.py:248-304 (get_samples, get_rays_from_uv),
//                       slams/tracking.py:148-160 / slams/mapping.py:519-531 (far plane, points),
//                       utils/common.py:561-599 (sample_along_rays)
//   dns_feature_gather  utils/common.py:632-670 (projection, rounding, masks, gather) with the
//                       bilinear up-sampling of :646 evaluated on the fly at the rounded pixel
//
// Integer / index work is bit exact against the oracle: every fp32 / fp64 operation is issued
// in the reference's order with explicit round-to-nearest intrinsics (no FMA contraction).
#include "common.cuh"

namespace dns {

__device__ __forceinline__ void sample_gather_body(const dns_sample_args& a, int block) {
  int r = block * blockDim.x + threadIdx.x;
  if (r >= a.n) return;
  int64_t idx = a.index[r];
  // class-balanced draws (utils/common.py:315-330): the draw is an offset into the pixels of the slot's class
  if (a.order && r >= a.n_direct) idx = a.order[(int64_t)a.slot_base[r - a.n_direct] + idx];
  if (a.pixel) a.pixel[r] = idx;
  int hh = a.H0 + (int)(idx / a.Ww), ww = a.W0 + (int)(idx % a.Ww);
  int64_t pix = (int64_t)hh * a.W + ww;
  float d = a.depth[pix];
  a.gt_color[3 * r + 0] = a.color[3 * pix + 0];
  a.gt_color[3 * r + 1] = a.color[3 * pix + 1];
  a.gt_color[3 * r + 2] = a.color[3 * pix + 2];
  a.gt_depth[r] = d;
  a.gt_label[r] = a.label[pix];
  float i = (float)ww, j = (float)hh;
  float dir[3] = {__fdiv_rn(__fsub_rn(i, a.cx), a.fx), __fdiv_rn(-__fsub_rn(j, a.cy), a.fy), -1.0f};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v = __fadd_rn(__fadd_rn(__fmul_rn(dir[0], a.R[3 * c]), __fmul_rn(dir[1], a.R[3 * c + 1])),
                        __fmul_rn(dir[2], a.R[3 * c + 2]));
    a.rays_d[3 * r + c] = v;
    a.rays_o[3 * r + c] = a.T[c];
  }
  // batch-global max depth (common.py:581,591); depths are >= 0 so the int ordering is the float ordering
  atomicMax(reinterpret_cast<int*>(a.scratch), __float_as_int(fmaxf(d, 0.f)));
}

__global__ void k_sample_gather(dns_sample_args a) { sample_gather_body(a, blockIdx.x); }

// Several frames of one mapping iteration in ONE launch (blockIdx.y = frame): at SLAM batch sizes the per-frame launches
// of the sampler were a chain of ~7 us kernels, 60 us of a 0.8 ms iteration.
constexpr int kMaxSampleFrames = 8;
struct SampleBatch {
  dns_sample_args f[kMaxSampleFrames];
};
__global__ void k_sample_gather_batch(const __grid_constant__ SampleBatch b) {
  const dns_sample_args& a = b.f[blockIdx.y];
  if ((int)(blockIdx.x * blockDim.x) >= a.n) return;
  sample_gather_body(a, blockIdx.x);
}

// kRaysZ rays per block, 128 threads: per-ray scalars by one thread each, then the S values, their ranks (S
// comparisons per value) and the optional points spread over the whole block -- a single thread per ray made the
// O(S^2) rank sort a 30 us latency chain at tracking-sized batches.
constexpr int kRaysZ = 16;
__device__ __forceinline__ void sample_z_body(const dns_sample_args& a, int block) {
  extern __shared__ float sm[];  // [S][kRaysZ] values
  __shared__ double s_far[kRaysZ];
  __shared__ float s_d[kRaysZ];
  const int tid = threadIdx.x, r0 = block * kRaysZ;
  const int S = a.n_uniform + a.n_surface;
  const int nr = min(kRaysZ, a.n - r0);
  const float maxd = a.scratch[0];
  if (tid < nr) {
    const int r = r0 + tid;
    const float d = a.gt_depth[r];
    // far plane in float64 (tracking.py:151-156)
    double far_bb = INFINITY;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double o = (double)a.rays_o[3 * r + c], dd = (double)a.rays_d[3 * r + c];
      double t0 = (a.bound[c][0] - o) / dd, t1 = (a.bound[c][1] - o) / dd;
      double m = fmax(t0, t1);
      far_bb = fmin(far_bb, m);
    }
    const bool ins = far_bb >= (double)d;
    a.inside[r] = ins ? 1 : 0;
    if (!ins) atomicAdd(a.scratch + 1, 1.0f);   // rays that leave the bound before their depth (mapping.py:525 drops them)
    far_bb += 0.01;
    double hi = (double)__fmul_rn(maxd, 1.2f);
    s_far[tid] = fmin(fmax(far_bb, 0.0), hi);
    s_d[tid] = d;
  }
  __syncthreads();
  for (int e = tid; e < nr * S; e += blockDim.x) {
    const int lr = e / S, k = e - lr * S;
    const float d = s_d[lr];
    float v;
    if (k < a.n_uniform) {
      // uniform samples: near fp32, far fp64 (clamped), blended in fp64, rounded to fp32 at the end
      const float near = __fmul_rn(d, 0.001f), t = a.t_lin[k];
      v = (float)((double)__fmul_rn(near, __fsub_rn(1.0f, t)) + s_far[lr] * (double)t);
    } else if (d > 0.f) {   // depth-guided samples
      const float lo = __fmul_rn(0.95f, d), hi = __fmul_rn(1.05f, d), t = a.t_surface[k - a.n_uniform];
      v = __fadd_rn(__fmul_rn(lo, __fsub_rn(1.0f, t)), __fmul_rn(hi, t));
    } else {
      const float t = a.t_zero[k - a.n_uniform];
      v = __fadd_rn(__fmul_rn(0.001f, __fsub_rn(1.0f, t)), __fmul_rn(maxd, t));
    }
    sm[k * kRaysZ + lr] = v;
  }
  __syncthreads();
  // ascending sort by rank (ties broken by position; equal values are interchangeable)
  for (int e = tid; e < nr * S; e += blockDim.x) {
    const int lr = e / S, i = e - lr * S;
    const float* col = sm + lr;
    const float v = col[i * kRaysZ];
    int rank = 0;
    for (int j = 0; j < S; ++j) {
      float u = col[j * kRaysZ];
      rank += (u < v) || (u == v && j < i);
    }
    a.z_vals[(int64_t)(r0 + lr) * S + rank] = v;
  }
  if (a.pts) {
    __syncthreads();   // the block's z values are complete (written by this block only)
    for (int e = tid; e < nr * S; e += blockDim.x) {
      const int lr = e / S;
      const int64_t r = r0 + lr, q = (int64_t)r0 * S + e;
      const float z = a.z_vals[q];
#pragma unroll
      for (int c = 0; c < 3; ++c) a.pts[q * 3 + c] = __fadd_rn(a.rays_o[3 * r + c], __fmul_rn(a.rays_d[3 * r + c], z));
    }
  }
}

__global__ void __launch_bounds__(128) k_sample_z(dns_sample_args a) { sample_z_body(a, blockIdx.x); }
__global__ void __launch_bounds__(128) k_sample_z_batch(const __grid_constant__ SampleBatch b) {
  const dns_sample_args& a = b.f[blockIdx.y];
  if ((int)(blockIdx.x * kRaysZ) >= a.n) return;   // uniform over the block
  sample_z_body(a, blockIdx.x);
}

// one warp per (view, point); feats are channels-last [R][h][w][C], C == 64
__global__ void k_feature_gather(const float* __restrict__ pts, int64_t P, const float* __restrict__ w2c, int R,
                                 const float* __restrict__ K, int H, int W, const float* __restrict__ feats, int C,
                                 int h, int w, float* __restrict__ code, int64_t* __restrict__ uv_out,
                                 uint8_t* __restrict__ mask_out) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= (int64_t)R * P) return;
  const int v = (int)(wid / P);
  const int64_t p = wid - (int64_t)v * P;
  const float* M = w2c + 16 * v;
  const float px = pts[3 * p], py = pts[3 * p + 1], pz = pts[3 * p + 2];
  float cam[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) cam[c] = fmaf(M[4 * c + 2], pz, fmaf(M[4 * c + 1], py, fmaf(M[4 * c], px, M[4 * c + 3])));
  cam[1] = -cam[1];
  cam[2] = -cam[2];
  float img[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) img[c] = fmaf(K[3 * c + 2], cam[2], fmaf(K[3 * c + 1], cam[1], K[3 * c] * cam[0]));
  float den = img[2] + 1e-5f;
  float u = rintf(img[0] / den), vv = rintf(img[1] / den);
  bool m = (u > 0.f) && (u < (float)(W - 1)) && (vv > 0.f) && (vv < (float)(H - 1)) && (cam[2] > 0.f);
  int64_t ui = m ? (int64_t)u : 0, vi = m ? (int64_t)vv : 0;
  if (lane == 0) {
    if (uv_out) {
      uv_out[2 * wid] = ui;
      uv_out[2 * wid + 1] = vi;
    }
    if (mask_out) mask_out[wid] = m ? 1 : 0;
  }
  float* dst = code + wid * C;
  if (!m) {
    for (int c = lane; c < C; c += 32) dst[c] = 0.f;
    return;
  }
  // F.interpolate(..., mode='bilinear', align_corners=True) sampled at integer pixel (vi, ui)
  float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  float fy = sy * (float)vi, fx = sx * (float)ui;
  int y0 = (int)fy, x0 = (int)fx;
  int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  float ly1 = fy - (float)y0, ly0 = 1.f - ly1, lx1 = fx - (float)x0, lx0 = 1.f - lx1;
  const float* f = feats + (int64_t)v * h * w * C;
  const float* f00 = f + ((int64_t)y0 * w + x0) * C;
  const float* f01 = f + ((int64_t)y0 * w + x1) * C;
  const float* f10 = f + ((int64_t)y1 * w + x0) * C;
  const float* f11 = f + ((int64_t)y1 * w + x1) * C;
  for (int c = lane; c < C; c += 32)
    dst[c] = ly0 * (lx0 * __ldg(f00 + c) + lx1 * __ldg(f01 + c)) + ly1 * (lx0 * __ldg(f10 + c) + lx1 * __ldg(f11 + c));
}

}  // namespace dns

using namespace dns;

extern "C" {

int dns_sample_rays(const dns_sample_args* a, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int S = a->n_uniform + a->n_surface;
  if (a->n <= 0) return DNS_OK;
  if (S < 1 || S > 256 || a->n_uniform < 0 || a->n_surface < 0) {
    set_error("sample: need 1 <= n_uniform + n_surface <= 256");
    return DNS_ERR_UNSUPPORTED;
  }
  if (a->order && (!a->slot_base || a->n_direct < 0 || a->n_direct > a->n)) {
    set_error("sample: class-balanced draws need slot_base and 0 <= n_direct <= n");
    return DNS_ERR_ARG;
  }
  PhaseScope ph(phSample, st, 3);
  if (a->phase != 2) {   // pixel gather, rays, batch max depth -> scratch[0]
    cudaMemsetAsync(a->scratch, 0, 2 * sizeof(float), st);
    k_sample_gather<<<(a->n + 127) / 128, 128, 0, st>>>(*a);
  }
  if (a->phase != 1)     // far plane, inside mask (outside count -> scratch[1]), z values
    k_sample_z<<<(a->n + kRaysZ - 1) / kRaysZ, 128, sizeof(float) * S * kRaysZ, st>>>(*a);
  return check_launch("sample_rays");
}

int dns_sample_rays_batch(const dns_sample_args* frames, int n_frames, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n_frames <= 0) return DNS_OK;
  if (!frames) {
    set_error("sample batch: frames == NULL");
    return DNS_ERR_ARG;
  }
  for (int f0 = 0; f0 < n_frames; f0 += kMaxSampleFrames) {
    const int nf = n_frames - f0 < kMaxSampleFrames ? n_frames - f0 : kMaxSampleFrames;
    SampleBatch b;
    memset(&b, 0, sizeof(b));
    int n_max = 0;
    const int S = frames[f0].n_uniform + frames[f0].n_surface, phase = frames[f0].phase;
    bool contiguous = true;
    for (int f = 0; f < nf; ++f) {
      const dns_sample_args& a = frames[f0 + f];
      if (a.n_uniform + a.n_surface != S || a.phase != phase || a.n_uniform != frames[f0].n_uniform) {
        set_error("sample batch: the frames of one call share n_uniform, n_surface and phase");
        return DNS_ERR_ARG;
      }
      if (S < 1 || S > 256 || a.n_uniform < 0 || a.n_surface < 0) {
        set_error("sample: need 1 <= n_uniform + n_surface <= 256");
        return DNS_ERR_UNSUPPORTED;
      }
      if (a.n > 0 && a.order && (!a.slot_base || a.n_direct < 0 || a.n_direct > a.n)) {
        set_error("sample: class-balanced draws need slot_base and 0 <= n_direct <= n");
        return DNS_ERR_ARG;
      }
      b.f[f] = a;
      if (a.n > n_max) n_max = a.n;
      contiguous = contiguous && a.scratch == frames[f0].scratch + 2 * f;
    }
    if (n_max <= 0) continue;
    PhaseScope ph(phSample, st, 3);
    if (phase != 2) {
      if (contiguous) cudaMemsetAsync(frames[f0].scratch, 0, 2 * sizeof(float) * nf, st);
      else
        for (int f = 0; f < nf; ++f)
          if (b.f[f].n > 0) cudaMemsetAsync(b.f[f].scratch, 0, 2 * sizeof(float), st);
      k_sample_gather_batch<<<dim3((n_max + 127) / 128, nf), 128, 0, st>>>(b);
    }
    if (phase != 1)
      k_sample_z_batch<<<dim3((n_max + kRaysZ - 1) / kRaysZ, nf), 128, sizeof(float) * S * kRaysZ, st>>>(b);
  }
  return check_launch("sample_rays_batch");
}

int dns_feature_gather(const float* pts, int64_t P, const float* w2c, int R, const float* K, int H, int W,
                       const float* feats, int C, int h, int w, float* code, int64_t* uv, uint8_t* mask, void* stream) {
  if (P <= 0 || R <= 0) return DNS_OK;
  int64_t warps = (int64_t)R * P;
  int64_t blocks = (warps * 32 + 255) / 256;
  PhaseScope ph(phFeature, (cudaStream_t)stream, 1);
  k_feature_gather<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pts, P, w2c, R, K, H, W, feats, C, h, w, code, uv,
                                                                       mask);
  return check_launch("feature_gather");
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// Per-frame class tables of the class-balanced draw (utils/common.py:312-322: unique(label) + nonzero per class): a
// STABLE counting sort of the window pixels by label.  One warp owns kClsChunk consecutive pixels and walks them 32 at
// a time in order (lanes = consecutive pixels, __match_any_sync groups equal labels), so the pixel order inside a class
// is ascending exactly like torch.nonzero.
// ---------------------------------------------------------------------------------------
namespace dns {
constexpr int kClsChunk = 4096;

__global__ void __launch_bounds__(32) k_cls_hist(const int64_t* __restrict__ label, int64_t n, int n_ids, int* __restrict__ hist,
                                                 int* __restrict__ err) {
  extern __shared__ int cnt[];   // [n_ids]
  const int lane = threadIdx.x;
  for (int c = lane; c < n_ids; c += 32) cnt[c] = 0;
  __syncwarp();
  const int64_t base = (int64_t)blockIdx.x * kClsChunk;
  for (int k = 0; k < kClsChunk; k += 32) {
    const int64_t i = base + k + lane;
    int64_t c = i < n ? label[i] : -1;
    if (i < n && (c < 0 || c >= n_ids)) {
      *err = 1;
      c = 0;
    }
    const unsigned m = __match_any_sync(0xffffffffu, c);
    if (c >= 0 && lane == __ffs(m) - 1) cnt[c] += __popc(m);
    __syncwarp();
  }
  for (int c = lane; c < n_ids; c += 32) hist[(int64_t)blockIdx.x * n_ids + c] = cnt[c];
}

// counts[c], starts[c] and the per-chunk offsets (in place over hist): one block, ids spread over the threads
__global__ void k_cls_scan(int* __restrict__ hist, int n_blocks, int n_ids, int* __restrict__ counts, int* __restrict__ starts) {
  for (int c = threadIdx.x; c < n_ids; c += blockDim.x) {
    int run = 0;
    for (int b = 0; b < n_blocks; ++b) {
      const int v = hist[(int64_t)b * n_ids + c];
      hist[(int64_t)b * n_ids + c] = run;
      run += v;
    }
    counts[c] = run;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int c = 0; c < n_ids; ++c) {
      starts[c] = run;
      run += counts[c];
    }
  }
}

__global__ void __launch_bounds__(32) k_cls_scatter(const int64_t* __restrict__ label, int64_t n, int n_ids,
                                                    const int* __restrict__ offs, const int* __restrict__ starts,
                                                    int64_t* __restrict__ order) {
  extern __shared__ int cur[];   // [n_ids] next free position of the class for this chunk
  const int lane = threadIdx.x;
  for (int c = lane; c < n_ids; c += 32) cur[c] = starts[c] + offs[(int64_t)blockIdx.x * n_ids + c];
  __syncwarp();
  const int64_t base = (int64_t)blockIdx.x * kClsChunk;
  for (int k = 0; k < kClsChunk; k += 32) {
    const int64_t i = base + k + lane;
    int64_t c = i < n ? label[i] : -1;
    if (i < n && (c < 0 || c >= n_ids)) c = 0;
    const unsigned m = __match_any_sync(0xffffffffu, c);
    int pos = -1;
    if (c >= 0) pos = cur[c] + __popc(m & ((1u << lane) - 1u));
    __syncwarp();
    if (c >= 0 && lane == __ffs(m) - 1) cur[c] += __popc(m);
    __syncwarp();
    if (pos >= 0) order[pos] = i;
  }
}
}  // namespace dns

extern "C" {

int64_t dns_class_tables_workspace_bytes(int64_t n_pixels, int n_ids) {
  const int64_t blocks = (n_pixels + dns::kClsChunk - 1) / dns::kClsChunk;
  return blocks * (int64_t)n_ids * 4 + 256;
}

int dns_class_tables(const int64_t* label, int64_t n_pixels, int n_ids, int64_t* order, int32_t* counts, int32_t* starts,
                     int32_t* err, void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace dns;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_pixels <= 0 || n_ids <= 0 || n_ids > 8192 || !label || !order || !counts || !starts || !err) {
    set_error("class_tables: bad arguments (1 <= n_ids <= 8192)");
    return DNS_ERR_ARG;
  }
  if (!workspace || workspace_bytes < dns_class_tables_workspace_bytes(n_pixels, n_ids)) {
    set_error("class_tables: workspace too small");
    return DNS_ERR_ARG;
  }
  const int blocks = (int)((n_pixels + kClsChunk - 1) / kClsChunk);
  int* hist = (int*)workspace;
  PhaseScope ph(phSample, st, 4);
  cudaMemsetAsync(err, 0, sizeof(int), st);
  k_cls_hist<<<blocks, 32, n_ids * sizeof(int), st>>>(label, n_pixels, n_ids, hist, err);
  k_cls_scan<<<1, 256, 0, st>>>(hist, blocks, n_ids, counts, starts);
  k_cls_scatter<<<blocks, 32, n_ids * sizeof(int), st>>>(label, n_pixels, n_ids, hist, starts, order);
  return check_launch("class_tables");
}

}  // extern "C"
