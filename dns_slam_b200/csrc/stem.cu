// ResNet stem of the pixel-feature branch: what models/encoder.py:9-17 runs once per frame,
//   conv1 7x7 / stride 2 / pad 3, 3 -> 64, no bias          (models/layers.py:55-56, 98)
//   bn1   BatchNorm2d(64)                                   (models/layers.py:57, 99)
//   ReLU                                                    (models/layers.py:58, 100)
// The reference never calls .eval() on the encoder (slams/tracking.py:30, slams/mapping.py:33), so bn1 normalises
// with the statistics of the CURRENT batch (all views of the call, biased variance) and updates its running
// statistics (momentum 0.1, unbiased variance); training = 0 gives the eval-mode form on the running statistics.
//
// Output is written channels-last [n][h][w][64] -- the layout dns_feature_gather reads -- so the NCHW tensor of the
// reference and its per-iteration 64 x H x W up-sample (utils/common.py:646) never exist.
//
//   k_stem_prep      zero the statistics, weights [co][ci][ky][kx] -> [ky][kx][ci][co]
//   k_stem_conv      4 x 64 output pixels x 64 channels per CTA; a thread = 4 pixels x 16 channels (64 fp32
//                    accumulators); input tile staged in shared memory with even / odd columns de-interleaved
//                    (stride-2 taps become unit-stride, conflict-free reads), weights read as warp-uniform
//                    broadcasts; per-channel sum / sum of squares reduced per CTA, one fp64 atomic per value
//   k_stem_finalize  mean / variance -> scale / shift, running statistics
//   k_stem_bn_relu   in-place normalise + ReLU (float4)
#include "common.cuh"

namespace dns {

constexpr int kSoR = 4, kSoC = 64;                   // output tile
constexpr int kSiR = 2 * kSoR + 5;                   // 13 input rows
constexpr int kSiC = 2 * kSoC + 5;                   // 133 input columns
constexpr int kSiH = 72;                             // padded half-row (67 used): two rows apart = 16 banks
constexpr int kStemW = 147 * 64;                     // weights
constexpr int kStemIn = 2 * 3 * kSiR * kSiH;         // [parity][ci][row][half col]

struct StemWs {
  double stats[128];   // sum[64] | sum of squares[64]
  float scale[64];
  float shift[64];
  float wt[kStemW];    // [ky][kx][ci][co]
};

__global__ void k_stem_prep(const float* __restrict__ conv_w, StemWs* ws) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < 128) ws->stats[g] = 0.0;
  if (g < kStemW) {
    const int co = g / 147, rem = g % 147, ci = rem / 49, k = rem % 49;   // k = ky * 7 + kx
    ws->wt[(k * 3 + ci) * 64 + co] = conv_w[g];
  }
}

__global__ void __launch_bounds__(256, 2) k_stem_conv(const float* __restrict__ img, int H, int W, int h, int w,
                                                      StemWs* ws, float* __restrict__ out) {
  extern __shared__ __align__(16) float smf[];
  float* wt = smf;               // [147][64]
  float* in = smf + kStemW;      // [2][3][13][72]
  __shared__ float sred[8][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ox0 = blockIdx.x * kSoC, oy0 = blockIdx.y * kSoR, n = blockIdx.z;
  for (int e = tid; e < kStemW / 4; e += 256)
    reinterpret_cast<float4*>(wt)[e] = reinterpret_cast<const float4*>(ws->wt)[e];
  {
    const int gx0 = 2 * ox0 - 3, gy0 = 2 * oy0 - 3;
    const float* src = img + (int64_t)n * H * W * 3;
    for (int e = tid; e < kSiR * kSiC * 3; e += 256) {
      const int r = e / (kSiC * 3), rem = e - r * (kSiC * 3), c = rem / 3, ci = rem - 3 * c;
      const int gy = gy0 + r, gx = gx0 + c;
      float v = 0.f;
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(src + ((int64_t)gy * W + gx) * 3 + ci);
      in[(((c & 1) * 3 + ci) * kSiR + r) * kSiH + (c >> 1)] = v;
    }
  }
  __syncthreads();
  const int cg = tid >> 6, pq = tid & 63, orow = pq >> 4, q = pq & 15;
  float acc[4][16];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[j][k] = 0.f;
#pragma unroll 1
  for (int ci = 0; ci < 3; ++ci) {
#pragma unroll 1
    for (int ky = 0; ky < 7; ++ky) {
      const float* irow_e = in + ((0 * 3 + ci) * kSiR + 2 * orow + ky) * kSiH + q;
      const float* irow_o = in + ((1 * 3 + ci) * kSiR + 2 * orow + ky) * kSiH + q;
      const float* wrow = wt + ((ky * 7) * 3 + ci) * 64 + cg * 16;
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const float* ir = ((kx & 1) ? irow_o : irow_e) + (kx >> 1);
        float x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = ir[16 * j];
        const float4* w4 = reinterpret_cast<const float4*>(wrow + kx * 3 * 64);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const float4 wv = w4[k4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[j][4 * k4 + 0] = fmaf(x[j], wv.x, acc[j][4 * k4 + 0]);
            acc[j][4 * k4 + 1] = fmaf(x[j], wv.y, acc[j][4 * k4 + 1]);
            acc[j][4 * k4 + 2] = fmaf(x[j], wv.z, acc[j][4 * k4 + 2]);
            acc[j][4 * k4 + 3] = fmaf(x[j], wv.w, acc[j][4 * k4 + 3]);
          }
        }
      }
    }
  }
  // pre-normalisation values + statistics of the valid pixels
  const int oy = oy0 + orow;
  float s1[16], s2[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) s1[k] = s2[k] = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ox = ox0 + q + 16 * j;
    if (oy < h && ox < w) {
      float4* d4 = reinterpret_cast<float4*>(out + (((int64_t)n * h + oy) * w + ox) * 64 + cg * 16);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4)
        d4[k4] = make_float4(acc[j][4 * k4], acc[j][4 * k4 + 1], acc[j][4 * k4 + 2], acc[j][4 * k4 + 3]);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        s1[k] += acc[j][k];
        s2[k] = fmaf(acc[j][k], acc[j][k], s2[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], o);
      s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], o);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      sred[warp][k] = s1[k];
      sred[warp][16 + k] = s2[k];
    }
  }
  __syncthreads();
  if (tid < 128) {   // channel group g = warps 2g and 2g + 1
    const int g = tid >> 5, v = tid & 31;
    const float t = sred[2 * g][v] + sred[2 * g + 1][v];
    atomicAdd(ws->stats + (v >> 4) * 64 + g * 16 + (v & 15), (double)t);
  }
}

__global__ void k_stem_finalize(StemWs* ws, double count, const float* __restrict__ gamma, const float* __restrict__ beta,
                                float eps, float momentum, int training, float* running_mean, float* running_var) {
  const int c = threadIdx.x;
  if (c >= 64) return;
  double mean, var;
  if (training) {
    mean = ws->stats[c] / count;
    var = ws->stats[64 + c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    if (running_var) {
      const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const double inv = 1.0 / sqrt(var + (double)eps);
  const double sc = (double)gamma[c] * inv;
  ws->scale[c] = (float)sc;
  ws->shift[c] = (float)((double)beta[c] - mean * sc);
}

__global__ void k_stem_bn_relu(float4* __restrict__ out, int64_t n4, const StemWs* __restrict__ ws) {
  __shared__ float4 sc4[16], sh4[16];
  if (threadIdx.x < 16) {
    sc4[threadIdx.x] = reinterpret_cast<const float4*>(ws->scale)[threadIdx.x];
    sh4[threadIdx.x] = reinterpret_cast<const float4*>(ws->shift)[threadIdx.x];
  }
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;   // a multiple of 16: the channel group of a thread is fixed
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const float4 sc = sc4[i & 15], sh = sh4[i & 15];
  for (; i < n4; i += stride) {
    float4 v = out[i];
    v.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f);
    v.y = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
    v.z = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f);
    v.w = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
    out[i] = v;
  }
}

}  // namespace dns

using namespace dns;

extern "C" {

int64_t dns_stem_workspace_bytes(void) { return (int64_t)sizeof(StemWs); }

int dns_stem_fwd(const float* images, int n, int H, int W, const float* conv_w, const float* bn_weight,
                 const float* bn_bias, float eps, float momentum, int training, float* running_mean,
                 float* running_var, float* out, void* workspace, int64_t workspace_bytes, void* stream) {
  if (n <= 0 || H <= 0 || W <= 0) return DNS_OK;
  if (!images || !conv_w || !bn_weight || !bn_bias || !out || !workspace) {
    set_error("dns_stem_fwd: null pointer");
    return DNS_ERR_ARG;
  }
  if (workspace_bytes < (int64_t)sizeof(StemWs)) {
    set_error("dns_stem_fwd: workspace of %lld bytes, need %lld", (long long)workspace_bytes, (long long)sizeof(StemWs));
    return DNS_ERR_ARG;
  }
  if (!training && (!running_mean || !running_var)) {
    set_error("dns_stem_fwd: eval mode needs the running statistics");
    return DNS_ERR_ARG;
  }
  const int h = (H - 1) / 2 + 1, w = (W - 1) / 2 + 1;
  const int gy = (h + kSoR - 1) / kSoR;
  if (n > 65535 || gy > 65535) {
    set_error("dns_stem_fwd: %d views of %d rows exceed the grid limits", n, h);
    return DNS_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  StemWs* ws = (StemWs*)workspace;
  static unsigned long long seen = 0;
  const int smem = (kStemW + kStemIn) * (int)sizeof(float);
  if (first_call_on_device(seen)) cudaFuncSetAttribute(k_stem_conv, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  PhaseScope ph(phFeature, st, 4);
  k_stem_prep<<<(kStemW + 255) / 256, 256, 0, st>>>(conv_w, ws);
  dim3 grid((w + kSoC - 1) / kSoC, gy, n);
  k_stem_conv<<<grid, 256, smem, st>>>(images, H, W, h, w, ws, out);
  k_stem_finalize<<<1, 64, 0, st>>>(ws, (double)n * h * w, bn_weight, bn_bias, eps, momentum, training, running_mean,
                                    running_var);
  const int64_t n4 = (int64_t)n * h * w * 16;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_stem_bn_relu<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<float4*>(out), n4, ws);
  return check_launch("stem_fwd");
}

}  // extern "C"
