// Camera poses of an iteration on the device: quaternion -> rotation, reference-view poses, and the
// closed-form backward of the ray construction.
//
//   dns_pose_prepare   utils/common.py:406-429 (quad2rotation, un-normalised quaternion (w,x,y,z), two_s = 2/|q|^2),
//                      slams/mapping.py:534-551 (reference views that follow a target frame use its CURRENT pose,
//                      detached) with the rigid inverse [R^T | -R^T t] in place of torch.inverse (common.py:672)
//   dns_pose_grad      autograd of  rays_d = R(q) dirs,  rays_o = T  (common.py:257-263):
//                        dT = sum_r d_rays_o[r],   G[a][b] = sum_r d_rays_d[r][a] dirs[r][b]   (= dL/dR)
//                        R = I + s A(q), s = 2/|q|^2:  dL/dq_m = s (G : dA/dq_m) - s^2 q_m (G : A)
#include "common.cuh"

namespace dns {

__device__ __forceinline__ void quat_to_rot(const float* q, float R[9]) {
  // element order and operations of common.py:419-428 (fp32, no contraction)
  const float qr = q[0], qi = q[1], qj = q[2], qk = q[3];
  const float n2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(qr, qr), __fmul_rn(qi, qi)), __fmul_rn(qj, qj)), __fmul_rn(qk, qk));
  const float two_s = __fdiv_rn(2.0f, n2);
  R[0] = __fsub_rn(1.f, __fmul_rn(two_s, __fadd_rn(__fmul_rn(qj, qj), __fmul_rn(qk, qk))));
  R[1] = __fmul_rn(two_s, __fsub_rn(__fmul_rn(qi, qj), __fmul_rn(qk, qr)));
  R[2] = __fmul_rn(two_s, __fadd_rn(__fmul_rn(qi, qk), __fmul_rn(qj, qr)));
  R[3] = __fmul_rn(two_s, __fadd_rn(__fmul_rn(qi, qj), __fmul_rn(qk, qr)));
  R[4] = __fsub_rn(1.f, __fmul_rn(two_s, __fadd_rn(__fmul_rn(qi, qi), __fmul_rn(qk, qk))));
  R[5] = __fmul_rn(two_s, __fsub_rn(__fmul_rn(qj, qk), __fmul_rn(qi, qr)));
  R[6] = __fmul_rn(two_s, __fsub_rn(__fmul_rn(qi, qk), __fmul_rn(qj, qr)));
  R[7] = __fmul_rn(two_s, __fadd_rn(__fmul_rn(qj, qk), __fmul_rn(qi, qr)));
  R[8] = __fsub_rn(1.f, __fmul_rn(two_s, __fadd_rn(__fmul_rn(qi, qi), __fmul_rn(qj, qj))));
}

__global__ void k_pose_prepare(const float* __restrict__ quats, const float* __restrict__ trans, int F,
                               const int* __restrict__ view_src, const float* __restrict__ fixed_w2c,
                               const float* __restrict__ fixed_cam_o, int V, float* __restrict__ R_out,
                               float* __restrict__ w2c, float* __restrict__ cam_o) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < F && R_out) {
    float R[9];
    quat_to_rot(quats + 4 * t, R);
    for (int k = 0; k < 9; ++k) R_out[9 * t + k] = R[k];
  }
  if (t < V && w2c) {
    const int src = view_src ? view_src[t] : -1;
    if (src >= 0 && src < F) {
      float R[9];
      quat_to_rot(quats + 4 * src, R);
      const float* T = trans + 3 * src;
      float* M = w2c + 16 * t;
      for (int a = 0; a < 3; ++a) {   // [R^T | -R^T t]
        float acc = 0.f;
        for (int b = 0; b < 3; ++b) {
          M[4 * a + b] = R[3 * b + a];
          acc += R[3 * b + a] * T[b];
        }
        M[4 * a + 3] = -acc;
        cam_o[3 * t + a] = T[a];
      }
      M[12] = M[13] = M[14] = 0.f;
      M[15] = 1.f;
    } else {
      for (int k = 0; k < 16; ++k) w2c[16 * t + k] = fixed_w2c[16 * t + k];
      for (int k = 0; k < 3; ++k) cam_o[3 * t + k] = fixed_cam_o[3 * t + k];
    }
  }
}

// per frame: partial sums of dT (3) and G = dL/dR (9) over the frame's rays
// all frames in one launch (blockIdx.y = frame; ray ranges by value)
constexpr int kMaxPoseFrames = 32;
struct PoseRanges {
  int start[kMaxPoseFrames + 1];
};
__global__ void __launch_bounds__(256) k_pose_reduce(const float* __restrict__ d_o, const float* __restrict__ d_d,
                                                     const int64_t* __restrict__ pixel, PoseRanges rg, int f0, int H0, int W0,
                                                     int Ww, float fx, float fy, float cx, float cy, float* __restrict__ out) {
  const int r0 = rg.start[blockIdx.y], r1 = rg.start[blockIdx.y + 1];
  float* const out12 = out + 12 * (f0 + blockIdx.y);
  __shared__ float red[32];
  float acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.f;
  for (int r = r0 + blockIdx.x * blockDim.x + threadIdx.x; r < r1; r += gridDim.x * blockDim.x) {
    const int64_t idx = pixel[r];
    const float i = (float)(W0 + (int)(idx % Ww)), j = (float)(H0 + (int)(idx / Ww));
    const float dir[3] = {__fdiv_rn(__fsub_rn(i, cx), fx), __fdiv_rn(-__fsub_rn(j, cy), fy), -1.0f};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      acc[a] += d_o[3 * r + a];
      const float g = d_d[3 * r + a];
#pragma unroll
      for (int b = 0; b < 3; ++b) acc[3 + 3 * a + b] = fmaf(g, dir[b], acc[3 + 3 * a + b]);
    }
  }
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    const float v = block_reduce_sum(acc[k], red);
    if (threadIdx.x == 0 && v != 0.f) atomicAdd(out12 + k, v);
  }
}

__global__ void k_pose_finish(const float* __restrict__ sums, const float* __restrict__ quats, int F, float* __restrict__ d_quats,
                              float* __restrict__ d_trans) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const float* s12 = sums + 12 * f;
  const float* G = s12 + 3;
  if (d_trans)
    for (int a = 0; a < 3; ++a) d_trans[3 * f + a] = s12[a];
  if (!d_quats) return;
  const float r = quats[4 * f], i = quats[4 * f + 1], j = quats[4 * f + 2], k = quats[4 * f + 3];
  const float s = 2.0f / (r * r + i * i + j * j + k * k);
  // A(q): R = I + s A
  const float A[9] = {-(j * j + k * k), i * j - k * r, i * k + j * r, i * j + k * r, -(i * i + k * k), j * k - i * r,
                      i * k - j * r, j * k + i * r, -(i * i + j * j)};
  float GA = 0.f;
  for (int e = 0; e < 9; ++e) GA += G[e] * A[e];
  // dA/dq_m contracted with G
  const float dr = G[1] * (-k) + G[2] * j + G[3] * k + G[5] * (-i) + G[6] * (-j) + G[7] * i;
  const float di = G[1] * j + G[2] * k + G[3] * j + G[4] * (-2.f * i) + G[5] * (-r) + G[6] * k + G[7] * r + G[8] * (-2.f * i);
  const float dj = G[0] * (-2.f * j) + G[1] * i + G[2] * r + G[3] * i + G[5] * k + G[6] * (-r) + G[7] * k + G[8] * (-2.f * j);
  const float dk = G[0] * (-2.f * k) + G[1] * (-r) + G[2] * i + G[3] * r + G[4] * (-2.f * k) + G[5] * j + G[6] * i + G[7] * j;
  const float q[4] = {r, i, j, k}, dq[4] = {dr, di, dj, dk};
  for (int m = 0; m < 4; ++m) d_quats[4 * f + m] = s * dq[m] - s * s * q[m] * GA;
}


// best-pose bookkeeping of the tracking loop (slams/tracking.py:331-338), one thread: the reference compares the loss
// on the host every iteration; here the comparison, the copy of the current pose, the loss history and the running
// error flag stay on the device
// loss / result bookkeeping of one native mapping iteration (dns_map_step_result)
__global__ void k_map_step_result(int phase, const float* __restrict__ losses8, const float* __restrict__ tv_loss, float tv_w,
                                  float tv_lambda_w, float nvalid_scale, const float* __restrict__ scratch, int F,
                                  float* loss_vec, float* result) {
  const int t = threadIdx.x;
  if (phase & 1) {
    if (t < 8) {
      float v = losses8[t];
      if (t == 6 && tv_loss) v += tv_lambda_w * *tv_loss;
      loss_vec[t] = v;
    }
    if (t == 8 && tv_loss) loss_vec[8] = tv_w * *tv_loss;
    __syncthreads();
  }
  if (phase & 2) {
    if (t == 7) loss_vec[7] *= nvalid_scale;
    __syncthreads();
    if (t < 9) result[t] = loss_vec[t];
    for (int e = t; e < 2 * F; e += blockDim.x) result[9 + e] = scratch[e];
    if (t == 0) {
      float outside = 0.f;
      for (int f = 0; f < F; ++f) outside += scratch[2 * f + 1];
      result[9 + 2 * F] += outside;
      result[10 + 2 * F] = fminf(result[10 + 2 * F], loss_vec[7]);
    }
  }
}
__global__ void k_track_best(const float* __restrict__ losses, const float* __restrict__ quat, const float* __restrict__ trans,
                             float* best7, float* best_loss, float* hist, int* slot, int hist_len, float* err_min) {
  const float l = losses[6];
  if (l < *best_loss) {
    *best_loss = l;
    for (int k = 0; k < 4; ++k) best7[k] = quat[k];
    for (int k = 0; k < 3; ++k) best7[4 + k] = trans[k];
  }
  const int s = *slot;
  if (hist && s < hist_len) hist[s] = l;
  *slot = s + 1;
  if (losses[7] < *err_min) *err_min = losses[7];
}

}  // namespace dns

using namespace dns;

extern "C" {

int dns_pose_prepare(const float* quats, const float* trans, int n_frames, const int32_t* view_src, const float* fixed_w2c,
                     const float* fixed_cam_o, int n_views_total, float* R, float* w2c, float* cam_o, void* stream) {
  if (n_frames <= 0 || !quats || !trans) {
    set_error("pose_prepare: bad arguments");
    return DNS_ERR_ARG;
  }
  if (n_views_total > 0 && (!w2c || !cam_o || !fixed_w2c || !fixed_cam_o)) {
    set_error("pose_prepare: view outputs / fixed poses missing");
    return DNS_ERR_ARG;
  }
  const int n = n_frames > n_views_total ? n_frames : n_views_total;
  PhaseScope ph(phSample, (cudaStream_t)stream, 1);
  k_pose_prepare<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(quats, trans, n_frames, view_src, fixed_w2c, fixed_cam_o,
                                                                n_views_total, R, n_views_total > 0 ? w2c : nullptr, cam_o);
  return check_launch("pose_prepare");
}

int dns_pose_grad(const float* d_rays_o, const float* d_rays_d, const int64_t* pixel, int n_frames, const int32_t* ray_start,
                  int H0, int W0, int Ww, float fx, float fy, float cx, float cy, const float* quats, float* d_quats,
                  float* d_trans, float* scratch, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n_frames <= 0 || !d_rays_o || !d_rays_d || !pixel || !ray_start || !scratch || (d_quats && !quats)) {
    set_error("pose_grad: bad arguments");
    return DNS_ERR_ARG;
  }
  PhaseScope ph(phFinalize, st, 3);
  cudaMemsetAsync(scratch, 0, sizeof(float) * 12 * n_frames, st);
  for (int f0 = 0; f0 < n_frames; f0 += kMaxPoseFrames) {
    const int nf = n_frames - f0 < kMaxPoseFrames ? n_frames - f0 : kMaxPoseFrames;
    PoseRanges rg;
    int n_max = 0;
    for (int f = 0; f <= nf; ++f) rg.start[f] = ray_start[f0 + f];
    for (int f = 0; f < nf; ++f) n_max = rg.start[f + 1] - rg.start[f] > n_max ? rg.start[f + 1] - rg.start[f] : n_max;
    if (n_max <= 0) continue;
    int blocks = (n_max + 255) / 256;
    if (blocks > 148) blocks = 148;
    k_pose_reduce<<<dim3(blocks, nf), 256, 0, st>>>(d_rays_o, d_rays_d, pixel, rg, f0, H0, W0, Ww, fx, fy, cx, cy, scratch);
  }
  k_pose_finish<<<(n_frames + 31) / 32, 32, 0, st>>>(scratch, quats, n_frames, d_quats, d_trans);
  return check_launch("pose_grad");
}

int dns_track_best(const float* losses, const float* quat, const float* trans, float* best7, float* best_loss, float* hist,
                   int32_t* slot, int hist_len, float* err_min, void* stream) {
  if (!losses || !quat || !trans || !best7 || !best_loss || !slot || !err_min) {
    set_error("track_best: bad arguments");
    return DNS_ERR_ARG;
  }
  PhaseScope ph(phFinalize, (cudaStream_t)stream, 1);
  k_track_best<<<1, 1, 0, (cudaStream_t)stream>>>(losses, quat, trans, best7, best_loss, hist, slot, hist_len, err_min);
  return check_launch("track_best");
}

int dns_map_step_result(int phase, const float* losses8, const float* tv_loss, float tv_w, float tv_lambda_w,
                        float nvalid_scale, const float* scratch, int n_frames, float* loss_vec9, float* result,
                        void* stream) {
  if (!(phase & 3) || !loss_vec9 || ((phase & 1) && !losses8) || ((phase & 2) && (!scratch || !result || n_frames < 0))) {
    set_error("map_step_result: bad arguments");
    return DNS_ERR_ARG;
  }
  PhaseScope ph(phFinalize, (cudaStream_t)stream, 1);
  k_map_step_result<<<1, 64, 0, (cudaStream_t)stream>>>(phase, losses8, tv_loss, tv_w, tv_lambda_w, nvalid_scale, scratch,
                                                        n_frames, loss_vec9, result);
  return check_launch("map_step_result");
}

}  // extern "C"
