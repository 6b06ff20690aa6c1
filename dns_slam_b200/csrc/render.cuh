// Shared argument structs / enums of the fused render kernels (render.cu, ray_tc.cu).
#pragma once
#include "common.cuh"

namespace dns {

enum { kTrack = 0, kMap = 1, kTv = 2 };
// counts[] slots
enum { cMask = 0, cDpos = 1, cFront = 2, cBand = 3, cTiles = 4, cErr = 5 };
// raw loss sums
enum { rP = 0, rD = 1, rL = 2, rLt = 3, rFs = 4, rOp = 5 };

struct RayArgs {
  int mode;
  int S, T, RPC, C, C4;
  int64_t N_total, ray0, Nc;  // chunk of rays [ray0, ray0 + Nc)
  Bound B;
  const float* rays_o;
  const float* rays_d;
  const float* z;
  const float* gt_color;
  const float* gt_depth;
  const int64_t* gt_label;
  const uint8_t* mask;
  const float* features;
  const float* fine36;
  const float* W1T2;
  const float* W2cT;
  const float* logit;  // tcnn layout; W2l = logit + 32*112, [Cpad][32]
  const int* counts;
  float lam_p, lam_d, lam_l;
  float* pred_color;
  float* pred_depth;
  float* pred_var;
  float* pred_logits;
  float* raw;
  float* dfine36;
  float* d_features;
  float* d_rays_o;
  float* d_rays_d;
  // stashes (point / ray order of the chunk)
  float* X2;     // [Pc][112]
  float* dH2;    // [Pc][64]
  float* Hcol;   // [Pc][32]
  float* dpre;   // [Pc][4]
  float* dlogit; // [Nc][C4]
  float* Hbar;   // [Nc][32]
  int need_dparams, need_drays, need_dfeat;
};

// tcgen05 ray kernel (ray_tc.cu)
void pick_ray_block_tc(int S, int& T, int& RPC);
size_t ray_tc_smem_bytes(int T, int RPC, int C4);
int launch_ray_tc(const RayArgs& ra, const float* color, const float* logit, uint4* w1_hi, uint4* w1_lo, bool prep,
                  int64_t n_rays_chunk, cudaStream_t st);

}  // namespace dns
