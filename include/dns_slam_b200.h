/*
 * dns_slam_b200 -- C ABI of the B200-native (sm_100a) DNS-SLAM render-and-optimise hot path.
 *
 * Plain C, raw DEVICE pointers + explicit cudaStream_t (passed as void*), int error codes
 * (0 = ok; dns_last_error() gives the text), caller-owned memory, no hidden allocations
 * (scratch comes from a caller-provided workspace sized by dns_render_workspace_bytes /
 * dns_tv_workspace_bytes), no global state.  There is NO CPU fallback: every entry point
 * launches CUDA kernels.
 *
 * The reference (li-kunyi/dns-slam) has no FFI of its own; its native boundary is the
 * un-vendored `tinycudann` torch binding.  Each entry point below cites the reference
 * interface it replaces (paths relative to the reference root).
 */
#ifndef DNS_SLAM_B200_H
#define DNS_SLAM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DNS_MAX_LEVELS 16
#define DNS_HIDDEN 32      /* n_neurons of every MLP on the path (configs/slam.yaml:18) */
#define DNS_PE_DIM 48      /* OneBlob 3 x 16 bins  (models/pos_encoding.py:61-71) */
#define DNS_GRID_DIM 32    /* 16 levels x 2 features (models/pos_encoding.py:31-46) */
#define DNS_LATENT 33      /* occupancy logit + 32 latent channels (models/decoder.py:85) */
#define DNS_FEAT_DIM 32    /* merged pixel feature width (models/decoder.py:59) */

enum { DNS_OK = 0, DNS_ERR_ARG = 1, DNS_ERR_CUDA = 2, DNS_ERR_UNSUPPORTED = 3 };
enum { DNS_MODE_TRACK = 0, DNS_MODE_MAP = 1 };

/* Multi-resolution hash grid geometry; computed ONCE on the host (fp32 emulation of
 * tiny-cuda-nn grid_scale / grid_resolution) and shared with the oracle so that the hash
 * indices are bit exact.  Replaces the encoding_config dict of
 * models/pos_encoding.py:34-45. */
typedef struct dns_grid {
  int32_t n_levels;                     /* <= DNS_MAX_LEVELS */
  int32_t n_features;                   /* 2 */
  float scale[DNS_MAX_LEVELS];          /* pos = fma(scale, x, 0.5) */
  uint32_t res[DNS_MAX_LEVELS];         /* grid resolution of the level */
  uint32_t size[DNS_MAX_LEVELS];        /* entries in the level (multiple of 8, <= 2^log2_T) */
  uint32_t offset[DNS_MAX_LEVELS + 1];  /* running sum of size[], in entries */
  uint32_t hashed[DNS_MAX_LEVELS];      /* 1: coherent-prime hash, 0: dense index */
} dns_grid;

const char* dns_last_error(void);
int dns_version(void);

/* ---------------------------------------------------------------------------------------
 * Operator surface == the tinycudann module API the reference calls
 * ------------------------------------------------------------------------------------- */

/* tcnn.Encoding(otype=OneBlob).forward / backward  (models/pos_encoding.py:61-71,
 * called at models/decoder.py:46,71).  x [P,D] f32 -> out [P,D*n_bins] f32. */
int dns_oneblob_fwd(const float* x, int64_t P, int D, int n_bins, float* out, void* stream);
int dns_oneblob_bwd(const float* x, const float* d_out, int64_t P, int D, int n_bins, float* d_x,
                    void* stream);

/* tcnn.Encoding(otype=HashGrid).forward / backward (models/pos_encoding.py:31-46, called at
 * models/decoder.py:47).  x [P,3] f32 in [0,1], table [n_entries,2] f32 -> out [P,2L].
 * bwd ACCUMULATES into d_table (atomics) and, if d_x != NULL, writes dL/dx [P,3]. */
int dns_hashgrid_fwd(const dns_grid* g, const float* x, const float* table, int64_t P, float* out,
                     void* stream);
int dns_hashgrid_bwd(const dns_grid* g, const float* x, const float* table, const float* d_out,
                     int64_t P, float* d_table, float* d_x, void* stream);
/* bit-exactness probe used by the parity tests: table indices of the 8 corners,
 * idx [P, L, 8] uint32 (absolute entry index incl. level offset). */
int dns_hashgrid_indices(const dns_grid* g, const float* x, int64_t P, uint32_t* idx, void* stream);

/* tcnn.Network(CutlassMLP, 1 hidden layer of 32, ReLU, no bias).forward / backward
 * (models/decoder.py:58-65,84-91,101-117; slams/mapping.py:737-744).
 * params = [W1 (32 x n_in) | W2 (out_pad x 32)] row-major, n_in % 16 == 0.
 * hidden [P,32] (post-ReLU) is written by fwd and consumed by bwd.
 * bwd writes d_hidden [P,32] (caller-owned scratch), ACCUMULATES into d_params (may be NULL)
 * and writes d_x [P,n_in] (may be NULL). */
int dns_mlp_fwd(const float* x, const float* params, int64_t P, int n_in, int n_out, float* out,
                float* hidden, void* stream);
int dns_mlp_bwd(const float* x, const float* params, const float* hidden, const float* d_out,
                int64_t P, int n_in, int n_out, float* d_hidden, float* d_x, float* d_params,
                void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused path == the bodies of Tracker.renderer + losses + backward
 * (slams/tracking.py:188-214,85-96,326-338) and Mapper.renderer + losses + backward
 * (slams/mapping.py:590-635,110-126,891-909; utils/common.py:506-537,769-802)
 * ------------------------------------------------------------------------------------- */
typedef struct dns_render_args {
  int32_t mode;          /* DNS_MODE_TRACK | DNS_MODE_MAP */
  int32_t n_rays;        /* N (after the inside-mask compaction in MAP mode) */
  int32_t n_samples;     /* S = n_samples_ray + n_surface_ray */
  int32_t n_class;       /* C: width of the semantic head */
  int32_t n_experts;     /* MAP: rows of `experts` */
  int32_t need_dparams;  /* accumulate hash-table / MLP gradients */
  int32_t need_drays;    /* write d_rays_o / d_rays_d (pose gradients) */
  int32_t need_dfeat;    /* write d_features (gradient into the Merge branch) */
  double bound[3][2];    /* float64 scene bound (slams/dns_slam.py:100-107) */
  /* loss weights: lambda_color, lambda_depth, lambda_label, lambda_lt, lambda_fs,
   * lambda_opacity (slams/mapping.py:906-907, slams/tracking.py:329) */
  float lambda_p, lambda_d, lambda_l, lambda_lt, lambda_fs, lambda_op;
  float opacity_trunc;   /* = cfg opacity_sigma: lands in the truncation slot (mapping.py:896) */
  float opacity_sigma;   /* = 0.05 default of utils/common.py:769 */
  dns_grid grid;
  /* inputs (device) */
  const float* rays_o;     /* [N,3] */
  const float* rays_d;     /* [N,3] */
  const float* z_vals;     /* [N,S] */
  const float* gt_color;   /* [N,3] */
  const float* gt_depth;   /* [N] */
  const int64_t* gt_label; /* [N] */
  const uint8_t* mask;     /* TRACK: [N] valid-ray mask (tracking.py:174-175); NULL = all */
  const float* features;   /* [N,S,32] merged pixel features, may be NULL (= zeros) */
  /* parameters (device) */
  const float* table;      /* [n_entries,2] */
  const float* coarse;     /* 80->32->33(48)   W1[32][80] | W2[48][32] */
  const float* color;      /* 112->32->3(16) */
  const float* logit;      /* 112->32->C(pad16) */
  const float* experts;    /* MAP: [n_experts][3616+...] same layout as coarse */
  const int32_t* class_to_expert; /* MAP: [n_class_ids] expert row of a label id, -1 = none */
  int32_t n_class_ids;
  /* outputs (device) */
  float* pred_color;   /* [N,3] */
  float* pred_depth;   /* [N] */
  float* pred_var;     /* [N] */
  float* pred_logits;  /* [N,C] */
  float* fine;         /* [P,33] (MAP; TRACK: the coarse latents), may be NULL */
  float* coarse_out;   /* [P,33] (MAP), may be NULL */
  float* losses;       /* [8]: p, d, l, lt, fs, op, total, n_valid */
  /* gradients of `total` (device); accumulated (+=) unless noted */
  float* d_table;
  float* d_coarse;
  float* d_color;
  float* d_logit;
  float* d_experts;    /* [n_experts][...] */
  float* d_rays_o;     /* [N,3] overwritten */
  float* d_rays_d;     /* [N,3] overwritten */
  float* d_features;   /* [N,S,32] overwritten */
  /* scratch */
  void* workspace;
  int64_t workspace_bytes;
  /* ray sharding across GPUs; all zero / NULL for a single-GPU call.  The n_rays
   * local rays are rays [ray_offset, ray_offset + n_rays) of a batch of n_rays_total: loss
   * denominators, the count_nonzero guards and the class rule class(p) = label[p mod N]
   * (mapping.py:612-613) use the GLOBAL batch, so per-rank losses / gradients SUM to the
   * single-GPU result. */
  int64_t n_rays_total;
  int64_t ray_offset;
  const int64_t* gt_label_all;   /* [n_rays_total] labels of the whole batch */
  const int32_t* global_counts;  /* [4] all-reduced dns_render_counts output */
  /* inference (slams/mapping.py:638-724 frame_vis, slams/meshing.py:461-498 eval_points): predictions and latents
   * only -- no loss gradients, no backward kernels, gradient outputs untouched.
   * 1: rays (compositing as in training); 2: free points (n_samples must be 1: rays_o holds the points, the
   * colour / logits of the single sample are returned without the occupancy weight). */
  int32_t forward_only;
  /* 0 (default): every MLP contraction on tcgen05 tensor cores (bf16 hi+lo split, three products, fp32 accumulation in
   * TMEM); 1: the fp32 SIMT kernels -- kept as the A/B reference of the parity tests.  Per call: the library holds
   * no mode state. */
  int32_t use_simt;
  /* 1: `features` comes from dns_featmerge_fwd with no_zero_fill = 1 and `d_features` goes to dns_featmerge_bwd: rows of
   * samples OUTSIDE the truncation band (slams/tracking.py:167-170: z within +-5 % of a positive gt depth; two thirds of
   * the samples) are taken as zero without being read, and their d_features rows are not written.  tcgen05 path only. */
  int32_t features_band_only;
  int32_t reserved_;
} dns_render_args;

/* table and d_table must be 16-byte aligned (the kernels fetch / reduce the x, x+1 corner pair of a cell edge
 * with one 16-byte access when both entries share an aligned pair). */
int64_t dns_render_workspace_bytes(int mode, int n_rays, int n_samples, int n_class, int n_class_ids);
int dns_render_fwd_bwd(const dns_render_args* a, void* stream);
/* Local batch counts {n_mask, n_depth>0, n_front, n_band} (int32[4], device) that the loss
 * denominators / guards of dns_render_fwd_bwd use; all-reduce (sum) them across ranks and pass
 * the result as global_counts. */
int dns_render_counts(const dns_render_args* a, int32_t* counts4, void* stream);

/* Mapper.smoothness forward + backward (slams/mapping.py:129-159, oracle patch P1): TV of the
 * coarse occupancy on an n^3 lattice (n = smooth_pts - 1).  Point (i,j,k) is
 *   x_a = ((((double)idx_a + jitter[a]) * voxel + bound_lo[a]) + offset[a] - bound_lo[a]) / extent[a]
 * loss = sum of squared forward differences / smooth_pts^3, gradients scaled by lambda. */
typedef struct dns_tv_args {
  int32_t n;              /* lattice points per axis = smooth_pts - 1 */
  int32_t smooth_pts;
  double voxel;           /* 0.1 */
  double bound[3][2];
  double offset[3];       /* rand(3) * offset_max + margin, float64 */
  double jitter[3];       /* rand(1,1,1,3) as float64 */
  float lambda_sm;
  int32_t need_dparams;
  dns_grid grid;
  const float* table;
  const float* coarse;
  float* loss;            /* [1] unweighted TV loss */
  float* d_table;         /* += */
  float* d_coarse;        /* += */
  void* workspace;
  int64_t workspace_bytes;
  /* optional: offset[3] | jitter[3] as float64 in DEVICE memory (overrides the by-value fields), so that a
   * captured CUDA graph can be replayed with new draws */
  const double* offset_jitter_dev;
  int32_t use_simt;       /* as in dns_render_args */
  int32_t reserved_;
} dns_tv_args;

int64_t dns_tv_workspace_bytes(int n);
int dns_tv_fwd_bwd(const dns_tv_args* a, void* stream);

/* ---------------------------------------------------------------------------------------
 * Sampling == get_samples / get_rays_from_uv / far plane / sample_along_rays / point build
 * (utils/common.py:248-304,561-599; slams/tracking.py:137-160; slams/mapping.py:497-531)
 * ------------------------------------------------------------------------------------- */
typedef struct dns_sample_args {
  int32_t n;              /* rays */
  int32_t H, W;           /* image size */
  int32_t H0, W0, Ww;     /* window origin and width: flat index -> (H0 + idx / Ww, W0 + idx % Ww) */
  int32_t n_uniform;      /* linspace samples (n_samples_ray) */
  int32_t n_surface;      /* depth-guided samples (n_surface_ray) */
  float fx, fy, cx, cy;
  double bound[3][2];
  const float* color;     /* [H,W,3] */
  const float* depth;     /* [H,W] */
  const int64_t* label;   /* [H,W] */
  const int64_t* index;   /* [n] flat window indices (draws hoisted, oracle patch P5) */
  const float* R;         /* [3,3] row-major c2w rotation */
  const float* T;         /* [3] */
  const float* t_lin;     /* [n_uniform] torch.linspace(0,1) */
  const float* t_surface; /* [n_surface] first draw (element n_surface/2+1 already forced to 0.5) */
  const float* t_zero;    /* [n_surface] second draw (zero-depth rays) */
  float* gt_color;        /* [n,3] */
  float* gt_depth;        /* [n] */
  int64_t* gt_label;      /* [n] */
  float* rays_o;          /* [n,3] */
  float* rays_d;          /* [n,3] */
  float* z_vals;          /* [n,S] ascending */
  float* pts;             /* [n,S,3] or NULL */
  uint8_t* inside;        /* [n] far_bb >= gt_depth */
  float* scratch;         /* [2] device scratch: [0] batch max depth, [1] number of rays with inside == 0 */
  /* class-balanced draws resolved on the device (utils/common.py:307-330, select_by_class): rays
   * [n_direct, n) take the window index  order[slot_base[r - n_direct] + index[r]]  where `order` lists the window
   * pixels sorted by label (stable) and slot_base[j] is the first entry of the class that slot j draws from; rays
   * [0, n_direct) use index[r] directly (the uniform draws of common.py:274).  order == NULL: n_direct = n. */
  const int64_t* order;
  const int32_t* slot_base;
  int32_t n_direct;
  /* 0: the whole call.  A frame whose rays are split over several GPUs needs the max depth of ALL its rays
   * (utils/common.py:581,591) between the two halves: 1 = pixel gather + rays only (scratch[0] = local max depth),
   * then max-all-reduce scratch[0] across the ranks, then 2 = far plane + z values only. */
  int32_t phase;
  int64_t* pixel;         /* [n] resolved flat window index of every ray, or NULL */
} dns_sample_args;

int dns_sample_rays(const dns_sample_args* a, void* stream);
/* The same for several frames of one iteration (the target frames of slams/mapping.py:519-531) in ONE pair of launches:
 * frames [n_frames] (host array); all frames share n_uniform, n_surface and phase.  Results equal n_frames calls of
 * dns_sample_rays. */
int dns_sample_rays_batch(const dns_sample_args* frames, int n_frames, void* stream);

/* Per-frame tables of the class-balanced draw (utils/common.py:312-322: torch.unique(label) + torch.nonzero per
 * class): a STABLE counting sort of the window pixels by label.  label [n_pixels] int64 with ids in [0, n_ids);
 * order [n_pixels]: pixel indices grouped by ascending label, ascending inside a label (== nonzero); counts / starts
 * [n_ids] int32: pixels of every id and the first position of its group in `order`; *err = 1 if a label is out of
 * range.  dns_sample_rays resolves a class-balanced draw as order[slot_base + offset] (see dns_sample_args). */
int64_t dns_class_tables_workspace_bytes(int64_t n_pixels, int n_ids);
int dns_class_tables(const int64_t* label, int64_t n_pixels, int n_ids, int64_t* order, int32_t* counts, int32_t* starts,
                     int32_t* err, void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Pixel-feature branch == utils/common.py:645-679 without the 209 MB/view upsample:
 * project P points into R views, round, mask, 4-tap bilinear fetch of the half-res feature
 * map at the rounded full-res pixel (align_corners=True).
 * ------------------------------------------------------------------------------------- */
int dns_feature_gather(const float* pts, int64_t P, const float* w2c /*[R,4,4]*/, int R,
                       const float* K /*[3,3]*/, int H, int W,
                       const float* feats /*[R,h,w,C] channels-last*/,
                       int C, int h, int w, float* code /*[R,P,C]*/, int64_t* uv /*[R,P,2] or NULL*/,
                       uint8_t* mask /*[R,P] or NULL*/, void* stream);

/* Merge MLP of the pixel-feature branch, fused (models/decoder.py:67-77):
 *   out[p] = mean_r  MLP_112->32->32( OneBlob((refer_p[r,p] - lo) / (hi - lo)) || code[r,p] )
 * refer_p [R,P,3], code [R,P,64] (dns_feature_gather output), params = the tinycudann vector of Merge.decoder
 * (W1[32][112] | W2[32][32]), out [P,32] overwritten.  With keep_for_backward the workspace keeps the activation
 * images and the SAME workspace must be passed to dns_merge_bwd, which writes d_refer_p [R,P,3] and ACCUMULATES
 * d_params (may be NULL).  The gathered features carry no gradient (utils/common.py:657 rounds the pixels). */
int64_t dns_merge_workspace_bytes(int64_t n_rows /* R * P */);
int dns_merge_fwd(const float* refer_p, const float* code, const float* params, int64_t P, int R,
                  const double bound[3][2], float* out, int keep_for_backward, void* workspace,
                  int64_t workspace_bytes, void* stream);
int dns_merge_bwd(const float* refer_p, const float* d_out, int64_t P, int R, const double bound[3][2],
                  float* d_refer_p, float* d_params, void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * The whole pixel-feature branch of one ray batch, fused == feature_matching + Merge + truncation mask as the
 * iteration bodies use them (slams/tracking.py:163-171, slams/mapping.py:549-557 over utils/common.py:645-679 and
 * models/decoder.py:67-77):
 *   features[n,s,:] = trunc(n,s) * mean_v MLP_112->32->32( OneBlob((pt - cam_o[v] - lo) / (hi - lo)) || feat_v(pt) )
 * with pt = rays_o[n] + rays_d[n] * z[n,s] rebuilt in the kernel, feat_v(pt) the 4-tap bilinear fetch of view v's
 * half-resolution map at the ROUNDED projection of pt (zero when it falls outside the image or behind the camera) and
 * trunc = 1 for samples within +-5 % of a positive gt depth.  Only those samples are evaluated (the reference
 * evaluates all and multiplies two thirds by zero).  Rays [ray_start[f], ray_start[f+1]) belong to target frame f,
 * whose n_views reference views are w2c / cam_o rows f*n_views .. and feats[f] ([n_views,h,w,64] channels-last).
 * apply_trunc = 0 evaluates every sample (Mapper.decoder_init, slams/mapping.py:807-809, has no mask).
 * dns_featmerge_bwd recomputes the activations (no stash): it ACCUMULATES d_params (Merge.decoder.params layout) and
 * ADDS the gradient that reaches the points through OneBlob to d_rays_o / d_rays_d (the gathered features carry no
 * gradient: common.py:657 rounds).  It must get the workspace of the matching forward call (band row list).
 * ------------------------------------------------------------------------------------- */
#define DNS_MAX_FRAMES 8
typedef struct dns_featmerge_args {
  int32_t n_rays, n_samples;
  int32_t n_frames, n_views;
  int32_t ray_start[DNS_MAX_FRAMES + 1];
  int32_t H, W, h, w;       /* image size, feature-map size */
  int32_t apply_trunc;
  int32_t need_dparams, need_drays;   /* backward only */
  /* forward with apply_trunc: 1 = rows of samples outside the band are left untouched instead of being zeroed (the consumer
   * is dns_render_fwd_bwd with features_band_only = 1, which never reads them): saves a pass over [N,S,32] */
  int32_t no_zero_fill;
  double bound[3][2];
  const float* K;           /* [3,3] */
  const float* w2c;         /* [n_frames*n_views,4,4] */
  const float* cam_o;       /* [n_frames*n_views,3] camera centres of the views (inverse(w2c)[:3,3], common.py:672-673) */
  const float* feats[DNS_MAX_FRAMES];
  const float* rays_o;      /* [N,3] */
  const float* rays_d;      /* [N,3] */
  const float* z_vals;      /* [N,S] */
  const float* gt_depth;    /* [N] */
  const float* params;      /* W1[32][112] | W2[32][32] */
  float* features;          /* [N,S,32] overwritten (forward) */
  const float* d_features;  /* [N,S,32] (backward) */
  float* d_params;          /* += (backward), may be NULL */
  float* d_rays_o;          /* [N,3] += (backward), may be NULL */
  float* d_rays_d;
  void* workspace;
  int64_t workspace_bytes;
  /* optional: room for the operand tiles of the forward pass (57 344 B per 128 / n_views band samples).  When the band of
   * the call fits (decided on the device), the backward pass bulk-copies them instead of gathering and encoding again;
   * NULL, or a band that does not fit: the backward recomputes.  Must be the same buffer in both calls. */
  void* stash;
  int64_t stash_bytes;
} dns_featmerge_args;
int64_t dns_featmerge_workspace_bytes(int n_rays, int n_samples);
int dns_featmerge_fwd(const dns_featmerge_args* a, void* stream);
int dns_featmerge_bwd(const dns_featmerge_args* a, void* stream);

/* ---------------------------------------------------------------------------------------
 * Camera poses of a mapping / tracking iteration on the device (utils/common.py:406-458 quad2rotation /
 * get_camera_from_tensor; slams/mapping.py:534-551 reference-view poses; the pose part of autograd's backward).
 * dns_pose_prepare: quats [F,4] (w,x,y,z, un-normalised), trans [F,3]  ->  R [F,3,3] (the sampler's input) and, for
 * every view v < n_views_total, w2c[v] / cam_o[v]: view_src[v] >= 0 follows target frame view_src[v] (rigid inverse of
 * its CURRENT pose, detached as in the reference), view_src[v] < 0 copies fixed_w2c[v] / fixed_cam_o[v].
 * dns_pose_grad: d_rays_o / d_rays_d [N,3] of rays whose camera-frame directions follow from pixel[n] (flat window
 * index as dns_sample_rays resolves it) -> d_quats [F,4], d_trans [F,3] (overwritten): the chain
 * rays_d = R(q) dirs, rays_o = T (common.py:262-263) differentiated in closed form.  scratch: 12*F floats.
 * ------------------------------------------------------------------------------------- */
int dns_pose_prepare(const float* quats, const float* trans, int n_frames, const int32_t* view_src, const float* fixed_w2c,
                     const float* fixed_cam_o, int n_views_total, float* R, float* w2c, float* cam_o, void* stream);
int dns_pose_grad(const float* d_rays_o, const float* d_rays_d, const int64_t* pixel, int n_frames,
                  const int32_t* ray_start /* host, [n_frames+1] */, int H0, int W0, int Ww, float fx, float fy, float cx,
                  float cy, const float* quats, float* d_quats, float* d_trans, float* scratch, void* stream);
/* Loss / result bookkeeping of one native mapping iteration (the scalar arithmetic of slams/mapping.py:896-907 around the
 * fused calls, on the device so that the loop holds no library kernel).  phase bit 0, before the gradient all-reduce:
 * loss_vec[0..7] <- losses8 (dns_render_fwd_bwd), and with tv_loss != NULL loss_vec[8] <- tv_w * *tv_loss,
 * loss_vec[6] += tv_lambda_w * *tv_loss (the smoothness term joins the total, mapping.py:895-897).  phase bit 1, after it:
 * loss_vec[7] *= nvalid_scale (n_valid is a batch constant, not a partial sum), result[0..8] <- loss_vec,
 * result[9 .. 9 + 2F) <- scratch [F][2] (per frame: max depth, rays outside the bound), result[9 + 2F] += sum_f scratch[f][1],
 * result[10 + 2F] <- min(result[10 + 2F], loss_vec[7]) (running error flag over all steps). */
int dns_map_step_result(int phase, const float* losses8, const float* tv_loss, float tv_w, float tv_lambda_w,
                        float nvalid_scale, const float* scratch, int n_frames, float* loss_vec9, float* result,
                        void* stream);
/* Best-pose bookkeeping of the tracking loop (slams/tracking.py:331-338: `if loss < current_min_loss` + the copy of the
 * candidate pose, a host comparison per iteration in the reference): losses = the [8] vector of dns_render_fwd_bwd
 * (total at 6, n_valid / error flag at 7); if total < *best_loss, best7 <- [quat | trans] and *best_loss <- total;
 * hist[*slot] <- total (when hist != NULL and *slot < hist_len), *slot += 1, *err_min <- min(*err_min, losses[7]). */
int dns_track_best(const float* losses, const float* quat, const float* trans, float* best7, float* best_loss, float* hist,
                   int32_t* slot, int hist_len, float* err_min, void* stream);


/* ---------------------------------------------------------------------------------------
 * ResNet stem of the pixel-feature branch == models/encoder.py:9-17 over models/layers.py:52-114 (what is left of
 * ResNet18 there): conv1 7x7 / stride 2 / pad 3, 3 -> 64, no bias -> bn1 -> ReLU, once per frame
 * (slams/tracking.py:295-296, slams/mapping.py:666,768,846).  images [n,H,W,3] (the gt_color frames as the
 * reference passes them), out [n,h,w,64] CHANNELS-LAST (the layout dns_feature_gather reads), h = (H-1)/2+1,
 * w = (W-1)/2+1.  training = 1 (the reference never calls .eval()): bn1 normalises with the statistics of this
 * batch of n views (biased variance) and, when the running_* pointers are given, updates them like
 * torch.nn.BatchNorm2d (momentum, unbiased variance); training = 0: normalises with running_mean / running_var.
 * conv_w [64,3,7,7]; bn_weight, bn_bias, running_mean, running_var [64].
 * ------------------------------------------------------------------------------------- */
int64_t dns_stem_workspace_bytes(void);
int dns_stem_fwd(const float* images, int n, int H, int W, const float* conv_w, const float* bn_weight,
                 const float* bn_bias, float eps, float momentum, int training, float* running_mean,
                 float* running_var, float* out, void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Adam == torch.optim.Adam defaults over a flat fp32 buffer
 * (slams/tracking.py:119-124,339; slams/mapping.py:464-466,910)
 * ------------------------------------------------------------------------------------- */
int dns_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                  float lr, float beta1, float beta2, float eps, int step, void* stream);

/* The same update over several segments in one launch, each with its own learning rate: the optimiser groups of
 * slams/tracking.py:119-124 (translation, quaternion) and slams/mapping.py:464-466 (decoder, quaternions,
 * translations).  segs_dev [n_segs] and step_dev live in DEVICE memory; the call first increments *step_dev and
 * uses the new value for the bias corrections, so a captured CUDA graph advances the step on every replay.
 * max_n = the longest segment (sizes the grid). */
typedef struct {
  float* p;        /* parameters, updated in place */
  const float* g;  /* gradients */
  float* m;        /* exp_avg */
  float* v;        /* exp_avg_sq */
  int64_t n;
  float lr;
  /* row_len > 0: the segment is n / row_len independent parameter tensors (the class experts of slams/mapping.py:445-446,
   * one tinycudann network each).  torch.optim.Adam skips a tensor whose gradient is None -- no moment decay, no step
   * count -- which is what happens to the expert of a class that no sample of the iteration belongs to: a row whose
   * gradient is all zero is skipped the same way, and every row keeps its own step count in row_steps[row] (int32,
   * device, zero-initialised by the caller) for the bias corrections. */
  int32_t row_len;
  int32_t* row_steps;
} dns_adam_seg;
int dns_adam_multi(const dns_adam_seg* segs_dev, int n_segs, int64_t max_n, int* step_dev, float beta1, float beta2,
                   float eps, void* stream);

/* Test entry of the tcgen05 GEMM: C[m][n] (row stride N) += sum_p A[p][m] * B[p][n]. */
int dns_debug_gemm_tc(const float* A, int lda, int M, const float* B, int ldb, int N, int64_t rows,
                      float* C, void* stream);

/* Same with the operand halves as fp16 (1: 11 + 11 mantissa bits, for bounded values that feed a ReLU decision) or bf16
 * (0: 8 + 8 bits, full fp32 range, for gradients).  a_f16 must equal b_f16: a mixed-format tcgen05.mma kind::f16 traps
 * on B200 (DNS_ERR_UNSUPPORTED).  M >= N. */
int dns_debug_gemm_fmt(const float* A, int lda, int M, const float* B, int ldb, int N, int64_t rows, int a_f16, int b_f16,
                       float* C, void* stream);

/* Test entry of the tile-image GEMM pipeline (cp.async.bulk -> tcgen05.mma): same contract as
 * dns_debug_gemm_tc, operands converted to bf16 hi/lo tile images of RS rows first. */
int dns_debug_gemm_img(const float* A, int lda, int M, const float* B, int ldb, int N, int64_t rows,
                       int RS, float* C, void* stream);

/* Phase accounting for benchmarks: per-phase kernel-launch counters (always on) and, when
 * enabled, CUDA-event timing of each phase on the launching stream.  Phases: 0 prep, 1 class
 * prep, 2 point_fwd, 3 ray, 4 point_bwd, 5 dw_gemm, 6 finalize, 7 adam, 8 tv_fwd, 9 tv_bwd,
 * 10 sample, 11 feature, 12 operator kernels.  dns_profile_read returns the number of phases. */
void dns_profile_enable(int on);
int dns_profile_read(double* ms /*[16]*/, long long* launches /*[16]*/, int reset);

/* sizeof(dns_grid), sizeof(dns_render_args), sizeof(dns_tv_args), sizeof(dns_sample_args),
 * sizeof(dns_featmerge_args): lets a foreign-language binding assert that its struct mirror matches this header. */
void dns_struct_sizes(int64_t out[5]);

#ifdef __cplusplus
}
#endif
#endif /* DNS_SLAM_B200_H */
