/*
 * itermvs_b200 -- C ABI of the B200 (sm_100a) implementation of the IterMVS hot path.
 *
 * The reference (FangjinhuaWang/IterMVS @ 453e9c7) is pure Python/PyTorch and has no FFI; its
 * boundary is the Python call surface of models/module.py, models/itermvs.py and models/net.py.
 * Each entry point below replaces one stock ATen/cuDNN call chain of that surface; the comment
 * on every function names the reference lines it stands in for.  The Python binding a
 * maintainer would add is shown in INTEGRATION.md (ctypes, exactly what itermvs_b200/_lib.py does).
 *
 * Conventions
 *   - plain C: device pointers to dense fp32 buffers, int shapes, a cudaStream_t passed as void*.
 *   - every function returns 0 on success, non-zero on error; imvs_last_error() gives the text
 *     (thread-local).  Nothing synchronises, allocates or frees device memory: the caller owns all
 *     buffers (outputs and scratch included), so every call is CUDA-graph capturable.
 *   - the library uses the caller's CURRENT device; re-entrant across host threads.
 *
 * Layouts  ("P_l" = H_l*W_l, level 1 = 1/2 res C=16, level 2 = 1/4 res C=32, level 3 = 1/8 res C=48)
 *   activations / features  [N][H][W][C]      channels-last ("NHWC"); pyramids are [B][V][H_l][W_l][C_l],
 *                                             view 0 = reference view
 *   projection matrices     [B][V][4][4]      row-major, as produced by the reference loaders
 *   composed projections    [B][S][12]        rot (3x3 row-major) then trans (3)
 *   correlation volumes     [B][slice][P][8]  8 = group-wise correlation channels, innermost
 *   GRU input x             [B][H2][W2][16]   ch 0 = normalized depth, 1..10 = correlation, 11..15 = 0
 *   conv weights            [tap][CinP][CoutP] CinP/CoutP = channels padded to a multiple of 8 with zeros; each
 *                                             weight is given twice: TF32-rounded (1-pass mode) and plain
 *                                             fp32 (3-pass fp32-grade mode) (packed by itermvs_b200/_pack.py)
 */
#ifndef ITERMVS_B200_H_
#define ITERMVS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IMVS_ABI_VERSION 5
#define IMVS_GROUPS 8          /* reference models/itermvs.py:28 */
#define IMVS_OUT_BINS 256      /* reference models/itermvs.py:134 */
#define IMVS_RADIUS 4          /* reference models/itermvs.py:135 */
#define IMVS_HIDDEN 32         /* reference models/net.py:72 */
#define IMVS_ITER_SLICES 10    /* 4 + 4 + 2 refinement samples, models/itermvs.py:231-235 */
#define IMVS_XCH 16            /* stored channels of the GRU input x (11 used) */
#define IMVS_MAX_VIEWS 16      /* source views per reference view supported by the fused kernels */

/* Library plumbing -- no reference counterpart (the reference reports errors as Python exceptions, e.g. the
 * assert at module.py:83,87): ABI version checked by the binding at load time, text of the last failed call on
 * this thread, number of kernels launched by this library so far (bench.py's gpu_launches). */
int imvs_abi_version(void);
const char* imvs_last_error(void);
long long imvs_launches_total(void);

/* Tensor-core convolution precision ("mode"): 1 = single-pass TF32 (what the reference's own cuDNN path
 * uses on Ampere+ GPUs, torch.backends.cudnn.allow_tf32 = True by default), 3 = error-compensated
 * 3-product TF32 split (fp32-grade, relative 2^-21), 4 = the same 3-product compensation on the FP16
 * tensor-core path (x = fp16 hi + fp16 lo: 22 bits for |x| < 65504, absolute floor 2^-25; twice the
 * TF32 MMA rate).  Process-wide; default 4. */
int imvs_set_conv_passes(int passes);
int imvs_get_conv_passes(void);
/* tcgen05/TMEM kernels for the 32..64-channel stride-1 convolutions in the 1-pass TF32 mode (default on).
 * imvs_tcgen05_status() synchronises the device and returns 0, or 1 if a tcgen05 kernel timed out on its
 * mbarrier (never expected; guards against a hung GPU). */
int imvs_set_tcgen05(int enabled);
int imvs_tcgen05_status(void);
/* Device status word of the library's kernels (synchronises the device; clear != 0 resets it): bit 0 = a tcgen05 kernel timed
 * out on its mbarrier (never expected); bit 1 = FP16 RANGE: in the fp32-grade mode 4 a convolution produced a value beyond
 * +-65504 -- the fp16 hi/lo split of the next layer saturates there (cvt.rn.satfinite) and the result would be a finite wrong
 * number; the flag makes that loud.  Checked on every convolution's accumulators (mma.sync and tcgen05 kernels);
 * 4 = CUDA error while reading the word. */
int imvs_device_status(int clear);

typedef struct imvs_wpair {      /* packed conv weight [tap][CinP][CoutP] */
    const float* tf32;           /* values rounded to TF32 (round-to-nearest): operand of the 1-pass mode */
    const float* fp32;           /* plain fp32: the 3-pass mode splits hi/lo in registers */
    const float* umma;           /* TF32-rounded, tcgen05 K-major canonical order [cout block][tap][CinP/4][NB][4]
                                    with NB = min(CoutP, 64); may be NULL (then the mma.sync kernels run) */
    const void* f16x3;           /* fp16 hi/lo split for mode 4: [tap][CinK/2][CoutP] of 8-byte entries
                                    {half2 hi(k, k+1), half2 lo(k, k+1)}, CinK = CinP rounded up to 16 (zeros) */
    const void* f16umma;         /* fp16 hi/lo split in the tcgen05 K-major canonical order (mode 4 on the 5th-generation
                                    tensor core, csrc/tc5conv.cuh:tc5h_conv_kernel): [tap][hi | lo][CinK/8][CoutP][8 halves];
                                    NULL when the shape is not served (CoutP % 16 != 0 or CoutP > 64) */
    const void* f16ummai;        /* the same with hi and lo interleaved per K chunk: [tap][CinK/8][hi | lo][CoutP][8 halves]
                                    (csrc/tc5pconv.cuh: A_hi x [B_hi | B_lo] is one N = 2 CoutP instruction); NULL alike */
} imvs_wpair;

typedef struct imvs_corrnet_weights {   /* models/itermvs.py:352-381, one CorrNet */
    imvs_wpair conv0;      /* conv0.conv.weight  8 -> 8            [9][8][8]   */
    imvs_wpair conv1;      /* conv1.conv.weight  8 -> 16, stride 2 [9][8][16]  */
    imvs_wpair conv2;      /* conv2.conv.weight 16 -> 32, stride 2 [9][16][32] */
    imvs_wpair conv3;      /* conv3.weight      32 -> 16, transposed stride 2 [9][32][16] */
    imvs_wpair conv4;      /* conv4.weight      16 -> 8,  transposed stride 2 [9][16][8]  */
    imvs_wpair conv5;      /* conv5.weight       8 -> 1   [9][8][8] (cout padded) */
    const float* conv5_b;  /* conv5.bias [1] */
} imvs_corrnet_weights;

typedef struct imvs_weights {
    /* evaluation.pixel_view_weight (models/itermvs.py:333-350) */
    imvs_wpair pvw_conv0;          /* conv.0.conv.weight 8 -> 16  [9][8][16] */
    const float* pvw_conv1;        /* conv.1.weight [16] (fp32) */
    const float* pvw_conv1_b;      /* conv.1.bias [1] */
    /* evaluation.corr_conv1[0..2] (level 1, 2, 3) */
    imvs_corrnet_weights corrnet[3];
    /* update.gru (models/module.py:52-66): input channels = [h(32), x(16 stored, 11 used)] -> 48 */
    imvs_wpair gru_zr;             /* convz|convr stacked on Cout: [9][48][64] */
    const float* gru_zr_b;         /* [64] */
    imvs_wpair gru_q;              /* convq [9][48][32] */
    const float* gru_q_b;          /* [32] */
    /* update.depth_head.0 | update.confidence_head.0 stacked on Cout (itermvs.py:139-151) */
    imvs_wpair head_conv0;         /* [9][32][64] : cout 0..31 depth head, 32..63 confidence head */
    imvs_wpair head_fc1;           /* depth_head.2.weight  [1][32][64]  */
    imvs_wpair head_fc2;           /* depth_head.4.weight  [1][64][256] */
    const float* head_fc2_b;       /* depth_head.4.bias    [256] */
    const float* conf_fc;          /* confidence_head.2.weight [32] */
    const float* conf_fc_b;        /* confidence_head.2.bias [1] */
    /* update.hidden_init_head (itermvs.py:153-157) */
    imvs_wpair hinit_conv0;        /* [9][D][64] */
    imvs_wpair hinit_fc;           /* [1][64][32] */
    const float* hinit_fc_b;       /* [32] */
    /* iter_mvs.upsample (itermvs.py:246-250) */
    imvs_wpair ups_conv0;          /* [9][32][64] */
    const float* ups_fc;           /* [64][144] (fp32) */
    /* depth_head.2 / depth_head.4 (+ bias) for the fused tcgen05 head (csrc/headfused.cuh): fp16 hi / lo split in the UMMA
     * K-major canonical order [K/8][N][8 halves]: W1 hi | W1 lo (K = 32, N = 64) | W2 hi | W2 lo (K = 64, N = 256), then
     * the fp32 bias [256] = 74 752 bytes, 16-byte aligned.  May be NULL: then the unfused kernels run. */
    const void* head_fused;
} imvs_weights;

/* ------------------------------------------------------------------------------------------
 * Single operators (drop-in for the reference's module-level functions / modules)
 * ---------------------------------------------------------------------------------------- */

/* module.py:78-90 -- proj = src_proj @ inverse(ref_proj), rot/trans split, for every source view.
 * proj: [B][V][4][4] (view 0 = reference).  out: [B][V-1][12].  nan_flag (device int, may be NULL)
 * is set to 1 if a composed matrix contains NaN (the reference asserts, module.py:83,87). */
int imvs_compose_projections(const float* proj, int B, int V, float* out, int* nan_flag, void* stream);

/* module.py:68-125 -- differentiable_warping, the reference's own layouts:
 * src_fea [B][C][H1][W1], src_proj/ref_proj [B][4][4], depth_samples [B][D][H][W] -> out [B][C][D][H][W].
 * rt_scratch: B*12 floats (receives the composed rot|trans). */
int imvs_differentiable_warping(const float* src_fea, const float* src_proj, const float* ref_proj,
                                const float* depth_samples, float* out, int B, int C, int H1, int W1,
                                int D, int H, int W, float* rt_scratch, int* nan_flag, void* stream);

/* Backward of differentiable_warping with respect to src_fea -- what autograd derives from F.grid_sample at
 * module.py:118 (the sampling grid is built under torch.no_grad(), module.py:77, so nothing flows to the
 * projections or the depth samples):  grad_src_fea[b,c,y0+dy,x0+dx] += w_tap * grad_out[b,c,d,y,x] over the four
 * in-range taps.  grad_out [B][C][D][H][W] -> grad_src_fea [B][C][H1][W1] (zeroed by the call, accumulated with
 * fp32 atomics: the summation order is not deterministic, as in ATen's CUDA grid_sampler backward).
 * rt_scratch: B*12 floats. */
int imvs_differentiable_warping_backward(const float* grad_out, const float* src_proj, const float* ref_proj,
                                         const float* depth_samples, float* grad_src_fea, int B, int C, int H1, int W1,
                                         int D, int H, int W, float* rt_scratch, int* nan_flag, void* stream);

/* layout helpers: [N][C][H][W] <-> [N][H][W][C].  The reference keeps every tensor NCHW (e.g. the feature lists
 * returned by FeatureNet.forward, net.py:56-66); the single-operator mirrors convert at their boundary. */
int imvs_nchw_to_nhwc(const float* in, float* out, int N, int C, int H, int W, void* stream);
int imvs_nhwc_to_nchw(const float* in, float* out, int N, int C, int H, int W, void* stream);

/* itermvs.py:11-19 + 45-51 -- fused plane sweep at level 3: generates the D inverse-depth
 * hypotheses, warps + bilinearly samples every source view and writes the group-wise correlation
 * per view.  fea3 [B][V][H3][W3][48], rt3 [B][S][12], out corr [B][S][D][P3][8].
 * depth_samples: NULL (hypotheses generated in-kernel from depth_min/max [B], the estimator's path)
 * or explicit [B][D][P3] hypotheses (Evaluation.forward's depth_sample argument, itermvs.py:33). */
int imvs_warpcorr_init(const float* fea3, const float* rt3, const float* depth_min, const float* depth_max,
                       const float* depth_samples, float* corr, int B, int V, int H3, int W3, int D, void* stream);

/* itermvs.py:341-350 + 53-57 -- PixelViewWeight on every (view, hypothesis) slice, softmax over D,
 * max over D, and the x2 bilinear upsampling.  corr [B][S][D][P3][8];
 * logits scratch [B][S][D][P3]; vw3 [B][S][P3]; vw2 [B][S][4*P3]. */
int imvs_pixel_view_weight(const imvs_weights* w, const float* corr, float* logits, float* vw3, float* vw2,
                           int B, int S, int D, int H3, int W3, void* stream);

/* itermvs.py:59-69 -- view-weighted aggregation of the init volume: agg [B][D][P3][8]. */
int imvs_aggregate_init(const float* corr, const float* vw3, float* agg, int B, int S, int D, int P3, void* stream);

/* itermvs.py:289-293 + 86-120 -- fused iteration kernel: hypotheses from the normalized depth,
 * warp + sample of the three pyramids, group-wise correlation, pixel-wise view-weighted
 * aggregation.  nd: normalized depth of pixel p of batch b at nd[b*nd_batch_stride + p*nd_pixel_stride];
 * vw2 [B][S][P2]; agg [B][10][P2][8] (slices 0-3 level 1, 4-7 level 2, 8-9 level 3).
 * samples1/2/3: all NULL (hypotheses from nd) or all given as explicit depths [B][4][P2], [B][4][P2],
 * [B][2][P2] (Evaluation.forward's dict argument). */
int imvs_warpcorr_iter(const float* fea1, const float* fea2, const float* fea3,
                       const float* rt1, const float* rt2, const float* rt3,
                       const float* nd, size_t nd_batch_stride, size_t nd_pixel_stride, const float* vw2,
                       const float* depth_min, const float* depth_max,
                       const float* samples1, const float* samples2, const float* samples3, float* agg,
                       int B, int V, int H2, int W2, void* stream);

/* The same two kernels on a level-3 pyramid PADDED to 64 floats per texel (written by imvs_pad_level3 from the
 * [B][V][H3][W3][48] pyramid: correlation group g's six channels at floats 4g..4g+3 and 32+4g, 32+4g+1 of the texel, the other
 * 16 floats zero): a tap is then one aligned 128-byte line + one 64-byte piece per lane group instead of three 64-byte pieces,
 * and a lane owns its group (no regrouping).  Same arguments, results equal up to fp32 summation order.  Used by
 * imvs_itermvs_forward with IMVS_TUNE_WC_PAD3=1 (the padded copy lives in the workspace).
 * Reference: the same lines as imvs_warpcorr_init / imvs_warpcorr_iter (itermvs.py:11-19, 45-51, 86-120). */
int imvs_pad_level3(const float* fea3, float* fea3p, int B, int V, int H3, int W3, void* stream);
int imvs_warpcorr_init_padded(const float* fea3p, const float* rt3, const float* depth_min, const float* depth_max,
                              const float* depth_samples, float* corr, int B, int V, int H3, int W3, int D, void* stream);
int imvs_warpcorr_iter_padded(const float* fea1, const float* fea2, const float* fea3p,
                              const float* rt1, const float* rt2, const float* rt3,
                              const float* nd, size_t nd_batch_stride, size_t nd_pixel_stride, const float* vw2,
                              const float* depth_min, const float* depth_max,
                              const float* samples1, const float* samples2, const float* samples3, float* agg,
                              int B, int V, int H2, int W2, void* stream);

/* Backward of imvs_warpcorr_init with respect to the feature pyramid (the sampling grid carries no
 * gradient, module.py:77; source views receive grid_sample's input gradient, the reference view the gradient
 * through the product of itermvs.py:50).  Same inputs as the forward (the hypotheses are recomputed, nothing is
 * saved between the passes); grad_corr [B][S][D][P3][8]; grad_fea3 [B][V][H3][W3][48] is overwritten.
 * No counterpart call in the reference: there torch.autograd walks the saved [C,D,H,W] volumes. */
int imvs_warpcorr_init_backward(const float* fea3, const float* rt3, const float* depth_min, const float* depth_max,
                                const float* depth_samples, const float* grad_corr, float* grad_fea3,
                                int B, int V, int H3, int W3, int D, void* stream);

/* Backward of imvs_warpcorr_iter with respect to the three feature pyramids (reference view incl. its resampling of
 * itermvs.py:95-98, and source views).  The view weights and the hypotheses are constants here, as in the reference
 * (itermvs.py:295 view_weights.detach(); 282-283 normalized_depth.detach()).  grad_agg [B][10][P2][8];
 * grad_fea1/2/3 have the shapes of fea1/2/3 and are overwritten. */
int imvs_warpcorr_iter_backward(const float* fea1, const float* fea2, const float* fea3,
                                const float* rt1, const float* rt2, const float* rt3,
                                const float* nd, size_t nd_batch_stride, size_t nd_pixel_stride, const float* vw2,
                                const float* depth_min, const float* depth_max,
                                const float* samples1, const float* samples2, const float* samples3,
                                const float* grad_agg, float* grad_fea1, float* grad_fea2, float* grad_fea3,
                                int B, int V, int H2, int W2, void* stream);

/* itermvs.py:367-381 -- CorrNet on N slices of a [N][P][8] volume.  Slice n uses weight set
 * sets[(n % period) < split1 ? 0 : (n % period) < split2 ? 1 : 2].  The scalar output of slice n, pixel p
 * goes to out[(n / period) * out_batch_stride + p * out_pixel_stride + (n % period)].
 * scratch: imvs_corrnet_scratch_floats(N,H,W) floats. */
size_t imvs_corrnet_scratch_floats(int N, int H, int W);
int imvs_corrnet(const imvs_corrnet_weights* sets, int period, int split1, int split2, const float* vol,
                 float* out, size_t out_batch_stride, size_t out_pixel_stride, float* scratch,
                 int N, int H, int W, void* stream);

/* itermvs.py:159-164 -- hidden_init: corr [B][H3][W3][D] -> hidden [B][2*H3][2*W3][32].
 * scratch: B*(64+32)*H3*W3 floats.  D must be a multiple of 8. */
int imvs_hidden_init(const imvs_weights* w, const float* corr, float* hidden, float* scratch,
                     int B, int D, int H3, int W3, void* stream);

/* module.py:59-66 -- ConvGRU.forward(h, x), in place on h.
 * h [B][H][W][32], x [B][H][W][16] (channels 11..15 must be zero); scratch 4*B*32*H*W floats (z, and the [h|x] and
 * [r*h|x] operands as fp16 hi / lo split planes for the TMA + tcgen05 path). */
int imvs_conv_gru(const imvs_weights* w, float* h, const float* x, float* scratch, int B, int H, int W, void* stream);

/* itermvs.py:171-190 / 196-219 -- depth_head (+ confidence_head when conf or conf_logit != NULL),
 * softmax over 256 bins, arg-max, clamped +-4 window regression.  hidden [B][H][W][32].
 *   nd_out        normalized depth of pixel p, batch b -> nd_out[b*nd_batch_stride + p*nd_pixel_stride]
 *   probability   [B][256][H][W] (the reference's layout) or NULL (training needs it, itermvs.py:282,302)
 *   conf / conf_logit [B][H][W] or NULL (sigmoid / raw)
 *   depth_out     [B][H][W] or NULL: depth_unnormalization of nd_out (module.py:148-152)
 * scratch: B*384*H*W floats (conv output 64 + hidden 64 + logits 256 per pixel). */
int imvs_depth_head(const imvs_weights* w, const float* hidden, float* nd_out, size_t nd_batch_stride,
                    size_t nd_pixel_stride, float* probability, float* conf, float* conf_logit, float* depth_out,
                    const float* depth_min, const float* depth_max, float* scratch,
                    int B, int H, int W, void* stream);

/* itermvs.py:262-264 + module.py:127-140 + itermvs.py:321-324 -- output stage: upsampling-weight
 * net on the level-2 reference feature (NHWC [H2][W2][32], image b at ref_fea2 + b*ref_batch_stride),
 * softmax over the 9 taps, convex x4 upsampling of nd, depth_unnormalization; bilinear x4 of the confidence.
 * depth_up / conf_up [B][4*H2][4*W2]; conf may be NULL (then conf_up is not written).
 * scratch: B*64*H2*W2 floats. */
int imvs_upsample_outputs(const imvs_weights* w, const float* ref_fea2, size_t ref_batch_stride, const float* nd,
                          size_t nd_batch_stride, size_t nd_pixel_stride, const float* conf, const float* depth_min,
                          const float* depth_max, float* depth_up, float* conf_up, float* scratch,
                          int B, int H2, int W2, void* stream);

/* ------------------------------------------------------------------------------------------
 * Depth-map filtering for fusion (eval.py:154-309) -- the caller after the hot path
 * ---------------------------------------------------------------------------------------- */

/* cams68 (HOST pointer, 68 floats per (reference, source) pair, row-major float32 exactly as numpy forms them in
 * eval.py:162-190):  inv(K_ref)[9] | E_src @ inv(E_ref) [16] | K_src [9] | inv(K_src) [9] | E_ref @ inv(E_src) [16] |
 * K_ref [9].  Depth maps are [H][W] float32 on the device.
 *
 * eval.py:199-215 check_geometric_consistency for one pair: mask [H][W] bytes (1 = consistent), depth_reprojected
 * (0 where inconsistent), x2d_src / y2d_src; any output may be NULL.  acc_sum / acc_count (may be NULL) additionally
 * accumulate depth_reprojected and mask (eval.py:259-260); successive calls on one stream add in call order. */
int imvs_check_geometric_consistency(const float* depth_ref, const float* depth_src, const float* cams68_host,
                                     float geo_pixel_thres, float geo_depth_thres, unsigned char* mask,
                                     float* depth_reprojected, float* x2d_src, float* y2d_src,
                                     float* acc_sum, int* acc_count, int H, int W, void* stream);

/* eval.py:243-265 for one reference view: S sources (depth_srcs [S][H][W], cams68_host [S][68]) checked in order,
 * depth_averaged = (sum of masked reprojections + depth_ref) / (count + 1) as float64, photo_mask = confidence >
 * photo_thres, geo_mask = count >= geo_mask_thres, final_mask = both.  acc_sum [H][W] float and acc_count [H][W]
 * int are caller-owned scratch (zeroed by the call); photo_mask / geo_mask may be NULL. */
int imvs_filter_depth_view(const float* depth_ref, const float* confidence, const float* depth_srcs,
                           const float* cams68_host, int S, float geo_pixel_thres, float geo_depth_thres,
                           float photo_thres, int geo_mask_thres, float* acc_sum, int* acc_count,
                           double* depth_averaged, unsigned char* photo_mask, unsigned char* geo_mask,
                           unsigned char* final_mask, int H, int W, void* stream);

/* ------------------------------------------------------------------------------------------
 * Whole estimator: IterMVS.forward in test mode (itermvs.py:253-329), one call = all launches.
 * ---------------------------------------------------------------------------------------- */
typedef struct imvs_problem {
    int B;            /* reference views in this call */
    int V;            /* views per reference view (1 + source views) */
    int H, W;         /* full-resolution image size, multiples of 32 */
    int D;            /* initial hypotheses (32 in the reference, itermvs.py:237); multiple of 8 */
    int iterations;   /* GRU iterations (>= 1) */
} imvs_problem;

/* a7, models/itermvs.py:74-81 -- the init branch's own depth estimate (training loss term depths["initial"]): softmax over the
 * D hypotheses of the init CorrNet output, expectation, depth_unnormalization, bilinear x2.
 * corr element (b, d, p) at b*batch_stride + d*slice_stride + p*pixel_stride; scratch: B*H3*W3 floats;
 * depth_out: [B][2*H3][2*W3]. */
int imvs_init_depth(const float* corr, size_t batch_stride, size_t slice_stride, size_t pixel_stride,
                    const float* depth_min, const float* depth_max, float* scratch, float* depth_out,
                    int B, int D, int H3, int W3, void* stream);

size_t imvs_forward_workspace_bytes(const imvs_problem* pb);

/* fea1/2/3: channels-last pyramids [B][V][H_l][W_l][C_l]; proj1/2/3: [B][V][4][4] fp32.
 * Outputs (any may be NULL): depth [B][H2][W2] (pre-last-update value, itermvs.py:319),
 * depth_up [B][H][W], conf [B][H2][W2], conf_up [B][H][W].
 * nan_flag: device int, set to 1 on NaN projections (checked by the caller when it next syncs). */
int imvs_itermvs_forward(const imvs_problem* pb, const imvs_weights* w,
                         const float* fea1, const float* fea2, const float* fea3,
                         const float* proj1, const float* proj2, const float* proj3,
                         const float* depth_min, const float* depth_max,
                         void* workspace, size_t workspace_bytes,
                         float* depth, float* depth_up, float* conf, float* conf_up,
                         int* nan_flag, void* stream);

/* number of kernel launches one imvs_itermvs_forward issues for this problem (for bench.py's gpu_launches) */
int imvs_forward_launch_count(const imvs_problem* pb);

/* ------------------------------------------------------------------------------------------
 * FeatureNet (net.py:7-66), eval mode, BatchNorm folded into the convolutions by the packer.
 * imgs [N][3][H][W] (N = B*V views) -> channels-last pyramids fea1 [N][H/2][W/2][16],
 * fea2 [N][H/4][W/4][32], fea3 [N][H/8][W/8][48].
 * ---------------------------------------------------------------------------------------- */
#define IMVS_FNET_CONVS 24
typedef struct imvs_featurenet_weights {
    imvs_wpair w[IMVS_FNET_CONVS];      /* order documented in itermvs_b200/_pack.py:FNET_LAYERS; slots 21..23 =
                                           [layerK.0.conv1 | layerK.0.downsample] stacked on Cout, K = 1..3 */
    const float* b[IMVS_FNET_CONVS];    /* per-layer bias (folded BN shift or conv bias) */
} imvs_featurenet_weights;
size_t imvs_featurenet_workspace_bytes(int N, int H, int W);
int imvs_featurenet_forward(const imvs_featurenet_weights* w, const float* imgs, float* fea1, float* fea2, float* fea3,
                            void* workspace, size_t workspace_bytes, int N, int H, int W, void* stream);
/* Same with the raw 8-bit images [N][3][H][W]: the first layer normalises them as the reference's loaders do
 * (2 * np.array(img, dtype=np.float32) / 255. - 1, datasets/dtu_yao_eval.py:63-64) -- f-4: a quarter of the host-to-device bytes. */
int imvs_featurenet_forward_u8(const imvs_featurenet_weights* w, const unsigned char* imgs, float* fea1, float* fea2, float* fea3,
                               void* workspace, size_t workspace_bytes, int N, int H, int W, void* stream);
int imvs_featurenet_launch_count(void);

/* f-4, datasets/dtu_yao_eval.py:61-76 (`read_img`) after the image decode, on the device: raw 8-bit image [H0][W0][3] ->
 * level0 [3][H][W] = cv2.resize(2 * img / 255. - 1, (W, H), INTER_LINEAR) and, where non-NULL, level k = cv2.resize(level0,
 * (W >> k, H >> k), INTER_LINEAR), k = 1..3 (float32, planar: the loader's transpose([0, 3, 1, 2]) of one view).
 * Agrees with OpenCV to 1 ulp of the interpolated value (its SIMD paths may fuse one product of the interpolation). */
int imvs_image_pyramid_u8(const unsigned char* img, int H0, int W0, float* level0, float* level1, float* level2, float* level3,
                          int H, int W, void* stream);

/* Stride-1 3x3 convolution (dilation dil, zero padding dil) + bias (+ residual) (+ ReLU) on fp32 NHWC tensors through the
 * persistent TMA + tcgen05 kernel (csrc/tc5pconv.cuh) in the fp32-grade mode: operands as fp16 hi / lo "split planes"
 * [N][C/8][H][W][8 halves] loaded by cp.async.bulk.tensor, three products per tap into a TMEM accumulator.  The operator
 * behind nn.Conv2d(Cin, Cout, 3, padding=dil, dilation=dil) of models/module.py:6-50 (ConvBnReLU / ResidualBlock with the
 * BatchNorm folded) and models/net.py:18-20 (output convolutions) inside imvs_featurenet_forward, exposed for tests.
 * w_f16ummai: imvs_wpair::f16ummai of the layer; (Cin, Cout) in {(16,16), (32,32), (48,48), (48,32), (48,16)} with dil = 1,
 * (32,32) also with dil = 2;
 * x: [N][H][W][Cin], residual (may be NULL) and out: [N][H][W][Cout]; via_split_output != 0 stores the result as split
 * planes first (the layout the next layer's TMA loads read) and converts back. */
/* Persistent (TMA + tcgen05) launches use sm_count / share CTAs from now on (share >= 1; default 1 = the whole GPU).  A serving
 * loop with several reference views in flight on different streams sets 2 while it captures / enqueues them, so that launches
 * of different streams run side by side instead of one after the other (graph.StreamingPipeline does).  Process-wide. */
int imvs_set_sm_share(int share);

size_t imvs_conv3x3_tcgen05_workspace_bytes(int N, int H, int W, int Cin, int Cout);
int imvs_conv3x3_tcgen05(const float* x, const void* w_f16ummai, const float* bias, const float* residual, float* out,
                         void* workspace, size_t workspace_bytes, int N, int H, int W, int Cin, int Cout, int dil, int relu,
                         int via_split_output, void* stream);

/* profiling taps (bench.py): CUDA-event pair around every stage of the two forward functions */
int imvs_profile_begin(int capacity);
int imvs_profile_end(float* ms_out, int* tags_out, int capacity);

#ifdef __cplusplus
}
#endif
#endif /* ITERMVS_B200_H_ */
