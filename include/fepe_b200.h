/* fepe_b200 -- C ABI of the B200-native weighted 8-point / relative-pose path of deepFEPE.
 *
 * The reference (eric-yyjau/pytorch-deepFEPE @ 7f3e775) has no FFI / operator registry; its
 * boundary is the Python nn.Module surface (deepFEPE/utils/loader.py:117-129 resolves
 * "GoodCorresNet_layers_deepF" to deepFEPE/models/DeepFNet.py::DeepFNet).  This header is the
 * layer BELOW that surface: plain pointers and sizes, no torch types, caller-owned device buffers,
 * no host allocation, asynchronous on the given CUDA stream.  Each entry point names the reference
 * code it replaces.  INTEGRATION.md shows the ctypes binding the Python modules use.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless stated otherwise; float = fp32 (the reference's dtype,
 *     deepFEPE/datasets/kitti_odo_corr.py:444), double = fp64 scratch the kernels produce themselves;
 *   - B = number of image pairs, N = correspondences per pair (reference: 1000 / 2000,
 *     deepFEPE/configs/kitti_corr_baseline.yaml:12-13); every pair is independent;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - return value: 0 on success, a positive cudaError_t from the launch, or a negative FEPE_E_* code.
 *     Nothing throws across the ABI.
 */
#ifndef FEPE_B200_H
#define FEPE_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FEPE_E_BADARG   (-1)  /* null pointer / non-positive size                              */
#define FEPE_E_TOOLARGE (-2)  /* N does not fit the shared-memory staging of one pair           */
#define FEPE_E_NODEVICE (-3)  /* no CUDA device / not an sm_100 device                          */

/* doubles of per-pair state fepe_fit_fwd saves for fepe_fit_bwd:
 *   [0..5]   raw means and Hartley scales (m1x,m1y,s1,m2x,m2y,s2)
 *   [6..14]  f, unit eigenvector (= vec of the normalised F before the rank-2 projection)
 *   [15]     lambda_min          [16..51] the 36 distinct Gram entries (fepe_math.cuh order)
 *   [52]     factorisations used [53..55] v3, right singular vector of reshape(f) removed by rank 2
 *   [56..60] SM cycles this pair spent waiting for its copy / in Hartley / Gram / solve / residual
 *   [61],[62] of the solve cycles: Gram reduce / eigen iteration   [63] sigma_3 of reshape(f) */
#define FEPE_SAVED_DOUBLES 64

/* library / build identification: returns e.g. "fepe_b200 0.1 sm_100a" (host pointer, static). */
const char* fepe_version(void);

/* Test / tuning hook (host call, process-wide, thread-safe; no environment variable is ever read): force a kernel
 * variant the library otherwise picks from the problem size, so that every path can be exercised at small sizes.
 *   FEPE_DISPATCH_FIT        0 by size | 1 one-CTA-per-pair latency kernel | 2 persistent pair ring | 3 split pipeline
 *   FEPE_DISPATCH_GRAM_TEAM  0 by size | 1, 2, 3 = teams of 1, 2, 4 warps per pair in the split pipeline's Gram kernel
 *   FEPE_DISPATCH_MLP_GEMM   0 by size | 1 one tile per CTA | 2 persistent kernel with 128-column tiles (bf16 path)
 *   FEPE_DISPATCH_MLP_FUSE   0 default | 2 fused-norm variant with 8 transform warps (bf16 path)
 *   FEPE_DISPATCH_WGRAD      0 by shape | 1 weight-gradient GEMM tiles of at most 128 x 128 | 2 128 x 256 where Ci % 256 == 0
 *   FEPE_DISPATCH_NN_DIST    0 by shape | 1 descriptor distances on CUDA cores (plain fp32) | 2 on tcgen05 (split fp16)
 * A forced variant that cannot run the problem falls back to the automatic choice.  Returns the previous value, or
 * FEPE_E_BADARG. */
#define FEPE_DISPATCH_FIT       0
#define FEPE_DISPATCH_GRAM_TEAM 1
#define FEPE_DISPATCH_MLP_GEMM  2
#define FEPE_DISPATCH_MLP_FUSE  3
#define FEPE_DISPATCH_WGRAD     4
#define FEPE_DISPATCH_NN_DIST   5
#define FEPE_DISPATCH_COUNT     6
int fepe_set_dispatch(int which, int value);

/* Diagnostics of the split pipeline (throughput path of fepe_fit_fwd): while `buf` (device memory, 32 bytes per record)
 * is set, every CTA of its three kernels appends {kernel 1 Gram | 2 solve | 3 residual, SM id, start ns, end ns} as four
 * u64 (globaltimer); pass NULL to stop.  fepe_debug_trace_count returns the number of records requested so far (may
 * exceed the capacity; the surplus was dropped).  Host calls, synchronous with the device; not for production use. */
int fepe_debug_trace(void* buf, int capacity_records);
int fepe_debug_trace_count(void);

/* Largest N one launch can stage (depends on the device's opt-in shared memory). Host call. */
int fepe_max_correspondences(void);

/* ---- Fit.forward + compute_epi_residual, fused ------------------------------------------------
 * Replaces, per pair:  Fit.normalize x2 (deepFEPE/models/DeepFNet.py:148-179, called :198-199),
 * the N x 9 constraint build + row normalisation + weighting (:203-214), torch.svd(X[b]) -> V[:,-1]
 * (:233-235), torch.svd(F) rank-2 projection (:236-237), residual = X @ f (:251),
 * out = T2^T F_ T1 (:256) and, when `epi` != NULL, utils_F.compute_epi_residual(pts1, pts2, out,
 * clamp_at) (deepFEPE/dsac_tools/utils_F.py:400-413; called DeepFNet.py:479).
 *
 *   matches  [B,N,4]  (x1,y1,x2,y2) -- the reference's data_batch['matches_xy_ori'] (pixels) or any
 *                     other coordinates; the affine x' = ax*x+bx, y' = ay*y+by is applied on load to
 *                     both images.  With ax=2/W,bx=-1,ay=2/H,by=-1 this is NormalizeAndExpand_HW
 *                     (DeepFNet.py:108-114); ax=ay=1,bx=by=0 gives plain Fit.forward semantics.
 *   weights  [B,N]    per-correspondence weights (the reference's [B,1,N] squeezed).
 *   F_out    [B,9]    rank-2 F, row major, in the primed (post-affine) coordinates, like the reference.
 *   resid    [B,N]    signed algebraic residual w_i * p_hat_i . f   (sign follows f; f is returned with
 *                     its largest-magnitude entry positive -- LAPACK's sign in the reference is arbitrary).
 *   epi      [B,N]    clamped symmetric epipolar distance, or NULL to skip.
 *   saved    [B,FEPE_SAVED_DOUBLES] state for the backward pass, or NULL (inference).
 */
int fepe_fit_fwd(const float* matches, const float* weights, int B, int N,
                 float ax, float bx, float ay, float by, float clamp_at,
                 float* F_out, float* resid, float* epi, double* saved, void* stream);

/* ---- backward of the above w.r.t. the weights ---------------------------------------------------
 * Replaces autograd's SvdBackward x2 + the elementwise chain (Train_model_pipeline.py:595 walking
 * DeepFNet.py:198-256 and utils_F.py:400-413).  Inputs are the forward inputs, `saved`, and the
 * upstream gradients gF [B,9], gresid [B,N], gepi [B,N] (either of the last two may be NULL = zero).
 *   gweights [B,N]   d loss / d weights.
 */
int fepe_fit_bwd(const float* matches, const float* weights, int B, int N,
                 float ax, float bx, float ay, float by, float clamp_at,
                 const double* saved, const float* gF, const float* gresid, const float* gepi,
                 float* gweights, void* stream);

/* Same, plus the gradient w.r.t. the COORDINATES: gmatches [B,N,4] (16-byte aligned) receives
 * d loss / d (x1,y1,x2,y2) of every correspondence, through compute_epi_residual's direct dependence
 * (utils_F.py:400-413), the constraint rows and their L2 normalisation (DeepFNet.py:203-212), both Hartley
 * transforms including their mean and mean-distance terms (Fit.normalize, DeepFNet.py:148-179; T enters
 * out = T2^T F_ T1 as well, :256) and the affine (ax,bx,ay,by).  The reference needs it when learned offsets
 * are added to the matches (if_learn_offsets, DeepFNet.py:373,489-505) or the keypoint front-end is trained
 * (Train_model_pipeline.py:384).  gmatches == NULL is fepe_fit_bwd.  */
int fepe_fit_bwd_coords(const float* matches, const float* weights, int B, int N,
                        float ax, float bx, float ay, float by, float clamp_at,
                        const double* saved, const float* gF, const float* gresid, const float* gepi,
                        float* gweights, float* gmatches, void* stream);

/* ---- pose / loss head -----------------------------------------------------------------------------
 * Replaces, per layer and pair: E = K^T T2^T F T1 K (deepFEPE/train_good_utils.py:356-358), the host
 * loop of get_Rt_loss (train_good_utils.py:96-188: E^T -> _get_M2s (dsac_tools/utils_F.py:478-498),
 * _R_to_q (dsac_tools/utils_geo.py:58-86), L2 errors to the GT pose, min-select, rot12_to_angle_error /
 * vector_angle metrics) and the per-pair F-loss over the virtual correspondences
 * (train_good_utils.py:325-354; clamp_at = 0.02 in the shipped configs).
 *
 *   F        [L,B,9]   F of each layer in primed coordinates (fepe_fit_fwd output)
 *   K        [B,9]     intrinsics;  (ax,bx,ay,by) the NormalizeAndExpand_HW transform (= outs['T1'] = ['T2'])
 *   q_gt     [B,4]     qs_cam (w,x,y,z);  t_gt [B,3] ts_cam (normalised inside, F.normalize semantics)
 *   Rt_scene [B,16]    delta_Rtijs_4_4 (scene motion) for the rotation-angle metric, or NULL
 *   virt1/2  [B,V,3]   homogeneous pixel virtual correspondences, or both NULL (no F-loss)
 *   out      [L,B,FEPE_POSE_OUT_FLOATS]:
 *            [0..8] E   [9..17] selected R   [18..20] selected t   [21] q L2 error   [22] t L2 error
 *            [23] rotation angle error (deg)  [24] translation angle error (deg)
 *            [25] F-loss (mean clamped epipolar residual of the V virtual points)
 *            [26],[27] which of the two candidates won for q / t   [28..30] singular values of E
 */
#define FEPE_POSE_OUT_FLOATS 32
int fepe_pose_fwd(const float* F, const float* K, int L, int B, float ax, float bx, float ay, float by,
                  const float* q_gt, const float* t_gt, const float* Rt_scene,
                  const float* virt1, const float* virt2, int V, float clamp_at,
                  float* out, void* stream);

/* fepe_fit_fwd followed by fepe_pose_fwd (L = 1) on the fitted F of every pair, as ONE call: the evaluation path
 * "F and R,t per image pair" (train_good_utils.py:298 get_all_loss_DeepF + :64 get_Rt_loss on the last layer).  For
 * batches that take the one-CTA-per-pair latency kernel the pose head runs inside that kernel (one launch, F never
 * leaves the SM in between); larger batches run the two kernels back to back.  Arguments as in the two calls;
 * `clamp_at` clamps the epipolar residual rows, `virt_clamp_at` the F-loss of the virtual correspondences. */
int fepe_fit_pose_fwd(const float* matches, const float* weights, int B, int N, float ax, float bx, float ay,
                      float by, float clamp_at, float* F_out, float* resid, float* epi, double* saved,
                      const float* K, const float* q_gt, const float* t_gt, const float* Rt_scene,
                      const float* virt1, const float* virt2, int V, float virt_clamp_at, float* pose_out,
                      void* stream);

/* Backward of fepe_pose_fwd: dL/dF [L,B,9] from the upstream gradients g_q, g_t, g_loss [L,B] (any may be NULL)
 * of out[..,21], out[..,22], out[..,25]; `pose_out` is the forward output (it records which candidates won).
 * Replaces autograd through torch.svd / _get_M2s / _R_to_q / _l2_error / compute_epi_residual on the host
 * (train_good_utils.py:96-188, :325-354). */
int fepe_pose_bwd(const float* F, const float* K, int L, int B, float ax, float bx, float ay, float by,
                  const float* q_gt, const float* t_gt, const float* virt1, const float* virt2, int V, float clamp_at,
                  const float* pose_out, const float* g_q, const float* g_t, const float* g_loss, float* dF,
                  void* stream);

/* ---- per-correspondence weight MLP on tensor cores (inference path) -----------------------------
 * Replaces ErrorEstimator.forward (deepFEPE/models/ErrorEstimators.py:46-68: five Conv1d(k=1) ->
 * InstanceNorm1d(affine) -> LeakyReLU(0.01) blocks and a final Conv1d) and the softmax over the N
 * correspondences (deepFEPE/models/DeepFNet.py:443,512).  Activations are bf16 row-major [B*Npad, C]
 * with Npad = N rounded up to 128 (padded rows are zero and excluded from the statistics); weights are
 * bf16 [Co, Ci] (Conv1d weight squeezed); accumulation is fp32 in tensor memory (tcgen05.mma).
 *   fepe_mlp_first  layer 1 from fp32 features X0 [B,N,Ci<=8]                -> Y [B*Npad,Co] + stats
 *   fepe_mlp_gemm   Y = X W^T + b (K % 64 == 0, Co % 64 == 0; b may be NULL, else 16-byte aligned) -> Y + stats [B,Co,2] (stats may be NULL)
 *   fepe_mlp_norm   X' = LeakyReLU(gamma (Y - mean) rstd + beta) from stats (biased variance, eps)
 *   fepe_mlp_last   logits = X w + b (Co = 1), weights = softmax over the N rows of each pair
 * `stats` must be zeroed by the caller before fepe_mlp_first / fepe_mlp_gemm (they accumulate with atomics).
 */
int fepe_mlp_first(const float* X0, const float* W, const float* bias, void* Y, float* stats, int B, int N, int Npad,
                   int Ci, int Co, void* stream);
int fepe_mlp_gemm(const void* X, const void* W, const float* bias, void* Y, float* stats, int B, int Npad,
                  int Nvalid, int K, int Co, void* stream);
int fepe_mlp_norm(const void* Y, const float* stats, const float* gamma, const float* beta, void* X, int B, int Npad,
                  int Nvalid, int Co, float eps, float slope, void* stream);
/* Inference with the InstanceNorm + LeakyReLU of layer l fused into the GEMM of layer l+1 (the normalised activations
 * X' are never written to memory): fepe_mlp_scale_shift turns the statistics of layer l into per-(pair, channel)
 * (a, d) with x' = LeakyReLU(a y + d) -- ss [B, Co/2, 4] fp32 = (a_c, a_c+1, d_c, d_c+1), 16-byte aligned; with
 * clear_stats != 0 it also zeroes `stats` for the next accumulation -- and fepe_mlp_gemm_norm computes
 * Y = LeakyReLU(a Yprev + d) W^T + b from the PRE-norm output Yprev [B*Npad, K] of layer l (Co % 128 == 0,
 * 0 < slope < 1).  Same arithmetic as fepe_mlp_norm followed by fepe_mlp_gemm (identical bf16 results). */
int fepe_mlp_scale_shift(float* stats, const float* gamma, const float* beta, float* ss, int B, int Co, int Nvalid,
                         float eps, int clear_stats, void* stream);
int fepe_mlp_gemm_norm(const void* Yprev, const float* ss, float slope, const void* W, const float* bias, void* Y,
                       float* stats, int B, int Npad, int Nvalid, int K, int Co, void* stream);
/* fepe_mlp_last on the PRE-norm output Y [B*Npad, Ci] of the last block (its InstanceNorm + LeakyReLU fused in). */
int fepe_mlp_last_norm(const void* Y, const float* ss, float slope, const float* W, float bias, float* logits,
                       float* weights, int B, int N, int Npad, int Ci, void* stream);
int fepe_mlp_last(const void* X, const float* W, float bias, float* logits, float* weights, int B, int N, int Npad,
                  int Ci, void* stream);
/* training path: weight gradient of one Conv1d(k=1): dW[Co,Ci] += dY[M,Co]^T X[M,Ci] on tcgen05 (bf16 operands read
 * in place as MN-major tiles, fp32 accumulation, split over row slabs with atomics; dW zeroed by the caller). */
int fepe_mlp_wgrad(const void* dY, const void* X, float* dW, int M, int Co, int Ci, void* stream);
/* backward of InstanceNorm(affine) + LeakyReLU: from dX' (gradient of the block output), the saved block output
 * X' (sign), the saved pre-norm Y and the forward statistics: A[B,Co,2] = (sum dZ, sum dZ Yhat) (zeroed by the
 * caller; dgamma = sum_b A[..,1], dbeta = sum_b A[..,0]) and dY [B*Npad,Co] bf16 (padded rows zero).  The data
 * gradient dX = dY W then is fepe_mlp_gemm with W^T as the weight operand and stats = NULL. */
int fepe_mlp_normbwd(const void* dX, const void* Xp, const void* Y, const float* stats, const float* gamma, float* A,
                     void* dY, int B, int Npad, int Nvalid, int Co, float eps, float slope, void* stream);
int fepe_mlp_last_bwd(const float* dlogits, const void* X, const float* W, void* dX, float* dW, float* db, int B, int N,
                      int Npad, int Ci, void* stream);
int fepe_mlp_first_bwd(const void* dY, const float* X0, const float* W, float* dX0, float* dW, int B, int N, int Npad,
                       int Ci, int Co, void* stream);

/* ---- the same MLP at the reference's fp32 accuracy, still on tensor cores (the DEFAULT path of the modules) ----------
 * Replaces ErrorEstimator.forward (deepFEPE/models/ErrorEstimators.py:46-68, fp32 in the reference) and the softmax
 * (DeepFNet.py:443,512).  tcgen05 has no fp32 operands: every operand is split into two fp16 numbers (hi + lo, 22
 * significant bits) and a product is three kind::f16 MMAs accumulated in fp32 (x w ~= hi hi + lo hi + hi lo), which
 * keeps the logits within ~1e-5 of an fp32 evaluation.  Activations are fp32 row-major [B*Npad, C] in memory
 * (Npad = N rounded up to 128; padded rows are zero and excluded from the statistics); every GEMM applies the previous
 * block's InstanceNorm + LeakyReLU to its operand on the fly, so normalised activations are never stored.
 *   fepe_mlp32_prepare_weights  W [Co,K] fp32 -> Whi, Wlo [Co,K] fp16 of W * s and wscale [4] = (s, 1/s, scratch, -),
 *                               s = the power of two that puts max |W| into [2^13, 2^14)
 *   fepe_mlp32_first            layer 1 straight from the model's inputs (DeepFNet.get_input :359-404 and the torch.cat
 *                               of :487 are folded in): channels = [((ax x1+bx)+1)/2, ((ay y1+by)+1)/2, same for x2,y2]
 *                               when matches [B,N,4] != NULL, then extra0..3, each [B,N,c_i] fp32 or NULL; at most 16
 *                               channels in total; W [64,Ci] fp32, bias [64] or NULL -> Y [B*Npad,64] + stats
 *   fepe_mlp32_scale_shift      stats [B,Co,2] fp64 (sum, sum of squares; zeroed by the caller before the producing
 *                               launch) -> ss [B,Co,2] fp32 = (a, d), x' = LeakyReLU(a y + d); biased variance, eps
 *   fepe_mlp32_gemm             Y = LeakyReLU(a Yprev + d) W^T (+ bias) with K % 64 == 0, Co % 64 == 0 -> Y [B*Npad,Co] +
 *                               stats (or NULL).  ss == NULL: Yprev is used as is (data-gradient GEMM dX = dY W with
 *                               Whi/Wlo of W^T, Nvalid = Npad)
 *   fepe_mlp32_last             logits [B,Co,N] = LeakyReLU(a Y + d) W[Co,256]^T + bias for Co = 1 (then weights [B,N] =
 *                               softmax over N, or NULL) or Co = 4 (the offsets network, DeepFNet.py:341-342)
 */
int fepe_mlp32_prepare_weights(const float* W, void* Whi, void* Wlo, float* wscale, int Co, int K, void* stream);
/* X0_out [B,N,Ci] or NULL: the assembled input features, kept for fepe_mlp32_first_bwd */
int fepe_mlp32_first(const float* matches, float ax, float bx, float ay, float by, const float* extra0, int c0,
                     const float* extra1, int c1, const float* extra2, int c2, const float* extra3, int c3,
                     const float* W, const float* bias, float* Y, double* stats, float* X0_out, int B, int N, int Npad,
                     int Co, void* stream);
/* mean_rstd [B,Co,2] fp32 or NULL: (mean, 1/sqrt(var + eps)) per (pair, channel), kept for fepe_mlp32_normbwd */
int fepe_mlp32_scale_shift(double* stats, const float* gamma, const float* beta, float* ss, float* mean_rstd, int B, int Co,
                           int Nvalid, float eps, int clear_stats, void* stream);
/* a_amax (ss == NULL only; device pointer or NULL): bit pattern of max |Yprev| -- the operand is multiplied by the power
 * of two that brings it to 2^13..2^14 before the fp16 split (gradients can lie far below fp16's range) and the result is
 * divided by it again */
int fepe_mlp32_gemm(const float* Yprev, const float* ss, float slope, const unsigned* a_amax, const void* Whi,
                    const void* Wlo, const float* wscale, const float* bias, float* Y, double* stats, int B, int Npad,
                    int Nvalid, int K, int Co, void* stream);
int fepe_mlp32_last(const float* Y, const float* ss, float slope, const float* W, const float* bias, float* logits,
                    float* weights, int B, int N, int Npad, int Ci, int Co, void* stream);
/* ---- backward of the same (training: what autograd does at Train_model_pipeline.py:595 for ErrorEstimators.py:46-64) --
 * Everything fp32 in memory; the block outputs x' = LeakyReLU(a y + d) are recomputed from the saved pre-norm y.
 *   fepe_mlp32_last_bwd   dlogits [B,Co,N] -> dX [B*Npad,Ci] (gradient of the last block's output), dW [Co,Ci] and db [Co]
 *                         (accumulated: zero them first)
 *   fepe_mlp32_normbwd    dX -> dY [B*Npad,C] through LeakyReLU + InstanceNorm(affine); A [B,C,2] fp64 = (sum dZ, sum dZ
 *                         yhat) (zeroed by the caller; dbeta = sum_b A[..,0], dgamma = sum_b A[..,1]); dy_amax = bit
 *                         pattern of max |dY| (zeroed by the caller).  C / 4 a power of two <= 256 or a multiple of 256.
 *   fepe_mlp32_wgrad      dW [Co,Ci] += dY^T LeakyReLU(a Yprev + d) on tcgen05 (Co % 128 == 0, Ci % 64 == 0; dW zeroed by
 *                         the caller); the data gradient dX = dY W is fepe_mlp32_gemm(dY, NULL, 1, dy_amax, (W^T)hi/lo, ...)
 *   fepe_mlp32_first_bwd  layer 1: dW [64,Ci] += dY^T X0 and, when dX0 != NULL, dX0 [B,N,Ci] = dY W
 *   fepe_mlp32_affine_grads  dgamma_s[c] += sum_b A_s[b,c,1], dbeta_s[c] += sum_b A_s[b,c,0] for nseg <= 8 blocks whose A
 *                         arrays [B,C_s,2] lie back to back at A (C, dgamma, dbeta: HOST arrays of nseg entries; the
 *                         device pointers in them are accumulated into) -- one launch for a whole estimator
 * All d* outputs are ACCUMULATED, so a caller may hand in its gradient buffers directly (fepe_b200.dist.FlatGradients).
 */
int fepe_mlp32_last_bwd(const float* dlogits, const float* Y, const float* ss, float slope, const float* W, float* dX,
                        float* dW, float* db, int B, int N, int Npad, int Ci, int Co, void* stream);
int fepe_mlp32_normbwd(const float* dX, const float* Y, const float* ss, const float* mean_rstd, const float* gamma,
                       float slope, double* A, float* dY, unsigned* dy_amax, int B, int Npad, int Nvalid, int C,
                       void* stream);
int fepe_mlp32_wgrad(const float* dY, const unsigned* dy_amax, const float* Yprev, const float* ss_prev, float slope,
                     float* dW, int M, int Npad, int Co, int Ci, void* stream);
int fepe_mlp32_affine_grads(const double* A, int B, int nseg, const int* C, float* const* dgamma, float* const* dbeta,
                            void* stream);
int fepe_mlp32_first_bwd(const float* dY, const float* X0, const float* W, float* dX0, float* dW, int B, int N, int Npad,
                         int Ci, int Co, void* stream);

/* ---- validation pose recovery ----------------------------------------------------
 * Replaces, per (layer, pair), the host work of deepFEPE/dsac_tools/utils_F.py:909-954 goodCorr_eval_nondecompose
 * (called per sample from train_good_utils.py:553-646 val_rt through a process pool):
 *   cv2.recoverPose(E_hat, p1s, p2s, focal=K[0,0], pp=(K[0,2],K[1,2]))  -- decomposeEssentialMat, DLT triangulation of
 *   every correspondence for the four (R,t) candidates, cheirality + 50-unit distance test, most-points-in-front wins --
 *   then utils_geo.invert_Rt / rot12_to_angle_error / vector_angle against the ground-truth motion.
 *   E        [L,B,9]  essential matrices (row-major), any scale
 *   K        [B,9]    intrinsics; like the reference call, focal = K[0,0] serves both axes
 *   matches  [B,N,4]  (x1,y1,x2,y2) pixels, 16-byte aligned
 *   n_valid  [B] or NULL: only the first n_valid[b] correspondences of pair b take part
 *   Rt_scene [B,16] or NULL: delta_Rtijs_4_4 (scene motion); errors are 180 / 90 without it or with < 5 points
 *   out      [L,B,FEPE_RECOVER_OUT_FLOATS]: [0..8] R  [9..11] t  [12] points in front of the winner (cv2's return value)
 *            [13] winning candidate 0..3 = (R1,t),(R2,t),(R1,-t),(R2,-t)  [14..17] the four counts
 *            [18] err_q  [19] err_t (degrees)  [20] correspondences used
 *   mask     [L,B,N] or NULL: 255 where the winner's point passed (cv2's mask), else 0
 */
#define FEPE_RECOVER_OUT_FLOATS 24
int fepe_recover_pose(const float* E, const float* K, const float* matches, const int* n_valid,
                      int L, int B, int N, float distance_thresh, const float* Rt_scene,
                      float* out, unsigned char* mask, void* stream);

/* ---- mutual nearest-neighbour descriptor matching ---------------------------------
 * Replaces, per sample, SP_tracker.nn_match_two_way(desc1.T, desc2.T, nn_thresh) at
 * deepFEPE/train_good_utils.py:683-691 (numpy on the host after a D2H copy of both descriptor sets; PointTracker of the
 * un-vendored `superpoint` package): distances sqrt(2 - 2 clip(<d1_i, d2_j>, -1, 1)), row-wise nearest neighbour below
 * nn_thresh whose column-wise nearest neighbour points back; matches ordered by the first index.
 *   desc1    [B,N1,D], desc2 [B,N2,D]  fp32, L2-normalised rows, 16-byte aligned, D a multiple of 16
 *   n1, n2   [B] valid keypoints per image, or NULL (= N1 / N2)
 *   workspace  fepe_nn_match_workspace_bytes(B,N1,N2) bytes of device memory, 8-byte aligned (caller-owned)
 *   idx1, idx2 [B,N1] int32, score [B,N1] fp32: the first count[b] entries of row b are the matches
 *              (matching_mask[0], [1], [2] of the reference); count [B] int32
 */
size_t fepe_nn_match_workspace_bytes(int B, int N1, int N2);
int fepe_nn_match(const float* desc1, const float* desc2, const int* n1, const int* n2,
                  int B, int N1, int N2, int D, float nn_thresh, void* workspace,
                  int* idx1, int* idx2, float* score, int* count, void* stream);

/* ---- ground truth + virtual correspondences of a batch ----------------------------
 * Replaces, per sample, the host work of the reference's dataset (deepFEPE/datasets/kitti_odo_corr.py:290-302, :526-566):
 *   E, F = utils_F.E_F_from_Rt_np(R, t, K)  (dsac_tools/utils_F.py:835-846);
 *   utils_misc.get_virt_x1x2_np(image_size, F, K, pts1_virt_b, pts2_virt_b)  (dsac_tools/utils_misc.py:173-199) =
 *   cv2.correctMatches(F, pts2_virt_b, pts1_virt_b) with OpenCV's NaNs replaced by 0, homogeneous, and K^-1 applied;
 *   q_cam, t_cam of the inverse motion and q_scene, t_scene  (dsac_tools/utils_geo.py:88-117 R_to_q_np).
 *   K        [B,9]    intrinsics
 *   Rt_scene [B,16]   scene motion x2 = R x1 + t (relative_scene_poses[1] / delta_Rtijs_4_4), or NULL with F_in
 *   F_in     [B,9]    or NULL: fundamental matrix to correct the grid onto instead of the one built from Rt_scene
 *   grid1, grid2 [P,2] pixel grids pts1_virt_b, pts2_virt_b (utils_misc.py:163-171: 10x10 points, P = 100)
 *   gt       [B,FEPE_GT_FLOATS] or NULL: [0..8] E  [9..17] F  [18..21] q_cam (w,x,y,z; w >= 0)  [22..24] t_cam
 *            [25..28] q_scene  [29..31] t_scene      (needs Rt_scene)
 *   pts1_virt, pts2_virt [B,P,3] homogeneous pixels with pts2^T F pts1 = 0
 *   pts_virt_normalized  [B,P,3] or NULL: K^-1 pts1_virt -- the reference returns this for BOTH normalised keys
 */
#define FEPE_GT_FLOATS 32
int fepe_gt_virt(const float* K, const float* Rt_scene, const float* F_in, const float* grid1, const float* grid2,
                 int B, int P, float* gt, float* pts1_virt, float* pts2_virt, float* pts_virt_normalized, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FEPE_B200_H */
