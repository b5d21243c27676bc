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
 *   [52]     factorisations used [53..55] singular values of reshape(f)
 *   [56..60] SM cycles this pair spent waiting for its copy / in Hartley / Gram / solve / residual
 *   [61..63] reserved */
#define FEPE_SAVED_DOUBLES 64

/* library / build identification: returns e.g. "fepe_b200 0.1 sm_100a" (host pointer, static). */
const char* fepe_version(void);

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

#ifdef __cplusplus
}
#endif
#endif /* FEPE_B200_H */
