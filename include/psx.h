/*
 * psx.h -- C ABI of libpsx, the sm_100a posterior-sampling step library.
 *
 * This is the drop-in boundary of samplers_b200.  The reference
 * (thomashirtz/samplers) is pure Python and has no FFI of its own; each entry
 * point below replaces the eager-torch arithmetic of the cited reference lines
 * and is what a maintainer of the reference would bind with ctypes (see
 * INTEGRATION.md for the stub).
 *
 * Conventions
 *   - plain C, no torch / C++ types in any signature;
 *   - every pointer named d_* is DEVICE memory owned by the caller (allocated by
 *     torch in the host package); h_* is HOST memory read during the call only;
 *   - all tensors are fp32, contiguous, "flat" layout (L, C, H, W) = (L, n);
 *   - every launch is asynchronous on `stream` (a cudaStream_t passed as void*;
 *     NULL = legacy default stream); no entry point synchronises the device;
 *   - return value: PSX_OK or an error code; psx_last_error() gives the text
 *     (thread-local).  Nothing throws across the boundary.
 *   - there is NO CPU implementation behind these calls: without a CUDA device
 *     they fail with PSX_ERR_CUDA.
 */
#ifndef PSX_H_
#define PSX_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define PSX_API __attribute__((visibility("default")))
#else
#define PSX_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define PSX_OK 0
#define PSX_ERR_INVALID 1     /* bad argument (shape, null pointer, range)   */
#define PSX_ERR_CUDA 2        /* CUDA runtime / launch failure               */
#define PSX_ERR_UNSUPPORTED 3 /* valid request this build has no kernel for  */

#define PSX_MAX_TAPS 127  /* longest 1-D tap vector of a separable blur       */
#define PSX_ABI_VERSION 5

/* Operator kinds (psx_op_kind). */
#define PSX_OP_IDENTITY 0
#define PSX_OP_MASK 1
#define PSX_OP_BOX 2
#define PSX_OP_SEPBLUR 3
#define PSX_OP_CONV2D 4

typedef struct psx_op psx_op; /* opaque host-side operator descriptor */

PSX_API int psx_abi_version(void);
PSX_API const char* psx_last_error(void);
/* The kernel-selection switches (environment: PSX_SPLIT, PSX_FUSED, PSX_NO_TC, PSX_NO_PIPE, PSX_NO_FAST16; all
 * default to the measured-best path) are read ONCE, at the first launch.  A process that changes them afterwards
 * (the tests and measurement tools do) calls this to have them read again.  No reference counterpart: the
 * reference has one code path per operator. */
PSX_API void psx_reload_env(void);
/* Number of CUDA kernels this library has launched in this process so far (a launch recorded into a CUDA graph
 * counts once, when it is captured).  Measurement aid: bench.py's "gpu_launches" is a difference of two readings. */
PSX_API long long psx_kernel_launches(void);

/* ---------------------------------------------------------------- operators
 * Degradation operators A (forward) / A^T (adjoint).  Replaces
 *   samplers/operators/identity.py:36-66      (identity)
 *   samplers/operators/inpainting.py:98-131,141-187 + linear.py:141-165 (mask)
 * and adds the operators BASELINE.json names that the reference lacks
 * (box super-resolution, separable Gaussian blur, 2-D motion blur); their
 * definition is oracle/operators.py.
 */
PSX_API int psx_op_create_identity(int64_t n, psx_op** out);
/* d_keep: n bytes, 1 = pixel observed, 0 = missing (the reference's mask is the
 * negation: True = missing, inpainting.py:27).  Dense form: y has n entries,
 * zeros at missing pixels. */
PSX_API int psx_op_create_mask(int64_t n, const uint8_t* d_keep, psx_op** out);
/* mean over non-overlapping factor x factor blocks; y is (C, H/f, W/f). */
PSX_API int psx_op_create_box(int C, int H, int W, int factor, psx_op** out);
/* BASELINE config 3 read as ONE composed operator, y = keep * box_f(x) (random-mask inpainting of the 4x
 * box-downsampled image; no counterpart in the reference, defined by oracle/operators.py: OracleMaskedBox):
 * d_keep has C * (H/f) * (W/f) bytes, 1 = coarse pixel observed.  Dense form: y is (C, H/f, W/f) with zeros at
 * the dropped pixels.  Not served by the bf16-state entry points. */
PSX_API int psx_op_create_box_masked(int C, int H, int W, int factor, const uint8_t* d_keep, psx_op** out);
/* y = V(H(x)): zero-padded "same" cross-correlation of every row with
 * h_taps_h (length kh, odd) then of every column with h_taps_v (length kv, odd). */
PSX_API int psx_op_create_sepblur(int C, int H, int W, const float* h_taps_h, int kh,
                          const float* h_taps_v, int kv, psx_op** out);
/* depthwise 2-D zero-padded "same" cross-correlation with a kh x kw PSF (odd
 * sizes, row-major host array); zero taps are skipped. */
PSX_API int psx_op_create_conv2d(int C, int H, int W, const float* h_kernel, int kh, int kw,
                         psx_op** out);
PSX_API int psx_op_destroy(psx_op* op);

PSX_API int psx_op_kind(const psx_op* op);
PSX_API int64_t psx_op_x_numel(const psx_op* op); /* n   = elements of one x */
PSX_API int64_t psx_op_y_numel(const psx_op* op); /* n_y = elements of one y */
/* number of per-sample partial sums psx_dps_pre writes for this operator */
PSX_API int psx_op_err_parts(const psx_op* op);
/* bytes of device scratch psx_dps_pre / psx_op_apply / psx_op_adjoint need for L samples */
PSX_API size_t psx_op_workspace_bytes(const psx_op* op, int64_t L);

/* y = A x  (Operator.apply, operators/base.py:51-61)  */
PSX_API int psx_op_apply(const psx_op* op, const float* d_x, float* d_y, int64_t L,
                 void* d_workspace, size_t workspace_bytes, void* stream);
/* x = A^T y  (Operator.apply_transpose, operators/base.py:63-75) */
PSX_API int psx_op_adjoint(const psx_op* op, const float* d_y, float* d_x, int64_t L,
                   void* d_workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ data formats either side of the loop
 * psx_observe -- simulate an observation: d_y = A d_x + eps,  eps = noise_scale * d_noise + noise_shift
 * (each op rounded as torch does).  Replaces InverseProblem.from_clean_data (samplers/inverse_problem.py:35-67):
 *   GaussianNoise.sample (noise.py:81-92):  d_noise ~ N(0,1),            noise_scale = sigma, noise_shift = 0
 *   PoissonNoise.sample  (noise.py:125-138): d_noise = k ~ Poisson(rate), noise_scale = 1,     noise_shift = -rate
 * The random draw itself stays with the caller's generator.  d_noise (L, n_y) may be NULL (noise-free). */
PSX_API int psx_observe(const psx_op* op, const float* d_x, const float* d_noise, float noise_scale,
                        float noise_shift, float* d_y, int64_t L, void* d_workspace, size_t workspace_bytes,
                        void* stream);
/* the noise half alone, in place: d_y += noise_scale * d_noise + noise_shift (observations whose layout is not the
 * operator descriptor's, e.g. the gathered (L, m) form of inpainting) */
PSX_API int psx_add_noise(float* d_y, const float* d_noise, int64_t numel, float noise_scale, float noise_shift,
                          void* stream);
/* Image tensors <-> bytes on the device, so that only uint8 crosses PCIe.  Replaces samplers/utils/image.py:
 *   psx_image_to_u8  : tensor_to_pil (:9-32)  clamp to [-1,1], (x+1)*0.5, then torchvision's mul(255).byte();
 *                      (images, C, H, W) float -> (images, H, W, C) uint8
 *   psx_image_from_u8: pil_to_tensor (:35-64) u8/255 (to_tensor), *2 - 1;  (images, H, W, C) -> (images, C, H, W) */
PSX_API int psx_image_to_u8(const float* d_chw, uint8_t* d_hwc, int64_t images, int C, int H, int W, void* stream);
PSX_API int psx_image_from_u8(const uint8_t* d_hwc, float* d_chw, int64_t images, int C, int H, int W,
                              void* stream);

/* Gather / scatter form of the inpainting operator (flatten=True in the reference:
 * inpainting.py:141-145 gather of the kept pixels, :178-187 scatter into zeros).
 *   psx_gather : d_out[l, j] = d_in[l, d_idx[j]]           (L, n) -> (L, m)
 *   psx_scatter: d_out = 0; d_out[l, d_idx[j]] = d_in[l, j] (L, m) -> (L, n)
 * d_idx: m int64 indices in [0, n), strictly increasing (kept pixels). */
PSX_API int psx_gather(const float* d_in, const int64_t* d_idx, float* d_out, int64_t L, int64_t n,
                       int64_t m, void* stream);
PSX_API int psx_scatter(const float* d_in, const int64_t* d_idx, float* d_out, int64_t L, int64_t n,
                        int64_t m, void* stream);

/* ------------------------------------------------------------------ K1
 * psx_dps_pre -- everything between the network's eps output and the network
 * VJP, in one pass.  Replaces
 *   samplers/networks/base.py:41-43        Tweedie  x0 = (x_t - s1*eps)/sa
 *   samplers/inverse_problem.py:17-21      r = y - A(x0)
 *   samplers/noise.py:77-79 / 121-123      Gaussian / Poisson log-likelihood
 *   the autograd mirror of those lines     (samplers/samplers/dps.py:102-103)
 *   samplers/samplers/dps.py:117-120       per-sample |r|_2  (as partial sums)
 *
 *   d_cot[l]      = lik_weight * A^T r_l / sqrt_acp          (L, n)
 *   d_err_part[l] = psx_op_err_parts(op) partial sums of |r_l|^2
 *
 * d_cot is d(sum log-lik)/d(x0)/sqrt_acp: the direct term of the gradient and,
 * times -sqrt_1m_acp, the cotangent of eps.  The caller runs
 * v = VJP_eps(d_cot) in torch and hands both to psx_dps_post.
 *
 * Observation indexing: sample l uses observation l / obs_repeat (so
 * obs_repeat = L broadcasts one observation, obs_repeat = 1 gives one per
 * sample) -- BatchView.repeat_observation, batch_view.py:128-137.
 * lik_weight = 1/sigma^2 (Gaussian) or 2/(rate + 1e-3) (Poisson).
 * d_x0_out may be NULL.
 */
PSX_API int psx_dps_pre(const psx_op* op, const float* d_x_t, const float* d_eps, const float* d_y,
                int64_t L, int64_t obs_repeat, float sqrt_acp, float sqrt_1m_acp,
                float lik_weight, float* d_cot, float* d_err_part, float* d_x0_out,
                void* d_workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ K2
 * psx_dps_post -- bridge (DDIM/DDPM ancestral) update fused with the guidance
 * step and the injected noise.  Replaces
 *   samplers/samplers/utils/bridge_kernels.py:41,44-45,59   mean + std * z
 *   samplers/samplers/dps.py:116-122                        + gamma/(|r|+1e-9) * grad
 * with grad = d_cot - sqrt_1m_acp * d_vjp  (dps.py:103 written out).
 * c_ell, c_s, std are the fp64 bridge statistics of bridge_kernels.py:28-39
 * rounded to fp32 by the caller.  d_z may be NULL when std == 0.
 * d_err_out (L,) receives |r_l|_2 when not NULL.  d_x_next may alias d_x_t.
 * Fixed-scale mode (d_err_part = NULL, err_parts = 0): the guidance term is gamma * grad, which is the
 * PGDM update samplers/samplers/pgdm.py:130-135 with gamma = guidance_weight * sqrt(1 - acp_t) and
 * d_cot = 2c A^T r / sqrt_acp (c = the operator's pseudo-inverse gain, A^+ = c A^T).
 */
PSX_API int psx_dps_post(const float* d_x_t, const float* d_eps, const float* d_cot,
                 const float* d_vjp, const float* d_z, const float* d_err_part, int err_parts,
                 int64_t L, int64_t n, float sqrt_acp, float sqrt_1m_acp, float c_ell,
                 float c_s, float std, float gamma, float* d_x_next, float* d_err_out,
                 void* stream);

/* ------------------------------------------------------------------ graph-replayable K1 / K2
 * Same kernels, but the per-timestep scalars are read from DEVICE memory, so that one captured CUDA graph of
 * a whole timestep (network forward, K1, network VJP, K2) can be replayed for every timestep of the loop
 * (dps.py:91-122) without re-recording launch parameters: the caller keeps the table of all steps on the device
 * and moves row k into d_step_row with a device-side copy inside the graph.
 *
 *   d_step_row[PSX_STEP_ROW] = { sqrt_acp, sqrt_1m_acp, lik_weight / sqrt_acp (one fp32 division),
 *                                c_ell, c_s, std, gamma, unused }
 *
 * psx_dps_pre_dev reads entries 0-2, psx_dps_post_dev entries 0, 1, 3-6; d_z is required (std may be 0 in the
 * row; the product std * z is then an exact zero).  Results are bit-identical to the by-value entry points. */
#define PSX_STEP_ROW 8
PSX_API int psx_dps_pre_dev(const psx_op* op, const float* d_x_t, const float* d_eps, const float* d_y,
                            int64_t L, int64_t obs_repeat, const float* d_step_row, float* d_cot,
                            float* d_err_part, float* d_x0_out, void* d_workspace, size_t workspace_bytes,
                            void* stream);
PSX_API int psx_dps_post_dev(const float* d_x_t, const float* d_eps, const float* d_cot, const float* d_vjp,
                             const float* d_z, const float* d_err_part, int err_parts, int64_t L, int64_t n,
                             const float* d_step_row, float* d_x_next, float* d_err_out, void* stream);

/* ------------------------------------------------------------------ K1 that also emits the bridge mean
 * The bridge mean  mean = c_ell * x_t + c_s * x0  (samplers/samplers/utils/bridge_kernels.py:41) depends on x_t and
 * eps only.  Where K1 is not HBM-bound AND leaves SMs idle -- the tensor-core blur at small batches (config 2: 96 of
 * 148 SMs hold planes): psx_op_fuses_mean(op, L) == 1 -- the same launch carries extra CTAs on the idle SMs that
 * write the mean under K1's span, and K2 then reads ONE array instead of x_t and eps: 4 B per element leave the
 * HBM-bound kernel (K2 24 -> 20 B per element).
 *   psx_dps_pre_mean[_dev]   = psx_dps_pre[_dev] + d_mean_out (L, n); c_ell / c_s / std by value or from row entries
 *                              3-5.  d_z (the step's N(0,1) field, drawn before K1) may be NULL; when given,
 *                              d_mean_out = mean + std * z (bridge_kernels.py:59) and K2 takes d_z = NULL
 *   psx_dps_post_mean[_dev]  = psx_dps_post[_dev] with d_mean in place of (d_x_t, d_eps); d_z = NULL: no noise term
 *                              (std == 0, or already folded into d_mean); d_x_next is usually the state x_t itself
 *                              (in-place step).  K2 then moves 16 (z folded) or 20 B per element instead of 24
 * The roundings are those of psx_dps_post (Tweedie, mul, mul, add; then + std*z; then + scale*grad), so the pair is
 * bit-identical to psx_dps_pre + psx_dps_post.  Other operators / larger batches: PSX_ERR_UNSUPPORTED (their K1 is
 * HBM-bound or fills the GPU; moving the 4 bytes would gain nothing). */
PSX_API int psx_op_fuses_mean(const psx_op* op, int64_t L);
PSX_API int psx_dps_pre_mean(const psx_op* op, const float* d_x_t, const float* d_eps, const float* d_y, int64_t L,
                             int64_t obs_repeat, float sqrt_acp, float sqrt_1m_acp, float lik_weight, float c_ell,
                             float c_s, const float* d_z, float std, float* d_cot, float* d_err_part,
                             float* d_mean_out, void* d_workspace, size_t workspace_bytes, void* stream);
PSX_API int psx_dps_pre_mean_dev(const psx_op* op, const float* d_x_t, const float* d_eps, const float* d_y,
                                 int64_t L, int64_t obs_repeat, const float* d_step_row, const float* d_z,
                                 float* d_cot, float* d_err_part, float* d_mean_out, void* d_workspace,
                                 size_t workspace_bytes, void* stream);
PSX_API int psx_dps_post_mean(const float* d_mean, const float* d_cot, const float* d_vjp, const float* d_z,
                              const float* d_err_part, int err_parts, int64_t L, int64_t n, float sqrt_1m_acp,
                              float std, float gamma, float* d_x_next, float* d_err_out, void* stream);
PSX_API int psx_dps_post_mean_dev(const float* d_mean, const float* d_cot, const float* d_vjp, const float* d_z,
                                  const float* d_err_part, int err_parts, int64_t L, int64_t n,
                                  const float* d_step_row, float* d_x_next, float* d_err_out, void* stream);

/* ------------------------------------------------------------------ K2 with in-kernel noise (production mode)
 * Same update as psx_dps_post, but the N(0,1) field is drawn inside the kernel instead of being read: no noise
 * tensor is written by a generator kernel and read back (36 -> 32 algorithmic B/element with the generator's
 * write).  Philox4x32-10, key = seed, counter = (flat element index / 4, step); four outputs ->
 * four normals by Box-Muller; restated in oracle/philox.py.  The field is a pure function of (seed, step, element
 * index): eager launches and graph replays agree bit for bit.  It replaces `torch.randn_like` of
 * bridge_kernels.py:59 and is NOT bit-comparable with torch's stream -- parity runs inject z through psx_dps_post.
 *   psx_dps_post_philox      scalars, seed and step by value
 *   psx_dps_post_philox_dev  scalars from d_step_row, {seed, step} from the device pair d_seed_step (graph replay)
 *   psx_philox_normal        writes the field itself: d_out[i], i < numel                                   */
PSX_API int psx_dps_post_philox(const float* d_x_t, const float* d_eps, const float* d_cot, const float* d_vjp,
                                const float* d_err_part, int err_parts, int64_t L, int64_t n, float sqrt_acp,
                                float sqrt_1m_acp, float c_ell, float c_s, float std, float gamma, uint64_t seed,
                                uint64_t step, float* d_x_next, float* d_err_out, void* stream);
PSX_API int psx_dps_post_philox_dev(const float* d_x_t, const float* d_eps, const float* d_cot, const float* d_vjp,
                                    const float* d_err_part, int err_parts, int64_t L, int64_t n,
                                    const float* d_step_row, const uint64_t* d_seed_step, float* d_x_next,
                                    float* d_err_out, void* stream);
PSX_API int psx_philox_normal(float* d_out, int64_t numel, uint64_t seed, uint64_t step, void* stream);

/* ------------------------------------------------------------------ bf16 state (production mode)
 * K1 / K2 with the sampler state, eps, the cotangent, the network VJP and the noise tensor stored as bf16
 * (d_* void pointers below are __nv_bfloat16 arrays; d_y, d_err_part, d_err_out stay fp32).  Values are widened
 * to fp32 on load, the arithmetic is that of psx_dps_pre / psx_dps_post, results are rounded to bf16 on store:
 *     K_bf16(inputs) == bf16_rn(K_fp32(float(inputs)))   bit for bit.
 * 18 B/element per step with in-kernel noise against 40 B/element in fp32; the tolerance against
 * the fp32 reference is bf16's own rounding, 2^-9 relative per stored value.  Identity, mask, 4x box and separable-blur
 * operators (the blur's intermediates in d_workspace stay fp32).
 * d_step_row (nullable) overrides the by-value scalars as in the *_dev entry points; use_philox != 0 draws the
 * noise in the kernel from (seed, step) or, when d_seed_step is not NULL, from that device pair. */
PSX_API int psx_dps_pre_bf16(const psx_op* op, const void* d_x_t, const void* d_eps, const float* d_y, int64_t L,
                             int64_t obs_repeat, float sqrt_acp, float sqrt_1m_acp, float lik_weight,
                             const float* d_step_row, void* d_cot, float* d_err_part, void* d_workspace,
                             size_t workspace_bytes, void* stream);
PSX_API int psx_dps_post_bf16(const void* d_x_t, const void* d_eps, const void* d_cot, const void* d_vjp,
                              const void* d_z, const float* d_err_part, int err_parts, int64_t L, int64_t n,
                              float sqrt_acp, float sqrt_1m_acp, float c_ell, float c_s, float std, float gamma,
                              const float* d_step_row, int use_philox, uint64_t seed, uint64_t step,
                              const uint64_t* d_seed_step, void* d_x_next, float* d_err_out, void* stream);

/* ------------------------------------------------------------- latent samplers (PSLD)
 * psx_bridge_update -- bridge (DDIM/DDPM) update with an additive correction, the tail of a PSLD step:
 *   x_next = c_ell*x + c_s*x0 + std*z + grad_scale*grad,   x0 = (x - s1*eps)/sa
 * Replaces samplers/samplers/psld.py:144-153 (ddim_step(..., e_t=z0) followed by `z_t - gradient`:
 * grad_scale = -1) and bridge_kernels.py:41,44-45,59.  d_z may be NULL when std == 0; d_grad may be
 * NULL when grad_scale == 0 (plain ddim_step).  d_x_next may alias d_x.
 */
PSX_API int psx_bridge_update(const float* d_x, const float* d_eps, const float* d_z, const float* d_grad,
                              int64_t numel, float sqrt_acp, float sqrt_1m_acp, float c_ell, float c_s,
                              float std, float grad_scale, float* d_x_next, void* stream);

/* psx_lincomb3 -- out = ca*a + cb*b + cc*c elementwise (c may be NULL).  Glue of the PSLD pixel-space
 * block around psx_dps_pre: x_eff = x0 + A^T(y - A x0) (psld.py:132-136) and its cotangent. */
PSX_API int psx_lincomb3(const float* d_a, float ca, const float* d_b, float cb, const float* d_c, float cc,
                         float* d_out, int64_t numel, void* stream);
/* psx_lincomb3_dev -- the same with the third coefficient completed ON THE DEVICE: cc * (*d_num) / (*d_den)
 * (d_den may be NULL = 1; a zero denominator switches the term off).  The cotangent of a batch-global norm,
 * -(c / ||r||) * A^T r (autograd mirror of torch.norm, psld.py:130,138, resample_kernels.py:27), has both factors
 * in device memory: the sampling loop stays free of host synchronisation. */
PSX_API int psx_lincomb3_dev(const float* d_a, float ca, const float* d_b, float cb, const float* d_c, float cc,
                             const float* d_num, const float* d_den, float* d_out, int64_t numel, void* stream);

/* ------------------------------------------------------------- ReSample
 * psx_ddim_eps_step -- DDIM update in the eps parameterisation (bridge_kernels.py:82-115), three outputs:
 *   pred_x0   = (x - sqrt_oma*eps) / sqrt_a_t
 *   pseudo_x0 = (x - oma*eps)      / sqrt_a_t            (oma = 1 - acp_t)
 *   x_prev    = sqrt_a_p*pred_x0 + dir*eps + sigma_t*z
 * The six scalars are evaluated by the caller in fp32 exactly as the reference does (0-dim tensor ops).
 * d_z may be NULL when sigma_t == 0; d_pred_x0 / d_pseudo_x0 may be NULL.
 */
PSX_API int psx_ddim_eps_step(const float* d_x, const float* d_eps, const float* d_z, int64_t numel,
                              float sqrt_a_t, float sqrt_oma, float oma, float sqrt_a_p, float dir,
                              float sigma_t, float* d_x_prev, float* d_pred_x0, float* d_pseudo_x0,
                              void* stream);

/* psx_stochastic_resample (resample_kernels.py:96-107):
 *   out = (c_p*pseudo_x0 + c_x*x_t) / den + k_n*noise
 * with c_p = sigma*sqrt(a), c_x = 1 - a, den = sigma + 1 - a, k_n = sqrt(1/(1/sigma + 1/(1-a))). */
PSX_API int psx_stochastic_resample(const float* d_pseudo_x0, const float* d_x_t, const float* d_noise,
                                    int64_t numel, float c_p, float c_x, float den, float k_n, float* d_out,
                                    void* stream);

/* psx_adamw_step -- one fused AdamW update (torch.optim.AdamW defaults' arithmetic: decoupled weight
 * decay, lerp first moment, bias-corrected step), used by the hard-consistency optimisers
 * (resample_kernels.py:32-54 pixel space, :57-93 latent space).  `step` is 1-based.
 * Optional on-device early stop for the pixel-space loop (no host sync per iteration):
 *   d_flags: int32[2] or NULL.  If d_flags[flag_in] != 0 the call is a no-op and sets d_flags[1-flag_in];
 *   otherwise the update is applied and d_flags[1-flag_in] = (sum(d_loss_parts[0..n_loss_parts)) *
 *   loss_scale < loss_threshold)  -- the reference's `if loss.item() < eps**2: break` evaluated after
 *   the step of the same iteration.
 */
PSX_API int psx_adamw_step(float* d_param, const float* d_grad, float* d_m, float* d_v, int64_t numel, float lr,
                           float beta1, float beta2, float adam_eps, float weight_decay, int step,
                           int* d_flags, int flag_in, const float* d_loss_parts, int64_t n_loss_parts,
                           float loss_scale, float loss_threshold, void* stream);

/* ------------------------------------------------------------- final estimate
 * psx_tweedie -- x0 = (x_t - s1*eps)/sa (dps.py:125-126), written straight into
 * the caller's gather slot, optionally accumulating the per-pixel sum and sum
 * of squares over the L local samples (posterior mean / variance send buffers).
 * d_sum / d_sumsq (n,) may be NULL; when given they are OVERWRITTEN.
 */
PSX_API int psx_tweedie(const float* d_x_t, const float* d_eps, int64_t L, int64_t n, float sqrt_acp,
                float sqrt_1m_acp, float* d_x0, float* d_sum, float* d_sumsq, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PSX_H_ */
