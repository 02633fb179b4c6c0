/*
 * laminr_b200 -- C ABI of the B200-native (sm_100a) LAMINR hot path.
 *
 * Drop-in boundary for the invariance-manifold optimisation loop of sinzlab/laminr v0.0.7.  The reference has no
 * FFI of its own (it is pure PyTorch); each entry point below replaces the PyTorch op sequence of one reference
 * callable, cited as `file:line` relative to the reference's `src/`.  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous row-major float32 (or int32) unless marked "host";
 *   - the caller (PyTorch's allocator) owns every buffer including workspaces (`*_workspace_bytes` queries);
 *   - no allocation, no host synchronisation, no RNG inside: calls are stream-ordered and CUDA-graph capturable;
 *   - return 0 on success, a negative lmr_status otherwise; `lmr_last_error()` gives a thread-local message;
 *   - `stream` is a cudaStream_t passed as void*.
 */
#ifndef LAMINR_B200_H
#define LAMINR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LMR_ABI_VERSION 1

typedef enum {
    LMR_OK = 0,
    LMR_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
    LMR_ERR_WORKSPACE = -2,   /* workspace too small */
    LMR_ERR_CUDA = -3,        /* CUDA runtime error (message in lmr_last_error) */
    LMR_ERR_NO_DEVICE = -4    /* not an sm_100 device */
} lmr_status;

/* arithmetic mode of the INR MLP hidden layers */
typedef enum {
    LMR_PREC_FP32 = 0,        /* CUDA-core fp32 FMA, tanhf: the bit-tight parity path */
    LMR_PREC_BF16X3_FAST = 1, /* as BF16X3 but tanh = MUFU tanh.approx.f32 (2^-11 relative error): the throughput mode */
    LMR_PREC_BF16X3 = 2       /* tcgen05 kind::f16, error-compensated split-bf16 operands (hi*hi + lo*hi + hi*lo), fp32 accumulate in
                                 TMEM, tanh from ex2/rcp (~1e-7): fp32-class results; layer 0 and the output layer stay on CUDA cores */
} lmr_precision;

/* Flag, OR-ed into `precision` of lmr_inr_forward AND of the matching lmr_inr_backward_after_forward (tensor-core modes only): the forward
 * keeps the split-bf16 operand tiles of the hidden activations h_1..h_3 and the output values y in its workspace (numTiles x (3 x 32 KB + 2 KB):
 * 163 MB for one 20 x 100 x 100 image set) and the backward loads them with bulk async copies instead of recomputing the forward per tile.  The forward workspace must
 * then have lmr_inr_forward_workspace_bytes_keep(...) bytes. */
#define LMR_PREC_KEEP_ACTIVATIONS 0x100
/* Flag, OR-ed into `precision` of lmr_inr_forward (tensor-core modes, learn stage: D == 1 and no affine transform; ignored otherwise): the sin/cos
 * table of the pixel coordinates that the previous lmr_inr_forward on this workspace left there is reused instead of recomputed.  The caller
 * guarantees the same desc / B / coords / coord_shift / P as that call (the pixel grid of a learn loop never changes; only the latents and the
 * parameters do).  Same table, same products: results are bit-identical to a call without the flag. */
#define LMR_PREC_REUSE_COORD_TABLE 0x200

typedef enum { LMR_CONSTRAIN_NORM = 0, LMR_CONSTRAIN_STD = 1 } lmr_constrain_mode;

int lmr_abi_version(void);
const char* lmr_last_error(void);
/* 0 if the current device is sm_100 (B200); LMR_ERR_NO_DEVICE otherwise. */
int lmr_check_device(void);

/* ------------------------------------------------------------------------------------------------------------
 * INR generator  -- replaces INRTemplates.forward (laminr/inr.py:53-100) incl. transform_auxiliary_input
 * (:371-386), apply_positional_encoding (:388-413), apply_coordinate_transform (:317-369) with
 * CoordinateTransform(only_affine=True) / AffineTransform (laminr/coordinate_transforms.py:226-309,351-354),
 * combine_multiple_tensors (:436-452, never materialised) and `func` (:156-174).
 *
 * Network: n_hidden layers Linear(+bias)+tanh of width `hidden`, then Linear(hidden, channels)+tanh.
 * `params` is the concatenation of func.parameters() in module order, torch Linear layout [out,in]:
 *     W0[hidden, 1+2*aux_pe+2*pe] b0[hidden]  W1[hidden,hidden] b1 ... W_L[channels,hidden] b_L[channels]
 * Input column order of W0: [template(1) | sin,cos(aux.auxB) (2*aux_pe) | sin,cos(coords.B) (2*pe)].
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
    int hidden;      /* width of every hidden layer (supported: 50) */
    int n_hidden;    /* number of hidden layers (2..8) */
    int pe_dim;      /* positional_encoding_dim: B is [2, pe_dim] */
    int aux_pe_dim;  /* aux_positional_encoding_dim: auxB is [2, aux_pe_dim] (periodic latent: [sin g, cos g]) */
    int channels;    /* out_channels, 1..4 */
} lmr_inr_desc;

/* Per-target affine coordinate transform (match stage).  All NULL => learn stage (identity; D must be 1).
 * M_d = R(angle_d) diag(scal_d) [[1,sh_d0],[sh_d1,1]];  c = clamp(M_d(c_p - tc) + t_d + tc + [M_d(shift_from_d + shift_to) + t_d])
 * with a straight-through clamp and the bracket constant in the backward pass (inr.py:345-369).  */
typedef struct {
    const float* angles;       /* [D] */
    const float* scalings;     /* [D, scal_cols] */
    int scal_cols;             /* 2, or 1 for uniform_scale */
    const float* shears;       /* [D,2] */
    const float* translation;  /* [D,2] (the reference's [D,1,2]) */
    const float* tc;           /* [2] template_center_in_coords_space (y,x); NULL => 0 */
    const float* shift_to;     /* [2] coords_shift_to_center; NULL => 0 */
    const float* shift_from;   /* [D,2] coords_shift_from_center; NULL => 0 */
    int clamp;                 /* coordinate_transform_clamp_boundaries */
} lmr_affine;

size_t lmr_inr_params_count(const lmr_inr_desc* desc);
size_t lmr_inr_forward_workspace_bytes(const lmr_inr_desc* desc, int N, int P, int D);
size_t lmr_inr_backward_workspace_bytes(const lmr_inr_desc* desc, int N, int P, int D);
size_t lmr_inr_forward_workspace_bytes_keep(const lmr_inr_desc* desc, int N, int P, int D);  /* with LMR_PREC_KEEP_ACTIVATIONS */

/* img[D,N,C,P] = INR(grid[N] latents, coords[P,2] (y,x) pixel coordinates).
 * coord_shift: optional [2] added to coords (centralize_template, inr.py:329-335), learn stage only.
 * Every pointer is device memory; `grid` alone may also point into pinned host memory (cudaHostAlloc: mapped at the same address under unified
 * addressing) -- its N floats are read once, by N blocks of the preparation kernel, so a host-fed step needs no H2D copy. */
int lmr_inr_forward(const lmr_inr_desc* desc, const float* params, const float* B, const float* auxB,
                    const float* grid, int N, const float* coords, int P, const float* coord_shift,
                    const lmr_affine* affine, int D, float* img, void* workspace, size_t workspace_bytes,
                    int precision, void* stream);

/* Backward of lmr_inr_forward given g_img[D,N,C,P] (recomputes the forward per tile; stores no activations).
 * g_params (nullable): gradient wrt `params`, same layout, overwritten.          [learn: Adam over template.func]
 * g_angles[D], g_scalings[D,scal_cols], g_shears[D,2], g_translation[D,2] (nullable as a group), overwritten.
 *                                                                       [match: Adam over coordinate_transform] */
int lmr_inr_backward(const lmr_inr_desc* desc, const float* params, const float* B, const float* auxB,
                     const float* grid, int N, const float* coords, int P, const float* coord_shift,
                     const lmr_affine* affine, int D, const float* g_img, float* g_params, float* g_angles,
                     float* g_scalings, float* g_shears, float* g_translation, void* workspace,
                     size_t workspace_bytes, int precision, void* stream);

/* lmr_inr_backward that takes the per-step tables (latent encoding, separable layer-0 parts) from `forward_workspace`, the workspace of the
 * lmr_inr_forward call that produced the images whose gradient g_img is.  The caller guarantees identical desc / params / B / auxB / grid /
 * coords / coord_shift / affine / D / precision and that the forward workspace has not been overwritten since; saves one launch per step.
 * LMR_PREC_FP32 ignores the forward workspace and recomputes. */
int lmr_inr_backward_after_forward(const lmr_inr_desc* desc, const float* params, const float* B, const float* auxB, const float* grid, int N,
                                   const float* coords, int P, const float* coord_shift, const lmr_affine* affine, int D, const float* g_img,
                                   float* g_params, float* g_angles, float* g_scalings, float* g_shears, float* g_translation,
                                   void* workspace, size_t workspace_bytes, const void* forward_workspace,
                                   size_t forward_workspace_bytes, int precision, void* stream);

/* The learn step's lmr_inr_backward_after_forward (parameter gradients of ONE image set, no affine transform) and the lmr_adam_step over the INR
 * parameters behind it -- `loss.backward(); optimizer.step()` of laminr/inv_temp_learn_match.py:206-207 with torch.optim.Adam over
 * template.func.parameters() (:162) -- as one call.  On the tensor-core precisions everything behind the tile kernel (layer-0 weight-gradient
 * partials, the cross-CTA reductions, the latent block of W0, Adam) is ONE cooperative launch instead of four; same summation orders and Adam
 * arithmetic as the two separate calls, so gradients and parameters are bit-identical to them.  params is updated in place; g_params [count]
 * still receives the gradient; adam_step: int32[2] = {steps taken, 0} as in lmr_adam_step.  LMR_PREC_FP32: runs the two calls. */
int lmr_inr_backward_adam(const lmr_inr_desc* desc, float* params, const float* B, const float* auxB, const float* grid, int N,
                          const float* coords, int P, const float* coord_shift, const float* g_img, float* g_params, float* adam_m,
                          float* adam_v, int* adam_step, float lr, float beta1, float beta2, float eps, void* workspace,
                          size_t workspace_bytes, const void* forward_workspace, size_t forward_workspace_bytes, int precision,
                          void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Image-constraint projection -- replaces FixEnergyNormClip / StandardizeClip (laminr/utils/img_transforms.py:5-65)
 * and, with eps = 1e-9, the non-differentiable ChangeNorm/ChangeStats + ClipRange (featurevis/ops.py:294-307,441-481).
 *   NORM: y = (x-mean)/(||x-mean||+eps)*target + mean          STD: y = target*(x-mu)/(std_unbiased(x)+eps) + mean
 *   z = clamp(y, pmin, pmax) (each bound optional)
 * x,z: [B, n]; stats: [B,2] = (||x-mean|| or std, mu), consumed by the backward.
 * ---------------------------------------------------------------------------------------------------------- */
size_t lmr_constrain_workspace_bytes(int B, int n);
int lmr_constrain_fwd(int mode, const float* x, int B, int n, float target, float mean, int has_min, float pmin,
                      int has_max, float pmax, float eps, float* z, float* stats, void* workspace,
                      size_t workspace_bytes, void* stream);
int lmr_constrain_bwd(int mode, const float* x, const float* g_z, const float* stats, int B, int n, float target,
                      float mean, int has_min, float pmin, int has_max, float pmax, float eps, float* g_x,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Gradient-ascent update with post-processing, batched -- replaces one iteration's `optimizer.step()` (SGD, no momentum) + `post_update`
 * of featurevis.gradient_ascent (featurevis/core.py:107-119) for any response model whose gradient g is already on the device:
 *     x <- clip(renorm(x + step_size * g))      renorm = ChangeNorm | ChangeStats (eps 1e-9 there), clip = ClipRange.
 * x (in/out), g: [B, n].  Workspace: lmr_constrain_workspace_bytes(B, n). */
int lmr_ascent_update(int mode, float* x, const float* g, float step_size, int B, int n, float target, float mean, int has_min,
                      float pmin, int has_max, float pmax, float eps, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Simulated response model -- replaces ArbitraryNeuronModel / ArbitraryMultiNeuronModel.forward
 * (laminr/neuron_models/utils/simulated_model.py:8-29) and SingleNeuronModel (laminr/utils/general.py:5-12).
 *   r_f = sum_{c,p} img[c,p] w_f[p];  m = max_f r_f (first maximum);  act = (elu(m)+1)/2 + eps
 * Work is organised in G groups: group g evaluates neuron `neuron_of_group[g]` on NB images starting at
 * img + g*img_group_stride (stride 0: every group sees the same NB images = "all neurons on all images").
 * filters: [n_neurons, Fmax, HW] (ragged banks padded by repeating a filter), nfilt[n_neurons] valid counts.
 * Outputs act, rmax: [G,NB]; argmax: [G,NB] (int32).
 * ---------------------------------------------------------------------------------------------------------- */
size_t lmr_simresp_workspace_bytes(int G, int NB, int Fmax, int HW);
int lmr_simresp_fwd(const float* img, long long img_group_stride, int NB, int G, int C, int HW,
                    const float* filters, int Fmax, const int* nfilt, const int* neuron_of_group, float eps,
                    float* act, float* rmax, int* argmax, void* workspace, size_t workspace_bytes, void* stream);
/* g_img (+)= sum over groups g_act[g,b] * 1/2 * elu'(rmax) * w_{argmax}; `accumulate` adds into g_img. */
int lmr_simresp_bwd(const float* g_act, const float* rmax, const int* argmax, long long img_group_stride, int NB,
                    int G, int C, int HW, const float* filters, int Fmax, const int* neuron_of_group, float* g_img,
                    int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Stacked-conv core + Gaussian readout response model -- replaces, for the neuron the loop selects, the chain
 * Stacked2dCore.forward (neuralpredictors/layers/cores/conv2d.py:246-253; layers :202-234) -> FullGaussian2d.forward
 * (neuralpredictors/layers/readouts/gaussian.py:532-581, eval-mode grid of sample_grid :396-429) -> FiringRateEncoder.forward
 * (neuralpredictors/layers/encoders/firing_rate.py:36-52) -> SingleNeuronModel (laminr/utils/general.py:5-12), and its autograd
 * gradient back to the image.  Plain stride-1 "same" convolutions (zero padding k//2), eval-mode batch norm folded to a per-channel
 * scale/shift, AdaptiveELU elu(v - xshift) + yshift, only the last layer read out (stack = -1), bilinear read at one position per neuron.
 * Only the receptive-field cone of the selected neuron is evaluated (identical arithmetic to the full-frame network on those pixels).
 * Work is organised like lmr_simresp_*: group g evaluates neuron `neuron_of_group[g]` on NB images [NB, c_in, H*W] starting at
 * img + g*img_group_stride (stride 0: all groups see the same images; forward only).
 *   act   [G,NB]             = elu(readout + elu_offset) + 1
 *   g_img (nullable, layout of img): (+)= d sum(g_act * act) / d img;  g_act [G,NB] nullable (= 1).  accumulate = 0 overwrites the whole
 *   gradient image (zero outside the cone); 1 adds the cone's window into it; 2 overwrites the window only (for a caller that keeps the rest
 *   of the image at zero and evaluates the same neurons again and again: the MEI ascent).
 * ---------------------------------------------------------------------------------------------------------- */
#define LMR_CONV_MAX_LAYERS 6
typedef struct {
    int n_layers;                      /* 1..LMR_CONV_MAX_LAYERS */
    int c_in;                          /* image channels */
    int hidden;                        /* output channels of every layer */
    int kernel[LMR_CONV_MAX_LAYERS];   /* odd kernel sizes, layer 0 first */
    int nonlin[LMR_CONV_MAX_LAYERS];   /* 1: AdaptiveELU after the layer, 0: linear */
    float elu_xshift, elu_yshift;
    int height, width;
    int align_corners;                 /* of the readout's grid_sample */
    float elu_offset;                  /* FiringRateEncoder.offset */
} lmr_convcore_desc;

typedef struct {
    /* channel-blocked: CB(n) = 4 if n % 4 == 0, else 2 if n is even, else 1 */
    const float* w_fwd[LMR_CONV_MAX_LAYERS];  /* [c_in_l][k][hidden / CB][k][CB], CB = CB(hidden): element [ci][ky][co / CB][kx][co % CB] = conv.weight[co][ci][ky][kx] */
    const float* w_bwd[LMR_CONV_MAX_LAYERS];  /* [hidden][k][c_in_l / CB][k][CB], CB = CB(c_in_l): element [co][ky][ci / CB][kx][ci % CB] = conv.weight[co][ci][k-1-ky][k-1-kx];
                                                 may be NULL for forward-only calls */
    const float* scale[LMR_CONV_MAX_LAYERS];  /* [hidden] gamma / sqrt(running_var + eps)  (1 without batch norm) */
    const float* shift[LMR_CONV_MAX_LAYERS];  /* [hidden] beta + (conv_bias - running_mean) * scale */
    const float* mu;                          /* [n_neurons, 2] readout positions (x -> width, y -> height); clamped to [-1,1] by the kernel */
    const float* features;                    /* [n_neurons, hidden] */
    const float* bias;                        /* [n_neurons] or NULL */
} lmr_convcore_weights;

int lmr_convcore_response(const lmr_convcore_desc* desc, const lmr_convcore_weights* weights, const float* img,
                          long long img_group_stride, int NB, int G, const int* neuron_of_group, const float* g_act, float* act,
                          float* g_img, int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Masked contrastive objective -- replaces apply_mask (laminr/utils/mask.py:7-22) + SimCLROnGrid.reg_term
 * (laminr/contrastive_loss.py:139-154) + cosine_similarity (laminr/utils/img_similarity_metrics.py:4-10).
 * img: [N, C, HW]; mask: [HW] (broadcast over channels); pos/neg: [N,N] 0/1 neighbour masks.
 * reg: [1];  K: [N,N] coefficients such that d reg/d img_i = mask * sum_j K_ij xm_j (consumed by the backward).
 * ---------------------------------------------------------------------------------------------------------- */
size_t lmr_simclr_workspace_bytes(int N, int C, int HW);
int lmr_simclr_fwd(const float* img, const float* mask, int N, int C, int HW, float pmin, float pmax,
                   const float* pos, const float* neg, float temperature, float eps, float* reg, float* K,
                   void* workspace, size_t workspace_bytes, void* stream);
/* g_img (+)= g_reg[0] * scale * d reg/d img.  g_reg is a device scalar (the autograd upstream). */
int lmr_simclr_bwd(const float* img, const float* mask, const float* K, const float* g_reg, float scale, int N,
                   int C, int HW, float pmin, float pmax, float* g_img, int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fused learn-stage objective, forward AND backward in ONE launch -- replaces, for a simulated template neuron, the chain
 * img_transf -> SingleNeuronModel -> apply_mask -> SimCLROnGrid.reg_term -> loss -> autograd back to the INR output
 * (laminr/inv_temp_learn_match.py:187-204; same maths as lmr_constrain_* + lmr_simresp_* + lmr_simclr_* above):
 *     z = constrain(x);  act_n = (elu(max_f <sum_c z_n, w_f>) + 1)/2 + act_eps;  reg = simclr(mask(z))
 *     loss = -mean_n(act_n / mei_act) + g_reg[0] * reg        (g_reg = -reg_scale, a device scalar)
 *     g_x  = d loss / d x
 * x, z, g_x: [N, C, HW];  bank: [F, HW] filters of the template neuron;  stats [N,2], rmax [N], argmax [N], Kmat [N,N] nullable.
 * `loss` (one float, written once) may point into pinned host memory: the host then has the step's loss without a D2H copy.
 * The workspace must be zero-initialised once before its first use (it holds the software grid-barrier state) and must not be shared by
 * launches that may run concurrently.  lmr_learn_objective_workspace_bytes returns 0 when the configuration does not fit the fused
 * kernel (use the unfused entry points then).
 * ---------------------------------------------------------------------------------------------------------- */
size_t lmr_learn_objective_workspace_bytes(int N, int C, int HW, int F);
int lmr_learn_objective(int mode, const float* x, int N, int C, int HW, float target, float mean, int has_min, float pmin, int has_max,
                        float pmax, float eps_con, const float* bank, int F, float act_eps, float mei_act, const float* mask,
                        const float* pos, const float* neg, float temperature, float eps_clr, const float* g_reg, float* z, float* stats,
                        float* act, float* rmax, int* argmax, float* reg, float* Kmat, float* loss, float* g_x, void* workspace,
                        size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Adam -- torch.optim.Adam(lr, betas, eps) single-tensor update (used on template.func / coordinate_transform
 * parameters, inv_temp_learn_match.py:162,335).  `step` is a device int32[2]: [0] holds the number of steps already
 * taken and is incremented by the call (graph-capturable), [1] is scratch and must be zero-initialised.  One flat tensor per call.
 * ---------------------------------------------------------------------------------------------------------- */
int lmr_adam_step(float* p, const float* g, float* exp_avg, float* exp_avg_sq, int n, float lr, float beta1,
                  float beta2, float eps, int* step, void* stream);

/* The same step over up to LMR_ADAM_MAX_GROUPS parameter tensors in ONE launch -- torch.optim.Adam over
 * coordinate_transform.parameters() (inv_temp_learn_match.py:335,366: angles, scalings, shears, translation): four launches per match step
 * otherwise, a tenth of the step's kernel time for two targets.  Every group keeps its own moments and its own int32[2] step word; element
 * arithmetic and step semantics are those of lmr_adam_step (bit-identical results).  n <= 262144 per group. */
#define LMR_ADAM_MAX_GROUPS 8
typedef struct {
    float* p;
    const float* g;
    float* exp_avg;
    float* exp_avg_sq;
    int* step;
    int n;
} lmr_adam_group;
int lmr_adam_step_multi(const lmr_adam_group* groups, int count, float lr, float beta1, float beta2, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Match epoch statistics -- replaces the host-driven reductions of the match loop (inv_temp_learn_match.py:361-362:
 * acts_d = mean_n act[d,n,d] / mei_act_d; :368-371: their mean / min / max for the stop rule and the progress bar).
 * act: [D,N], mei_act: [D]  ->  out3 = {sum_d acts_d, min_d acts_d, max_d acts_d} (the caller divides the sum by the
 * GLOBAL number of targets, which differs from D when the targets are sharded over GPUs).
 * ---------------------------------------------------------------------------------------------------------- */
int lmr_match_epoch_stats(const float* act, const float* mei_act, int D, int N, float* out3, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * MEI gradient ascent -- replaces featurevis.gradient_ascent with SGD (featurevis/core.py:54-136) driven as in
 * generate_mei (laminr/utils/mei.py:33-59), batched over G neurons (the reference loops neurons sequentially,
 * mei.py:98-119):   x <- post(x + step_size * d act/dx),  post = ChangeNorm|ChangeStats -> ClipRange (eps 1e-9).
 * x: [G,C,HW] in/out.  If apply_post_first, post() is applied once before the first evaluation (mei.py:42).
 * fevals: [G, iters+1]: f before every step and once after the last (core.py:81-82,122-126).
 * ---------------------------------------------------------------------------------------------------------- */
int lmr_mei_ascent(float* x, int G, int C, int HW, const float* filters, int Fmax, const int* nfilt,
                   const int* neuron_of_group, float step_size, int iters, int mode, float target, float mean,
                   float pmin, float pmax, float act_eps, int apply_post_first, float* fevals, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * RF masks on the device -- replaces create_mask (laminr/utils/mask.py:25-64; scipy.ndimage + skimage.morphology on
 * the host, once per neuron, mei.py:56), batched over G channel-averaged MEIs [G,H,W] (H*W < 65536):
 *   |z-score| > zscore_thresh -> binary closing (`closing_iters` dilations then erosions, 3x3 cross, outside = 0)
 *   -> largest 8-connected component (ties: first in raster order) -> convex hull of the pixel edge mid-points,
 *   tested at pixel centres (closed, exact integer arithmetic) -> separable Gaussian blur (scipy 'reflect',
 *   float64 accumulation in scipy's summation order, float32 between the axes).
 * blur_weights: device double[blur_radius+1], w[j] = weight of offset +-(blur_radius - j), w[blur_radius] = centre
 *   (scipy: radius = int(4*sigma + 0.5), normalised exp(-0.5 (k/sigma)^2)).
 * mask_out: [G,H,W] float32.  stats: uint32[G,4] = {hull pixel count, sum of hull row indices, sum of hull column
 *   indices, flag}; centroid (mask.py:61) = sums / count + 0.5; flag = 1 when nothing exceeds the threshold (the
 *   reference raises there: argmax of an empty sequence).
 * ---------------------------------------------------------------------------------------------------------- */
size_t lmr_create_mask_workspace_bytes(int G, int H, int W);
int lmr_create_mask(const float* mei, int G, int H, int W, float zscore_thresh, int closing_iters,
                    const double* blur_weights, int blur_radius, float* mask_out, unsigned* stats, void* workspace,
                    size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Gaussian blur -- replaces featurevis.ops.GaussianBlur (featurevis/ops.py:378-423), the gradient preconditioner
 * behind get_mei_dict(gb=) (laminr/utils/mei.py:44-45) and learn(gaussian_blur_sigma=) (inv_temp_learn_match.py:158):
 * F.pad(pad_mode) + depthwise correlation along y with wy, then along x with wx, result scaled by inv_norm.
 * x, out: [B, H, W] float32 planes (B = batch * channels).  wy: [2 hy + 1], wx: [2 hx + 1] float32 windows
 * (exp(-0.5 (k / sigma)^2), h = max(round(sigma * truncate), 1)); inv_norm = 1 / (sum(wy) * sum(wx)).
 * pad_mode: 0 constant (zeros), 1 reflect (torch semantics, needs h < size), 2 replicate.  workspace: B*H*W floats.
 * ---------------------------------------------------------------------------------------------------------- */
int lmr_gaussian_blur(const float* x, int B, int H, int W, const float* wy, int hy, const float* wx, int hx,
                      int pad_mode, float inv_norm, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Full-frame conv response model -- every neuron's response to every image, and the gradient of sum(g_act * act) back to the
 * images.  Replaces, in eval mode, the reference's per-call module chains
 *   SE2dCore.forward (laminr/neuron_models/utils/monkey_model.py:114-317: first convolution, DepthSeparableConv2d hidden layers
 *   neuralpredictors/layers/conv.py:4-30 with dilation, BatchNorm, ELU, SQ_EX_Block :95-108) or Stacked2dCore.forward
 *   (neuralpredictors/layers/cores/conv2d.py:246-253)
 *   -> PointPooled2d.forward (neuralpredictors/layers/readouts/point_pooled.py:121-169) or eval-mode FullGaussian2d.forward
 *   (readouts/gaussian.py:532-581; = one scale)
 *   -> elu(. + offset) + 1 (laminr/neuron_models/monkey_models.py:292-294, neuralpredictors/layers/encoders/firing_rate.py:36-52)
 * and torch autograd's gradient with respect to the images (parameters frozen).  This is the path for models whose layers pool over
 * the whole frame (squeeze-excitation) and for `model(x)` of all neurons; lmr_convcore_response stays the single-neuron cone path.
 *
 * Layers (stride 1 everywhere); after every convolution optionally  v*scale[c] + shift[c]  (eval-mode batch norm + conv bias folded)
 * and  elu(v - elu_xshift) + elu_yshift:
 *   LMR_FF_DENSE      conv2d, dilation 1, zero padding `pad`.  Weights packed for the kernel: CO = (c_out % 4 == 0 ? 4 : 1),
 *                     kp = k rounded up to 8,  w_fwd [c_in][c_out / CO][k][kp][CO] = conv.weight[co][ci][ky][kx] (0 for kx >= k);
 *                     w_bwd = the same packing of the transposed, flipped kernel  w'[ci][co][ky][kx] = w[co][ci][k-1-ky][k-1-kx]
 *                     (roles of c_in / c_out swapped).
 *   LMR_FF_DEPTHWISE  groups = channels, dilation 1 or 2, k in {3,5,7,9}; w_fwd [c][k][k], w_bwd = flipped.
 *   LMR_FF_POINTWISE  1 x 1 convolution; w_fwd [c_in][c_out] (= weight^T), w_bwd [c_out][c_in] (= weight).
 *   LMR_FF_SE         x * sigmoid(W2 relu(W1 mean_hw(x) + b1) + b2); se_w1 [se_hidden][c], se_w2 [c][se_hidden].
 * w_bwd may be NULL for forward-only use.
 * Readout: n_scales feature maps (the last layer's output and n_scales - 1 average poolings k = stride = pool_kern of it); neuron n
 * reads every map bilinearly at grid[n] (x -> width, y -> height, clamped to [-1,1]) and dots the n_scales * C values with
 * features[n][:] (scale-major), + bias[n];  act [B, n_neurons] = elu(. + elu_offset) + 1.
 * lmr_ff_forward leaves the layer outputs in the workspace; lmr_ff_backward must follow it with the same net, B and workspace:
 *   g_img [B, c_in, H, W] = d sum(g_act * act) / d img   (overwritten).
 * ---------------------------------------------------------------------------------------------------------- */
#define LMR_FF_MAX_LAYERS 16
enum { LMR_FF_DENSE = 0, LMR_FF_DEPTHWISE = 1, LMR_FF_POINTWISE = 2, LMR_FF_SE = 3 };
typedef struct {
    int kind;
    int c_in, c_out;
    int k, dilation, pad;
    int affine, elu;
    float elu_xshift, elu_yshift;
    int se_hidden;
    const float* w_fwd;
    const float* w_bwd;
    const float* scale;
    const float* shift;
    const float* se_w1;
    const float* se_b1;
    const float* se_w2;
    const float* se_b2;
} lmr_ff_layer;

typedef struct {
    int n_layers;
    int c_in, height, width;
    int n_neurons, n_scales, pool_kern, align_corners;
    float elu_offset;
    const float* grid;     /* [n_neurons, 2] */
    const float* features; /* [n_neurons, n_scales * C] */
    const float* bias;     /* [n_neurons] or NULL */
    lmr_ff_layer layers[LMR_FF_MAX_LAYERS];
} lmr_ff_net;

size_t lmr_ff_workspace_bytes(const lmr_ff_net* net, int B);
int lmr_ff_forward(const lmr_ff_net* net, const float* img, int B, float* act, void* workspace, size_t workspace_bytes, void* stream);
int lmr_ff_backward(const lmr_ff_net* net, const float* g_act, int B, float* g_img, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LAMINR_B200_H */
