/* uorf_b200.h — C ABI of libuorf_b200.so: the uORF per-scene render hot path on NVIDIA B200 (sm_100a).
 *
 * The reference (KovenYu/uORF) has no FFI of its own: the path sits behind Python nn.Module calls.
 * Each entry point below names the reference interface it replaces (paths relative to the reference
 * repository root).  The host-side mirror of those interfaces (same class / function names and
 * argument meaning) lives in uorf_b200/modules.py and binds these symbols with ctypes; INTEGRATION.md
 * shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 data unless stated; tensors are contiguous, row-major;
 *   - the caller allocates all outputs and workspaces (no ownership crosses the ABI);
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work, they never synchronise;
 *   - return value 0 = ok, negative = error (see UORF_ERR_*); uorf_last_error() gives the text of the
 *     last error on the calling thread; nothing throws across the boundary;
 *   - there is no CPU fallback: on a device that is not sm_100 every call returns UORF_ERR_DEVICE.
 */
#ifndef UORF_B200_H_
#define UORF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UORF_ABI_VERSION 2

#define UORF_OK 0
#define UORF_ERR_ARG (-1)
#define UORF_ERR_UNSUPPORTED (-2)
#define UORF_ERR_DEVICE (-3)
#define UORF_ERR_WORKSPACE (-4)

int uorf_abi_version(void);
const char* uorf_last_error(void);
/* Checks that `device` is an sm_100 part and configures the kernels (max dynamic shared memory).
 * Idempotent; every other call does it implicitly for the current device. */
int uorf_init(int device);
/* Number of kernels this library has launched since it was loaded (all devices, all threads).  bench.py reports the
 * difference across its timed region as `gpu_launches`; launches replayed from a CUDA graph are counted at capture time. */
int64_t uorf_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Kernel 1 — frustum sample generation.
 * Replaces Projection.construct_sampling_coor (models/projection.py:51-113, grid built by
 * construct_frus_coor :33-49).
 *   z_cam       [D] camera-space depths, torch.linspace(near, far, D) of projection.py:42
 *   cam2world   [N][4][4]
 *   coords      [n_chunks][N*D*Hc*Wc][3]   reference order ((n*D+d)*Hc+h)*Wc+w   (ray_major = 0)
 *                                          or ray-major   ((n*Hc+h)*Wc+w)*D+d   (ray_major = 1)
 *   z_vals      [n_chunks][N*Hc*Wc][D]     (may be NULL)
 *   ray_dir     [n_chunks][N*Hc*Wc][3]     (may be NULL)
 * scale = 1 gives the unpartitioned output; scale > 1 the reference's strided partition into scale^2
 * chunks (chunk j <-> (h0,w0) = divmod(j, scale), Hc = H/scale, Wc = W/scale).
 * With crop_h/crop_w >= 0 (scale must be 1) only the window [crop_h, crop_h+Hc) x [crop_w, crop_w+Wc)
 * is produced (fine-stage training crop, models/uorf_nogan_model.py:153-165); pass Hc = Wc = render size.
 * Numerics: z_vals and every index layout are bit-identical to the reference (ray_dir <= 1e-6); coords are evaluated as
 * origin + z * direction per ray (one FMA per coordinate) and agree with the reference's fp32 matmul chain to
 * rounding (<= 2e-6 abs on the NSS range; all layouts / kernels of this entry point agree bit for bit with each other).
 * ------------------------------------------------------------------------------------------- */
typedef struct uorf_frustum {
  int32_t W, H, D;          /* frustum_size */
  float near_plane, far_plane;
  float spixel2cam[16];     /* inverse intrinsics, row-major 4x4 (projection.py:24-31) */
  float nss_scale;          /* world2nss = diag(1/nss_scale) (projection.py:17-20) */
} uorf_frustum;

int uorf_sample_coords(const uorf_frustum* f, const float* z_cam, const float* cam2world, int32_t n_views, int32_t scale,
                       int32_t crop_h, int32_t crop_w, int32_t Hc, int32_t Wc, int32_t ray_major,
                       float* coords, float* z_vals, float* ray_dir, void* stream);
/* Same, for a crop window whose origin lives on the DEVICE: crop_origin -> int32 (h0, w0), read by the kernel and clamped to
 * [0, H-Hc] x [0, W-Wc].  The fine-stage training step draws its window with torch.randint on the device
 * (models/uorf_nogan_model.py:160-161); reading the two ints on the device keeps the step free of host synchronisation, so the
 * whole fine-stage step can be captured and replayed as one CUDA graph. */
int uorf_sample_coords_window(const uorf_frustum* f, const float* z_cam, const float* cam2world, int32_t n_views,
                              const int32_t* crop_origin, int32_t Hc, int32_t Wc, int32_t ray_major,
                              float* coords, float* z_vals, float* ray_dir, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Kernel 2 — fused K-slot decoder.
 * Replaces Decoder.forward (models/model.py:106-161) including sin_emb (:261-277).
 * Weights are the reference's state-dict tensors (fp32, [out][in] row-major).
 * ------------------------------------------------------------------------------------------- */
typedef struct uorf_decoder_weights {
  /* foreground MLP */
  const float *f_before_w[3], *f_before_b[3]; /* f_before.{0,2,4}: [Z][33+Z], [Z][Z], [Z][Z] */
  const float *f_after_w[3], *f_after_b[3];   /* f_after.{0,2,4}:  [Z][33+2Z], [Z][Z], [Z][Z] */
  const float *f_after_latent_w, *f_after_latent_b; /* [Z][Z] */
  const float *f_after_shape_w, *f_after_shape_b;   /* [1][Z] */
  const float *f_color0_w, *f_color0_b;             /* f_color.0 [Z/4][Z] */
  const float *f_color2_w, *f_color2_b;             /* f_color.2 [3][Z/4] */
  /* background MLP */
  const float *b_before_w[3], *b_before_b[3]; /* b_before.{0,2,4} */
  const float *b_after_w[4], *b_after_b[4];   /* b_after.{0,2,4,6}; .6 is [4][Z] */
} uorf_decoder_weights;

/* precision of the width-Z layers on the tensor cores (accumulation is always fp32):
 *   BF16 / FP16      one pass with 16-bit operands (fast modes);
 *   BF16X3 / FP16X3  three-term split a_hi*w_hi + a_lo*w_hi + a_hi*w_lo: ~16 (bf16) / ~22 (fp16) operand mantissa
 *                    bits.  FP16X3 is the fp32-parity mode; fp16 conversions saturate at +-65504 (never inf),
 *                    BF16X3 keeps the full fp32 range. */
#define UORF_PREC_BF16 1
#define UORF_PREC_FP16 2
#define UORF_PREC_BF16X3 3
#define UORF_PREC_FP16X3 4

/* Bytes of the prepared weight image for (K slots, precision). */
size_t uorf_decoder_image_bytes(int32_t K, int32_t precision);

/* Packs the weights into the tensor-core image and hoists the slot-latent columns of the two wide
 * layers into per-slot biases.  Must be re-run when weights or z_slots change (once per scene).
 *   z_slots [K][Z] (slot 0 = background), image: uorf_decoder_image_bytes(K, precision) bytes, 1024-aligned. */
int uorf_decoder_prepare(const uorf_decoder_weights* w, const float* z_slots, int32_t K, int32_t Z,
                         int32_t n_freq, int32_t n_layers, int32_t precision, void* image, void* stream);

typedef struct uorf_decoder_args {
  const float* coor_bg;     /* [P][3] */
  const float* coor_fg;     /* slot s (0..K-2), point p at coor_fg + s*fg_slot_stride + 3*p */
  int64_t fg_slot_stride;   /* in floats; 0 for the reference's expand() view of coor_bg */
  const float* fg_transform;/* device: 9 floats (3x3) or, with fixed_locality, 16 floats (4x4) */
  const float* noise;       /* [K][P] standard normals (model.py:155) or NULL when dens_noise == 0 */
  float dens_noise;
  float locality_ratio;
  int32_t locality;         /* Decoder.locality */
  int32_t fixed_locality;   /* Decoder.fixed_locality */
  int64_t P;
  int32_t K;
  int32_t precision;        /* UORF_PREC_* */
  const void* image;        /* from uorf_decoder_prepare (same K, precision) */
  float* workspace;         /* uorf_decoder_forward_workspace_floats(P, K) floats of scratch */
  /* outputs; any may be NULL to skip the write */
  float* raws;              /* [P][4] */
  float* masked_raws;       /* [K][P][4] */
  float* unmasked_raws;     /* [K][P][4] */
  float* masks;             /* [K][P] */
  uint8_t* outside;         /* [K-1][P] locality flags (outsider_idx, model.py:121/128) */
  const float* fg_slot_offset; /* device [K-1][3] or NULL.  Slot s evaluates coor_fg - fg_slot_offset[s] (fp32 subtract, then the
                                  reference's transform / locality test): the moved-object render of uorf_manip_model.py:201-204
                                  without materialising its (K-1)xPx3 clone of the coordinates. */
} uorf_decoder_args;

/* Floats of scratch uorf_decoder_forward needs (background sweep results + the per-tile head stash). */
size_t uorf_decoder_forward_workspace_floats(int64_t P, int32_t K);
int uorf_decoder_forward(const uorf_decoder_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Kernel 5 (decoder part) — backward of Decoder.forward w.r.t. every decoder weight and z_slots, given dL/d raws.
 * Replaces torch.autograd through models/model.py:106-161 as exercised by uorfNoGanModel.backward
 * (models/uorf_nogan_model.py:223-227; only `raws` carries gradient, masked/unmasked/masks are detached :185-196).
 * Activations are recomputed per tile on chip (fp16 activations, weights and activation gradients - the latter under a
 * power-of-two loss scale -, fp32 accumulation); nothing is stored by
 * the forward pass except its `unmasked_raws` output.
 * ------------------------------------------------------------------------------------------- */
typedef struct uorf_decoder_grads {   /* same fields / shapes as uorf_decoder_weights, written by the call */
  float *f_before_w[3], *f_before_b[3];
  float *f_after_w[3], *f_after_b[3];
  float *f_after_latent_w, *f_after_latent_b;
  float *f_after_shape_w, *f_after_shape_b;
  float *f_color0_w, *f_color0_b;
  float *f_color2_w, *f_color2_b;
  float *b_before_w[3], *b_before_b[3];
  float *b_after_w[4], *b_after_b[4];
} uorf_decoder_grads;

typedef struct uorf_decoder_bwd_args {
  const float* coor_bg;       /* as in uorf_decoder_args */
  const float* coor_fg;
  int64_t fg_slot_stride;
  const float* fg_transform;
  const float* noise;         /* the noise tensor the forward used, or NULL */
  float dens_noise;
  float locality_ratio;
  int32_t locality;
  int32_t fixed_locality;
  int64_t P;
  int32_t K;
  int32_t Z;
  const float* z_slots;       /* [K][Z] */
  const void* image;          /* uorf_decoder_prepare(..., UORF_PREC_FP16, ...) of the same weights / slots */
  const float* unmasked_raws; /* [K][P][4] forward output */
  const float* grad_raws;     /* [P][4] */
  float* workspace;           /* uorf_decoder_backward_workspace_floats(P, K) floats */
  float* grad_z_slots;        /* [K][Z] out */
} uorf_decoder_bwd_args;

size_t uorf_decoder_backward_workspace_floats(int64_t P, int32_t K);
int uorf_decoder_backward(const uorf_decoder_weights* w, const uorf_decoder_grads* g, const uorf_decoder_bwd_args* a,
                          void* stream);

/* ---------------------------------------------------------------------------------------------
 * Kernel 3 — alpha compositing.  Replaces raw2outputs (models/model.py:279-313).
 *   raw [R][D][4] (rgb, sigma), z_vals [geom_rays][D], rays_d [geom_rays][3]
 *   rgb_map [R][3], depth_map [R], weights_norm [R][D] (may be NULL), mask_map [R] (NULL unless render_mask)
 * geom_rays = R for the reference call; R = k*geom_rays composites k stacked ray sets that share one
 * geometry in a single launch (the K per-slot renders of uorf_eval_model.py:181-190); 0 means R.
 * ------------------------------------------------------------------------------------------- */
int uorf_composite_forward(const float* raw, const float* z_vals, const float* rays_d, int64_t R, int64_t geom_rays, int32_t D,
                           float* rgb_map, float* depth_map, float* weights_norm, float* mask_map,
                           void* stream);
/* Backward of rgb_map w.r.t. raw (weights_norm is detached in the reference, model.py:304, so depth
 * carries no gradient).  grad_rgb [R][3] -> grad_raw [R][D][4]. */
int uorf_composite_backward(const float* raw, const float* z_vals, const float* rays_d, const float* grad_rgb,
                            int64_t R, int64_t geom_rays, int32_t D, float* grad_raw, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Kernel 4 — slot attention.  Replaces SlotAttention.forward (models/model.py:205-258), batch 1.
 * ------------------------------------------------------------------------------------------- */
typedef struct uorf_slot_attention_weights {
  const float *slots_mu, *slots_logsigma, *slots_mu_bg, *slots_logsigma_bg; /* [C] */
  const float *norm_feat_w, *norm_feat_b;                                     /* LayerNorm [C] */
  const float *to_k_w, *to_v_w;                                               /* [C][C], no bias */
  const float *to_q_ln_w, *to_q_ln_b, *to_q_w;                                /* to_q.0 / to_q.1 */
  const float *to_q_bg_ln_w, *to_q_bg_ln_b, *to_q_bg_w;
  const float *gru_w_ih, *gru_w_hh, *gru_b_ih, *gru_b_hh;                     /* [3C][C], [3C] */
  const float *gru_bg_w_ih, *gru_bg_w_hh, *gru_bg_b_ih, *gru_bg_b_hh;
  const float *to_res_ln_w, *to_res_ln_b, *to_res_w1, *to_res_b1, *to_res_w2, *to_res_b2; /* [Hd][C], [C][Hd] */
  const float *to_res_bg_ln_w, *to_res_bg_ln_b, *to_res_bg_w1, *to_res_bg_b1, *to_res_bg_w2, *to_res_bg_b2;
} uorf_slot_attention_weights;

typedef struct uorf_slot_attention_grads {   /* same fields / shapes as uorf_slot_attention_weights, written by the backward */
  float *slots_mu, *slots_logsigma, *slots_mu_bg, *slots_logsigma_bg;
  float *norm_feat_w, *norm_feat_b;
  float *to_k_w, *to_v_w;
  float *to_q_ln_w, *to_q_ln_b, *to_q_w;
  float *to_q_bg_ln_w, *to_q_bg_ln_b, *to_q_bg_w;
  float *gru_w_ih, *gru_w_hh, *gru_b_ih, *gru_b_hh;
  float *gru_bg_w_ih, *gru_bg_w_hh, *gru_bg_b_ih, *gru_bg_b_hh;
  float *to_res_ln_w, *to_res_ln_b, *to_res_w1, *to_res_b1, *to_res_w2, *to_res_b2;
  float *to_res_bg_ln_w, *to_res_bg_ln_b, *to_res_bg_w1, *to_res_bg_b1, *to_res_bg_w2, *to_res_bg_b2;
} uorf_slot_attention_grads;

/* Floats of workspace uorf_slot_attention_forward needs. */
size_t uorf_slot_attention_workspace_floats(int32_t n_tokens, int32_t C, int32_t K);

/*   feat: element (token t, channel c) at feat + t*feat_token_stride + c*feat_chan_stride (floats), so the
 *         Encoder's NCHW feature map is consumed in place (token stride 1, channel stride n_tokens);
 *   noise_fg [K-1][C], noise_bg [C] (the randn draws of model.py:216,219)
 *   slots [K][C] (slot 0 = background), attn [K][n_tokens] (last iteration, includes +eps) */
int uorf_slot_attention_forward(const uorf_slot_attention_weights* w, const float* feat, int64_t feat_token_stride,
                                int64_t feat_chan_stride, const float* noise_fg, const float* noise_bg,
                                int32_t n_tokens, int32_t C, int32_t K, int32_t hidden, int32_t iters, float eps,
                                float* workspace, float* slots, float* attn, void* stream);

/* Kernel 5 (slot-attention part) - backward of SlotAttention.forward (torch.autograd through models/model.py:205-258 as
 * run by uorfNoGanModel.backward): gradient of every parameter and of feat, given d slots [K][C] (attn is detached by the
 * reference's driver).  fwd_workspace is the workspace the forward call filled (it keeps k, v and 3KC+K floats of state per
 * iteration; everything else is recomputed).  grad_feat [n_tokens][C] contiguous. */
size_t uorf_slot_attention_backward_workspace_floats(int32_t n_tokens, int32_t C, int32_t K, int32_t hidden);
int uorf_slot_attention_backward(const uorf_slot_attention_weights* w, const uorf_slot_attention_grads* g, const float* feat,
                                 int64_t feat_token_stride, int64_t feat_chan_stride, const float* noise_fg,
                                 const float* noise_bg, int32_t n_tokens, int32_t C, int32_t K, int32_t hidden, int32_t iters,
                                 float eps, const float* fwd_workspace, const float* grad_slots, float* workspace,
                                 float* grad_feat, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Encoder layer (inference path) — 3x3 convolution (padding 1, stride 1 or 2) + bias + ReLU, fp32.
 * Replaces one `nn.Sequential(nn.Conv2d(..., 3, stride, padding=1), nn.ReLU)` of Encoder (models/model.py:19-34) together with
 * the data movement around it in Encoder.forward (models/model.py:36-61): the input is the channel concatenation of src_a
 * [ca][*][*] and src_b [cb][h_in][w_in] (torch.cat, :49,:58,:60); with upsample_a != 0, src_a is stored at half resolution
 * [ca][h_in/2][w_in/2] and read through nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False) (:27-32).
 * weight_t is the Conv2d weight transposed to [ca+cb][3][3][c_out]; out [c_out][h_out][w_out], h_out = (h_in-1)/stride + 1.
 * Batch 1 (the drivers encode view 0 only, uorf_eval_model.py:121-123).
 * ------------------------------------------------------------------------------------------- */
int uorf_encoder_conv3x3(const float* src_a, int32_t ca, int32_t upsample_a, const float* src_b, int32_t cb, const float* weight_t,
                         const float* bias, int32_t c_out, int32_t h_in, int32_t w_in, int32_t stride, float* out, void* stream);

/* Backward of the same layer, fp32 - what `loss.backward()` (models/uorf_nogan_model.py:181-183) runs through autograd / cuDNN
 * for each Conv2d + ReLU of the Encoder (models/model.py:19-34).  kept_out [c_out][h_out][w_out] is the layer's forward output: the
 * ReLU gate (grad_out counts where kept_out > 0) is applied while reading, no gated copy is made.
 *   _dgrad: grad_in [c_in][h_in][w_in] (all input channels of the layer: the caller splits the concatenation and takes the
 *           upsample backward), added to grad_in's contents when accumulate != 0 (second gradient of a skip connection); weight_g = the Conv2d weight as [c_out][3][3][c_in] (permute(0, 2, 3, 1)).
 *   _wgrad: grad_weight [c_out][ca+cb][3][3] and grad_bias [c_out]; the layer input is described as in the forward call (two
 *           sources, source A optionally through the upsample); workspace: uorf_encoder_wgrad_workspace_floats(...) floats.
 * Sums run in a fixed order (bit-reproducible); they differ from cuDNN's by fp32 rounding only. */
/* weight layouts of the layer calls above for up to 8 layers in one launch (`layers` is a HOST array; the training step refreshes
 * them once per step): weight [c_out][c_in][3][3] -> fwd_t [c_in][3][3][c_out] (uorf_encoder_conv3x3's weight_t) and
 * dgrad_t [c_out][3][3][c_in] (uorf_encoder_conv3x3_dgrad's weight_g); either output may be NULL. */
typedef struct uorf_encoder_prep_layer {
  const float* weight;
  float* fwd_t;
  float* dgrad_t;
  int32_t c_out;
  int32_t c_in;
} uorf_encoder_prep_layer;
int uorf_encoder_weight_prep(const uorf_encoder_prep_layer* layers, int32_t n_layers, void* stream);
int uorf_encoder_conv3x3_dgrad(const float* grad_out, const float* kept_out, int32_t c_out, const float* weight_g, int32_t c_in,
                               int32_t h_in, int32_t w_in, int32_t stride, int32_t accumulate, float* grad_in, void* stream);
/* backward of nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False) (models/model.py:27-32): grad_hi [c][2 h_lo][2 w_lo]
 * -> grad_lo [c][h_lo][w_lo], as a fixed-order gather (the library op scatters with atomics) */
int uorf_encoder_upsample2x_bwd(const float* grad_hi, int32_t c, int32_t h_lo, int32_t w_lo, float* grad_lo, void* stream);
int64_t uorf_encoder_wgrad_workspace_floats(int32_t c_in, int32_t c_out, int32_t h_in, int32_t w_in, int32_t stride);
int uorf_encoder_conv3x3_wgrad(const float* src_a, int32_t ca, int32_t upsample_a, const float* src_b, int32_t cb, const float* grad_out,
                               const float* kept_out, int32_t c_out, int32_t h_in, int32_t w_in, int32_t stride, float* workspace,
                               float* grad_weight, float* grad_bias, void* stream);

/* nn.MaxPool2d(2, 2) of the same network (models/model.py:316-324, VGG16 pools 1-3) folded into the operand split of its reader:
 *   uorf_conv_split_pool2x2:    x NHWC [n][h][w][c] fp32 -> hi / lo planes [n][h/2][w/2][c] of maxpool(x); the pooled tensor is not
 *                               materialised (the next convolution is its only reader);
 *   uorf_conv_split_unpool2x2:  backward: grad_pooled [n][h/2][w/2][c], x the pool's input (the ReLU output of the layer before it)
 *                               -> hi / lo planes [n][h][w][c] of the gradient routed to the first maximum of every window (torch's
 *                               choice) where x > 0: the input of that layer's data-gradient convolution.
 * h, w, c even. */
int uorf_conv_split_pool2x2(const float* x, void* hi, void* lo, int32_t n, int32_t h, int32_t w, int32_t c, int32_t precision, void* stream);
int uorf_conv_split_unpool2x2(const float* grad_pooled, const float* x, void* hi, void* lo, int32_t n, int32_t h, int32_t w, int32_t c,
                              int32_t precision, void* stream);

/* 3-channel stem of the same network (VGG16 conv1_1 + ReLU, models/model.py:316-324), fp32 FFMA:
 *   uorf_conv3x3_stem:        x NCHW [n][3][h][w], weight [c_out][3][3][3] (the Conv2d tensor), bias [c_out]
 *                             -> y NHWC [n][h][w][c_out] = relu(conv + bias): the layout uorf_conv_split / uorf_conv3x3_x3 read;
 *   uorf_conv3x3_stem_dgrad:  grad_y NHWC taken where y > 0 -> grad_x NCHW [n][3][h][w] (fixed summation order). */
int uorf_conv3x3_stem(const float* x, const float* weight, const float* bias, int32_t n, int32_t h, int32_t w, int32_t c_out, float* y,
                      void* stream);
int uorf_conv3x3_stem_dgrad(const float* grad_y, const float* y, const float* weight, int32_t n, int32_t h, int32_t w, int32_t c_out,
                            float* grad_x, void* stream);

/* Image-side losses of the training step (models/uorf_nogan_model.py:168-183: `x_recon = rendered * 2 - 1`, `L2_loss(x_recon, x)`,
 * `vgg_normalizer` of both images, `L2_loss` of the two feature maps), each direction one launch; means are block partial sums
 * added in a fixed order (bit-reproducible).  `scratch`: UORF_LOSS_SCRATCH_FLOATS floats, ZERO-INITIALISED ONCE by the caller and
 * then left alone (it holds the partial sums and a self-resetting ticket); one scratch per call site.
 *   uorf_image_loss_forward:   rgb_map [n*hw][3] (uorf_composite_forward's rgb), x [n][3][hw], mean / stdv [3] ->
 *                              x_recon, rendered, rendered_norm, x_norm [n][3][hw], *loss_recon = mean((x_recon - x)^2)
 *   uorf_image_loss_backward:  grad_rgb_map [n*hw][3] from grad_x_recon / grad_rendered_norm [n][3][hw] and *grad_loss (each may be NULL)
 *   uorf_feature_mse_forward:  features = [rendered half | ground-truth half], m floats each, same layout -> *loss = mean((a - b)^2)
 *   uorf_feature_mse_backward: grad_features [2 m]: first half *grad_loss * 2 (a - b) / m, second half zero (constants) */
#define UORF_LOSS_SCRATCH_FLOATS 72
int uorf_image_loss_forward(const float* rgb_map, const float* x, const float* mean, const float* stdv, int32_t n, int32_t hw, float* x_recon,
                            float* rendered, float* rendered_norm, float* x_norm, float* scratch, float* loss_recon, void* stream);
int uorf_image_loss_backward(const float* x_recon, const float* x, const float* stdv, const float* grad_x_recon, const float* grad_rendered_norm,
                             const float* grad_loss, int32_t n, int32_t hw, float* grad_rgb_map, void* stream);
int uorf_feature_mse_forward(const float* features, int64_t m, float* scratch, float* loss, void* stream);
int uorf_feature_mse_backward(const float* features, const float* grad_loss, int64_t m, float* grad_features, void* stream);

/* Inverse of `batch` row-major dim x dim matrices (dim 3 or 4), device to device, one launch: the `.inverse()` of the scene pose in
 * the drivers' set_input (models/uorf_nogan_model.py:117-126, uorf_eval_model.py:98-107) when the pose already lives on the device.
 * Gauss-Jordan with partial pivoting in fp32; a singular matrix gives inf / nan entries as the library does. */
int uorf_invert_small(const float* src, float* dst, int32_t batch, int32_t dim, void* stream);

/* -----------------------------------------------------------------------------------------------
 * Adam update of many small fp32 tensors in one launch - the `optimizer.step()` of the training drivers
 * (reference models/uorf_nogan_model.py:59-61 torch.optim.Adam over encoder + slot attention + decoder, :229-245 the step).
 * `chunks` is a DEVICE array: every entry is at most UORF_ADAM_CHUNK consecutive elements of one parameter with the matching
 * slices of its gradient and both moment buffers, and the parameter's step counter (a device float, already incremented for
 * this update - torch's capturable / fused state layout).  lr_dev != NULL: the learning rate is read from that device float
 * (graph replays with a scheduler), else lr_host.  Update rule of torch.optim.Adam (amsgrad off, maximize off).
 * --------------------------------------------------------------------------------------------- */
#define UORF_ADAM_CHUNK 1024
typedef struct uorf_adam_chunk {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  const float* step;
  int32_t n;
  int32_t reserved;
} uorf_adam_chunk;
int uorf_adam_step(const uorf_adam_chunk* chunks, int32_t n_chunks, const float* lr_dev, float lr_host, float beta1, float beta2, float eps,
                   float weight_decay, void* stream);

/* -----------------------------------------------------------------------------------------------
 * Kernel 6: 3x3 convolution (stride 1, pad 1) of the perceptual-loss network, fp32-parity on the tensor cores.
 * Replaces the cuDNN fp32 convolutions behind `get_perceptual_net` (reference models/model.py:316-324: VGG16 features[:23]) as
 * the training step uses them (models/uorf_nogan_model.py:181-183): forward on both images, data gradient of the rendered branch.
 *   uorf_conv_split:  x [n] fp32 (times [gate > 0] when gate != NULL) -> 16-bit hi / lo planes [n] (precision: UORF_PREC_FP16X3 /
 *                     UORF_PREC_BF16X3); n even.
 *   uorf_conv3x3_x3:  y[p, co] = bias[co] + sum_{tap, ci} x[p + tap, ci] w[co, ci, tap], NHWC, zero padding, optional ReLU.
 *                     x_hi / x_lo [n*h*w][c_in] from uorf_conv_split; c_in, c_out multiples of 64.
 *                     w_img: [c_out/64][9 taps][c_in/64][hi | lo] tiles of 64 x 64 16-bit values, rows = output channels, 128 bytes
 *                     per row, 16-byte chunk c of row r stored at chunk c ^ (r & 7)  (packed on the host side, modules.pack_conv3x3).
 *                     The data gradient is the same call on (grad_out * [y > 0]) with the flipped / transposed weight image.
 *                     workspace: uorf_conv3x3_workspace_floats(...) floats (0 for layers with enough pixel tiles), or NULL: layers
 *                     with few tiles divide their K iterations over several CTAs (split-K) and a second kernel sums the partial
 *                     results in a fixed order, then adds bias / ReLU.
 * ------------------------------------------------------------------------------------------- */
int uorf_conv_split(const float* x, const float* gate, void* hi, void* lo, int64_t n, int32_t precision, void* stream);
size_t uorf_conv3x3_workspace_floats(int32_t n, int32_t h, int32_t w, int32_t c_in, int32_t c_out);
int uorf_conv3x3_x3(const void* x_hi, const void* x_lo, const void* w_img, const float* bias, float* y, float* workspace, int32_t n, int32_t h,
                    int32_t w, int32_t c_in, int32_t c_out, int32_t relu, int32_t precision, void* stream);

/* Test hook: site code written by a kernel whose mbarrier wait timed out (0 = none). */
int uorf_debug_flag(void);

#ifdef __cplusplus
}
#endif
#endif /* UORF_B200_H_ */
