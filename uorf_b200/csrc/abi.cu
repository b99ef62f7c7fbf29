// extern "C" surface of libuorf_b200.so (declared in include/uorf_b200.h): argument validation, the
// per-device context, error strings.  Kernels live in the sibling .cu files.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <atomic>
#include <mutex>
#include <stdio.h>
#include <string.h>

#include "../../include/uorf_b200.h"
#include "common.cuh"
#include "decoder_layout.h"

namespace uorf {
int launch_sample_coords(const uorf_frustum* f, const float* z_cam, const float* cam2world, int n_views, int scale, int crop_h,
                         int crop_w, const int* crop_origin, int Hc, int Wc, int ray_major, float* coords, float* z_vals, float* ray_dir,
                         int num_sms, cudaStream_t stream);
int launch_encoder_dgrad3x3(const float* grad_out, const float* kept_out, int c_out, const float* weight_g, int c_in, int h_in, int w_in,
                            int stride, int accumulate, float* grad_in, cudaStream_t stream);
int launch_encoder_upsample2x_bwd(const float* grad_hi, int c, int h_lo, int w_lo, float* grad_lo, cudaStream_t stream);
size_t encoder_wgrad_workspace_floats(int c_in, int c_out, int h_in, int w_in, int stride);
int launch_encoder_wgrad3x3(const float* src_a, int ca, int up_a, const float* src_b, int cb, const float* grad_out, const float* kept_out,
                            int c_out, int h_in, int w_in, int stride, float* workspace, float* grad_weight, float* grad_bias,
                            cudaStream_t stream);
int launch_vgg_stem_fwd(const float* x, const float* weight, const float* bias, int n, int h, int w, int c_out, float* y, cudaStream_t stream);
int launch_vgg_stem_dgrad(const float* g, const float* y, const float* weight, int n, int h, int w, int c_out, float* gx, cudaStream_t stream);
int launch_encoder_weight_prep(const uorf_encoder_prep_layer* layers, int n_layers, cudaStream_t stream);
int launch_conv_split_pool(const float* x, const float* g, void* hi, void* lo, int n, int h, int w, int c, int f16, int num_sms, cudaStream_t stream);
int launch_invert_small(const float* src, float* dst, int batch, int dim, cudaStream_t stream);
int launch_image_loss_fwd(const float* rgb, const float* x, const float* mean, const float* stdv, int n, int hw, float* x_recon, float* rendered,
                          float* rendered_norm, float* x_norm, float* scratch, float* loss, cudaStream_t stream);
int launch_image_loss_bwd(const float* x_recon, const float* x, const float* stdv, const float* g_x_recon, const float* g_rendered_norm,
                          const float* g_loss, int n, int hw, float* grad_rgb, cudaStream_t stream);
int launch_feat_mse_fwd(const float* f, long long m, float* scratch, float* loss, cudaStream_t stream);
int launch_feat_mse_bwd(const float* f, const float* g_loss, long long m, float* grad, cudaStream_t stream);
int launch_adam_step(const uorf_adam_chunk* chunks, int n_chunks, const float* lr_dev, float lr_host, float beta1, float beta2, float eps,
                     float weight_decay, cudaStream_t stream);
int launch_encoder_conv3x3(const float* src_a, int ca, int up_a, const float* src_b, int cb, const float* weight_t, const float* bias,
                           int c_out, int h_in, int w_in, int stride, float* out, cudaStream_t stream);
int launch_decoder_prep(const uorf_decoder_weights* w, const float* z_slots, int K, int Z, int passes, int f16, void* image,
                        cudaStream_t stream);
int launch_decoder_fwd(const uorf_decoder_args* a, int num_sms, int* err_flag, cudaStream_t stream);
size_t decoder_fwd_workspace_floats(long long P, int K, int num_sms);
int launch_composite_fwd(const float* raw, const float* z_vals, const float* rays_d, long long R, long long Rg, int D, float* rgb_map,
                         float* depth_map, float* weights_norm, float* mask_map, int num_sms, cudaStream_t stream);
int launch_composite_bwd(const float* raw, const float* z_vals, const float* rays_d, const float* grad_rgb, long long R, long long Rg, int D,
                         float* grad_raw, int num_sms, cudaStream_t stream);
size_t slot_attention_ws_floats(int ntok, int C, int K);
int launch_slot_attention(const uorf_slot_attention_weights* w, const float* feat, long long feat_ts, long long feat_cs,
                          const float* noise_fg, const float* noise_bg, int ntok, int C, int K, int hidden, int iters, float eps,
                          float* workspace, float* slots, float* attn, cudaStream_t stream);
size_t slot_attention_bwd_ws_floats(int ntok, int C, int K, int hidden);
int launch_slot_attention_bwd(const uorf_slot_attention_weights* w, const uorf_slot_attention_grads* g, const float* feat, long long feat_ts,
                              long long feat_cs, const float* noise_fg, const float* noise_bg, int ntok, int C, int K, int hidden, int iters,
                              float eps, const float* fwd_workspace, const float* grad_slots, float* workspace, float* grad_feat,
                              cudaStream_t stream);
size_t decoder_bwd_workspace_floats(long long P, int K, int num_sms);
int launch_conv_split(const float* x, const float* gate, void* hi, void* lo, long long n, int f16, int num_sms, cudaStream_t stream);
int launch_conv3x3_x3(const void* x_hi, const void* x_lo, const void* w_img, const float* bias, float* y, float* partials, int splits, int n,
                      int h, int w, int c_in, int c_out, int relu, int f16, int num_sms, int* err_flag, cudaStream_t stream);
int conv3x3_splits(int n, int h, int w, int c_in, int c_out, int num_sms);
int launch_decoder_bwd(const uorf_decoder_weights* w, const uorf_decoder_grads* g, const uorf_decoder_bwd_args* a, int num_sms,
                       int* err_flag, cudaStream_t stream);
}  // namespace uorf

namespace {

thread_local char g_err[512] = "";
thread_local char g_cuda_err[256] = "";      // text of the CUDA error finish_launch() saw (cudaGetLastError clears it)
std::atomic<long long> g_launches{0};        // kernels launched by this library since load (uorf_launch_count)

// NVTX range around every launching entry point: names the C-ABI call in nsys / ncu timelines (no-op without a tool attached)
struct Range {
  explicit Range(const char* name) { nvtxRangePushA(name); }
  ~Range() { nvtxRangePop(); }
};

int fail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_err, sizeof(g_err), fmt, detail);
  return code;
}

struct DeviceCtx {
  std::once_flag once;
  int status = UORF_ERR_DEVICE;
  int num_sms = 0;
  int* err_flag_host = nullptr;  // pinned + mapped: still readable by the host after a device trap
  int* err_flag_dev = nullptr;
  char why[256] = "";
};
constexpr int kMaxDevices = 64;
DeviceCtx g_ctx[kMaxDevices];

void init_device(DeviceCtx* c, int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    snprintf(c->why, sizeof(c->why), "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    return;
  }
  if (prop.major != 10) {
    snprintf(c->why, sizeof(c->why), "device %d is sm_%d%d; libuorf_b200 runs on sm_100 (B200) only and has no fallback", device,
             prop.major, prop.minor);
    return;
  }
  c->num_sms = prop.multiProcessorCount;
  e = cudaHostAlloc(reinterpret_cast<void**>(&c->err_flag_host), sizeof(int), cudaHostAllocMapped);
  if (e == cudaSuccess) {
    *c->err_flag_host = 0;
    e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->err_flag_dev), c->err_flag_host, 0);
  }
  if (e != cudaSuccess) {
    snprintf(c->why, sizeof(c->why), "pinned flag allocation: %s", cudaGetErrorString(e));
    return;
  }
  c->status = UORF_OK;
}

// context of the CURRENT device (or error)
DeviceCtx* ctx() {
  int device = 0;
  if (cudaGetDevice(&device) != cudaSuccess || device < 0 || device >= kMaxDevices) {
    fail(UORF_ERR_DEVICE, "no current CUDA device%s");
    return nullptr;
  }
  DeviceCtx* c = &g_ctx[device];
  std::call_once(c->once, init_device, c, device);
  if (c->status != UORF_OK) {
    fail(c->status, "%s", c->why);
    return nullptr;
  }
  return c;
}

int check_launch(int rc, const char* what) {
  if (rc == UORF_OK) return rc;
  if (rc == UORF_ERR_DEVICE) {
    static thread_local char buf[400];
    if (g_cuda_err[0]) {                       // recorded by finish_launch (which consumed the CUDA error)
      snprintf(buf, sizeof(buf), "%s: %s", what, g_cuda_err);
      g_cuda_err[0] = 0;
    } else {                                   // e.g. a failed cudaFuncSetAttribute: the error is still pending
      snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(cudaGetLastError()));
    }
    return fail(rc, "%s", buf);
  }
  return fail(rc, "%s: unsupported shape", what);
}

}  // namespace

namespace uorf {
int finish_launch(int n_kernels) {
  g_launches.fetch_add(n_kernels, std::memory_order_relaxed);
  const cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return UORF_OK;
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s (%s)", cudaGetErrorString(e), cudaGetErrorName(e));
  return UORF_ERR_DEVICE;
}
}  // namespace uorf

extern "C" {

int uorf_abi_version(void) { return UORF_ABI_VERSION; }
const char* uorf_last_error(void) { return g_err; }
int64_t uorf_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int uorf_init(int device) {
  if (device < 0 || device >= kMaxDevices) return fail(UORF_ERR_ARG, "device index out of range%s");
  DeviceCtx* c = &g_ctx[device];
  std::call_once(c->once, init_device, c, device);
  if (c->status != UORF_OK) return fail(c->status, "%s", c->why);
  return UORF_OK;
}

static int split_pool_common(const char* what, const float* x, const float* g, void* hi, void* lo, int32_t n, int32_t h, int32_t w, int32_t c,
                             int32_t precision, void* stream) {
  DeviceCtx* dc = ctx();
  Range nvtx_range(what);
  if (!dc) return UORF_ERR_DEVICE;
  if (!x || !hi || !lo) return fail(UORF_ERR_ARG, "uorf_conv_split_pool2x2 / _unpool2x2: null pointer%s");
  if (n < 0 || h < 2 || w < 2 || c < 2 || ((h | w | c) & 1)) return fail(UORF_ERR_ARG, "uorf_conv_split_pool2x2 / _unpool2x2: h, w, c must be even%s");
  if (precision != UORF_PREC_FP16X3 && precision != UORF_PREC_BF16X3)
    return fail(UORF_ERR_ARG, "uorf_conv_split_pool2x2 / _unpool2x2: precision must be a x3 mode%s");
  return check_launch(uorf::launch_conv_split_pool(x, g, hi, lo, n, h, w, c, precision == UORF_PREC_FP16X3, dc->num_sms,
                                                   static_cast<cudaStream_t>(stream)), what);
}

int uorf_conv_split_pool2x2(const float* x, void* hi, void* lo, int32_t n, int32_t h, int32_t w, int32_t c, int32_t precision, void* stream) {
  return split_pool_common("uorf_conv_split_pool2x2", x, nullptr, hi, lo, n, h, w, c, precision, stream);
}

int uorf_conv_split_unpool2x2(const float* grad_pooled, const float* x, void* hi, void* lo, int32_t n, int32_t h, int32_t w, int32_t c,
                              int32_t precision, void* stream) {
  if (!grad_pooled) return fail(UORF_ERR_ARG, "uorf_conv_split_unpool2x2: null pointer%s");
  return split_pool_common("uorf_conv_split_unpool2x2", x, grad_pooled, hi, lo, n, h, w, c, precision, stream);
}

int uorf_conv3x3_stem(const float* x, const float* weight, const float* bias, int32_t n, int32_t h, int32_t w, int32_t c_out, float* y,
                      void* stream) {
  DeviceCtx* c = ctx();
  Range nvtx_range("uorf_conv3x3_stem");
  if (!c) return UORF_ERR_DEVICE;
  if (!x || !weight || !bias || !y) return fail(UORF_ERR_ARG, "uorf_conv3x3_stem: null pointer%s");
  if (n < 0 || h < 1 || w < 1 || c_out < 4 || c_out > 64 || (c_out & 3)) return fail(UORF_ERR_ARG, "uorf_conv3x3_stem: bad sizes (c_out: multiple of 4, <= 64)%s");
  return check_launch(uorf::launch_vgg_stem_fwd(x, weight, bias, n, h, w, c_out, y, static_cast<cudaStream_t>(stream)), "uorf_conv3x3_stem");
}

int uorf_conv3x3_stem_dgrad(const float* grad_y, const float* y, const float* weight, int32_t n, int32_t h, int32_t w, int32_t c_out,
                            float* grad_x, void* stream) {
  DeviceCtx* c = ctx();
  Range nvtx_range("uorf_conv3x3_stem_dgrad");
  if (!c) return UORF_ERR_DEVICE;
  if (!grad_y || !y || !weight || !grad_x) return fail(UORF_ERR_ARG, "uorf_conv3x3_stem_dgrad: null pointer%s");
  if (n < 0 || h < 1 || w < 1 || c_out < 1 || c_out > 64) return fail(UORF_ERR_ARG, "uorf_conv3x3_stem_dgrad: bad sizes (c_out <= 64)%s");
  return check_launch(uorf::launch_vgg_stem_dgrad(grad_y, y, weight, n, h, w, c_out, grad_x, static_cast<cudaStream_t>(stream)),
                      "uorf_conv3x3_stem_dgrad");
}

int uorf_invert_small(const float* src, float* dst, int32_t batch, int32_t dim, void* stream) {
  DeviceCtx* c = ctx();
  Range nvtx_range("uorf_invert_small");
  if (!c) return UORF_ERR_DEVICE;
  if (batch < 0 || (batch > 0 && (!src || !dst))) return fail(UORF_ERR_ARG, "uorf_invert_small: null pointer%s");
  if (dim != 3 && dim != 4) return fail(UORF_ERR_ARG, "uorf_invert_small: dim must be 3 or 4%s");
  return check_launch(uorf::launch_invert_small(src, dst, batch, dim, static_cast<cudaStream_t>(stream)), "uorf_invert_small");
}

int uorf_image_loss_forward(const float* rgb_map, const float* x, const float* mean, const float* stdv, int32_t n, int32_t hw, float* x_recon,
                            float* rendered, float* rendered_norm, float* x_norm, float* scratch, float* loss_recon, void* stream) {
  DeviceCtx* c = ctx();
  Range nvtx_range("uorf_image_loss_forward");
  if (!c) return UORF_ERR_DEVICE;
  if (!rgb_map || !x || !mean || !stdv || !x_recon || !rendered || !rendered_norm || !x_norm || !scratch || !loss_recon)
    return fail(UORF_ERR_ARG, "uorf_image_loss_forward: null pointer%s");
  if (n < 1 || hw < 1) return fail(UORF_ERR_ARG, "uorf_image_loss_forward: bad sizes%s");
  return check_launch(uorf::launch_image_loss_fwd(rgb_map, x, mean, stdv, n, hw, x_recon, rendered, rendered_norm, x_norm, scratch, loss_recon,
                                                  static_cast<cudaStream_t>(stream)), "uorf_image_loss_forward");
}

int uorf_image_loss_backward(const float* x_recon, const float* x, const float* stdv, const float* grad_x_recon, const float* grad_rendered_norm,
                             const float* grad_loss, int32_t n, int32_t hw, float* grad_rgb_map, void* stream) {
  DeviceCtx* c = ctx();
  Range nvtx_range("uorf_image_loss_backward");
  if (!c) return UORF_ERR_DEVICE;
  if (!x_recon || !x || !stdv || !grad_rgb_map) return fail(UORF_ERR_ARG, "uorf_image_loss_backward: null pointer%s");
  if (n < 1 || hw < 1) return fail(UORF_ERR_ARG, "uorf_image_loss_backward: bad sizes%s");
  return check_launch(uorf::launch_image_loss_bwd(x_recon, x, stdv, grad_x_recon, grad_rendered_norm, grad_loss, n, hw, grad_rgb_map,
                                                  static_cast<cudaStream_t>(stream)), "uorf_image_loss_backward");
}

int uorf_feature_mse_forward(const float* features, int64_t m, float* scratch, float* loss, void* stream) {
  DeviceCtx* c = ctx();
  Range nvtx_range("uorf_feature_mse_forward");
  if (!c) return UORF_ERR_DEVICE;
  if (!features || !scratch || !loss || m < 1) return fail(UORF_ERR_ARG, "uorf_feature_mse_forward: bad arguments%s");
  return check_launch(uorf::launch_feat_mse_fwd(features, m, scratch, loss, static_cast<cudaStream_t>(stream)), "uorf_feature_mse_forward");
}

int uorf_feature_mse_backward(const float* features, const float* grad_loss, int64_t m, float* grad_features, void* stream) {
  DeviceCtx* c = ctx();
  Range nvtx_range("uorf_feature_mse_backward");
  if (!c) return UORF_ERR_DEVICE;
  if (!features || !grad_loss || !grad_features || m < 1) return fail(UORF_ERR_ARG, "uorf_feature_mse_backward: bad arguments%s");
  return check_launch(uorf::launch_feat_mse_bwd(features, grad_loss, m, grad_features, static_cast<cudaStream_t>(stream)),
                      "uorf_feature_mse_backward");
}

int uorf_adam_step(const uorf_adam_chunk* chunks, int32_t n_chunks, const float* lr_dev, float lr_host, float beta1, float beta2, float eps,
                   float weight_decay, void* stream) {
  DeviceCtx* c = ctx();
  Range nvtx_range("uorf_adam_step");
  if (!c) return UORF_ERR_DEVICE;
  if (n_chunks < 0 || (n_chunks > 0 && !chunks)) return fail(UORF_ERR_ARG, "uorf_adam_step: bad chunk table%s");
  if (!(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f)) return fail(UORF_ERR_ARG, "uorf_adam_step: bad hyper-parameters%s");
  return check_launch(uorf::launch_adam_step(chunks, n_chunks, lr_dev, lr_host, beta1, beta2, eps, weight_decay, static_cast<cudaStream_t>(stream)),
                      "uorf_adam_step");
}

int uorf_encoder_weight_prep(const uorf_encoder_prep_layer* layers, int32_t n_layers, void* stream) {
  DeviceCtx* c = ctx();
  Range nvtx_range("uorf_encoder_weight_prep");
  if (!c) return UORF_ERR_DEVICE;
  if (n_layers < 0 || n_layers > 8 || (n_layers > 0 && !layers)) return fail(UORF_ERR_ARG, "uorf_encoder_weight_prep: 0..8 layers%s");
  for (int i = 0; i < n_layers; ++i)
    if (!layers[i].weight || layers[i].c_out < 1 || layers[i].c_in < 1) return fail(UORF_ERR_ARG, "uorf_encoder_weight_prep: bad layer%s");
  return check_launch(uorf::launch_encoder_weight_prep(layers, n_layers, static_cast<cudaStream_t>(stream)), "uorf_encoder_weight_prep");
}

int uorf_encoder_conv3x3_dgrad(const float* grad_out, const float* kept_out, int32_t c_out, const float* weight_g, int32_t c_in,
                               int32_t h_in, int32_t w_in, int32_t stride, int32_t accumulate, float* grad_in, void* stream) {
  DeviceCtx* c = ctx();
  Range nvtx_range("uorf_encoder_conv3x3_dgrad");
  if (!c) return UORF_ERR_DEVICE;
  if (!grad_out || !kept_out || !weight_g || !grad_in) return fail(UORF_ERR_ARG, "uorf_encoder_conv3x3_dgrad: null pointer%s");
  if (c_out < 1 || c_in < 1 || h_in < 1 || w_in < 1 || (stride != 1 && stride != 2))
    return fail(UORF_ERR_ARG, "uorf_encoder_conv3x3_dgrad: bad sizes (stride must be 1 or 2)%s");
  return check_launch(uorf::launch_encoder_dgrad3x3(grad_out, kept_out, c_out, weight_g, c_in, h_in, w_in, stride, accumulate != 0, grad_in,
                                                    static_cast<cudaStream_t>(stream)), "uorf_encoder_conv3x3_dgrad");
}

int uorf_encoder_upsample2x_bwd(const float* grad_hi, int32_t c, int32_t h_lo, int32_t w_lo, float* grad_lo, void* stream) {
  DeviceCtx* dc = ctx();
  Range nvtx_range("uorf_encoder_upsample2x_bwd");
  if (!dc) return UORF_ERR_DEVICE;
  if (!grad_hi || !grad_lo) return fail(UORF_ERR_ARG, "uorf_encoder_upsample2x_bwd: null pointer%s");
  if (c < 1 || h_lo < 1 || w_lo < 1) return fail(UORF_ERR_ARG, "uorf_encoder_upsample2x_bwd: bad sizes%s");
  return check_launch(uorf::launch_encoder_upsample2x_bwd(grad_hi, c, h_lo, w_lo, grad_lo, static_cast<cudaStream_t>(stream)),
                      "uorf_encoder_upsample2x_bwd");
}

int64_t uorf_encoder_wgrad_workspace_floats(int32_t c_in, int32_t c_out, int32_t h_in, int32_t w_in, int32_t stride) {
  if (c_in < 1 || c_out < 1 || h_in < 1 || w_in < 1 || (stride != 1 && stride != 2)) return 0;
  return (int64_t)uorf::encoder_wgrad_workspace_floats(c_in, c_out, h_in, w_in, stride);
}

int uorf_encoder_conv3x3_wgrad(const float* src_a, int32_t ca, int32_t upsample_a, const float* src_b, int32_t cb, const float* grad_out,
                               const float* kept_out, int32_t c_out, int32_t h_in, int32_t w_in, int32_t stride, float* workspace,
                               float* grad_weight, float* grad_bias, void* stream) {
  DeviceCtx* c = ctx();
  Range nvtx_range("uorf_encoder_conv3x3_wgrad");
  if (!c) return UORF_ERR_DEVICE;
  if (!src_a || !grad_out || !kept_out || !workspace || !grad_weight || !grad_bias || (cb > 0 && !src_b))
    return fail(UORF_ERR_ARG, "uorf_encoder_conv3x3_wgrad: null pointer%s");
  if (ca < 1 || cb < 0 || c_out < 1 || h_in < 1 || w_in < 1 || (stride != 1 && stride != 2))
    return fail(UORF_ERR_ARG, "uorf_encoder_conv3x3_wgrad: bad sizes (stride must be 1 or 2)%s");
  if (upsample_a && ((h_in | w_in) & 1)) return fail(UORF_ERR_ARG, "uorf_encoder_conv3x3_wgrad: upsampled source needs even h_in, w_in%s");
  return check_launch(uorf::launch_encoder_wgrad3x3(src_a, ca, upsample_a, src_b, cb, grad_out, kept_out, c_out, h_in, w_in, stride,
                                                    workspace, grad_weight, grad_bias, static_cast<cudaStream_t>(stream)),
                      "uorf_encoder_conv3x3_wgrad");
}

int uorf_encoder_conv3x3(const float* src_a, int32_t ca, int32_t upsample_a, const float* src_b, int32_t cb, const float* weight_t,
                         const float* bias, int32_t c_out, int32_t h_in, int32_t w_in, int32_t stride, float* out, void* stream) {
  DeviceCtx* c = ctx();
  Range nvtx_range("uorf_encoder_conv3x3");
  if (!c) return UORF_ERR_DEVICE;
  if (!src_a || !weight_t || !bias || !out || (cb > 0 && !src_b)) return fail(UORF_ERR_ARG, "uorf_encoder_conv3x3: null pointer%s");
  if (ca < 1 || cb < 0 || c_out < 1 || h_in < 1 || w_in < 1 || (stride != 1 && stride != 2))
    return fail(UORF_ERR_ARG, "uorf_encoder_conv3x3: bad sizes (stride must be 1 or 2)%s");
  if (upsample_a && ((h_in | w_in) & 1)) return fail(UORF_ERR_ARG, "uorf_encoder_conv3x3: upsampled source needs even h_in, w_in%s");
  return check_launch(uorf::launch_encoder_conv3x3(src_a, ca, upsample_a, src_b, cb, weight_t, bias, c_out, h_in, w_in, stride, out,
                                                   static_cast<cudaStream_t>(stream)), "uorf_encoder_conv3x3");
}

int uorf_conv_split(const float* x, const float* gate, void* hi, void* lo, int64_t n, int32_t precision, void* stream) {
  DeviceCtx* c = ctx();
  Range nvtx_range("uorf_conv_split");
  if (!c) return UORF_ERR_DEVICE;
  if (n < 0 || (n & 1)) return fail(UORF_ERR_ARG, "uorf_conv_split: n must be even and >= 0%s");
  if (n > 0 && (!x || !hi || !lo)) return fail(UORF_ERR_ARG, "uorf_conv_split: null pointer%s");
  if (precision != UORF_PREC_FP16X3 && precision != UORF_PREC_BF16X3) return fail(UORF_ERR_ARG, "uorf_conv_split: precision must be a x3 mode%s");
  return check_launch(uorf::launch_conv_split(x, gate, hi, lo, n, precision == UORF_PREC_FP16X3, c->num_sms, static_cast<cudaStream_t>(stream)),
                      "uorf_conv_split");
}

size_t uorf_conv3x3_workspace_floats(int32_t n, int32_t h, int32_t w, int32_t c_in, int32_t c_out) {
  DeviceCtx* c = ctx();
  if (!c || n < 1 || h < 1 || w < 1 || c_in < 64 || c_out < 64) return 0;
  const int s = uorf::conv3x3_splits(n, h, w, c_in, c_out, c->num_sms);
  return s > 1 ? (size_t)s * n * h * w * c_out : 0;
}

int uorf_conv3x3_x3(const void* x_hi, const void* x_lo, const void* w_img, const float* bias, float* y, float* workspace, int32_t n, int32_t h,
                    int32_t w, int32_t c_in, int32_t c_out, int32_t relu, int32_t precision, void* stream) {
  DeviceCtx* c = ctx();
  Range nvtx_range("uorf_conv3x3_x3");
  if (!c) return UORF_ERR_DEVICE;
  if (!x_hi || !x_lo || !w_img || !y) return fail(UORF_ERR_ARG, "uorf_conv3x3_x3: null pointer%s");
  if (n < 0 || h < 1 || w < 1 || c_in < 64 || c_out < 64 || (c_in & 63) || (c_out & 63))
    return fail(UORF_ERR_ARG, "uorf_conv3x3_x3: bad sizes (c_in, c_out must be multiples of 64)%s");
  if (precision != UORF_PREC_FP16X3 && precision != UORF_PREC_BF16X3) return fail(UORF_ERR_ARG, "uorf_conv3x3_x3: precision must be a x3 mode%s");
  const int splits = workspace ? uorf::conv3x3_splits(n, h, w, c_in, c_out, c->num_sms) : 1;
  return check_launch(uorf::launch_conv3x3_x3(x_hi, x_lo, w_img, bias, y, workspace, splits, n, h, w, c_in, c_out, relu,
                                              precision == UORF_PREC_FP16X3, c->num_sms, c->err_flag_dev, static_cast<cudaStream_t>(stream)),
                      "uorf_conv3x3_x3");
}

int uorf_debug_flag(void) {
  DeviceCtx* c = ctx();
  return c && c->err_flag_host ? *c->err_flag_host : -1;
}

static int sample_coords_impl(const char* what, const uorf_frustum* f, const float* z_cam, const float* cam2world, int32_t n_views,
                              int32_t scale, int32_t crop_h, int32_t crop_w, const int32_t* crop_origin, int32_t Hc, int32_t Wc,
                              int32_t ray_major, float* coords, float* z_vals, float* ray_dir, void* stream) {
  DeviceCtx* c = ctx();
  if (!c) return UORF_ERR_DEVICE;
  Range nvtx_range(what);
  if (!f || !z_cam || !cam2world || !coords) return fail(UORF_ERR_ARG, "%s: null pointer", what);
  if (f->W <= 0 || f->H <= 0 || f->D <= 0 || n_views < 0 || scale < 1 || Hc <= 0 || Wc <= 0) return fail(UORF_ERR_ARG, "%s: bad sizes", what);
  if (f->W != f->H) return fail(UORF_ERR_UNSUPPORTED, "%s: W != H (the reference's ray_dir view needs a square frustum)", what);
  const bool crop = crop_origin != nullptr || crop_h >= 0 || crop_w >= 0;
  if (crop_origin != nullptr) {
    if (scale != 1 || Hc > f->H || Wc > f->W) return fail(UORF_ERR_ARG, "%s: bad crop window", what);
    crop_h = crop_w = 0;   // the kernel reads the origin from the device and clamps it to [0, H-Hc] x [0, W-Wc]
  } else if (crop) {
    if (scale != 1 || crop_h < 0 || crop_w < 0 || crop_h + Hc > f->H || crop_w + Wc > f->W) return fail(UORF_ERR_ARG, "%s: bad crop window", what);
  } else if (Hc * scale != f->H || Wc * scale != f->W) {
    return fail(UORF_ERR_ARG, "%s: Hc*scale != H or Wc*scale != W", what);
  }
  return check_launch(uorf::launch_sample_coords(f, z_cam, cam2world, n_views, scale, crop ? crop_h : -1, crop ? crop_w : -1, crop_origin, Hc, Wc,
                                                 ray_major, coords, z_vals, ray_dir, c->num_sms, static_cast<cudaStream_t>(stream)), what);
}

int uorf_sample_coords(const uorf_frustum* f, const float* z_cam, const float* cam2world, int32_t n_views, int32_t scale,
                       int32_t crop_h, int32_t crop_w, int32_t Hc, int32_t Wc, int32_t ray_major, float* coords, float* z_vals,
                       float* ray_dir, void* stream) {
  return sample_coords_impl("uorf_sample_coords", f, z_cam, cam2world, n_views, scale, crop_h, crop_w, nullptr, Hc, Wc, ray_major, coords,
                            z_vals, ray_dir, stream);
}

int uorf_sample_coords_window(const uorf_frustum* f, const float* z_cam, const float* cam2world, int32_t n_views,
                              const int32_t* crop_origin, int32_t Hc, int32_t Wc, int32_t ray_major, float* coords, float* z_vals,
                              float* ray_dir, void* stream) {
  if (!crop_origin) return fail(UORF_ERR_ARG, "%s: null crop_origin", "uorf_sample_coords_window");
  return sample_coords_impl("uorf_sample_coords_window", f, z_cam, cam2world, n_views, 1, 0, 0, crop_origin, Hc, Wc, ray_major, coords,
                            z_vals, ray_dir, stream);
}

static bool prec_ok(int precision) { return precision >= 1 && precision <= 4; }
static int prec_passes(int precision) { return precision >= 3 ? 3 : 1; }
static int prec_f16(int precision) { return (precision % 2) == 0 ? 1 : 0; }

size_t uorf_decoder_image_bytes(int32_t K, int32_t precision) {
  if (K < 2 || !prec_ok(precision)) return 0;
  return (size_t)uorf::blob_bytes_total(prec_passes(precision), K);
}

int uorf_decoder_prepare(const uorf_decoder_weights* w, const float* z_slots, int32_t K, int32_t Z, int32_t n_freq,
                         int32_t n_layers, int32_t precision, void* image, void* stream) {
  DeviceCtx* c = ctx();
  Range nvtx_range("uorf_decoder_prepare");
  if (!c) return UORF_ERR_DEVICE;
  if (!w || !z_slots || !image) return fail(UORF_ERR_ARG, "uorf_decoder_prepare: null pointer%s");
  if ((reinterpret_cast<uintptr_t>(image) & 1023) != 0) return fail(UORF_ERR_ARG, "uorf_decoder_prepare: image must be 1024-byte aligned%s");
  if (!prec_ok(precision)) return fail(UORF_ERR_ARG, "uorf_decoder_prepare: precision must be one of UORF_PREC_*%s");
  if (K < 2 || K > 16) return fail(UORF_ERR_UNSUPPORTED, "uorf_decoder_prepare: K must be in [2,16]%s");
  if (Z < 4 || Z > uorf::kZPad || (Z % 4) != 0) return fail(UORF_ERR_UNSUPPORTED, "uorf_decoder_prepare: z_dim must be a multiple of 4, <= 64%s");
  if (n_freq != 5 || n_layers != 3) return fail(UORF_ERR_UNSUPPORTED, "uorf_decoder_prepare: only n_freq=5, n_layers=3 (all shipped configs)%s");
  return check_launch(uorf::launch_decoder_prep(w, z_slots, K, Z, prec_passes(precision), prec_f16(precision), image, static_cast<cudaStream_t>(stream)),
                      "uorf_decoder_prepare");
}

size_t uorf_decoder_forward_workspace_floats(int64_t P, int32_t K) {
  DeviceCtx* c = ctx();
  if (!c || P < 0 || K < 2) return 0;
  return uorf::decoder_fwd_workspace_floats(P, K, c->num_sms);
}

int uorf_decoder_forward(const uorf_decoder_args* a, void* stream) {
  DeviceCtx* c = ctx();
  Range nvtx_range("uorf_decoder_forward");
  if (!c) return UORF_ERR_DEVICE;
  if (!a || !a->coor_bg || !a->coor_fg || !a->fg_transform || !a->image || !a->workspace)
    return fail(UORF_ERR_ARG, "uorf_decoder_forward: null pointer%s");
  if (!prec_ok(a->precision)) return fail(UORF_ERR_ARG, "uorf_decoder_forward: precision must be one of UORF_PREC_*%s");
  if (a->K < 2 || a->K > 16) return fail(UORF_ERR_UNSUPPORTED, "uorf_decoder_forward: K must be in [2,16]%s");
  if (a->P < 0 || a->P > (1LL << 37)) return fail(UORF_ERR_ARG, "uorf_decoder_forward: bad P%s");
  if (a->P == 0) return UORF_OK;
  if (a->dens_noise != 0.f && !a->noise) return fail(UORF_ERR_ARG, "uorf_decoder_forward: dens_noise != 0 needs a noise tensor%s");
  const uintptr_t al = reinterpret_cast<uintptr_t>(a->workspace) | reinterpret_cast<uintptr_t>(a->raws) |
                       reinterpret_cast<uintptr_t>(a->masked_raws) | reinterpret_cast<uintptr_t>(a->unmasked_raws) |
                       reinterpret_cast<uintptr_t>(a->image);
  if (al & 15) return fail(UORF_ERR_ARG, "uorf_decoder_forward: buffers must be 16-byte aligned%s");
  return check_launch(uorf::launch_decoder_fwd(a, c->num_sms, c->err_flag_dev, static_cast<cudaStream_t>(stream)),
                      "uorf_decoder_forward");
}

size_t uorf_decoder_backward_workspace_floats(int64_t P, int32_t K) {
  DeviceCtx* c = ctx();
  if (!c || P < 0 || K < 2) return 0;
  return uorf::decoder_bwd_workspace_floats(P, K, c->num_sms);
}

int uorf_decoder_backward(const uorf_decoder_weights* w, const uorf_decoder_grads* g, const uorf_decoder_bwd_args* a, void* stream) {
  DeviceCtx* c = ctx();
  Range nvtx_range("uorf_decoder_backward");
  if (!c) return UORF_ERR_DEVICE;
  if (!w || !g || !a || !a->coor_bg || !a->coor_fg || !a->fg_transform || !a->image || !a->workspace || !a->unmasked_raws ||
      !a->grad_raws || !a->z_slots || !a->grad_z_slots)
    return fail(UORF_ERR_ARG, "uorf_decoder_backward: null pointer%s");
  if (a->K < 2 || a->K > 16) return fail(UORF_ERR_UNSUPPORTED, "uorf_decoder_backward: K must be in [2,16]%s");
  if (a->Z < 4 || a->Z > uorf::kZPad || (a->Z % 4) != 0) return fail(UORF_ERR_UNSUPPORTED, "uorf_decoder_backward: bad z_dim%s");
  if (a->P <= 0) return fail(UORF_ERR_ARG, "uorf_decoder_backward: bad P%s");
  if (a->dens_noise != 0.f && !a->noise) return fail(UORF_ERR_ARG, "uorf_decoder_backward: dens_noise != 0 needs the noise tensor%s");
  const uintptr_t al = reinterpret_cast<uintptr_t>(a->workspace) | reinterpret_cast<uintptr_t>(a->unmasked_raws) |
                       reinterpret_cast<uintptr_t>(a->grad_raws) | reinterpret_cast<uintptr_t>(a->image);
  if (al & 15) return fail(UORF_ERR_ARG, "uorf_decoder_backward: buffers must be 16-byte aligned%s");
  return check_launch(uorf::launch_decoder_bwd(w, g, a, c->num_sms, c->err_flag_dev, static_cast<cudaStream_t>(stream)),
                      "uorf_decoder_backward");
}

int uorf_composite_forward(const float* raw, const float* z_vals, const float* rays_d, int64_t R, int64_t geom_rays, int32_t D, float* rgb_map,
                           float* depth_map, float* weights_norm, float* mask_map, void* stream) {
  DeviceCtx* c = ctx();
  Range nvtx_range("uorf_composite_forward");
  if (!c) return UORF_ERR_DEVICE;
  if (R == 0) return UORF_OK;
  if (!raw || !z_vals || !rays_d || !rgb_map || !depth_map) return fail(UORF_ERR_ARG, "uorf_composite_forward: null pointer%s");
  if (geom_rays <= 0) geom_rays = R;
  if (R < 0 || D < 1 || (R > 0 && R % geom_rays != 0)) return fail(UORF_ERR_ARG, "uorf_composite_forward: bad sizes%s");
  if (D > 256) return fail(UORF_ERR_UNSUPPORTED, "uorf_composite_forward: D > 256%s");
  if (reinterpret_cast<uintptr_t>(raw) & 15) return fail(UORF_ERR_ARG, "uorf_composite_forward: raw must be 16-byte aligned%s");
  return check_launch(uorf::launch_composite_fwd(raw, z_vals, rays_d, R, geom_rays > 0 ? geom_rays : 1, D, rgb_map, depth_map, weights_norm, mask_map, c->num_sms,
                                                 static_cast<cudaStream_t>(stream)),
                      "uorf_composite_forward");
}

int uorf_composite_backward(const float* raw, const float* z_vals, const float* rays_d, const float* grad_rgb, int64_t R, int64_t geom_rays, int32_t D,
                            float* grad_raw, void* stream) {
  DeviceCtx* c = ctx();
  Range nvtx_range("uorf_composite_backward");
  if (!c) return UORF_ERR_DEVICE;
  if (R == 0) return UORF_OK;
  if (!raw || !z_vals || !rays_d || !grad_rgb || !grad_raw) return fail(UORF_ERR_ARG, "uorf_composite_backward: null pointer%s");
  if (geom_rays <= 0) geom_rays = R;
  if (R < 0 || D < 1 || (R > 0 && R % geom_rays != 0)) return fail(UORF_ERR_ARG, "uorf_composite_backward: bad sizes%s");
  if (D > 256) return fail(UORF_ERR_UNSUPPORTED, "uorf_composite_backward: D > 256%s");
  if ((reinterpret_cast<uintptr_t>(raw) | reinterpret_cast<uintptr_t>(grad_raw)) & 15)
    return fail(UORF_ERR_ARG, "uorf_composite_backward: raw/grad_raw must be 16-byte aligned%s");
  return check_launch(uorf::launch_composite_bwd(raw, z_vals, rays_d, grad_rgb, R, geom_rays > 0 ? geom_rays : 1, D, grad_raw, c->num_sms,
                                                 static_cast<cudaStream_t>(stream)),
                      "uorf_composite_backward");
}

size_t uorf_slot_attention_workspace_floats(int32_t n_tokens, int32_t C, int32_t K) {
  return uorf::slot_attention_ws_floats(n_tokens, C, K);
}

int uorf_slot_attention_forward(const uorf_slot_attention_weights* w, const float* feat, int64_t feat_token_stride,
                                int64_t feat_chan_stride, const float* noise_fg, const float* noise_bg, int32_t n_tokens,
                                int32_t C, int32_t K, int32_t hidden, int32_t iters, float eps, float* workspace, float* slots,
                                float* attn, void* stream) {
  DeviceCtx* c = ctx();
  Range nvtx_range("uorf_slot_attention_forward");
  if (!c) return UORF_ERR_DEVICE;
  if (!w || !feat || !noise_fg || !noise_bg || !workspace || !slots || !attn)
    return fail(UORF_ERR_ARG, "uorf_slot_attention_forward: null pointer%s");
  if (n_tokens < 1 || iters < 1) return fail(UORF_ERR_ARG, "uorf_slot_attention_forward: bad sizes%s");
  return check_launch(uorf::launch_slot_attention(w, feat, feat_token_stride, feat_chan_stride, noise_fg, noise_bg, n_tokens, C, K,
                                                  hidden, iters, eps, workspace, slots, attn, static_cast<cudaStream_t>(stream)),
                      "uorf_slot_attention_forward");
}

size_t uorf_slot_attention_backward_workspace_floats(int32_t n_tokens, int32_t C, int32_t K, int32_t hidden) {
  return uorf::slot_attention_bwd_ws_floats(n_tokens, C, K, hidden);
}

int uorf_slot_attention_backward(const uorf_slot_attention_weights* w, const uorf_slot_attention_grads* g, const float* feat,
                                 int64_t feat_token_stride, int64_t feat_chan_stride, const float* noise_fg, const float* noise_bg,
                                 int32_t n_tokens, int32_t C, int32_t K, int32_t hidden, int32_t iters, float eps,
                                 const float* fwd_workspace, const float* grad_slots, float* workspace, float* grad_feat, void* stream) {
  DeviceCtx* c = ctx();
  Range nvtx_range("uorf_slot_attention_backward");
  if (!c) return UORF_ERR_DEVICE;
  if (!w || !g || !feat || !noise_fg || !noise_bg || !fwd_workspace || !grad_slots || !workspace || !grad_feat)
    return fail(UORF_ERR_ARG, "uorf_slot_attention_backward: null pointer%s");
  if (n_tokens < 1 || iters < 1) return fail(UORF_ERR_ARG, "uorf_slot_attention_backward: bad sizes%s");
  return check_launch(uorf::launch_slot_attention_bwd(w, g, feat, feat_token_stride, feat_chan_stride, noise_fg, noise_bg, n_tokens, C, K,
                                                      hidden, iters, eps, fwd_workspace, grad_slots, workspace, grad_feat,
                                                      static_cast<cudaStream_t>(stream)),
                      "uorf_slot_attention_backward");
}

}  // extern "C"
