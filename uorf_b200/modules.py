"""Host-side mirror of the reference's hot-path module API (reference models/model.py, models/projection.py).

Same class / function names, constructor arguments, forward signatures, return shapes and state-dict keys as
the reference, so a reference driver can swap its imports for these.  All arithmetic below the module
boundary runs in libuorf_b200.so (hand-written sm_100a kernels, C ABI in include/uorf_b200.h); PyTorch is used
for device memory, streams and autograd plumbing only.  There is no CPU / eager fallback: tensors must live on
a CUDA device and the shared object must be loadable, otherwise the call raises.

    Encoder            models/model.py:11-61     -> uorf_encoder_conv3x3 (one fp32 image; other shapes: the torch layer stack)
    SlotAttention      models/model.py:164-258   -> uorf_slot_attention_forward
    Decoder            models/model.py:64-161    -> uorf_decoder_prepare + uorf_decoder_forward
    sin_emb            models/model.py:261-277   (fused into the decoder kernel; standalone version for callers)
    raw2outputs        models/model.py:279-313   -> uorf_composite_forward / _backward
    Projection         models/projection.py:6-113 -> uorf_sample_coords
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import torch
from torch import nn
import torch.nn.functional as F
from torch.nn import init

from . import _lib
from ._lib import lib, check


def _stream(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


_NULL_GUARD = contextlib.nullcontext()


def _on(t: torch.Tensor):
    """Device guard for a C-ABI call: the library takes its context (SM count, error flag) and launches on the CURRENT CUDA
    device (csrc/abi.cu ctx()), while streams and pointers come from the tensor's device.  When they differ (opt.gpu_ids=[n]
    in a process whose current device is another one) the call runs under torch.cuda.device(t.device), like the
    at::DeviceGuard of the reference's own native ops (models/op/fused_bias_act.cpp:25)."""
    return _NULL_GUARD if t.device.index == torch.cuda.current_device() else torch.cuda.device(t.device)


def _require_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.UorfError("uorf_b200 modules run on CUDA tensors only (there is no CPU fallback)")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _grad_buffers(named_params, sink):
    """Where a backward kernel writes its parameter gradients (each is written in full, never accumulated).
    Default: fresh tensors handed to autograd.  With `module.grad_sink = {name: tensor}` (the data-parallel driver points
    every `.grad` at a slice of ONE flat buffer) the kernels write straight into the caller's storage and autograd receives
    None for the parameters: the step then all-reduces that buffer with a single NCCL call, no gather / scatter copies."""
    if sink is None:
        return {n: torch.empty_like(p_, dtype=torch.float32, memory_format=torch.contiguous_format) for n, p_ in named_params.items()}
    for n, p_ in named_params.items():
        t = sink[n]
        if t.shape != p_.shape or t.dtype != torch.float32 or not t.is_contiguous() or t.device != p_.device:
            raise _lib.UorfError(f"grad_sink[{n!r}] must be a contiguous fp32 tensor of shape {tuple(p_.shape)} on {p_.device}")
    return {n: sink[n] for n in named_params}


# ======================================================================================================
# Encoder — identical structure and state-dict keys.  Inference (no gradient, batch 1): direct-convolution kernels
# (csrc/encoder.cu) with torch.cat and the bilinear upsamples folded into the input read; training: torch / cuDNN autograd.
# ======================================================================================================
class Encoder(nn.Module):
    """reference models/model.py:11-61"""

    def __init__(self, input_nc=3, z_dim=64, bottom=False):
        super().__init__()
        self.bottom = bottom
        if self.bottom:
            self.enc_down_0 = nn.Sequential(nn.Conv2d(input_nc + 4, z_dim, 3, stride=1, padding=1), nn.ReLU(True))
        self.enc_down_1 = nn.Sequential(nn.Conv2d(z_dim if bottom else input_nc + 4, z_dim, 3,
                                                  stride=2 if bottom else 1, padding=1), nn.ReLU(True))
        self.enc_down_2 = nn.Sequential(nn.Conv2d(z_dim, z_dim, 3, stride=2, padding=1), nn.ReLU(True))
        self.enc_down_3 = nn.Sequential(nn.Conv2d(z_dim, z_dim, 3, stride=2, padding=1), nn.ReLU(True))
        self.enc_up_3 = nn.Sequential(nn.Conv2d(z_dim, z_dim, 3, stride=1, padding=1), nn.ReLU(True),
                                      nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False))
        self.enc_up_2 = nn.Sequential(nn.Conv2d(z_dim * 2, z_dim, 3, stride=1, padding=1), nn.ReLU(True),
                                      nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False))
        self.enc_up_1 = nn.Sequential(nn.Conv2d(z_dim * 2, z_dim, 3, stride=1, padding=1), nn.ReLU(True))
        self._planes = {}
        self._wt = {}              # transposed conv weights of the inference path, keyed by weight version
        self._train_wt = None      # (weight addresses, {layer: (forward layout, data-gradient layout)}, launch table) of the training path
        self.use_kernels = True    # False: always the torch / cuDNN layers
        self.use_kernels_in_training = True   # forward by the kernels under autograd too (_EncoderFn)
        # ... and the backward by the fp32 gradient kernels; False (UORF_ENCODER_BACKWARD=cudnn, A-B runs): cuDNN convolution_backward per layer
        self.backward_kernels = os.environ.get("UORF_ENCODER_BACKWARD", "kernels") != "cudnn"
        self.grad_sink = None                 # {parameter name: tensor}: the training path's backward writes gradients there (see _GradSink)
        self.require_kernels = False          # True: raise instead of running the torch layer stack on a non-kernel shape
        self.last_path = None                 # 'kernels' | 'torch': which path the last forward took

    def _pixel_planes(self, H, W, device):
        key = (H, W, str(device))
        if key not in self._planes:
            # built on the CPU exactly like the reference (model.py:43-48), then moved: bit-identical planes
            X = torch.linspace(-1, 1, W)
            Y = torch.linspace(-1, 1, H)
            y1, x1 = torch.meshgrid([Y, X], indexing="ij")
            self._planes[key] = torch.stack([x1, 2 - x1, y1, 2 - y1]).unsqueeze(0).to(device)
        return self._planes[key]

    # ---- inference path: six (seven) direct-convolution launches, no cat / upsample tensors -------------------------
    def invalidate_weight_cache(self):
        """drop the cached transposed weights: call after anything that rewrites the weights WITHOUT bumping their version
        counters - torch's fused Adam (`torch._fused_adam_`) and CUDA-graph replays of an optimizer step both do"""
        self._wt.clear()

    def _training_weights(self):
        """{layer name: (forward layout [in][3][3][out], data-gradient layout [out][3][3][in])} of the CURRENT weights, written by one
        launch into buffers that stay (graph-capturable; the training step calls this once, before its forward)"""
        names = (["enc_down_0"] if self.bottom else []) + ["enc_down_1", "enc_down_2", "enc_down_3", "enc_up_3", "enc_up_2", "enc_up_1"]
        convs = [getattr(self, n)[0] for n in names]
        dev = convs[0].weight.device
        if any(c.weight.dtype != torch.float32 or not c.weight.is_contiguous() for c in convs):
            raise _lib.UorfError("Encoder kernels need contiguous fp32 weights")
        key = tuple(c.weight.data_ptr() for c in convs)
        if self._train_wt is None or self._train_wt[0] != key:
            bufs = {n: (torch.empty(c.in_channels, 3, 3, c.out_channels, device=dev, dtype=torch.float32),
                        torch.empty(c.out_channels, 3, 3, c.in_channels, device=dev, dtype=torch.float32) if i > 0 else None)
                    for i, (n, c) in enumerate(zip(names, convs))}       # the first layer's input needs no gradient
            table = (_lib.EncoderPrepLayer * len(names))()
            for i, (n, c) in enumerate(zip(names, convs)):
                table[i].weight, table[i].fwd_t, table[i].dgrad_t = c.weight.data_ptr(), bufs[n][0].data_ptr(), _ptr(bufs[n][1])
                table[i].c_out, table[i].c_in = c.out_channels, c.in_channels
            self._train_wt = (key, bufs, table)
        _, bufs, table = self._train_wt
        w0 = convs[0].weight
        with _on(w0):
            check(lib().uorf_encoder_weight_prep(table, len(table), _stream(w0)), "uorf_encoder_weight_prep")
        return bufs

    def _conv(self, layer, src_a, up_a, src_b, h_in, w_in, use_cache=True, wt=None):
        conv = layer[0]
        if wt is not None:
            return self._conv_launch(conv, wt, src_a, up_a, src_b, h_in, w_in)
        key = (conv.weight.data_ptr(), conv.weight._version)
        cache = self._wt.get(id(conv)) if use_cache else None
        if cache is None or cache[0] != key:       # [out][in][3][3] -> [in][3][3][out], once per weight version
            cache = (key, conv.weight.detach().permute(1, 2, 3, 0).contiguous().float())
            # a copy made while a CUDA graph is being captured lives in the graph's pool and is rewritten by every replay: it is
            # never cached.  The training path never caches either (use_cache=False): the weights move every step, and fused
            # Adam / graph replays move them without bumping `_version` (invalidate_weight_cache covers inference after those).
            if use_cache and not torch.cuda.is_current_stream_capturing():
                self._wt[id(conv)] = cache
        return self._conv_launch(conv, cache[1], src_a, up_a, src_b, h_in, w_in)

    def _conv_launch(self, conv, weight_t, src_a, up_a, src_b, h_in, w_in):
        stride = conv.stride[0]
        ca, cb = src_a.shape[0], (0 if src_b is None else src_b.shape[0])
        out = torch.empty(conv.out_channels, (h_in - 1) // stride + 1, (w_in - 1) // stride + 1, device=src_a.device, dtype=torch.float32)
        with _on(src_a):
            check(lib().uorf_encoder_conv3x3(src_a.data_ptr(), ca, int(up_a), _ptr(src_b), cb, weight_t.data_ptr(),
                                             _f32c(conv.bias.detach()).data_ptr(), conv.out_channels, h_in, w_in, stride, out.data_ptr(),
                                             _stream(src_a)), "uorf_encoder_conv3x3")
        return out

    def _forward_inference(self, x, keep=None):
        """`keep`: dict that receives every layer's input / output activations (the training path's backward needs them)"""
        H, W = x.shape[2], x.shape[3]
        img = _f32c(x[0])
        planes = self._pixel_planes(H, W, x.device)[0]
        d0 = None
        uc = keep is None      # training forward: always the current weights ...
        wts = self._training_weights() if keep is not None else None      # ... both layouts of all layers from one launch
        wt_of = lambda name: None if wts is None else wts[name][0]
        if self.bottom:
            d0 = self._conv(self.enc_down_0, img, False, planes, H, W, uc, wt_of("enc_down_0"))
            d1 = self._conv(self.enc_down_1, d0, False, None, H, W, uc, wt_of("enc_down_1"))
        else:
            d1 = self._conv(self.enc_down_1, img, False, planes, H, W, uc, wt_of("enc_down_1"))
        h1, w1 = d1.shape[1], d1.shape[2]
        d2 = self._conv(self.enc_down_2, d1, False, None, h1, w1, uc, wt_of("enc_down_2"))
        h2, w2 = d2.shape[1], d2.shape[2]
        d3 = self._conv(self.enc_down_3, d2, False, None, h2, w2, uc, wt_of("enc_down_3"))
        u3 = self._conv(self.enc_up_3, d3, False, None, d3.shape[1], d3.shape[2], uc, wt_of("enc_up_3"))      # kept at low resolution
        u2 = self._conv(self.enc_up_2, u3, True, d2, h2, w2, uc, wt_of("enc_up_2"))                           # cat[upsample(u3), d2]
        out = self._conv(self.enc_up_1, u2, True, d1, h1, w1, uc, wt_of("enc_up_1"))                          # cat[upsample(u2), d1]
        if keep is not None:
            keep.update(wts=wts, img=img, planes=planes, d0=d0, d1=d1, d2=d2, d3=d3, u3=u3, u2=u2, out=out)
        return out[None]

    def _kernels_apply(self, x):
        """The direct-convolution kernels cover what every reference driver feeds the Encoder: ONE fp32 CUDA image (the
        drivers encode view 0 only, uorf_eval_model.py:121) whose sides survive the stride-2 levels (multiples of 4, 8 with
        `bottom`), at most 65,536 pixels and a channel count that is a multiple of 4 (launch_encoder_conv3x3's limits).
        Anything else - a batch, another dtype, odd sizes, > 256x256 inputs - is NOT a kernel shape and takes the
        reference's own layer stack below (cuDNN); `last_path` records which one ran, `require_kernels` turns the second
        case into an error."""
        H, W = x.shape[2], x.shape[3]
        z = self.enc_down_1[0].out_channels
        return (x.is_cuda and x.shape[0] == 1 and x.dtype == torch.float32 and self.use_kernels
                and H % (8 if self.bottom else 4) == 0 and W % (8 if self.bottom else 4) == 0
                and H * W <= 65536 and z % 4 == 0)

    def forward(self, x):
        W, H = x.shape[3], x.shape[2]
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p_.requires_grad for p_ in self.parameters()))
        if self._kernels_apply(x):
            if not needs_grad:
                self.last_path = "kernels"
                return self._forward_inference(x)
            if self.use_kernels_in_training:
                self.last_path = "kernels"
                return _EncoderFn.apply(self, x, *self.parameters())
        elif self.require_kernels:
            raise _lib.UorfError(f"Encoder: input {tuple(x.shape)} {x.dtype} on {x.device} is outside the kernel shapes "
                                 "(one fp32 CUDA image, sides multiples of 4 (8 with bottom), <= 65536 pixels) and require_kernels is set")
        self.last_path = "torch"
        x_ = torch.cat([x, self._pixel_planes(H, W, x.device).expand(x.shape[0], -1, -1, -1)], dim=1)
        if self.bottom:
            x_down_1 = self.enc_down_1(self.enc_down_0(x_))
        else:
            x_down_1 = self.enc_down_1(x_)
        x_down_2 = self.enc_down_2(x_down_1)
        x_down_3 = self.enc_down_3(x_down_2)
        x_up_3 = self.enc_up_3(x_down_3)
        x_up_2 = self.enc_up_2(torch.cat([x_up_3, x_down_2], dim=1))
        return self.enc_up_1(torch.cat([x_up_2, x_down_1], dim=1))


class _EncoderFn(torch.autograd.Function):
    """Training path of the Encoder (batch 1): the forward is the direct-convolution kernels (every activation kept); the backward
    walks the U-Net in reverse with cuDNN's convolution_backward on the kept activations (the same op autograd would run), the
    ReLU gates taken from the kept outputs, `torch.cat` / the bilinear upsamples re-materialised only here."""

    @staticmethod
    def forward(ctx, enc, x, *params):
        keep = {}
        with torch.no_grad():
            out = enc._forward_inference(x, keep)
        ctx.enc = enc
        ctx.x_needs_grad = x.requires_grad
        ctx.wts = keep["wts"]      # the step's weight layouts (the weights do not move between this forward and its backward)
        ctx.names = [n for n in ("img", "planes", "d0", "d1", "d2", "d3", "u3", "u2", "out") if keep[n] is not None]
        ctx.save_for_backward(*[keep[n] for n in ctx.names])
        return out

    @staticmethod
    def backward(ctx, g_out):
        enc = ctx.enc
        a = dict(zip(ctx.names, ctx.saved_tensors))
        names = [n for n, _ in enc.named_parameters()]
        sink = enc.grad_sink
        if enc.backward_kernels and g_out.is_cuda and not ctx.x_needs_grad:
            grads = _encoder_backward_kernels(enc, a, g_out, sink, ctx.wts)     # writes straight into the sink's buffers
            return (None, None) + ((None,) * len(names) if sink is not None else tuple(grads[n] for n in names))
        gx, grads = _encoder_backward(enc, a, g_out, ctx.x_needs_grad)
        if sink is not None:       # gradients land in caller-owned buffers (one flat all-reduce bucket); autograd sees None
            torch._foreach_copy_([sink[n] for n in names], [grads[n] for n in names])
            return (None, gx) + (None,) * len(names)
        return (None, gx) + tuple(grads[n] for n in names)


def _encoder_backward_kernels(enc, a, g_out, sink=None, wts=None):
    """Backward of Encoder.forward (batch 1, no image gradient) on the fp32 kernels of csrc/encoder.cu: per layer one weight / bias
    gradient launch (+ its fixed-order reduce) that reads the layer input the way the forward did (no cat / upsample tensors) and one
    data-gradient launch; the ReLU gates are taken from the kept outputs inside the loads.  The two bilinear-upsample backwards are fixed-order gathers, the
    skip connections' second gradients are added by the data-gradient launch itself: the whole backward is bit-reproducible.  Replaces autograd / cuDNN for models/model.py:36-61 under
    models/uorf_nogan_model.py:181-183.  -> {parameter name: gradient} (the sink's own tensors when `sink` is given)."""
    dev = g_out.device
    grads = {}
    stream = _stream(g_out)

    def buf(name, like):
        if sink is not None:
            t = sink[name]
            if not (t.is_contiguous() and t.dtype == torch.float32):
                raise _lib.UorfError("Encoder.grad_sink buffers must be contiguous fp32")
            return t
        return torch.empty_like(like, dtype=torch.float32, memory_format=torch.contiguous_format)

    def wgrad(name, src_a, up_a, src_b, g, out, h_in, w_in):
        conv = getattr(enc, name)[0]
        ca, cb = src_a.shape[0], (0 if src_b is None else src_b.shape[0])
        stride = conv.stride[0]
        gw, gb = buf(name + ".0.weight", conv.weight), buf(name + ".0.bias", conv.bias)
        ws = torch.empty(lib().uorf_encoder_wgrad_workspace_floats(ca + cb, conv.out_channels, h_in, w_in, stride), device=dev, dtype=torch.float32)
        check(lib().uorf_encoder_conv3x3_wgrad(src_a.data_ptr(), ca, int(up_a), _ptr(src_b), cb, g.data_ptr(), out.data_ptr(), conv.out_channels,
                                               h_in, w_in, stride, ws.data_ptr(), gw.data_ptr(), gb.data_ptr(), stream), "uorf_encoder_conv3x3_wgrad")
        grads[name + ".0.weight"], grads[name + ".0.bias"] = gw, gb

    def dgrad(name, g, out, h_in, w_in, into=None):
        """-> gradient of the layer input [c_in][h_in][w_in]; `into`: added to that tensor instead (skip connection)"""
        conv = getattr(enc, name)[0]
        wg = wts[name][1] if wts is not None else conv.weight.detach().float().permute(0, 2, 3, 1).contiguous()   # [c_out][3][3][c_in]
        gi = torch.empty(conv.in_channels, h_in, w_in, device=dev, dtype=torch.float32) if into is None else into
        check(lib().uorf_encoder_conv3x3_dgrad(g.data_ptr(), out.data_ptr(), conv.out_channels, wg.data_ptr(), conv.in_channels, h_in, w_in,
                                               conv.stride[0], int(into is not None), gi.data_ptr(), stream), "uorf_encoder_conv3x3_dgrad")
        return gi

    def up_bwd(g, low):     # backward of nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False), model.py:27-32
        out = torch.empty_like(low)
        check(lib().uorf_encoder_upsample2x_bwd(g.data_ptr(), low.shape[0], low.shape[1], low.shape[2], out.data_ptr(), stream),
              "uorf_encoder_upsample2x_bwd")
        return out

    z = a["out"].shape[0]
    h1, w1 = a["d1"].shape[1], a["d1"].shape[2]
    h2, w2 = a["d2"].shape[1], a["d2"].shape[2]
    h3, w3 = a["d3"].shape[1], a["d3"].shape[2]
    with _on(g_out):
        g = _f32c(g_out[0])
        wgrad("enc_up_1", a["u2"], True, a["d1"], g, a["out"], h1, w1)
        gi = dgrad("enc_up_1", g, a["out"], h1, w1)
        g_d1 = gi[z:]
        g = up_bwd(gi[:z], a["u2"])
        wgrad("enc_up_2", a["u3"], True, a["d2"], g, a["u2"], h2, w2)
        gi = dgrad("enc_up_2", g, a["u2"], h2, w2)
        g_d2 = gi[z:]
        g = up_bwd(gi[:z], a["u3"])
        wgrad("enc_up_3", a["d3"], False, None, g, a["u3"], h3, w3)
        g_d3 = dgrad("enc_up_3", g, a["u3"], h3, w3)
        wgrad("enc_down_3", a["d2"], False, None, g_d3, a["d3"], h2, w2)
        dgrad("enc_down_3", g_d3, a["d3"], h2, w2, into=g_d2)
        wgrad("enc_down_2", a["d1"], False, None, g_d2, a["d2"], h1, w1)
        dgrad("enc_down_2", g_d2, a["d2"], h1, w1, into=g_d1)
        H, W = a["img"].shape[1], a["img"].shape[2]
        if enc.bottom:
            wgrad("enc_down_1", a["d0"], False, None, g_d1, a["d1"], H, W)
            g_d0 = dgrad("enc_down_1", g_d1, a["d1"], H, W)
            wgrad("enc_down_0", a["img"], False, a["planes"], g_d0, a["d0"], H, W)
        else:
            wgrad("enc_down_1", a["img"], False, a["planes"], g_d1, a["d1"], H, W)
    return grads


def _encoder_backward(enc, a, g_out, x_needs_grad):
    """Backward of Encoder.forward (batch 1) from the kept activations `a` (CxHxW each: img, planes, [d0], d1, d2, d3, u3 and u2
    at low resolution, out): -> (grad of the image or None, {parameter name: gradient}).  Plain torch ops, device-agnostic."""
    grads = {}

    def up(t):        # nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False), model.py:27-32
        return F.interpolate(t[None], scale_factor=2, mode="bilinear", align_corners=False)

    def up_bwd(g, low):
        return torch.ops.aten.upsample_bilinear2d_backward(g, [g.shape[2], g.shape[3]], [1, low.shape[0], low.shape[1], low.shape[2]],
                                                           False, 2.0, 2.0)

    def conv_bwd(name, g, out, inp, need_input=True):
        conv = getattr(enc, name)[0]
        g = g * (out[None] > 0)                                   # ReLU gate from the kept output
        gi, gw, gb = torch.ops.aten.convolution_backward(g, inp, conv.weight, [conv.out_channels], list(conv.stride), list(conv.padding),
                                                         list(conv.dilation), False, [0, 0], 1, [need_input, True, True])
        grads[name + ".0.weight"], grads[name + ".0.bias"] = gw, gb
        return gi

    z = a["out"].shape[0]
    gi = conv_bwd("enc_up_1", g_out, a["out"], torch.cat([up(a["u2"]), a["d1"][None]], dim=1))
    g_d1 = gi[:, z:]
    gi = conv_bwd("enc_up_2", up_bwd(gi[:, :z].contiguous(), a["u2"]), a["u2"], torch.cat([up(a["u3"]), a["d2"][None]], dim=1))
    g_d2 = gi[:, z:]
    g_d3 = conv_bwd("enc_up_3", up_bwd(gi[:, :z].contiguous(), a["u3"]), a["u3"], a["d3"][None])
    g_d2 = g_d2 + conv_bwd("enc_down_3", g_d3, a["d3"], a["d2"][None])
    g_d1 = g_d1 + conv_bwd("enc_down_2", g_d2, a["d2"], a["d1"][None])
    x_in = torch.cat([a["img"], a["planes"]], dim=0)[None]
    if enc.bottom:
        g_d0 = conv_bwd("enc_down_1", g_d1, a["d1"], a["d0"][None])
        gx = conv_bwd("enc_down_0", g_d0, a["d0"], x_in, x_needs_grad)
    else:
        gx = conv_bwd("enc_down_1", g_d1, a["d1"], x_in, x_needs_grad)
    return (gx[:, :a["img"].shape[0]] if x_needs_grad else None), grads


_LOSS_SCRATCH_FLOATS = 72        # UORF_LOSS_SCRATCH_FLOATS (include/uorf_b200.h)


def loss_scratch(device) -> torch.Tensor:
    """the zero-initialised scratch a fused-loss call site keeps for its whole life (partial sums + a self-resetting ticket)"""
    return torch.zeros(_LOSS_SCRATCH_FLOATS, device=device, dtype=torch.float32)


class _ImageLossFn(torch.autograd.Function):
    """rgb_map (N*H*W) x 3 from raw2outputs, x N x 3 x H x W -> x_recon, rendered, rendered_norm, x_norm (N x 3 x H x W) and loss_recon =
    mse(x_recon, x): `uorf_image_loss_forward / _backward` (reference models/uorf_nogan_model.py:168-183, ~17 torch launches)."""

    @staticmethod
    def forward(ctx, rgb_map, x, mean, std, scratch):
        n, _, h, w = x.shape
        rgb_c, x_c = _f32c(rgb_map), _f32c(x)
        outs = [torch.empty(n, 3, h, w, device=x.device, dtype=torch.float32) for _ in range(4)]
        loss = torch.empty((), device=x.device, dtype=torch.float32)
        with _on(x_c):
            check(lib().uorf_image_loss_forward(rgb_c.data_ptr(), x_c.data_ptr(), mean.data_ptr(), std.data_ptr(), n, h * w, outs[0].data_ptr(),
                                                outs[1].data_ptr(), outs[2].data_ptr(), outs[3].data_ptr(), scratch.data_ptr(), loss.data_ptr(),
                                                _stream(x_c)), "uorf_image_loss_forward")
        ctx.save_for_backward(outs[0], x_c, std)
        ctx.shape = (n, h, w)
        ctx.mark_non_differentiable(outs[1], outs[3])
        ctx.set_materialize_grads(False)
        return outs[0], outs[1], outs[2], outs[3], loss

    @staticmethod
    def backward(ctx, g_x_recon, g_rendered, g_rn, g_x_norm, g_loss):
        x_recon, x_c, std = ctx.saved_tensors
        n, h, w = ctx.shape
        if g_x_recon is None and g_rn is None and g_loss is None:
            return None, None, None, None, None
        gxr = _f32c(g_x_recon) if g_x_recon is not None else None
        grn = _f32c(g_rn) if g_rn is not None else None
        gl = _f32c(g_loss) if g_loss is not None else None
        g_rgb = torch.empty(n * h * w, 3, device=x_c.device, dtype=torch.float32)
        with _on(x_c):
            check(lib().uorf_image_loss_backward(x_recon.data_ptr(), x_c.data_ptr(), std.data_ptr(), _ptr(gxr), _ptr(grn), _ptr(gl), n, h * w,
                                                 g_rgb.data_ptr(), _stream(x_c)), "uorf_image_loss_backward")
        return g_rgb, None, None, None, None


class _FeatureMseFn(torch.autograd.Function):
    """features of [rendered images | ground-truth images] in one batch (any layout, the same for both halves) -> mse of the two halves;
    the gradient reaches the rendered half only (the reference detaches nothing but its ground-truth branch is constant)."""

    @staticmethod
    def forward(ctx, feats, scratch):
        base = feats if feats.is_contiguous() else (feats.permute(0, 2, 3, 1) if feats.dim() == 4 and feats.permute(0, 2, 3, 1).is_contiguous()
                                                    else feats.contiguous())
        if base.dtype != torch.float32 or base.shape[0] % 2:
            raise _lib.UorfError("feature mse: fp32 features of an even batch (rendered half, ground-truth half)")
        m = base.numel() // 2
        loss = torch.empty((), device=feats.device, dtype=torch.float32)
        with _on(base):
            check(lib().uorf_feature_mse_forward(base.data_ptr(), m, scratch.data_ptr(), loss.data_ptr(), _stream(base)), "uorf_feature_mse_forward")
        ctx.save_for_backward(base)
        ctx.perm = base.data_ptr() == feats.data_ptr() and not feats.is_contiguous()      # feats is the NCHW view of an NHWC buffer
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        (base,) = ctx.saved_tensors
        m = base.numel() // 2
        g = torch.empty_like(base)
        gl = _f32c(g_loss)
        with _on(base):
            check(lib().uorf_feature_mse_backward(base.data_ptr(), gl.data_ptr(), m, g.data_ptr(), _stream(base)), "uorf_feature_mse_backward")
        return (g.permute(0, 3, 1, 2) if ctx.perm else g), None


def invert_pose(m: torch.Tensor) -> torch.Tensor:
    """`m.inverse()` for the drivers' pose matrices (... x 3 x 3 or ... x 4 x 4): one launch of uorf_invert_small for fp32 CUDA
    tensors (torch's route is an LU factorisation + triangular solves: 13 launches for 16 numbers), torch's own inverse otherwise"""
    if not (m.is_cuda and m.dtype == torch.float32 and m.dim() >= 2 and m.shape[-1] == m.shape[-2] and m.shape[-1] in (3, 4)
            and not m.requires_grad):
        return m.inverse()
    src = m.contiguous()
    dst = torch.empty_like(src)
    with _on(src):
        check(lib().uorf_invert_small(src.data_ptr(), dst.data_ptr(), src.numel() // (src.shape[-1] ** 2), src.shape[-1], _stream(src)),
              "uorf_invert_small")
    return dst


# ======================================================================================================
# sin_emb (standalone; the decoder kernel computes it on chip)
# ======================================================================================================
def sin_emb(x, n_freq=5, keep_ori=True):
    """reference models/model.py:261-277.  Px3 -> Px(3+6*n_freq)."""
    embedded = [x] if keep_ori else []
    for i in range(n_freq):
        f = float(2 ** i)
        embedded += [torch.sin(f * x), torch.cos(f * x)]
    return torch.cat(embedded, dim=1)


# ======================================================================================================
# raw2outputs
# ======================================================================================================
class _Composite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, z_vals, rays_d, render_mask):
        _require_cuda(raw, z_vals, rays_d)
        raw_c, z_c, d_c = _f32c(raw), _f32c(z_vals), _f32c(rays_d)
        R, D = raw_c.shape[0], raw_c.shape[1]
        Rg = z_c.shape[0]
        if R % max(Rg, 1) != 0 or d_c.shape[0] != Rg:
            raise _lib.UorfError("raw2outputs: raw rows must be a multiple of the z_vals / rays_d rows")
        rgb = torch.empty(R, 3, device=raw.device, dtype=torch.float32)
        depth = torch.empty(R, device=raw.device, dtype=torch.float32)
        wn = torch.empty(R, D, device=raw.device, dtype=torch.float32)
        mm = torch.empty(R, device=raw.device, dtype=torch.float32) if render_mask else None
        with _on(raw_c):
            check(lib().uorf_composite_forward(raw_c.data_ptr(), z_c.data_ptr(), d_c.data_ptr(), R, Rg, D, rgb.data_ptr(),
                                               depth.data_ptr(), wn.data_ptr(), _ptr(mm), _stream(raw)),
                  "uorf_composite_forward")
        ctx.save_for_backward(raw_c, z_c, d_c)
        ctx.mark_non_differentiable(depth, wn)
        if render_mask:
            ctx.mark_non_differentiable(mm)
            return rgb, depth, wn, mm
        return rgb, depth, wn

    @staticmethod
    def backward(ctx, g_rgb, *unused):
        raw_c, z_c, d_c = ctx.saved_tensors
        R, D = raw_c.shape[0], raw_c.shape[1]
        g = _f32c(g_rgb)
        g_raw = torch.empty_like(raw_c)
        with _on(raw_c):
            check(lib().uorf_composite_backward(raw_c.data_ptr(), z_c.data_ptr(), d_c.data_ptr(), g.data_ptr(), R, z_c.shape[0], D,
                                                g_raw.data_ptr(), _stream(raw_c)), "uorf_composite_backward")
        return g_raw, None, None, None


def raw2outputs(raw, z_vals, rays_d, render_mask=False):
    """reference models/model.py:279-313.  raw RxDx4, z_vals RxD, rays_d Rx3 ->
    rgb_map Rx3, depth_map R, weights_norm RxD [, mask_map R].
    Gradients flow from rgb_map to raw (depth uses detached weights in the reference; mask_map is only used
    under no_grad by the reference's drivers).
    Extension: raw may hold k stacked ray sets ((k*R)xDx4) sharing one geometry (z_vals RxD, rays_d Rx3); they are
    composited in one launch (the per-slot renders of uorf_eval_model.py:181-190)."""
    return _Composite.apply(raw, z_vals, rays_d, bool(render_mask))


# ======================================================================================================
# Projection
# ======================================================================================================
class Projection(object):
    """reference models/projection.py:6-113 (same constructor and construct_sampling_coor signature)."""

    def __init__(self, focal_ratio=(350. / 320., 350. / 240.), near=5, far=16, frustum_size=[128, 128, 128],
                 device='cpu', nss_scale=7, render_size=(64, 64)):
        self.render_size = render_size
        self.device = torch.device(device)
        self.focal_ratio = focal_ratio
        self.near = near
        self.far = far
        self.frustum_size = frustum_size
        self.nss_scale = nss_scale
        # host-side constants computed exactly as the reference does (projection.py:17-31,42)
        self.world2nss = torch.tensor([[1 / nss_scale, 0, 0, 0], [0, 1 / nss_scale, 0, 0],
                                       [0, 0, 1 / nss_scale, 0], [0, 0, 0, 1]]).unsqueeze(0).to(self.device)
        focal_x = self.focal_ratio[0] * self.frustum_size[0]
        focal_y = self.focal_ratio[1] * self.frustum_size[1]
        bias_x = (self.frustum_size[0] - 1.) / 2.
        bias_y = (self.frustum_size[1] - 1.) / 2.
        intrinsic_mat = torch.tensor([[focal_x, 0, bias_x, 0], [0, focal_y, bias_y, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
        self.cam2spixel = intrinsic_mat.to(self.device)
        spixel2cam = intrinsic_mat.inverse()
        self.spixel2cam = spixel2cam.to(self.device)
        W, H, D = self.frustum_size
        self._z_cam = torch.linspace(self.near, self.far, D).to(self.device)
        f = _lib.Frustum()
        f.W, f.H, f.D = int(W), int(H), int(D)
        f.near_plane, f.far_plane = float(near), float(far)
        for i, v in enumerate(spixel2cam.flatten().tolist()):
            f.spixel2cam[i] = v
        f.nss_scale = float(nss_scale)
        self._frustum = f

    def construct_sampling_coor(self, cam2world, partitioned=False, ray_major=False, crop=None):
        """cam2world Nx4x4 -> frus_nss_coor (N*D*H*W)x3, z_vals (N*H*W)xD, ray_dir (N*H*W)x3
        (leading scale^2 chunk dim when partitioned), reference projection.py:51-113.

        Extensions used by the B200 drivers (default off, so the reference call is unchanged):
          ray_major=True  points ordered (n, h, w, d): the decoder output is then directly the RxDx4 tensor
                          raw2outputs wants (no permute copy, reference uorf_eval_model.py:147);
          crop=(h0, w0)   only the render_size window at (h0, w0) (fine-stage training crop,
                          reference uorf_nogan_model.py:153-165), never materialising the full frustum;
          crop=(h0, w0, Hc, Wc)  an arbitrary window (row blocks of one view when rays are sharded over GPUs);
          crop=<CUDA int32 tensor (h0, w0)>  the render_size window at an origin that stays on the device
                          (uorf_sample_coords_window: no host sync, so the fine-stage step is graph-capturable).
        """
        _require_cuda(cam2world)
        cam2world = _f32c(cam2world)
        N = cam2world.shape[0]
        W, H, D = self.frustum_size
        dev = cam2world.device
        if self._z_cam.device != dev:
            self._z_cam = self._z_cam.to(dev)
        crop_dev = None
        if isinstance(crop, torch.Tensor):
            # window origin (h0, w0) held on the device (int32, 2 elements): read and clamped by the kernel, no host sync
            if partitioned:
                raise ValueError("crop and partitioned are mutually exclusive")
            if not crop.is_cuda or crop.dtype != torch.int32 or crop.numel() != 2 or not crop.is_contiguous():
                raise ValueError("a device crop origin must be a contiguous CUDA int32 tensor (h0, w0)")
            crop_dev = crop
            scale, Hc, Wc, ch, cw = 1, int(self.render_size[0]), int(self.render_size[1]), 0, 0
            if Hc > H or Wc > W:
                raise ValueError(f"crop window {Hc}x{Wc} does not fit the {H}x{W} frustum")
        elif crop is not None:
            if partitioned:
                raise ValueError("crop and partitioned are mutually exclusive")
            scale, Hc, Wc = 1, int(self.render_size[0]), int(self.render_size[1])
            ch, cw = int(crop[0]), int(crop[1])
            if len(crop) == 4:
                Hc, Wc = int(crop[2]), int(crop[3])
            if ch < 0 or cw < 0 or Hc < 1 or Wc < 1 or ch + Hc > H or cw + Wc > W:
                raise ValueError(f"crop window {tuple(crop)} does not fit the {H}x{W} frustum")
        else:
            scale = H // self.render_size[0] if partitioned else 1
            Hc, Wc = H // scale, W // scale
            ch = cw = -1
        nchunk = scale * scale
        P, R = N * D * Hc * Wc, N * Hc * Wc
        coords = torch.empty(nchunk, P, 3, device=dev, dtype=torch.float32)
        z_vals = torch.empty(nchunk, R, D, device=dev, dtype=torch.float32)
        ray_dir = torch.empty(nchunk, R, 3, device=dev, dtype=torch.float32)
        with _on(cam2world):
            if crop_dev is not None:
                check(lib().uorf_sample_coords_window(C.byref(self._frustum), self._z_cam.data_ptr(), cam2world.data_ptr(), N,
                                                      crop_dev.data_ptr(), Hc, Wc, 1 if ray_major else 0, coords.data_ptr(),
                                                      z_vals.data_ptr(), ray_dir.data_ptr(), _stream(cam2world)),
                      "uorf_sample_coords_window")
            else:
                check(lib().uorf_sample_coords(C.byref(self._frustum), self._z_cam.data_ptr(), cam2world.data_ptr(), N, scale, ch, cw,
                                               Hc, Wc, 1 if ray_major else 0, coords.data_ptr(), z_vals.data_ptr(),
                                               ray_dir.data_ptr(), _stream(cam2world)), "uorf_sample_coords")
        if partitioned:
            return coords, z_vals, ray_dir
        return coords[0], z_vals[0], ray_dir[0]


# ======================================================================================================
# SlotAttention
# ======================================================================================================
class SlotAttention(nn.Module):
    """reference models/model.py:164-258 (same parameters / state-dict keys)."""

    def __init__(self, num_slots, in_dim=64, slot_dim=64, iters=3, eps=1e-8, hidden_dim=128):
        super().__init__()
        self.num_slots = num_slots
        self.iters = iters
        self.eps = eps
        self.scale = slot_dim ** -0.5
        self.slots_mu = nn.Parameter(torch.randn(1, 1, slot_dim))
        self.slots_logsigma = nn.Parameter(torch.zeros(1, 1, slot_dim))
        init.xavier_uniform_(self.slots_logsigma)
        self.slots_mu_bg = nn.Parameter(torch.randn(1, 1, slot_dim))
        self.slots_logsigma_bg = nn.Parameter(torch.zeros(1, 1, slot_dim))
        init.xavier_uniform_(self.slots_logsigma_bg)
        self.to_k = nn.Linear(in_dim, slot_dim, bias=False)
        self.to_v = nn.Linear(in_dim, slot_dim, bias=False)
        self.to_q = nn.Sequential(nn.LayerNorm(slot_dim), nn.Linear(slot_dim, slot_dim, bias=False))
        self.to_q_bg = nn.Sequential(nn.LayerNorm(slot_dim), nn.Linear(slot_dim, slot_dim, bias=False))
        self.gru = nn.GRUCell(slot_dim, slot_dim)
        self.gru_bg = nn.GRUCell(slot_dim, slot_dim)
        hidden_dim = max(slot_dim, hidden_dim)
        self.hidden_dim = hidden_dim
        self.to_res = nn.Sequential(nn.LayerNorm(slot_dim), nn.Linear(slot_dim, hidden_dim), nn.ReLU(inplace=True),
                                    nn.Linear(hidden_dim, slot_dim))
        self.to_res_bg = nn.Sequential(nn.LayerNorm(slot_dim), nn.Linear(slot_dim, hidden_dim), nn.ReLU(inplace=True),
                                       nn.Linear(hidden_dim, slot_dim))
        self.norm_feat = nn.LayerNorm(in_dim)
        self.slot_dim = slot_dim
        self.in_dim = in_dim
        self.grad_sink = None     # {parameter name: tensor}: the backward kernels write parameter gradients there (see _grad_buffers)
        self.noise_override = None   # (noise_fg, noise_bg) used when forward() gets no `noise` (a reference driver calls forward(feat))

    def _weights(self) -> _lib.SlotAttentionWeights:
        sd = dict(self.named_parameters())
        w = _lib.SlotAttentionWeights()
        keep = []
        for field, key in _lib.SA_STATE_KEYS.items():
            t = _f32c(sd[key].detach())
            keep.append(t)
            setattr(w, field, t.data_ptr())
        w._keep = keep
        return w

    def _launch_forward(self, feat, K, noise_fg, noise_bg):
        B, N, Cin = feat.shape
        Cs = self.slot_dim
        dev = feat.device
        f0 = feat[0]
        ws = torch.empty(lib().uorf_slot_attention_workspace_floats(N, Cs, K), device=dev, dtype=torch.float32)
        slots = torch.empty(1, K, Cs, device=dev, dtype=torch.float32)
        attn = torch.empty(1, K, N, device=dev, dtype=torch.float32)
        w = self._weights()
        with _on(feat):
            check(lib().uorf_slot_attention_forward(C.byref(w), f0.data_ptr(), f0.stride(0), f0.stride(1), noise_fg.data_ptr(),
                                                    noise_bg.data_ptr(), N, Cs, K, self.hidden_dim, self.iters, float(self.eps),
                                                    ws.data_ptr(), slots.data_ptr(), attn.data_ptr(), _stream(feat)),
                  "uorf_slot_attention_forward")
        return slots, attn, ws

    def _launch_backward(self, feat, K, noise_fg, noise_bg, fwd_ws, grad_slots):
        """-> (grad_feat 1xNxC, {param name: grad})"""
        B, N, Cin = feat.shape
        Cs = self.slot_dim
        dev = feat.device
        f0 = feat[0]
        named = dict(self.named_parameters())
        # every parameter gradient is written in full by sa_bwd_reduce_kernel: no zero fill (34 launches per step)
        grads = _grad_buffers(named, self.grad_sink)
        g = _lib.SlotAttentionWeights()
        for field, key in _lib.SA_STATE_KEYS.items():
            setattr(g, field, grads[key].data_ptr())
        w = self._weights()
        ws = torch.empty(lib().uorf_slot_attention_backward_workspace_floats(N, Cs, K, self.hidden_dim), device=dev, dtype=torch.float32)
        gfeat = torch.empty(1, N, Cs, device=dev, dtype=torch.float32)
        gs = _f32c(grad_slots.reshape(K, Cs))
        with _on(feat):
            check(lib().uorf_slot_attention_backward(C.byref(w), C.byref(g), f0.data_ptr(), f0.stride(0), f0.stride(1),
                                                     noise_fg.data_ptr(), noise_bg.data_ptr(), N, Cs, K, self.hidden_dim, self.iters,
                                                     float(self.eps), fwd_ws.data_ptr(), gs.data_ptr(), ws.data_ptr(), gfeat.data_ptr(),
                                                     _stream(feat)), "uorf_slot_attention_backward")
        return gfeat, grads

    def forward(self, feat, num_slots=None, noise=None):
        """feat BxNxC (B must be 1, as in every reference driver) -> slots BxKxC, attn BxKxN.
        `noise` = (noise_fg (K-1)xC, noise_bg 1xC) overrides the two randn draws (model.py:216,219);
        by default they are drawn from torch's CUDA generator in the reference's order.
        Under autograd, gradients flow from `slots` to `feat` and every parameter through the backward kernels; `attn`
        is non-differentiable (the reference's training driver detaches it, uorf_nogan_model.py:186)."""
        _require_cuda(feat)
        B, N, Cin = feat.shape
        if B != 1:
            raise _lib.UorfError("SlotAttention kernel handles batch 1 (one scene), as the reference drivers do")
        K = num_slots if num_slots is not None else self.num_slots
        Cs = self.slot_dim
        if Cin != Cs:
            raise _lib.UorfError("in_dim must equal slot_dim (true for every reference configuration)")
        dev = feat.device
        if noise is None:
            noise = self.noise_override
        if noise is None:
            noise_fg = torch.randn(1, K - 1, Cs, device=dev)
            noise_bg = torch.randn(1, 1, Cs, device=dev)
        else:
            noise_fg, noise_bg = noise
        noise_fg = _f32c(noise_fg.reshape(K - 1, Cs))
        noise_bg = _f32c(noise_bg.reshape(Cs))
        feat = feat if feat.dtype == torch.float32 else feat.float()
        needs_grad = torch.is_grad_enabled() and (feat.requires_grad or any(p_.requires_grad for p_ in self.parameters()))
        if not needs_grad:
            slots, attn, _ = self._launch_forward(feat, K, noise_fg, noise_bg)
            return slots, attn
        names = [n for n, _ in self.named_parameters()]
        params = [p_ for _, p_ in self.named_parameters()]
        return _SlotAttentionFn.apply(self, feat, noise_fg, noise_bg, K, names, *params)


class _SlotAttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, feat, noise_fg, noise_bg, K, names, *params):
        slots, attn, ws = module._launch_forward(feat, K, noise_fg, noise_bg)
        ctx.module, ctx.K, ctx.names = module, K, names
        ctx.save_for_backward(feat, noise_fg, noise_bg, ws)
        ctx.mark_non_differentiable(attn)
        return slots, attn

    @staticmethod
    def backward(ctx, g_slots, g_attn):
        feat, noise_fg, noise_bg, ws = ctx.saved_tensors
        gfeat, grads = ctx.module._launch_backward(feat, ctx.K, noise_fg, noise_bg, ws, g_slots)
        if ctx.module.grad_sink is not None:
            return (None, gfeat, None, None, None, None) + (None,) * len(ctx.names)
        return (None, gfeat, None, None, None, None) + tuple(grads[n] for n in ctx.names)


# ======================================================================================================
# Decoder
# ======================================================================================================
class Decoder(nn.Module):
    """reference models/model.py:64-161 (same parameters / state-dict keys, same forward signature).

    Extra, B200-specific knobs (attributes; defaults reproduce the reference call):
      precision   'fp16x3' (default, fp32-parity mode): three-term fp16 split on the tensor cores, fp32 accumulation;
                  'bf16x3': same with bf16 operands (full fp32 range, ~16 instead of ~22 operand bits);
                  'bf16' / 'fp16': one pass, ~3x less tensor work, tolerance stated in DESIGN.md.
      emit        which of ('raws', 'masked_raws', 'unmasked_raws', 'masks') to materialise; skipped outputs are
                  returned as None (the API-mandated per-slot tensors are 36*K bytes per point of HBM writes).
      match_rng_stream  draw the reference's unconditional randn_like (model.py:155) even when dens_noise == 0,
                  so later draws from the same generator line up with the reference.
      fg_slot_offset    (K-1)x3 tensor or None (inference only): slot s is evaluated at sampling_coor_fg[s] - fg_slot_offset[s],
                  bit-identical to passing the reference's shifted clone (uorf_manip_model.py:201-204) but without
                  materialising (K-1)xPx3 coordinates.
    """

    ALL_OUTPUTS = ("raws", "masked_raws", "unmasked_raws", "masks")

    def __init__(self, n_freq=5, input_dim=33 + 64, z_dim=64, n_layers=3, locality=True, locality_ratio=4 / 7,
                 fixed_locality=False):
        super().__init__()
        self.n_freq = n_freq
        self.locality = locality
        self.locality_ratio = locality_ratio
        self.fixed_locality = fixed_locality
        self.out_ch = 4
        self.z_dim = z_dim
        self.n_layers = n_layers
        before_skip = [nn.Linear(input_dim, z_dim), nn.ReLU(True)]
        after_skip = [nn.Linear(z_dim + input_dim, z_dim), nn.ReLU(True)]
        for i in range(n_layers - 1):
            before_skip += [nn.Linear(z_dim, z_dim), nn.ReLU(True)]
            after_skip += [nn.Linear(z_dim, z_dim), nn.ReLU(True)]
        self.f_before = nn.Sequential(*before_skip)
        self.f_after = nn.Sequential(*after_skip)
        self.f_after_latent = nn.Linear(z_dim, z_dim)
        self.f_after_shape = nn.Linear(z_dim, self.out_ch - 3)
        self.f_color = nn.Sequential(nn.Linear(z_dim, z_dim // 4), nn.ReLU(True), nn.Linear(z_dim // 4, 3))
        before_skip = [nn.Linear(input_dim, z_dim), nn.ReLU(True)]
        after_skip = [nn.Linear(z_dim + input_dim, z_dim), nn.ReLU(True)]
        for i in range(n_layers - 1):
            before_skip += [nn.Linear(z_dim, z_dim), nn.ReLU(True)]
            after_skip += [nn.Linear(z_dim, z_dim), nn.ReLU(True)]
        after_skip.append(nn.Linear(z_dim, self.out_ch))
        self.b_before = nn.Sequential(*before_skip)
        self.b_after = nn.Sequential(*after_skip)
        self.precision = "fp16x3"
        self.emit = self.ALL_OUTPUTS
        self.match_rng_stream = False
        self.return_outside = False
        self.last_outside = None
        self.events = None   # set to a list to collect (start, end) CUDA events around the fused kernel
        self.events_bwd = None   # same for the backward call (composition-backward, main, reduce and unfold kernels)
        self.fg_slot_offset = None
        self.grad_sink = None     # {parameter name: tensor}: the backward kernel writes parameter gradients there (see _grad_buffers)

    # ---- C-ABI plumbing ----
    def _weights(self) -> _lib.DecoderWeights:
        w = _lib.DecoderWeights()
        keep = []

        def p(t):
            t = _f32c(t.detach())
            keep.append(t)
            return t.data_ptr()

        for i in range(3):
            w.f_before_w[i], w.f_before_b[i] = p(self.f_before[2 * i].weight), p(self.f_before[2 * i].bias)
            w.f_after_w[i], w.f_after_b[i] = p(self.f_after[2 * i].weight), p(self.f_after[2 * i].bias)
            w.b_before_w[i], w.b_before_b[i] = p(self.b_before[2 * i].weight), p(self.b_before[2 * i].bias)
        for i in range(4):
            w.b_after_w[i], w.b_after_b[i] = p(self.b_after[2 * i].weight), p(self.b_after[2 * i].bias)
        w.f_after_latent_w, w.f_after_latent_b = p(self.f_after_latent.weight), p(self.f_after_latent.bias)
        w.f_after_shape_w, w.f_after_shape_b = p(self.f_after_shape.weight), p(self.f_after_shape.bias)
        w.f_color0_w, w.f_color0_b = p(self.f_color[0].weight), p(self.f_color[0].bias)
        w.f_color2_w, w.f_color2_b = p(self.f_color[2].weight), p(self.f_color[2].bias)
        w._keep = keep
        return w

    def _precision(self) -> int:
        try:
            return _lib.PRECISIONS[self.precision]
        except KeyError:
            raise ValueError(f"unknown precision {self.precision!r} (use one of {sorted(_lib.PRECISIONS)})") from None

    def _prepared_image(self, z, K, Cz, prec, dev, stream):
        nbytes = lib().uorf_decoder_image_bytes(K, prec)
        image = torch.empty(nbytes + 1024, device=dev, dtype=torch.uint8)
        image_ptr = image.data_ptr() + (-image.data_ptr()) % 1024
        w = self._weights()
        with _on(z):
            check(lib().uorf_decoder_prepare(C.byref(w), z.data_ptr(), K, Cz, self.n_freq, self.n_layers, prec, image_ptr,
                                             stream), "uorf_decoder_prepare")
        return image, image_ptr, w

    @staticmethod
    def _canon_inputs(sampling_coor_bg, sampling_coor_fg, z_slots, fg_transform):
        coor_bg = _f32c(sampling_coor_bg)
        fg = sampling_coor_fg if sampling_coor_fg.dtype == torch.float32 else sampling_coor_fg.float()
        if fg.stride(-1) != 1 or fg.stride(-2) != 3:
            fg = fg.contiguous()
        fg_stride = fg.stride(0) if fg.shape[0] > 1 else 0
        return coor_bg, fg, int(fg_stride), _f32c(z_slots.detach()), _f32c(fg_transform.detach().reshape(-1))

    def _launch_forward(self, sampling_coor_bg, sampling_coor_fg, z_slots, fg_transform, dens_noise, noise, emit):
        if self.n_layers != 3 or self.n_freq != 5:
            raise _lib.UorfError("decoder kernel is built for n_freq=5, n_layers=3 (every shipped configuration)")
        K, Cz = z_slots.shape
        P = sampling_coor_bg.shape[0]
        dev = sampling_coor_bg.device
        prec = self._precision()
        coor_bg, fg, fg_stride, z, T = self._canon_inputs(sampling_coor_bg, sampling_coor_fg, z_slots, fg_transform)
        if T.numel() != (16 if self.fixed_locality else 9):
            raise _lib.UorfError("fg_transform must be 1x4x4 with fixed_locality, else 1x3x3")
        stream = _stream(coor_bg)
        image, image_ptr, w = self._prepared_image(z, K, Cz, prec, dev, stream)
        noise_c = _f32c(noise.reshape(K, P)) if (noise is not None and dens_noise != 0.) else None
        mk = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.float32)
        raws = mk(P, 4) if "raws" in emit else None
        masked = mk(K, P, 4) if "masked_raws" in emit else None
        unmasked = mk(K, P, 4) if "unmasked_raws" in emit else None
        masks = mk(K, P, 1) if "masks" in emit else None
        outside = torch.empty(K - 1, P, device=dev, dtype=torch.uint8) if self.return_outside else None
        workspace = mk(lib().uorf_decoder_forward_workspace_floats(P, K))
        a = _lib.DecoderArgs()
        a.coor_bg, a.coor_fg, a.fg_slot_stride = coor_bg.data_ptr(), fg.data_ptr(), fg_stride
        a.fg_transform, a.noise = T.data_ptr(), _ptr(noise_c)
        a.dens_noise, a.locality_ratio = float(dens_noise), float(self.locality_ratio)
        a.locality, a.fixed_locality = int(bool(self.locality)), int(bool(self.fixed_locality))
        a.P, a.K, a.precision = P, K, prec
        a.image, a.workspace = image_ptr, workspace.data_ptr()
        a.raws, a.masked_raws, a.unmasked_raws, a.masks = _ptr(raws), _ptr(masked), _ptr(unmasked), _ptr(masks)
        a.outside = _ptr(outside)
        off = getattr(self, "fg_slot_offset", None)
        if off is not None:
            if tuple(off.shape) != (K - 1, 3) or not off.is_cuda:
                raise _lib.UorfError(f"fg_slot_offset must be a CUDA tensor of shape ({K - 1}, 3)")
            off = _f32c(off.detach())
        a.fg_slot_offset = _ptr(off)
        ev = self.events
        if ev is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream(dev))
        with _on(coor_bg):
            check(lib().uorf_decoder_forward(C.byref(a), stream), "uorf_decoder_forward")
        if ev is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record(torch.cuda.current_stream(dev))
            ev.append((e0, e1))
        self.last_outside = outside
        return raws, masked, unmasked, masks

    def _launch_backward(self, sampling_coor_bg, sampling_coor_fg, z_slots, fg_transform, dens_noise, noise, unmasked, grad_raws):
        """-> (grad_z_slots KxC, {param name: grad}) through uorf_decoder_backward."""
        K, Cz = z_slots.shape
        P = sampling_coor_bg.shape[0]
        dev = sampling_coor_bg.device
        coor_bg, fg, fg_stride, z, T = self._canon_inputs(sampling_coor_bg, sampling_coor_fg, z_slots, fg_transform)
        stream = _stream(coor_bg)
        image, image_ptr, w = self._prepared_image(z, K, Cz, _lib.PREC_FP16, dev, stream)
        named = dict(self.named_parameters())
        grads = _grad_buffers(named, self.grad_sink)
        g = _lib.DecoderWeights()
        for i in range(3):
            g.f_before_w[i], g.f_before_b[i] = grads[f"f_before.{2 * i}.weight"].data_ptr(), grads[f"f_before.{2 * i}.bias"].data_ptr()
            g.f_after_w[i], g.f_after_b[i] = grads[f"f_after.{2 * i}.weight"].data_ptr(), grads[f"f_after.{2 * i}.bias"].data_ptr()
            g.b_before_w[i], g.b_before_b[i] = grads[f"b_before.{2 * i}.weight"].data_ptr(), grads[f"b_before.{2 * i}.bias"].data_ptr()
        for i in range(4):
            g.b_after_w[i], g.b_after_b[i] = grads[f"b_after.{2 * i}.weight"].data_ptr(), grads[f"b_after.{2 * i}.bias"].data_ptr()
        g.f_after_latent_w, g.f_after_latent_b = grads["f_after_latent.weight"].data_ptr(), grads["f_after_latent.bias"].data_ptr()
        g.f_after_shape_w, g.f_after_shape_b = grads["f_after_shape.weight"].data_ptr(), grads["f_after_shape.bias"].data_ptr()
        g.f_color0_w, g.f_color0_b = grads["f_color.0.weight"].data_ptr(), grads["f_color.0.bias"].data_ptr()
        g.f_color2_w, g.f_color2_b = grads["f_color.2.weight"].data_ptr(), grads["f_color.2.bias"].data_ptr()
        ws = torch.empty(lib().uorf_decoder_backward_workspace_floats(P, K), device=dev, dtype=torch.float32)
        gz = torch.empty(K, Cz, device=dev, dtype=torch.float32)
        noise_c = _f32c(noise.reshape(K, P)) if (noise is not None and dens_noise != 0.) else None
        gr = _f32c(grad_raws)
        un = _f32c(unmasked)
        a = _lib.DecoderBwdArgs()
        a.coor_bg, a.coor_fg, a.fg_slot_stride = coor_bg.data_ptr(), fg.data_ptr(), fg_stride
        a.fg_transform, a.noise = T.data_ptr(), _ptr(noise_c)
        a.dens_noise, a.locality_ratio = float(dens_noise), float(self.locality_ratio)
        a.locality, a.fixed_locality = int(bool(self.locality)), int(bool(self.fixed_locality))
        a.P, a.K, a.Z = P, K, Cz
        a.z_slots, a.image, a.unmasked_raws, a.grad_raws = z.data_ptr(), image_ptr, un.data_ptr(), gr.data_ptr()
        a.workspace, a.grad_z_slots = ws.data_ptr(), gz.data_ptr()
        evb = getattr(self, "events_bwd", None)
        if evb is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream(dev))
        with _on(coor_bg):
            check(lib().uorf_decoder_backward(C.byref(w), C.byref(g), C.byref(a), stream), "uorf_decoder_backward")
        if evb is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record(torch.cuda.current_stream(dev))
            evb.append((e0, e1))
        return gz, grads

    def forward(self, sampling_coor_bg, sampling_coor_fg, z_slots, fg_transform, dens_noise=0., noise=None):
        """
        sampling_coor_bg Px3, sampling_coor_fg (K-1)xPx3 (an expand() view is consumed in place),
        z_slots KxC (slot 0 = background), fg_transform 1x3x3 (1x4x4 when fixed_locality)
        -> raws Px4, masked_raws KxPx4, unmasked_raws KxPx4, masks KxPx1
        Under autograd, gradients flow from `raws` to z_slots and every decoder parameter through the fused backward
        kernel; masked/unmasked/masks are non-differentiable (the reference detaches them, uorf_nogan_model.py:185-196).
        """
        _require_cuda(sampling_coor_bg, sampling_coor_fg, z_slots, fg_transform)
        K = z_slots.shape[0]
        P = sampling_coor_bg.shape[0]
        if noise is None and (dens_noise != 0. or self.match_rng_stream):
            noise = torch.randn(K, P, 1, device=sampling_coor_bg.device)  # model.py:155 (always drawn by the reference)
        needs_grad = torch.is_grad_enabled() and (z_slots.requires_grad or any(p.requires_grad for p in self.parameters()))
        if not needs_grad:
            return self._launch_forward(sampling_coor_bg, sampling_coor_fg, z_slots, fg_transform, dens_noise, noise, set(self.emit))
        if getattr(self, "fg_slot_offset", None) is not None:
            raise _lib.UorfError("fg_slot_offset is an inference-only knob (the backward kernel does not take it)")
        names = [n for n, _ in self.named_parameters()]
        params = [p_ for _, p_ in self.named_parameters()]
        raws, masked, unmasked, masks = _DecoderFn.apply(self, sampling_coor_bg, sampling_coor_fg, fg_transform, noise,
                                                         float(dens_noise), names, z_slots, *params)
        emit = set(self.emit)
        return (raws, masked if "masked_raws" in emit else None, unmasked if "unmasked_raws" in emit else None,
                masks if "masks" in emit else None)


class _DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, coor_bg, coor_fg, fg_transform, noise, dens_noise, names, z_slots, *params):
        emit = set(module.emit) | {"raws", "unmasked_raws"}   # unmasked is what the backward kernel re-reads
        raws, masked, unmasked, masks = module._launch_forward(coor_bg, coor_fg, z_slots, fg_transform, dens_noise, noise, emit)
        ctx.module, ctx.dens_noise, ctx.names = module, dens_noise, names
        ctx.save_for_backward(coor_bg, coor_fg, fg_transform, z_slots.detach(), unmasked)
        ctx.noise = noise
        dummy = lambda t, *shape: t if t is not None else raws.new_empty(0)
        outs = (raws, dummy(masked), unmasked, dummy(masks))
        ctx.mark_non_differentiable(outs[1], outs[2], outs[3])
        # autograd would hand the backward a zero tensor per non-differentiable output otherwise: two K x P x 4 fills per step
        ctx.set_materialize_grads(False)
        return outs

    @staticmethod
    def backward(ctx, g_raws, *unused):
        coor_bg, coor_fg, fg_transform, z_slots, unmasked = ctx.saved_tensors
        if g_raws is None:
            g_raws = torch.zeros(coor_bg.shape[0], 4, device=coor_bg.device, dtype=torch.float32)
        gz, grads = ctx.module._launch_backward(coor_bg, coor_fg, z_slots, fg_transform, ctx.dens_noise, ctx.noise, unmasked,
                                                g_raws)
        if ctx.module.grad_sink is not None:
            return (None, None, None, None, None, None, None, gz) + (None,) * len(ctx.names)
        return (None, None, None, None, None, None, None, gz) + tuple(grads[n] for n in ctx.names)


# ------------------------------------------------------------------------------------------------------------------------
# perceptual-loss network (reference models/model.py:316-324: frozen VGG16 features[:23], used at uorf_nogan_model.py:181-183)
# ------------------------------------------------------------------------------------------------------------------------
def pack_conv3x3(weight: torch.Tensor, precision: str = "fp16x3", dgrad: bool = False) -> torch.Tensor:
    """Conv2d weight [c_out][c_in][3][3] (fp32) -> the tile image uorf_conv3x3_x3 reads (include/uorf_b200.h):
    [c_out/64][9 taps][c_in/64][hi | lo] tiles of 64 x 64 16-bit values, 128-byte rows, chunk c of row r at chunk c ^ (r & 7).
    dgrad=True packs the weights of the data-gradient convolution: taps flipped, input / output channels exchanged."""
    w = weight.detach().float()
    if dgrad:
        w = w.flip(2, 3).permute(1, 0, 2, 3)
    co, ci = int(w.shape[0]), int(w.shape[1])
    if co % 64 or ci % 64 or tuple(w.shape[2:]) != (3, 3):
        raise _lib.UorfError("pack_conv3x3: 3x3 kernels with channel counts that are multiples of 64 only")
    dt = torch.float16 if precision == "fp16x3" else torch.bfloat16
    hi = w.to(dt)
    lo = (w - hi.float()).to(dt)

    def tiles(t):   # [co][ci][ky][kx] -> [co/64][tap][ci/64][row][col]
        return t.permute(2, 3, 0, 1).reshape(9, co // 64, 64, ci // 64, 64).permute(1, 0, 3, 2, 4)
    img = torch.stack([tiles(hi), tiles(lo)], dim=3).reshape(co // 64, 9, ci // 64, 2, 8, 8, 8, 8)   # [...][row group][row][chunk][elem]
    r8 = torch.arange(8, device=w.device).view(8, 1)
    pos = torch.arange(8, device=w.device).view(1, 8)
    idx = (pos ^ r8).view(1, 1, 1, 1, 1, 8, 8, 1).expand(co // 64, 9, ci // 64, 2, 8, 8, 8, 8)
    return img.gather(6, idx).contiguous().view(torch.int16).reshape(-1)


def _conv_split(x: torch.Tensor, gate: Optional[torch.Tensor], precision: str):
    hi = torch.empty(x.shape, dtype=torch.int16, device=x.device)
    lo = torch.empty(x.shape, dtype=torch.int16, device=x.device)
    with _on(x):
        check(lib().uorf_conv_split(x.data_ptr(), _ptr(gate), hi.data_ptr(), lo.data_ptr(), x.numel(), _lib.PRECISIONS[precision],
                                    _stream(x)), "uorf_conv_split")
    return hi, lo


def _conv3x3_run(hi: torch.Tensor, lo: torch.Tensor, shape, w_img: torch.Tensor, bias: Optional[torch.Tensor], c_out: int, relu: bool,
                 precision: str) -> torch.Tensor:
    """the convolution on operand planes that are already split (hi / lo of an n x h x w x c_in tensor) -> n x h x w x c_out fp32"""
    n, h, w, c_in = shape
    y = torch.empty((n, h, w, c_out), dtype=torch.float32, device=hi.device)
    with _on(hi):
        nws = int(lib().uorf_conv3x3_workspace_floats(n, h, w, c_in, c_out))
        ws = torch.empty(nws, dtype=torch.float32, device=hi.device) if nws else None
        check(lib().uorf_conv3x3_x3(hi.data_ptr(), lo.data_ptr(), w_img.data_ptr(), _ptr(bias), y.data_ptr(), _ptr(ws), n, h, w, c_in,
                                    c_out, int(relu), _lib.PRECISIONS[precision], _stream(hi)), "uorf_conv3x3_x3")
    return y


def _conv3x3_x3(x_nhwc: torch.Tensor, gate: Optional[torch.Tensor], w_img: torch.Tensor, bias: Optional[torch.Tensor], c_out: int,
                relu: bool, precision: str) -> torch.Tensor:
    hi, lo = _conv_split(x_nhwc, gate, precision)
    return _conv3x3_run(hi, lo, tuple(x_nhwc.shape), w_img, bias, c_out, relu, precision)


def _conv_split_pool(x_nhwc: torch.Tensor, grad_pooled: Optional[torch.Tensor], precision: str):
    """grad_pooled None: planes of maxpool2x2(x) (n x h/2 x w/2 x c); else planes of the pool's backward behind x's ReLU gate (n x h x w x c)"""
    n, h, w, c = x_nhwc.shape
    shape = (n, h // 2, w // 2, c) if grad_pooled is None else (n, h, w, c)
    hi = torch.empty(shape, dtype=torch.int16, device=x_nhwc.device)
    lo = torch.empty(shape, dtype=torch.int16, device=x_nhwc.device)
    with _on(x_nhwc):
        if grad_pooled is None:
            check(lib().uorf_conv_split_pool2x2(x_nhwc.data_ptr(), hi.data_ptr(), lo.data_ptr(), n, h, w, c, _lib.PRECISIONS[precision],
                                                _stream(x_nhwc)), "uorf_conv_split_pool2x2")
        else:
            check(lib().uorf_conv_split_unpool2x2(grad_pooled.data_ptr(), x_nhwc.data_ptr(), hi.data_ptr(), lo.data_ptr(), n, h, w, c,
                                                  _lib.PRECISIONS[precision], _stream(x_nhwc)), "uorf_conv_split_unpool2x2")
    return hi, lo, shape


def _pair(v):
    return [int(v), int(v)] if isinstance(v, int) else [int(t) for t in v]


class _Conv3x3ReluFn(torch.autograd.Function):
    """y = relu(conv3x3(x) + b) on NHWC tensors, frozen weights: the backward is the data gradient only (bf16 pairs: gradients of
    a weighted loss sit far below the fp16 range)."""

    @staticmethod
    def forward(ctx, x, w_fwd, w_bwd, bias, c_out, grad_rows):
        y = _conv3x3_x3(x, None, w_fwd, bias, c_out, True, "fp16x3")
        ctx.save_for_backward(y, w_bwd)
        ctx.c_in, ctx.grad_rows = x.shape[3], grad_rows
        return y

    @staticmethod
    def backward(ctx, g):
        y, w_bwd = ctx.saved_tensors
        r = ctx.grad_rows
        if r is not None and r < g.shape[0]:
            # only the first `grad_rows` images of the batch carry a gradient (the ground-truth half of the perceptual loss is
            # constant): the data gradient runs on those rows, the rest of grad_input is zero
            gx = g.new_zeros((g.shape[0], g.shape[1], g.shape[2], ctx.c_in))
            gx[:r] = _conv3x3_x3(_f32c(g[:r]), y[:r], w_bwd, None, ctx.c_in, False, "bf16x3")
        else:
            gx = _conv3x3_x3(_f32c(g), y, w_bwd, None, ctx.c_in, False, "bf16x3")
        return gx, None, None, None, None, None


class PerceptualVGG(nn.Module):
    """Drop-in for the nn.Sequential the reference's get_perceptual_net returns (models/model.py:316-324), built FROM it (same
    frozen weights): every 3x3 convolution + ReLU with at least 64 input channels runs on uorf_conv3x3_x3 (fp32-parity 3-term
    split on the tensor cores), the 3-channel stem and the max-pools stay torch.  Input N x 3 x H x W, output N x C x H' x W'
    (a channels-last view; the perceptual loss is a mean over all elements)."""

    def __init__(self, features: nn.Sequential):
        super().__init__()
        self.layers = []
        mods = list(features)
        i = 0
        self._bufs = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Conv2d):
                relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
                ok = (m.kernel_size == (3, 3) and m.stride == (1, 1) and m.padding == (1, 1) and m.in_channels % 64 == 0
                      and m.out_channels % 64 == 0 and relu and m.weight.is_cuda)
                if ok:
                    k = self._bufs
                    self.register_buffer(f"w_fwd{k}", pack_conv3x3(m.weight, "fp16x3"), persistent=False)
                    self.register_buffer(f"w_bwd{k}", pack_conv3x3(m.weight, "bf16x3", dgrad=True), persistent=False)
                    self.register_buffer(f"bias{k}", m.bias.detach().float().clone() if m.bias is not None else None, persistent=False)
                    self.layers.append(("x3", k, m.out_channels))
                    self._bufs += 1
                    i += 2
                    continue
                self.layers.append(("torch", m))
            else:
                self.layers.append(("torch", m))
            i += 1
        self._torch = nn.ModuleList([l[1] for l in self.layers if l[0] == "torch"])   # keeps them on the module's device
        self.use_stack = os.environ.get("UORF_VGG_STACK", "1") == "1"   # one autograd node for the whole network (_VggStackFn)
        self.fuse_pool = os.environ.get("UORF_VGG_FUSE_POOL", "1") == "1"   # ... with the 2x2 max-pools folded into the operand splits
        for p_ in self.parameters():
            p_.requires_grad = False

    def _stack_plan(self):
        """[("stem", conv) | ("x3", k, c_out) | ("pool", module)] when the network is stem conv + ReLU followed by tensor-core
        convolutions and max-pools only (VGG16 features[:23] is): the whole stack then runs under ONE autograd node"""
        L = self.layers
        if len(L) < 2 or L[0][0] != "torch" or L[1][0] != "torch" or not isinstance(L[0][1], nn.Conv2d) or not isinstance(L[1][1], nn.ReLU):
            return None
        c = L[0][1]
        if not (c.in_channels == 3 and c.kernel_size == (3, 3) and c.stride == (1, 1) and c.padding == (1, 1) and c.dilation == (1, 1)
                and c.groups == 1 and c.bias is not None and c.out_channels % 4 == 0 and c.out_channels <= 64 and c.weight.is_cuda
                and c.weight.dtype == torch.float32 and c.weight.is_contiguous() and c.bias.is_contiguous()):
            return None
        ops = [("stem", c)]
        for l in L[2:]:
            if l[0] == "x3":
                ops.append(l)
            elif isinstance(l[1], nn.MaxPool2d) and not l[1].return_indices:
                ops.append(("pool", l[1]))
            else:
                return None
        return ops

    def forward(self, x, grad_rows=None, x_const=None):
        """grad_rows: only the first `grad_rows` images need a gradient (the rest are constants evaluated in the same pass);
        x_const: further constant images appended to the batch (no torch.cat node, no zero gradient rows for them)"""
        _require_cuda(x)
        plan = self._stack_plan() if self.use_stack else None
        if plan is not None:
            r = x.shape[0] if grad_rows is None else int(grad_rows)
            return _VggStackFn.apply(self, plan, r, x, x_const).permute(0, 3, 1, 2)
        if x_const is not None:
            grad_rows = x.shape[0] if grad_rows is None else grad_rows
            x = torch.cat([x, x_const], dim=0)
        x = _f32c(x).permute(0, 2, 3, 1)                   # NHWC view of ...
        x = x.contiguous()                                 # ... a channels-last copy (3 channels: tiny)
        for l in self.layers:
            if l[0] == "x3":
                k, c_out = l[1], l[2]
                x = _Conv3x3ReluFn.apply(x, getattr(self, f"w_fwd{k}"), getattr(self, f"w_bwd{k}"), getattr(self, f"bias{k}"), c_out,
                                         grad_rows)
            else:
                y = l[1](x.permute(0, 3, 1, 2))            # torch layer on the channels-last NCHW view
                x = y.permute(0, 2, 3, 1)
                if not x.is_contiguous():
                    x = x.contiguous()
        return x.permute(0, 3, 1, 2)


class _VggStackFn(torch.autograd.Function):
    """The whole perceptual network under one autograd node: forward over the full batch (stem kernel, tensor-core layers, torch's
    channels-last max-pool with indices), backward over the first `rows` images only - the constant images of the batch never get
    gradient rows, so there are no per-layer zero fills / copies, and the 3-channel stem and its data gradient are two small FFMA
    launches instead of cuDNN's NHWC <-> NCHW conversions around a generic engine.  -> features NHWC."""

    @staticmethod
    def forward(ctx, vgg, plan, rows, x, x_const):
        xin = _f32c(x if x_const is None else torch.cat([x, x_const.to(x.dtype)], dim=0))
        n, _, h, w = xin.shape
        stream = _stream(xin)
        saved, meta = [], []
        cur = None
        with _on(xin):
            pending = None           # (hi, lo, shape): the next convolution's operand planes, produced by a fused pool + split
            for j, op in enumerate(plan):
                if op[0] == "stem":
                    conv = op[1]
                    cur = torch.empty(n, h, w, conv.out_channels, device=xin.device, dtype=torch.float32)
                    check(lib().uorf_conv3x3_stem(xin.data_ptr(), conv.weight.data_ptr(), conv.bias.data_ptr(), n, h, w, conv.out_channels,
                                                  cur.data_ptr(), stream), "uorf_conv3x3_stem")
                    meta.append(("stem", conv, len(saved)))
                    saved.append(cur)
                elif op[0] == "x3":
                    k, c_out = op[1], op[2]
                    if pending is not None:
                        c_in = pending[2][3]
                        cur = _conv3x3_run(pending[0], pending[1], pending[2], getattr(vgg, f"w_fwd{k}"), getattr(vgg, f"bias{k}"), c_out, True,
                                           "fp16x3")
                        pending = None
                    else:
                        c_in = cur.shape[3]
                        cur = _conv3x3_x3(cur, None, getattr(vgg, f"w_fwd{k}"), getattr(vgg, f"bias{k}"), c_out, True, "fp16x3")
                    meta.append(("x3", k, c_in, len(saved)))
                    saved.append(cur)
                else:
                    m = op[1]
                    ks, st, pd, dl = _pair(m.kernel_size), _pair(m.stride if m.stride is not None else m.kernel_size), _pair(m.padding), _pair(m.dilation)
                    fuse = (vgg.fuse_pool and ks == [2, 2] and st == [2, 2] and pd == [0, 0] and dl == [1, 1] and not m.ceil_mode
                            and j > 0 and plan[j - 1][0] == "x3" and j + 1 < len(plan) and plan[j + 1][0] == "x3"
                            and cur.shape[1] % 2 == 0 and cur.shape[2] % 2 == 0)
                    if fuse:         # the pooled tensor is never materialised: pool + split forward, unpool + gate + split backward
                        pending = _conv_split_pool(cur, None, "fp16x3")
                        meta.append(("pool_fused", None, len(saved) - 1))     # its input: the previous layer's output
                        cur = None
                        continue
                    out, idx = torch.ops.aten.max_pool2d_with_indices(cur.permute(0, 3, 1, 2), ks, st, pd, dl, m.ceil_mode)
                    meta.append(("pool", (ks, st, pd, dl, m.ceil_mode), len(saved)))
                    saved.append(idx)
                    cur = out.permute(0, 2, 3, 1)
                    if not cur.is_contiguous():
                        cur = cur.contiguous()
        ctx.vgg, ctx.meta, ctx.rows = vgg, meta, min(int(rows), n)
        ctx.x_shape = tuple(x.shape)
        ctx.save_for_backward(*saved)
        return cur

    @staticmethod
    def backward(ctx, g):
        vgg, r = ctx.vgg, ctx.rows
        saved = ctx.saved_tensors
        if not ctx.needs_input_grad[3] or r == 0:
            return None, None, None, None, None
        g = _f32c(g[:r])
        gs = None                    # (hi, lo, shape): operand planes of the next data-gradient convolution (fused unpool + gate + split)
        for i in range(len(ctx.meta) - 1, -1, -1):
            m = ctx.meta[i]
            if m[0] == "x3":
                if gs is not None:
                    g = _conv3x3_run(gs[0], gs[1], gs[2], getattr(vgg, f"w_bwd{m[1]}"), None, m[2], False, "bf16x3")
                    gs = None
                else:
                    g = _conv3x3_x3(g, saved[m[3]][:r], getattr(vgg, f"w_bwd{m[1]}"), None, m[2], False, "bf16x3")
            elif m[0] == "pool_fused":
                gs = _conv_split_pool(saved[m[2]][:r], g, "bf16x3")
            elif m[0] == "pool":
                ks, st, pd, dl, ceil = m[1]
                inp = saved[m[2] - 1][:r].permute(0, 3, 1, 2)          # a pool's input is the activation saved just before its indices
                gi = torch.ops.aten.max_pool2d_with_indices_backward(g.permute(0, 3, 1, 2), inp, ks, st, pd, dl, ceil, saved[m[2]][:r])
                g = _f32c(gi.permute(0, 2, 3, 1))
            else:
                conv = m[1]
                n_, h, w, c = g.shape
                gx = torch.empty(ctx.x_shape[0], 3, h, w, device=g.device, dtype=torch.float32)
                if r < ctx.x_shape[0]:
                    gx[r:].zero_()
                with _on(g):
                    check(lib().uorf_conv3x3_stem_dgrad(g.data_ptr(), saved[m[2]].data_ptr(), conv.weight.data_ptr(), r, h, w, c, gx.data_ptr(),
                                                        _stream(g)), "uorf_conv3x3_stem_dgrad")
                g = gx
        return None, None, None, g, None
