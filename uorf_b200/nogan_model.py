"""Training driver: the B200 mirror of the reference's uorfNoGanModel (models/uorf_nogan_model.py:15-285).

Same driver-level API (`set_input`, `forward(epoch)`, `backward`, `optimize_parameters(ret_grad, epoch)`, `compute_visuals`,
`save_networks` / `load_networks`, attributes `loss_recon`, `loss_perc`, `x_rec{i}`, `x{i}`, `masked_raws`, `unmasked_raws`,
`attn`) and the same option names.  One step = one scene of N views:

  Encoder (cuDNN, torch autograd) -> SlotAttention (kernel 4 / 5) -> frustum sampling (kernel 1; coarse 64^3 frustum or the
  fine-stage 64x64 crop of the 128^2 frustum, generated directly - the reference builds all 128^2 rays and slices) ->
  fused decoder (kernel 2 / 5) -> compositing (kernel 3 / 5) -> MSE + VGG16-relu4_3 perceptual loss (torch) -> Adam.

Everything runs in ray-major order, so no permute copies are made (reference :172).  Data-parallel training shards scenes
over ranks and averages gradients with one flat NCCL all-reduce (`uorf_b200.sharded.allreduce_gradients`).
"""
from __future__ import annotations

import argparse
import os
from itertools import chain
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn, optim
from torch.optim.lr_scheduler import LambdaLR

from . import optim as _optim
from .eval_model import init_weights_like_reference
from .modules import (Decoder, Encoder, PerceptualVGG, Projection, SlotAttention, _FeatureMseFn, _ImageLossFn, invert_pose, loss_scratch,
                      raw2outputs)


def default_train_options(**over) -> argparse.Namespace:
    """Option names and defaults of uorfNoGanModel.modify_commandline_options (uorf_nogan_model.py:28-55)."""
    opt = argparse.Namespace(num_slots=8, z_dim=64, attn_iter=3, warmup_steps=1000, nss_scale=7.0, render_size=64,
                             supervision_size=64, obj_scale=4.5, n_freq=5, n_samp=64, n_layer=3, weight_percept=0.006,
                             percept_in=100, no_locality_epoch=300, bottom=False, input_size=64, frustum_size=64,
                             frustum_size_fine=128, attn_decay_steps=2e5, coarse_epoch=600, near_plane=6.0, far_plane=20.0,
                             fixed_locality=False, dens_noise=1.0, lr=3e-4, stage="coarse", n_img_each_scene=4, gpu_ids=[0],
                             isTrain=True, vgg_weights=None, vgg_random_init=False)
    for k, v in over.items():
        setattr(opt, k, v)
    return opt


def exp_decay_schedule_with_warmup(optimizer, num_warmup_steps, num_decay_steps=2e5, decay_base=0.5):
    """reference models/networks.py:815-842 (lr_policy 'warmup')"""
    def lr_lambda(step: int):
        if step < num_warmup_steps:
            return float(step) / float(max(1, num_warmup_steps))
        return decay_base ** (float(step - num_warmup_steps) / float(num_decay_steps))
    return LambdaLR(optimizer, lr_lambda)


def perceptual_net(weights_path: Optional[str] = None, allow_random_init: bool = False) -> nn.Module:
    """reference models/model.py:316-324: frozen VGG16 features[:23] (through relu4_3).  The reference downloads the ImageNet
    weights (vgg16(pretrained=True)); there is no network here, so they are loaded from `weights_path` (a torchvision vgg16
    state dict).  Without a path the call raises - a randomly initialised VGG silently changes the training objective -
    unless `allow_random_init` is set (opt.vgg_random_init: benchmarks and parity tests, where only the arithmetic matters)."""
    from torchvision.models import vgg16
    vgg = vgg16(weights=None)
    if weights_path:
        vgg.load_state_dict(torch.load(weights_path, map_location="cpu"))
    elif not allow_random_init:
        raise RuntimeError("the perceptual loss needs the ImageNet VGG16 weights: set opt.vgg_weights to a torchvision vgg16 state dict "
                           "(the reference downloads them, models/model.py:320), pass with_perceptual=False / weight_percept=0, or opt in "
                           "to a randomly initialised network with opt.vgg_random_init=True (benchmarks and tests only)")
    net = nn.Sequential(*list(vgg.features)[:23]).eval()
    for p in net.parameters():
        p.requires_grad = False
    return net


class uorfNoGanModel:
    """reference models/uorf_nogan_model.py (forward :128-196, backward :223-227, optimize_parameters :229-245)."""

    def __init__(self, opt, precision: str = "fp16x3", with_perceptual: bool = True):
        self.opt = opt
        self.gpu_ids = list(getattr(opt, "gpu_ids", [0]))
        if not self.gpu_ids or not torch.cuda.is_available():
            raise RuntimeError("uorf_b200 drivers need a CUDA device (gpu_ids must not be empty); there is no CPU path")
        self.device = torch.device(f"cuda:{self.gpu_ids[0]}")
        self.isTrain = bool(getattr(opt, "isTrain", True))
        torch.backends.cudnn.allow_tf32 = False           # util/util.py:210-211 (set_seed)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.benchmark = True             # as the reference's set_seed does
        self.loss_names = ["recon", "perc"]
        self.model_names = ["Encoder", "SlotAttention", "Decoder"]
        self.with_perceptual = with_perceptual
        if with_perceptual and float(getattr(opt, "weight_percept", 0)) <= 0:
            self.with_perceptual = with_perceptual = False      # the term is multiplied by zero for the whole schedule (:181-183)
        if with_perceptual:
            self.perceptual_net = perceptual_net(getattr(opt, "vgg_weights", None),
                                                 bool(getattr(opt, "vgg_random_init", False))).to(self.device)
            if not bool(getattr(opt, "vgg_cudnn", False)):
                # the same frozen weights behind the tensor-core convolution kernel (uorf_conv3x3_x3); opt.vgg_cudnn keeps the
                # torch / cuDNN fp32 layers (A-B runs)
                self.perceptual_net = PerceptualVGG(self.perceptual_net)
            self._vgg_mean = torch.tensor([0.485, 0.456, 0.406], device=self.device).view(1, 3, 1, 1)
            self._vgg_std = torch.tensor([0.229, 0.224, 0.225], device=self.device).view(1, 3, 1, 1)
        render_size = (opt.render_size, opt.render_size)
        self.projection = Projection(device=self.device, nss_scale=opt.nss_scale,
                                     frustum_size=[opt.frustum_size, opt.frustum_size, opt.n_samp], near=opt.near_plane,
                                     far=opt.far_plane, render_size=render_size)
        self.projection_fine = Projection(device=self.device, nss_scale=opt.nss_scale,
                                          frustum_size=[opt.frustum_size_fine, opt.frustum_size_fine, opt.n_samp],
                                          near=opt.near_plane, far=opt.far_plane, render_size=render_size)
        z_dim = opt.z_dim
        self.num_slots = opt.num_slots
        self.netEncoder = init_weights_like_reference(Encoder(3, z_dim=z_dim, bottom=opt.bottom), "normal").to(self.device)
        self.netSlotAttention = init_weights_like_reference(
            SlotAttention(num_slots=opt.num_slots, in_dim=z_dim, slot_dim=z_dim, iters=opt.attn_iter), "normal").to(self.device)
        self.netDecoder = init_weights_like_reference(
            Decoder(n_freq=opt.n_freq, input_dim=6 * opt.n_freq + 3 + z_dim, z_dim=opt.z_dim, n_layers=opt.n_layer,
                    locality_ratio=opt.obj_scale / opt.nss_scale, fixed_locality=opt.fixed_locality), "xavier").to(self.device)
        self.netDecoder.precision = precision
        self.netDecoder.emit = ("raws", "masked_raws", "unmasked_raws")   # the driver keeps these for its visuals (:185-196)
        self.netDecoder.match_rng_stream = bool(getattr(opt, "match_rng_stream", False))
        # opt.cuda_graph: build the optimizer graph-capturable (device-side step counter, learning rate held in a device
        # scalar that the scheduler fills in place) so enable_cuda_graph() can replay the whole step as one graph launch
        self._capturable = bool(getattr(opt, "cuda_graph", False))
        self._graph_enabled, self._graph, self._graph_key, self._graph_warm, self._static = False, None, None, 0, None
        if self.isTrain:
            lr = torch.tensor(float(opt.lr), device=self.device) if self._capturable else opt.lr
            # torch's fused multi-tensor Adam (same update rule; one kernel per chunk of tensors instead of ~14 foreach launches
            # per step: 9.24 -> 8.94 ms).  opt.fused_adam=False / UORF_FUSED_ADAM=0 selects torch's default (foreach) - passed as
            # fused=None, never as an explicit False, which would select the slow single-tensor loop.
            fused = bool(getattr(opt, "fused_adam", os.environ.get("UORF_FUSED_ADAM", "1") == "1"))
            # ... and over it uorf_b200.optim.Adam: the same optimizer (state, state_dict) whose update is ONE launch of
            # uorf_adam_step with 1,024-element blocks (torch's fused kernel: 38 blocks of 65,536 elements, ~70 us); UORF_ADAM_KERNEL=0: torch's
            adam_cls = _optim.Adam if fused and os.environ.get("UORF_ADAM_KERNEL", "1") == "1" else optim.Adam
            self.optimizer = adam_cls(chain(self.netEncoder.parameters(), self.netSlotAttention.parameters(),
                                            self.netDecoder.parameters()), lr=lr, capturable=self._capturable,
                                      fused=True if fused else None)
            self.optimizers = [self.optimizer]
            self.schedulers = [exp_decay_schedule_with_warmup(self.optimizer, opt.warmup_steps, opt.attn_decay_steps)]
        self.L2_loss = nn.MSELoss()
        self.slot_noise = None        # (noise_fg, noise_bg) overrides the randn draws of SlotAttention
        self.crop_override = None     # (h0, w0) overrides the fine-stage randint draws
        self.process_group = None     # set (with world size > 1) for data-parallel training
        # every gradient in ONE persistent flat buffer the backward kernels write into (also on one GPU: the autograd path costs
        # an AccumulateGrad `add` launch per parameter - 83 per step - and per-tensor fills); False: plain autograd accumulation
        self.flat_grad_bucket = bool(getattr(opt, "flat_grad_bucket", True))
        self.fused_losses = bool(getattr(opt, "fused_losses", os.environ.get("UORF_FUSED_LOSSES", "1") == "1"))
        self._loss_scratch = None
        self._flat_grad = None        # data-parallel mode: the flat gradient bucket every .grad is a view of

    def parameters(self):
        return chain(self.netEncoder.parameters(), self.netSlotAttention.parameters(), self.netDecoder.parameters())

    def named_parameters(self):
        for pre, net in (("Encoder", self.netEncoder), ("SlotAttention", self.netSlotAttention), ("Decoder", self.netDecoder)):
            for n, p in net.named_parameters():
                yield pre + "." + n, p

    def set_input(self, input):
        """reference :117-126 (the fg transform is inverted on the host copy when the input arrives on the CPU)"""
        self.x = input["img_data"].to(self.device, non_blocking=True)
        c2w = input["cam2world"]
        self.cam2world = c2w.to(self.device, non_blocking=True)
        src = c2w if self.opt.fixed_locality else input["azi_rot"]
        if not self.opt.fixed_locality:
            self.cam2world_azi = input["azi_rot"].to(self.device, non_blocking=True)
        self.nss2cam0 = invert_pose(src[0:1]) if src.is_cuda else src[0:1].inverse().to(self.device, non_blocking=True)

    def forward(self, epoch=0):
        opt = self.opt
        self.weight_percept = opt.weight_percept if epoch >= opt.percept_in else 0
        dens_noise = opt.dens_noise if (epoch <= opt.percept_in and opt.fixed_locality) else 0
        dev = self.device
        rs = opt.render_size
        feature_map = self.netEncoder(F.interpolate(self.x[0:1], size=opt.input_size, mode="bilinear", align_corners=False))
        feat = feature_map.flatten(start_dim=2).permute([0, 2, 1])
        z_slots, attn = self.netSlotAttention(feat, noise=self.slot_noise)
        z_slots, attn = z_slots.squeeze(0), attn.squeeze(0)
        K = attn.shape[0]
        cam2world = self.cam2world
        N = cam2world.shape[0]
        D = opt.n_samp
        if opt.stage == "coarse":
            coords, z_vals, ray_dir = self.projection.construct_sampling_coor(cam2world, ray_major=True)
            x = F.interpolate(self.x, size=opt.supervision_size, mode="bilinear", align_corners=False)
        else:
            # The crop origin never visits the host: the two draws stay on the device (same generator calls, after the slot
            # noise, as the reference :160-161), the sampling kernel reads them there (uorf_sample_coords_window) and the
            # supervision patch is gathered with device indices - the reference's slicing by a device scalar (:162-165) is a
            # host sync per step, and the reason its fine stage cannot be graph-captured.
            if self.crop_override is not None:
                origin = torch.as_tensor(self.crop_override, dtype=torch.int32).reshape(2).to(dev, non_blocking=True)
            else:
                start_range = opt.frustum_size_fine - rs
                origin = torch.cat([torch.randint(low=0, high=start_range, size=(1,), device=dev),
                                    torch.randint(low=0, high=start_range, size=(1,), device=dev)]).to(torch.int32)
            self.crop_origin = origin
            span = self._crop_span(rs)
            x = self.x.index_select(2, span + origin[0]).index_select(3, span + origin[1])
            coords, z_vals, ray_dir = self.projection_fine.construct_sampling_coor(cam2world, ray_major=True, crop=origin)
        self.z_vals, self.ray_dir = z_vals, ray_dir
        W = H = opt.supervision_size
        raws, masked_raws, unmasked_raws, _ = self.netDecoder(coords, coords[None, ...].expand(K - 1, -1, -1), z_slots, self.nss2cam0,
                                                              dens_noise=dens_noise)
        rgb_map, _, _ = raw2outputs(raws.view(N * H * W, D, 4), z_vals, ray_dir)
        # image-side glue (x_recon, MSE, vgg_normalizer of both images) as one launch each way; opt.fused_losses=False /
        # UORF_FUSED_LOSSES=0: the reference's torch expressions (~25 tiny launches per step)
        fused = self.fused_losses and x.dtype == torch.float32 and rgb_map.dtype == torch.float32
        rendered_norm = x_norm = None
        if fused:
            if self._loss_scratch is None:
                self._loss_scratch = (loss_scratch(dev), loss_scratch(dev))
                self._vgg_mean3 = (self._vgg_mean if self.with_perceptual else torch.zeros(3, device=dev)).reshape(3).contiguous()
                self._vgg_std3 = (self._vgg_std if self.with_perceptual else torch.ones(3, device=dev)).reshape(3).contiguous()
            x_recon, rendered, rendered_norm, x_norm, self.loss_recon = _ImageLossFn.apply(rgb_map, x, self._vgg_mean3, self._vgg_std3,
                                                                                           self._loss_scratch[0])
        else:
            rendered = rgb_map.view(N, H, W, 3).permute([0, 3, 1, 2])
            x_recon = rendered * 2 - 1
            self.loss_recon = self.L2_loss(x_recon, x)
        self.loss_extra = self._extra_generator_loss(x_recon, x, epoch)   # None here; the adversarial term of the GAN driver
        if self.with_perceptual and self.weight_percept == 0:
            # the reference evaluates both VGG passes and multiplies by zero (:181-183): skipped, the loss is exactly 0
            self.loss_perc = torch.zeros((), device=dev)
        elif self.with_perceptual:
            if rendered_norm is None:
                rendered_norm = (rendered - self._vgg_mean) / self._vgg_std
                x_norm = (((x + 1) / 2 - self._vgg_mean) / self._vgg_std).detach()
            if isinstance(self.perceptual_net, PerceptualVGG) and fused:
                feats = self.perceptual_net(rendered_norm, x_const=x_norm)
                self.loss_perc = self.weight_percept * _FeatureMseFn.apply(feats, self._loss_scratch[1])
            elif isinstance(self.perceptual_net, PerceptualVGG):
                # both images in ONE pass (half the launches, twice the tiles per launch); the ground-truth half carries no
                # gradient: its features are detached and the convolution's backward skips those rows
                feats = self.perceptual_net(rendered_norm, x_const=x_norm)
                rendered_feat, x_feat = feats[:N], feats[N:].detach()
            else:
                rendered_feat = self.perceptual_net(rendered_norm)
                with torch.no_grad():      # the ground-truth branch carries no gradient (frozen VGG, constant images)
                    x_feat = self.perceptual_net(x_norm)
            if not (isinstance(self.perceptual_net, PerceptualVGG) and fused):
                self.loss_perc = self.weight_percept * self.L2_loss(rendered_feat, x_feat)
        else:
            self.loss_perc = torch.zeros((), device=dev)
        with torch.no_grad():
            self.x_recon, self.rendered, self.x_sup = x_recon.detach(), rendered.detach(), x.detach()
            for i in range(min(N, opt.n_img_each_scene)):
                setattr(self, f"x_rec{i}", x_recon[i].detach())
                setattr(self, f"x{i}", x[i])
            # reference shape [K,N,D,H,W,4] as views of the ray-major buffers; attn stays on the device (the reference's
            # .cpu() at :186 is a per-step host sync that only feeds the visualiser)
            self._masked_rm, self._unmasked_rm = masked_raws, unmasked_raws
            self.masked_raws = masked_raws.detach().view(K, N, H, W, D, 4).permute(0, 1, 4, 2, 3, 5)
            self.unmasked_raws = unmasked_raws.detach().view(K, N, H, W, D, 4).permute(0, 1, 4, 2, 3, 5)
            self.attn_raw, self._feat_hw = attn.detach(), (feature_map.shape[2], feature_map.shape[3])

    @property
    def attn(self):
        """K x 1 x H x W attention maps resized to the supervision size (reference :186-190), built on demand."""
        H_, W_ = self._feat_hw
        a = self.attn_raw.view(self.opt.num_slots, 1, H_, W_)
        S = self.opt.supervision_size
        return a if H_ == S else F.interpolate(a, size=[S, S], mode="bilinear")

    def compute_visuals(self):
        """reference :198-221: per-slot masked / unmasked renders, 2K compositing passes in two launches"""
        with torch.no_grad():
            K, N, D, H, W, _ = self.masked_raws.shape
            for name, buf in (("slot", self._masked_rm), ("unmasked_slot", self._unmasked_rm)):
                rgb_map, _, _ = raw2outputs(buf.detach().view(K * N * H * W, D, 4), self.z_vals, self.ray_dir)
                x_recon = rgb_map.view(K, N, H, W, 3).permute([0, 1, 4, 2, 3]) * 2 - 1
                for k in range(K):
                    for i in range(min(N, self.opt.n_img_each_scene)):
                        setattr(self, f"{name}{k}_view{i}", x_recon[k, i])
            a = self.attn
            for k in range(K):
                setattr(self, f"slot{k}_attn", a[k] * 2 - 1)

    def _crop_span(self, rs):
        if getattr(self, "_span", None) is None or self._span.numel() != rs:
            self._span = torch.arange(rs, device=self.device, dtype=torch.int32)
        return self._span

    def _extra_generator_loss(self, x_recon, x, epoch):
        return None

    def _graph_key_extra(self, epoch):
        """epoch-dependent constants of a subclass that are baked into a captured step"""
        return ()

    def backward(self):
        loss = self.loss_recon + self.loss_perc
        if self.loss_extra is not None:
            loss = loss + self.loss_extra
        loss.backward()
        self.loss_perc = self.loss_perc / self.weight_percept if self.weight_percept > 0 else self.loss_perc

    # ---- whole-step CUDA graph ------------------------------------------------------
    def enable_cuda_graph(self, enabled: bool = True):
        """After this, optimize_parameters() runs three more eager steps (cuDNN autotuning, optimizer state, .grad buffers must
        exist before capture), captures forward + backward + Adam into ONE CUDA graph and replays it from then on: ~460 small
        launches (Encoder, VGG, slot attention, Adam, losses) stop costing host time and inter-kernel gaps.  Inputs are
        copied into static buffers; a change of the epoch-dependent constants (perceptual weight, density noise, locality)
        or of the input shapes re-captures.  Coarse and fine stage alike: the fine-stage crop origin is drawn, read and applied
        on the device (forward()), and the generator's Philox offset advances with every replay, so each replayed step crops a
        new window.  With a process group set, the flat gradient all-reduce (NCCL) is captured inside the same graph."""
        if enabled and not self._capturable:
            raise RuntimeError("construct the model with opt.cuda_graph=True (capturable Adam) to use enable_cuda_graph()")
        self._graph_enabled, self._graph, self._graph_key, self._graph_warm = bool(enabled), None, None, 0
        return self

    def _step_graphed(self, epoch):
        opt = self.opt
        key = (opt.weight_percept if epoch >= opt.percept_in else 0,
               opt.dens_noise if (epoch <= opt.percept_in and opt.fixed_locality) else 0, bool(self.netDecoder.locality),
               tuple(self.x.shape), tuple(self.cam2world.shape), self.slot_noise is None) + tuple(self._graph_key_extra(epoch))
        st = self._static
        if st is None or st["x"].shape != self.x.shape or st["cam2world"].shape != self.cam2world.shape:
            st = self._static = {"x": self.x.clone(), "cam2world": self.cam2world.clone(), "nss2cam0": self.nss2cam0.clone()}
            self._graph = None
        for name in ("x", "cam2world", "nss2cam0"):
            cur = getattr(self, name)
            if cur.data_ptr() != st[name].data_ptr():
                st[name].copy_(cur, non_blocking=True)
                setattr(self, name, st[name])
        if self._graph is None or key != self._graph_key:
            if key != self._graph_key:
                self._graph_key, self._graph_warm, self._graph = key, 0, None
            if self._graph_warm < 3:
                self._graph_warm += 1
                return self._step_eager(False, epoch)
            # drop every reference to the previous step's autograd graph (its AccumulateGrad nodes belong to the eager
            # stream and would invalidate the capture) and let the captured backward allocate .grad in the graph's pool
            self.loss_recon = self.loss_perc = None
            if self.process_group is None and not self.flat_grad_bucket:     # (the flat bucket keeps its persistent gradient views)
                for p_ in self.parameters():
                    p_.grad = None
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            # thread-local capture mode: the NCCL watchdog thread of a data-parallel run may touch the CUDA API meanwhile
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._step_eager(False, epoch)
            self._graph = g
        self._graph.replay()
        self.netEncoder.invalidate_weight_cache()      # the replay moved the weights without bumping their version counters
        return [], []

    def optimize_parameters(self, ret_grad=False, epoch=0):
        if self._graph_enabled and not ret_grad:
            return self._step_graphed(epoch)
        return self._step_eager(ret_grad, epoch)

    def _flat_gradients(self):
        """Data-parallel mode: ONE persistent flat fp32 buffer holds every gradient; each `.grad` is a view of its slice and the
        backward kernels write there directly (`grad_sink`), so the step issues exactly one all-reduce (average) over the
        buffer - no torch.cat, no division pass, no per-tensor copy back (one flat-bucket ncclAllReduce)."""
        if self._flat_grad is None:
            from .sharded import flatten_gradients
            self._flat_grad = flatten_gradients(list(self.parameters()))
            for net in (self.netEncoder, self.netSlotAttention, self.netDecoder):
                net.grad_sink = {n: p_.grad for n, p_ in net.named_parameters()}
        return self._flat_grad

    def _step_eager(self, ret_grad=False, epoch=0):
        self.forward(epoch)
        if self.process_group is not None or self.flat_grad_bucket:
            flat = self._flat_gradients()
            flat.zero_()               # one fill; the kernels overwrite their slices, anything routed through autograd accumulates
            self.backward()
            if self.process_group is not None:
                from .sharded import allreduce_flat
                allreduce_flat(flat, self.process_group)
        else:
            # optimizer.zero_grad() of the reference (:235-236, zero in place): one multi-tensor launch instead of one fill per tensor
            grads = [p.grad for p in self.parameters() if p.grad is not None]
            if grads:
                torch._foreach_zero_(grads)
            self.backward()
        layers, avg_grads = [], []
        if ret_grad:
            for n, p in self.named_parameters():
                if p.grad is not None and "bias" not in n:
                    layers.append(n.split(".", 1)[1])
                    avg_grads.append(p.grad.abs().mean().item())
        for opm in self.optimizers:
            opm.step()
        # fused Adam rewrites the weights without bumping their version counters: inference-path encoder calls (visuals,
        # forward_disc) must not see transposed copies cached before the step
        self.netEncoder.invalidate_weight_cache()
        return layers, avg_grads

    def update_learning_rate(self):
        for s in self.schedulers:
            s.step()

    def get_current_losses(self):
        return {"recon": float(self.loss_recon), "perc": float(self.loss_perc)}

    # ---- checkpoints, reference base_model.py:142-201 + uorf_nogan_model.py:247-285 file names ----
    def save_networks(self, save_dir, suffix="latest"):
        os.makedirs(save_dir, exist_ok=True)
        for name in self.model_names:
            sd = {k: v.detach().cpu() for k, v in getattr(self, "net" + name).state_dict().items()}
            torch.save(sd, os.path.join(save_dir, f"{suffix}_net_{name}.pth"))
        for i, opm in enumerate(self.optimizers):
            torch.save(opm.state_dict(), os.path.join(save_dir, f"{suffix}_optimizer_{i}.pth"))
        for i, sch in enumerate(self.schedulers):
            torch.save(sch.state_dict(), os.path.join(save_dir, f"{suffix}_lr_scheduler_{i}.pth"))

    def load_networks(self, save_dir, suffix="latest", with_optimizer=True):
        # a captured step holds the old optimizer state / lr tensors: capture again after the load
        self._graph, self._graph_key, self._graph_warm = None, None, 0
        self.netEncoder.invalidate_weight_cache()
        for name in self.model_names:
            sd = torch.load(os.path.join(save_dir, f"{suffix}_net_{name}.pth"), map_location=str(self.device))
            getattr(self, "net" + name).load_state_dict(sd, strict=False)
        if with_optimizer and self.isTrain:
            for i, opm in enumerate(self.optimizers):
                p = os.path.join(save_dir, f"{suffix}_optimizer_{i}.pth")
                if os.path.exists(p):
                    opm.load_state_dict(torch.load(p, map_location=str(self.device)))
            for i, sch in enumerate(self.schedulers):
                p = os.path.join(save_dir, f"{suffix}_lr_scheduler_{i}.pth")
                if os.path.exists(p):
                    sch.load_state_dict(torch.load(p))

    def load_networks_from_state_dicts(self, enc=None, sa=None, dec=None):
        if enc is not None:
            self.netEncoder.load_state_dict(enc)
        if sa is not None:
            self.netSlotAttention.load_state_dict(sa)
        if dec is not None:
            self.netDecoder.load_state_dict(dec)
