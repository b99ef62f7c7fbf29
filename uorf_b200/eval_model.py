"""Eval driver: the B200 mirror of the reference's uorfEvalModel (models/uorf_eval_model.py:18-235).

Same driver-level API (`set_input`, `forward`, `test`, `compute_visuals`, attributes `x_rec{i}`, `masked_raws`,
`unmasked_raws`, `slot{k}_view{i}`, `loss_recon`, `loss_psnr`) and the same option names (`opt` is the argparse
Namespace the reference builds, see `default_eval_options`).  Two layouts:

  layout='reference'  the reference's ray-chunked loop, verbatim data movement (uorf_eval_model.py:135-157): strided
                      partition into scale^2 chunks, one decoder + compositing call per chunk, scatter into
                      [K,N,D,H,W,4] buffers.  Exists for index-for-index parity tests.
  layout='native'     (default) one pass over the frustum in ray-major order: the decoder output already is the
                      RxDx4 tensor raw2outputs wants and `masked_raws` is exposed as a permuted VIEW with the reference's
                      [K,N,D,H,W,4] shape - no permute copies, no scatter, no zero-filled buffers.

Out of scope: LPIPS / SSIM / ARI metrics (third-party packages), visdom/HTML output.
"""
from __future__ import annotations

import argparse
from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch import nn

from .modules import Decoder, Encoder, Projection, SlotAttention, invert_pose, raw2outputs


def default_eval_options(**over) -> argparse.Namespace:
    """Option names and defaults of uorfEvalModel.modify_commandline_options (uorf_eval_model.py:31-46)."""
    opt = argparse.Namespace(num_slots=8, z_dim=64, attn_iter=3, nss_scale=7.0, render_size=64, obj_scale=4.5, n_freq=5,
                             n_samp=64, n_layer=3, bottom=False, input_size=64, frustum_size=128, near_plane=6.0,
                             far_plane=20.0, fixed_locality=False, n_img_each_scene=4, gpu_ids=[0], isTrain=False)
    for k, v in over.items():
        setattr(opt, k, v)
    return opt


def init_weights_like_reference(net: nn.Module, init_type: str, init_gain: float = 0.02) -> nn.Module:
    """networks.init_weights (reference models/networks.py:67-95): Conv*/Linear weights N(0, gain) ('normal') or
    xavier_normal ('xavier'), zero bias; everything else keeps PyTorch defaults."""
    def f(m):
        name = m.__class__.__name__
        if hasattr(m, "weight") and (name.find("Conv") != -1 or name.find("Linear") != -1):
            if init_type == "normal":
                nn.init.normal_(m.weight.data, 0.0, init_gain)
            elif init_type == "xavier":
                nn.init.xavier_normal_(m.weight.data)
            else:
                raise NotImplementedError(init_type)
            if getattr(m, "bias", None) is not None:
                nn.init.constant_(m.bias.data, 0.0)
    net.apply(f)
    return net


class uorfEvalModel:
    """reference models/uorf_eval_model.py:18-235 (forward :116-157, compute_visuals :175-193)."""

    def __init__(self, opt, layout: str = "native", precision: str = "fp16x3"):
        self.opt = opt
        self.gpu_ids = list(getattr(opt, "gpu_ids", [0]))
        if not self.gpu_ids or not torch.cuda.is_available():
            raise RuntimeError("uorf_b200 drivers need a CUDA device (gpu_ids must not be empty); there is no CPU path")
        self.device = torch.device(f"cuda:{self.gpu_ids[0]}")
        self.isTrain = False
        # the reference runs fp32 convolutions / matmuls with TF32 disabled (util/util.py:210-211, set_seed)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.benchmark = True             # as the reference's set_seed does
        self.layout = layout
        self.model_names = ["Encoder", "SlotAttention", "Decoder"]
        render_size = (opt.render_size, opt.render_size)
        frustum_size = [opt.frustum_size, opt.frustum_size, opt.n_samp]
        self.projection = Projection(device=self.device, nss_scale=opt.nss_scale, frustum_size=frustum_size,
                                     near=opt.near_plane, far=opt.far_plane, render_size=render_size)
        z_dim = opt.z_dim
        self.num_slots = opt.num_slots
        # same construction order and init rules as the reference (uorf_eval_model.py:78-83)
        self.netEncoder = init_weights_like_reference(Encoder(3, z_dim=z_dim, bottom=opt.bottom), "normal").to(self.device)
        self.netSlotAttention = init_weights_like_reference(
            SlotAttention(num_slots=opt.num_slots, in_dim=z_dim, slot_dim=z_dim, iters=opt.attn_iter), "normal").to(self.device)
        self.netDecoder = init_weights_like_reference(
            Decoder(n_freq=opt.n_freq, input_dim=6 * opt.n_freq + 3 + z_dim, z_dim=opt.z_dim, n_layers=opt.n_layer,
                    locality=False, locality_ratio=opt.obj_scale / opt.nss_scale, fixed_locality=opt.fixed_locality),
            "xavier").to(self.device)
        self.netDecoder.precision = precision
        self.L2_loss = nn.MSELoss()
        # the reference's eval forward keeps raws + masked + unmasked per-slot tensors and discards `masks`
        # (uorf_eval_model.py:146-151)
        self.netDecoder.emit = ("raws", "masked_raws", "unmasked_raws")
        self.slot_noise = None                          # (noise_fg, noise_bg) overrides the randn draws

    # ---- BaseModel-style helpers -----------------------------------------------------------------
    def eval(self):
        for n in self.model_names:
            getattr(self, "net" + n).eval()
        return self

    def _weights_changed(self):
        """captured graphs bake the encoder's transposed weight copies (made once per weight version): re-capture on next use"""
        self._graph = None
        self._render_static = None

    def load_networks_from_state_dicts(self, enc=None, sa=None, dec=None):
        self._weights_changed()
        if enc is not None:
            self.netEncoder.load_state_dict(enc)
        if sa is not None:
            self.netSlotAttention.load_state_dict(sa)
        if dec is not None:
            self.netDecoder.load_state_dict(dec)

    def load_networks(self, save_dir: str, suffix: str = "latest"):
        """reference base_model.py:177-201: `{suffix}_net_{Name}.pth` state dicts (strict=False)."""
        import os
        self._weights_changed()
        for name in self.model_names:
            path = os.path.join(save_dir, f"{suffix}_net_{name}.pth")
            sd = torch.load(path, map_location=str(self.device))
            getattr(self, "net" + name).load_state_dict(sd, strict=False)

    def set_input(self, input):
        """reference uorf_eval_model.py:99-114.  The fg transform (inverse of view 0's azimuth rotation / pose) is
        taken on the host copy when the input arrives on the CPU: bit-identical to the reference's CPU path and no
        device LAPACK call + sync in the render loop."""
        self.x = input["img_data"].to(self.device, non_blocking=True)
        c2w = input["cam2world"]
        self.cam2world = c2w.to(self.device, non_blocking=True)
        src = c2w if self.opt.fixed_locality else input["azi_rot"]
        if not self.opt.fixed_locality:
            self.cam2world_azi = input["azi_rot"].to(self.device, non_blocking=True)
        if src.is_cuda:
            self.nss2cam0 = invert_pose(src[0:1])
        else:
            # staged through PINNED memory (a small ring, so that a caller may keep several steps in flight): an H2D copy from
            # pageable memory makes the driver wait for the stream's earlier work before it returns, which serialises the host
            # with the previous step's render
            inv = src[0:1].inverse()
            ring = getattr(self, "_pin_ring", None)
            if ring is None or ring[0].shape != inv.shape or ring[0].dtype != inv.dtype:
                ring = self._pin_ring = [torch.empty_like(inv).pin_memory() for _ in range(4)]
                self._pin_next = 0
            stage = ring[self._pin_next]
            self._pin_next = (self._pin_next + 1) % len(ring)
            stage.copy_(inv)
            self.nss2cam0 = stage.to(self.device, non_blocking=True)
        self.image_paths = input.get("paths", [])
        for k in ("masks", "mask_idx", "fg_idx", "obj_idxs"):
            if k in input:
                setattr(self, "gt_masks" if k == "masks" else k, input[k])

    # ---- forward ------------------------------------------------------------------------------------
    def encode(self):
        """Encoder -> SlotAttention on view 0 (uorf_eval_model.py:121-127): z_slots KxC, attn KxN."""
        opt = self.opt
        feature_map = self.netEncoder(F.interpolate(self.x[0:1], size=opt.input_size, mode="bilinear", align_corners=False))
        feat = feature_map.flatten(start_dim=2).permute([0, 2, 1])   # view; consumed in place by the kernel
        z_slots, attn = self.netSlotAttention(feat, noise=self.slot_noise)
        return z_slots.squeeze(0), attn.squeeze(0)

    def render(self, z_slots, cam2world, nss2cam0, fg_slot_offset=None, rows=None):
        """frustum sampling -> fused decoder -> compositing for the given views, ray-major layout.
        fg_slot_offset (K-1)x3: per-object translation in NSS units (the moved-object render of uorf_manip_model.py:196-204).
        rows=(h0, h1): only image rows h0..h1-1 of every view (one rank's block when the rays of a view are sharded).
        Returns dict(rendered Nx3xHxW, depth NxHxW, masked / unmasked ray-major KxPx4, z_vals, ray_dir); H = h1-h0 with rows."""
        K = z_slots.shape[0]
        N = cam2world.shape[0]
        W, H, D = self.projection.frustum_size
        if rows is None:
            coords, z_vals, ray_dir = self.projection.construct_sampling_coor(cam2world, ray_major=True)
        else:
            h0, h1 = int(rows[0]), int(rows[1])
            coords, z_vals, ray_dir = self.projection.construct_sampling_coor(cam2world, ray_major=True, crop=(h0, 0, h1 - h0, W))
            H = h1 - h0
        self.netDecoder.fg_slot_offset = fg_slot_offset
        try:
            raws, masked, unmasked, _ = self.netDecoder(coords, coords[None].expand(K - 1, -1, -1), z_slots, nss2cam0)
        finally:
            self.netDecoder.fg_slot_offset = None
        rgb_map, depth_map, _ = raw2outputs(raws.view(N * H * W, D, 4), z_vals, ray_dir)
        rendered = rgb_map.view(N, H, W, 3).permute([0, 3, 1, 2])
        return dict(rendered=rendered, depth=depth_map.view(N, H, W), masked=masked, unmasked=unmasked, z_vals=z_vals,
                    ray_dir=ray_dir)

    # ---- CUDA graph: the whole forward (about 60 small launches around the decoder) replayed as one graph launch ----
    def enable_cuda_graph(self, enabled: bool = True):
        """After this, forward() captures itself into a CUDA graph on its first call (shapes must then stay fixed) and
        replays the graph afterwards; inputs are copied into static buffers by forward(), outputs live in the graph's pool."""
        self._graph_enabled = enabled
        self._graph = None
        return self

    def _forward_graphed(self):
        cur = torch.cuda.current_stream(self.device)
        if self._graph is None:
            self._static = {"x": self.x.clone(), "cam2world": self.cam2world.clone(), "nss2cam0": self.nss2cam0.clone()}
            self.x, self.cam2world, self.nss2cam0 = self._static["x"], self._static["cam2world"], self._static["nss2cam0"]
            side = torch.cuda.Stream(self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._forward_impl()
            cur.wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._forward_impl()
            self._graph = g
        else:
            st = self._static
            if self.x.data_ptr() != st["x"].data_ptr():
                st["x"].copy_(self.x, non_blocking=True)
                st["cam2world"].copy_(self.cam2world, non_blocking=True)
                st["nss2cam0"].copy_(self.nss2cam0, non_blocking=True)
                self.x, self.cam2world, self.nss2cam0 = st["x"], st["cam2world"], st["nss2cam0"]
        self._graph.replay()

    def render_graphed(self, z_slots, cam2world, nss2cam0):
        """render() replayed as a CUDA graph (captured on first use; shapes must stay fixed).  Used by the ray-sharded
        multi-GPU path, where the slot latents arrive by broadcast and only the render is rank-local."""
        cur = torch.cuda.current_stream(self.device)
        st = getattr(self, "_render_static", None)
        if st is None:
            st = {"z": z_slots.clone(), "c2w": cam2world.clone(), "T": nss2cam0.clone()}
            side = torch.cuda.Stream(self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(2):
                    self.render(st["z"], st["c2w"], st["T"])
            cur.wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g), torch.no_grad():
                st["out"] = self.render(st["z"], st["c2w"], st["T"])
            st["graph"] = g
            self._render_static = st
        else:
            st["z"].copy_(z_slots, non_blocking=True)
            st["c2w"].copy_(cam2world, non_blocking=True)
            st["T"].copy_(nss2cam0, non_blocking=True)
        st["graph"].replay()
        return st["out"]

    def forward(self, epoch=0):
        if getattr(self, "_graph_enabled", False):
            return self._forward_graphed()
        return self._forward_impl()

    def _forward_impl(self, epoch=0):
        with torch.no_grad():
            opt = self.opt
            nss2cam0 = self.nss2cam0
            z_slots, attn = self.encode()
            self.z_slots, self.attn = z_slots, attn
            K = attn.shape[0]
            cam2world = self.cam2world
            N = cam2world.shape[0]
            W, H, D = self.projection.frustum_size
            if self.layout == "reference":
                x_recon, rendered = self._forward_reference_layout(z_slots, nss2cam0, cam2world, K, N, W, H, D)
            else:
                out = self.render(z_slots, cam2world, nss2cam0)
                rendered = out["rendered"]
                x_recon = rendered * 2 - 1
                self.depth, self.z_vals, self.ray_dir = out["depth"], out["z_vals"], out["ray_dir"]
                masked, unmasked = out["masked"], out["unmasked"]
                # reference shape [K,N,D,H,W,4] as a view of the ray-major buffers
                self.masked_raws = None if masked is None else masked.view(K, N, H, W, D, 4).permute(0, 1, 4, 2, 3, 5)
                self.unmasked_raws = None if unmasked is None else unmasked.view(K, N, H, W, D, 4).permute(0, 1, 4, 2, 3, 5)
                self._masked_ray_major = masked
            self.x_recon, self.rendered = x_recon, rendered
            x = self.x
            if x.shape[-1] == x_recon.shape[-1] and x.shape[0] == N and N > 1:
                x_recon_novel, x_novel = x_recon[1:], x[1:]
                self.loss_recon = self.L2_loss(x_recon_novel, x_novel)
                mse01 = ((x_recon_novel / 2 + 0.5).clamp(0, 1) - (x_novel / 2 + 0.5)).pow(2).mean()
                self.loss_psnr = -10.0 * torch.log10(mse01)
            for i in range(min(N, opt.n_img_each_scene, x.shape[0])):
                setattr(self, f"x_rec{i}", x_recon[i])
                setattr(self, "input_image" if i == 0 else f"gt_novel_view{i}", x[i])

    def _forward_reference_layout(self, z_slots, nss2cam0, cam2world, K, N, W, H, D):
        """the reference's chunk loop, uorf_eval_model.py:133-157"""
        opt, dev = self.opt, self.device
        scale = H // opt.render_size
        frus_nss_coor, z_vals, ray_dir = self.projection.construct_sampling_coor(cam2world, partitioned=True)
        x_recon = torch.zeros([N, 3, H, W], device=dev)
        rendered = torch.zeros([N, 3, H, W], device=dev)
        masked_raws = torch.zeros([K, N, D, H, W, 4], device=dev)
        unmasked_raws = torch.zeros([K, N, D, H, W, 4], device=dev)
        self.depth = torch.zeros([N, H, W], device=dev)
        for j, (coor_, z_vals_, ray_dir_) in enumerate(zip(frus_nss_coor, z_vals, ray_dir)):
            h, w = divmod(j, scale)
            H_, W_ = H // scale, W // scale
            raws_, masked_, unmasked_, _ = self.netDecoder(coor_, coor_[None, ...].expand(K - 1, -1, -1), z_slots, nss2cam0)
            raws_ = raws_.view([N, D, H_, W_, 4]).permute([0, 2, 3, 1, 4]).flatten(start_dim=0, end_dim=2)
            masked_raws[..., h::scale, w::scale, :] = masked_.view([K, N, D, H_, W_, 4])
            unmasked_raws[..., h::scale, w::scale, :] = unmasked_.view([K, N, D, H_, W_, 4])
            rgb_map_, depth_map_, _ = raw2outputs(raws_, z_vals_, ray_dir_)
            rendered_ = rgb_map_.view(N, H_, W_, 3).permute([0, 3, 1, 2])
            rendered[..., h::scale, w::scale] = rendered_
            x_recon[..., h::scale, w::scale] = rendered_ * 2 - 1
            self.depth[:, h::scale, w::scale] = depth_map_.view(N, H_, W_)
        self.masked_raws, self.unmasked_raws = masked_raws, unmasked_raws
        self._masked_ray_major = None
        return x_recon, rendered

    # ---- per-slot renders + segmentation (uorf_eval_model.py:175-193) ---------------------------------
    def compute_visuals(self):
        with torch.no_grad():
            K, N, D, H, W, _ = self.masked_raws.shape
            if self._masked_ray_major is not None:
                raws = self._masked_ray_major.view(K * N * H * W, D, 4)          # already (k, n, h, w, d)
                z_vals, ray_dir = self.z_vals, self.ray_dir
            else:
                raws = self.masked_raws.permute([0, 1, 3, 4, 2, 5]).reshape(K * N * H * W, D, 4)
                _, z_vals, ray_dir = self.projection.construct_sampling_coor(self.cam2world[:N])
            # one launch composites all K slots (geometry shared): reference loops K times and recomputes the frustum
            rgb_map, _, _, mask_map = raw2outputs(raws, z_vals, ray_dir, render_mask=True)
            self.mask_maps = mask_map.view(K, N, H, W)
            x_recon = rgb_map.view(K, N, H, W, 3).permute([0, 1, 4, 2, 3]) * 2 - 1
            for k in range(K):
                for i in range(min(N, self.opt.n_img_each_scene)):
                    setattr(self, f"slot{k}_view{i}", x_recon[k, i])
            self.mask_idx_pred = self.mask_maps.argmax(dim=0)      # NxHxW, stays on device

    def test(self):
        with torch.no_grad():
            self.forward()
            self.compute_visuals()

    def get_current_visuals(self):
        names = [f"x_rec{i}" for i in range(self.opt.n_img_each_scene)]
        return {n: getattr(self, n) for n in names if hasattr(self, n)}

    def get_current_losses(self):
        return {n: float(getattr(self, "loss_" + n)) for n in ("recon", "psnr") if hasattr(self, "loss_" + n)}
