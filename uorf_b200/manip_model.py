"""Scene-manipulation driver: reconstruct, pick the slot that explains a given object, render the scene with that object moved.

Mirrors the reference's models/uorf_manip_model.py (set_input :97-115, forward :117-232) on top of uorfEvalModel:
  * reconstruction and per-slot mask maps come from the fused render + one K-slot compositing launch (the reference loops
    over 4 ray chunks, then K times over construct_sampling_coor + raw2outputs, uorf_manip_model.py:140-170);
  * the slot to move is the one whose argmax mask has the best IoU with `fg_idx` (:181-193), decided on the device;
  * the moved render passes a (K-1)x3 offset table to the decoder kernel instead of cloning the (K-1)xPx3 coordinates and
    shifting one slice (:201-204) - bit-identical arithmetic (fp32 subtract before the transform), 12*(K-1) bytes per point
    less HBM traffic each way.
"""
from __future__ import annotations

import torch

from .eval_model import uorfEvalModel
from .modules import raw2outputs


class uorfManipModel(uorfEvalModel):
    def set_input(self, input):
        super().set_input(input)
        # world units.  The reference's dataset yields a flat (3,) tensor (multiscenes_manip_dataset.py: torch.tensor([x, y, 0]),
        # collated as flat_batch[0]['movement']; the 'Nx3' comment at uorf_manip_model.py:112 is not what arrives): accept both
        self.movement = input["movement"].to(self.device).float().reshape(-1, 3)
        for k, name in (("img_data_moved", "gt_moved"), ("img_data_changed", "gt_changed"), ("img_data_bg", "bg_img")):
            if k in input:
                setattr(self, name, input[k].to(self.device))

    def slot_mask_maps(self):
        """K x N x H x W: sum_i w_i * sigma_i of each slot's masked raws (uorf_manip_model.py:164-170), one launch."""
        K = self.z_slots.shape[0]
        N = self.cam2world.shape[0]
        W, H, D = self.projection.frustum_size
        raws = self._masked_ray_major.view(K * N * H * W, D, 4)
        _, _, _, mask_map = raw2outputs(raws, self.z_vals, self.ray_dir, render_mask=True)
        return mask_map.view(K, N, H, W)

    def select_slot(self, mask_idx):
        """index (0-based among fg slots) of the slot with the best IoU against fg_idx (uorf_manip_model.py:181-193)."""
        K = self.z_slots.shape[0]
        fg = self.fg_idx[0].to(self.device).bool()
        this = mask_idx[0:1][None] == torch.arange(1, K, device=self.device).view(K - 1, 1, 1, 1)     # (K-1)x1xHxW
        inter = (this & fg).flatten(1).sum(1).float()
        union = (this | fg).flatten(1).sum(1).float()
        self.ious = inter / union.clamp(min=1.0)     # an empty slot scores 0 (0/0 in the reference)
        return self.ious.argmax()

    def forward(self, epoch=0, move_slot_idx=None):
        with torch.no_grad():
            if self.layout != "native":
                raise ValueError("uorfManipModel renders in the native ray-major layout")
            self._forward_impl()
            self.our_recon = self.x_recon[0]
            self.input_view = self.x[0]
            K = self.z_slots.shape[0]
            self.mask_maps = self.slot_mask_maps()
            mask_idx = self.mask_maps.argmax(dim=0)                      # NxHxW, on the device
            self.mask_idx_pred = mask_idx
            if move_slot_idx is None:
                move_slot_idx = self.select_slot(mask_idx)
            self.move_slot_idx = move_slot_idx
            # per-slot offset table: zero except the moved object (movement / nss_scale, uorf_manip_model.py:203)
            offsets = torch.zeros(K - 1, 3, device=self.device)
            idx = torch.as_tensor(move_slot_idx, device=self.device).view(1)
            offsets.index_copy_(0, idx, (self.movement[0:1] / self.opt.nss_scale).float())
            out = self.render(self.z_slots, self.cam2world, self.nss2cam0, fg_slot_offset=offsets)
            self.rendered_moved = out["rendered"]
            self.x_moved = out["rendered"] * 2 - 1
            self.our_moved = self.x_moved[0]
            self.masked_moved = out["masked"]
            if hasattr(self, "gt_moved") and self.gt_moved.shape[-1] == self.x_moved.shape[-1]:
                self.loss_moved = self.L2_loss(self.x_moved, self.gt_moved)
        return self.x_moved
