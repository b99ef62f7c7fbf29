"""CPU oracle for the uORF per-scene render hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional, fp32, CPU restatement of the algorithm of the reference
(KovenYu/uORF) for the path

    Encoder -> SlotAttention -> Projection.construct_sampling_coor -> Decoder -> raw2outputs

It exists so the CUDA path can be checked on a machine where /root/reference does not
exist (the GPU box).  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product package
``uorf_b200`` never imports it.

Pinning: the reference ships no tests, no golden vectors and no checkpoints
(SURVEY.md section 4), so "parity unpinned" by reference-owned fixtures.  The pin used
instead is the live reference itself: ``tests/golden/make_golden.py`` imports the
unmodified modules from /root/reference in the build container, runs them on seeded
inputs and commits inputs+outputs under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against those vectors.

All parameter dictionaries use the reference's state-dict key names
(e.g. ``f_before.0.weight``), so a reference checkpoint can be fed in directly.

Each function cites the reference file:line it restates (paths relative to
/root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# positional encoding                                            models/model.py:261-277
# --------------------------------------------------------------------------------------
def posenc(x: Tensor, n_freq: int = 5, keep_ori: bool = True) -> Tensor:
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(n-1) x), cos(2^(n-1) x)] along dim 1."""
    cols = [x] if keep_ori else []
    for i in range(n_freq):
        f = float(2 ** i)
        cols += [torch.sin(f * x), torch.cos(f * x)]
    return torch.cat(cols, dim=1)


# --------------------------------------------------------------------------------------
# Encoder                                                         models/model.py:11-61
# --------------------------------------------------------------------------------------
def encoder_forward(p: Params, x: Tensor, bottom: bool = False) -> Tensor:
    """U-Net on image + 4 pixel-distance planes.  x: Bx3xHxW -> BxZxH'xW'."""
    B, _, H, W = x.shape
    xs = torch.linspace(-1, 1, W)
    ys = torch.linspace(-1, 1, H)
    y1, x1 = torch.meshgrid(ys, xs, indexing="ij")
    planes = torch.stack([x1, 2 - x1, y1, 2 - y1]).unsqueeze(0).to(x)  # model.py:46-48
    h = torch.cat([x, planes.expand(B, -1, -1, -1)], dim=1)

    def conv(name, t, stride):
        return F.relu(F.conv2d(t, p[name + ".0.weight"], p[name + ".0.bias"], stride=stride, padding=1))

    def up(t):
        return F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)

    if bottom:
        d1 = conv("enc_down_1", conv("enc_down_0", h, 1), 2)
    else:
        d1 = conv("enc_down_1", h, 1)
    d2 = conv("enc_down_2", d1, 2)
    d3 = conv("enc_down_3", d2, 2)
    u3 = up(conv("enc_up_3", d3, 1))
    u2 = up(conv("enc_up_2", torch.cat([u3, d2], 1), 1))
    return conv("enc_up_1", torch.cat([u2, d1], 1), 1)


# --------------------------------------------------------------------------------------
# SlotAttention                                                 models/model.py:205-258
# --------------------------------------------------------------------------------------
def _layer_norm(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def _gru_cell(p: Params, pre: str, x: Tensor, h: Tensor) -> Tensor:
    """torch.nn.GRUCell semantics (gate order r, z, n)."""
    gi = x @ p[pre + ".weight_ih"].t() + p[pre + ".bias_ih"]
    gh = h @ p[pre + ".weight_hh"].t() + p[pre + ".bias_hh"]
    ir, iz, inn = gi.chunk(3, dim=-1)
    hr, hz, hn = gh.chunk(3, dim=-1)
    r = torch.sigmoid(ir + hr)
    z = torch.sigmoid(iz + hz)
    n = torch.tanh(inn + r * hn)
    return (1 - z) * n + z * h


def _res_mlp(p: Params, pre: str, s: Tensor) -> Tensor:
    t = _layer_norm(s, p[pre + ".0.weight"], p[pre + ".0.bias"])
    t = F.relu(t @ p[pre + ".1.weight"].t() + p[pre + ".1.bias"])
    return t @ p[pre + ".3.weight"].t() + p[pre + ".3.bias"]


def slot_attention_forward(p: Params, feat: Tensor, num_slots: int, iters: int = 3,
                           noise_fg: Optional[Tensor] = None, noise_bg: Optional[Tensor] = None,
                           eps: float = 1e-8) -> Tuple[Tensor, Tensor]:
    """feat BxNxC -> (slots BxKxC [bg first], attn BxKxN of the last iteration).

    noise_fg (B,K-1,C) / noise_bg (B,1,C) stand in for the two ``randn_like`` draws at
    model.py:216,219 (drawn in that order)."""
    B, N, C = feat.shape
    K = num_slots
    if noise_fg is None:
        noise_fg = torch.randn(B, K - 1, C)
    if noise_bg is None:
        noise_bg = torch.randn(B, 1, C)
    slot_fg = p["slots_mu"] + p["slots_logsigma"].exp() * noise_fg
    slot_bg = p["slots_mu_bg"] + p["slots_logsigma_bg"].exp() * noise_bg
    scale = C ** -0.5

    f = _layer_norm(feat, p["norm_feat.weight"], p["norm_feat.bias"])
    k = f @ p["to_k.weight"].t()
    v = f @ p["to_v.weight"].t()

    attn = None
    for _ in range(iters):
        q_fg = _layer_norm(slot_fg, p["to_q.0.weight"], p["to_q.0.bias"]) @ p["to_q.1.weight"].t()
        q_bg = _layer_norm(slot_bg, p["to_q_bg.0.weight"], p["to_q_bg.0.bias"]) @ p["to_q_bg.1.weight"].t()
        q = torch.cat([q_bg, q_fg], dim=1)                       # B,K,C (bg is slot 0)
        dots = torch.einsum("bid,bjd->bij", q, k) * scale          # B,K,N
        attn = dots.softmax(dim=1) + eps                           # softmax over SLOTS
        w = attn / attn.sum(dim=-1, keepdim=True)                  # each slot normalised over tokens
        upd = torch.einsum("bjd,bij->bid", v, w)                   # B,K,C
        new_bg = _gru_cell(p, "gru_bg", upd[:, :1].reshape(-1, C), slot_bg.reshape(-1, C)).reshape(B, 1, C)
        slot_bg = new_bg + _res_mlp(p, "to_res_bg", new_bg)
        new_fg = _gru_cell(p, "gru", upd[:, 1:].reshape(-1, C), slot_fg.reshape(-1, C)).reshape(B, K - 1, C)
        slot_fg = new_fg + _res_mlp(p, "to_res", new_fg)
    return torch.cat([slot_bg, slot_fg], dim=1), attn


# --------------------------------------------------------------------------------------
# Projection                                                    models/projection.py:6-113
# --------------------------------------------------------------------------------------
class FrustumSpec:
    """Intrinsics of the reference's Projection object (projection.py:7-31)."""

    def __init__(self, frustum_size: Sequence[int], near: float = 6.0, far: float = 20.0,
                 nss_scale: float = 7.0, render_size: int = 64,
                 focal_ratio: Tuple[float, float] = (350.0 / 320.0, 350.0 / 240.0)):
        self.W, self.H, self.D = (int(s) for s in frustum_size)
        self.near, self.far, self.nss_scale = near, far, nss_scale
        self.render_size = int(render_size)
        fx, fy = focal_ratio[0] * self.W, focal_ratio[1] * self.H
        cx, cy = (self.W - 1.0) / 2.0, (self.H - 1.0) / 2.0
        K = torch.tensor([[fx, 0, cx, 0], [0, fy, cy, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=torch.float32)
        self.cam2spixel = K
        self.spixel2cam = K.inverse()
        s = 1.0 / nss_scale
        self.world2nss = torch.tensor([[s, 0, 0, 0], [0, s, 0, 0], [0, 0, s, 0], [0, 0, 0, 1]],
                                      dtype=torch.float32).unsqueeze(0)


def _strided_chunks(t: Tensor, scale: int, hdim: int, wdim: int) -> Tensor:
    """Stack the scale^2 strided sub-grids (chunk j <-> divmod(j, scale)) on a new dim 0."""
    parts = []
    for j in range(scale * scale):
        h0, w0 = divmod(j, scale)
        idx = [slice(None)] * t.dim()
        idx[hdim] = slice(h0, None, scale)
        idx[wdim] = slice(w0, None, scale)
        parts.append(t[tuple(idx)])
    return torch.stack(parts, dim=0)


def sampling_coords(spec: FrustumSpec, cam2world: Tensor, partitioned: bool = False):
    """projection.py:33-113.  cam2world Nx4x4 ->
    coords (N*D*H*W)x3 [flat index ((n*D+d)*H+h)*W+w], z_vals (N*H*W)xD, ray_dir (N*H*W)x3;
    with ``partitioned`` a leading dim of scale^2 strided chunks (scale = H // render_size)."""
    N = cam2world.shape[0]
    W, H, D = spec.W, spec.H, spec.D
    gx, gy, gz = torch.meshgrid(torch.arange(W), torch.arange(H), torch.arange(D), indexing="ij")
    gx, gy, gz = gx.reshape(-1), gy.reshape(-1), gz.reshape(-1)
    zc = torch.linspace(spec.near, spec.far, D)[gz]
    pix = torch.stack([gx * zc, gy * zc, zc, torch.ones_like(zc)])           # 4 x (W*H*D)
    cam = torch.matmul(spec.spixel2cam, pix.float())
    world = torch.matmul(cam2world, cam)                                       # N x 4 x WHD
    nss = torch.matmul(spec.world2nss, world)
    nss = nss.view(N, 4, W, H, D).permute(0, 4, 3, 2, 1)[..., :3]              # N D H W 3
    scale = H // spec.render_size

    zv = ((cam[2] - spec.near) / (spec.far - spec.near)).expand(N, W * H * D)
    zv = zv.view(N, W, H, D).permute(0, 2, 1, 3)                               # N H W D

    px, py = torch.meshgrid(torch.arange(W), torch.arange(H), indexing="ij")
    # note the reference stacks [Y, X, 1] built with 'ij' indexing and then views the
    # flattened (W*H) axis as (H, W): projection.py:96-102.  Valid because W == H.
    pixc = torch.stack([py, px, torch.ones_like(px)]).flatten(1).float()
    rd = torch.matmul(spec.spixel2cam[:3, :3], pixc).permute(1, 0).view(H, W, 3)
    rd = rd.expand(N, H, W, 3)

    if partitioned:
        coords = _strided_chunks(nss, scale, 2, 3).flatten(1, 4)
        zv = _strided_chunks(zv, scale, 1, 2).flatten(1, 3)
        rd = _strided_chunks(rd, scale, 1, 2).flatten(1, 3)
    else:
        coords = nss.flatten(0, 3)
        zv = zv.flatten(0, 2)
        rd = rd.flatten(0, 2)
    return coords, zv, rd


# --------------------------------------------------------------------------------------
# Decoder                                                       models/model.py:106-161
# --------------------------------------------------------------------------------------
def _lin(p: Params, name: str, x: Tensor) -> Tensor:
    return F.linear(x, p[name + ".weight"], p[name + ".bias"])


def _mlp_trunk(p: Params, pre: str, inp: Tensor, n_layers: int) -> Tensor:
    """before-skip stack, concat [input, tmp], after-skip stack (ReLU after each)."""
    t = inp
    for i in range(n_layers):
        t = F.relu(_lin(p, f"{pre}_before.{2 * i}", t))
    t = torch.cat([inp, t], dim=1)
    for i in range(n_layers):
        t = F.relu(_lin(p, f"{pre}_after.{2 * i}", t))
    return t


def decoder_forward(p: Params, coor_bg: Tensor, coor_fg: Tensor, z_slots: Tensor, fg_transform: Tensor,
                    *, n_freq: int = 5, n_layers: int = 3, locality: bool = True,
                    locality_ratio: float = 4.0 / 7.0, fixed_locality: bool = False,
                    dens_noise: float = 0.0, noise: Optional[Tensor] = None, return_aux: bool = False):
    """coor_bg Px3, coor_fg (K-1)xPx3, z_slots KxC, fg_transform 1x3x3 (or 1x4x4 if
    fixed_locality) -> raws Px4, masked_raws KxPx4, unmasked_raws KxPx4, masks KxPx1.
    ``noise`` (KxPx1) replaces the randn_like draw at model.py:155."""
    K, C = z_slots.shape
    P = coor_bg.shape[0]
    if fixed_locality:                                                       # model.py:120-124
        outside = (coor_fg.abs() > locality_ratio).any(dim=-1)
        hom = torch.cat([coor_fg, torch.ones_like(coor_fg[..., :1])], dim=-1)
        coor_fg = torch.matmul(fg_transform[None], hom[..., None]).squeeze(-1)[..., :3]
    else:                                                                    # model.py:125-128
        coor_fg = torch.matmul(fg_transform[None], coor_fg[..., None]).squeeze(-1)
        outside = (coor_fg.abs() > locality_ratio).any(dim=-1)

    in_bg = torch.cat([posenc(coor_bg, n_freq), z_slots[0:1].expand(P, -1)], dim=1)
    in_fg = torch.cat([posenc(coor_fg.reshape(-1, 3), n_freq),
                       z_slots[1:, None, :].expand(-1, P, -1).reshape(-1, C)], dim=1)

    t_bg = _mlp_trunk(p, "b", in_bg, n_layers)
    bg_raw = _lin(p, f"b_after.{2 * n_layers}", t_bg).view(1, P, 4)

    t_fg = _mlp_trunk(p, "f", in_fg, n_layers)
    lat = _lin(p, "f_after_latent", t_fg)
    rgb = _lin(p, "f_color.2", F.relu(_lin(p, "f_color.0", lat))).view(K - 1, P, 3)
    sig = _lin(p, "f_after_shape", t_fg).view(K - 1, P)
    if locality:
        sig = torch.where(outside, torch.zeros_like(sig), sig)
    all_raw = torch.cat([bg_raw, torch.cat([rgb, sig[..., None]], dim=-1)], dim=0)   # K P 4

    dens = F.relu(all_raw[..., 3:])
    masks = dens / (dens.sum(dim=0) + 1e-5)
    rgb01 = (all_raw[..., :3].tanh() + 1) / 2
    if noise is None:
        noise = torch.randn_like(dens)
    sigma = dens + dens_noise * noise
    unmasked = torch.cat([rgb01, sigma], dim=2)
    masked = unmasked * masks
    raws = masked.sum(dim=0)
    if return_aux:
        return raws, masked, unmasked, masks, outside
    return raws, masked, unmasked, masks


# --------------------------------------------------------------------------------------
# raw2outputs                                                   models/model.py:279-313
# --------------------------------------------------------------------------------------
def composite(raw: Tensor, z_vals: Tensor, rays_d: Tensor, render_mask: bool = False):
    """raw RxDx4, z_vals RxD, rays_d Rx3 -> rgb_map Rx3, depth_map R, weights_norm RxD [, mask_map R]."""
    R = raw.shape[0]
    delta = torch.cat([z_vals[:, 1:] - z_vals[:, :-1], torch.full((R, 1), 1e-2, dtype=raw.dtype, device=raw.device)], dim=1)
    delta = delta * rays_d.norm(dim=-1, keepdim=True)
    alpha = 1.0 - torch.exp(-raw[..., 3] * delta)
    trans = torch.cumprod(torch.cat([torch.ones(R, 1, dtype=raw.dtype, device=raw.device), 1.0 - alpha + 1e-10], dim=1), dim=1)[:, :-1]
    w = alpha * trans
    rgb_map = (w[..., None] * raw[..., :3]).sum(dim=1)
    wn = w.detach() + 1e-5
    wn = wn / wn.sum(dim=-1, keepdim=True)
    depth = (wn * z_vals).sum(dim=-1)
    if render_mask:
        return rgb_map, depth, wn, (w * raw[..., 3]).sum(dim=1)
    return rgb_map, depth, wn


# --------------------------------------------------------------------------------------
# eval driver                                        models/uorf_eval_model.py:116-157,181-193
# --------------------------------------------------------------------------------------
def eval_render(enc_p: Params, sa_p: Params, dec_p: Params, x: Tensor, cam2world: Tensor, azi_rot: Tensor,
                spec: FrustumSpec, *, num_slots: int, input_size: int = 64, bottom: bool = False,
                n_freq: int = 5, n_layers: int = 3, attn_iter: int = 3, obj_scale: float = 4.5,
                fixed_locality: bool = False, noise_fg=None, noise_bg=None, with_slot_outputs: bool = True):
    """One scene through the ray-chunked eval path.  Returns dict with x_recon Nx3xHxW,
    depth NxHxW, z_slots, attn and (optionally) masked_raws / unmasked_raws K,N,D,H,W,4."""
    with torch.no_grad():
        nss2cam0 = (cam2world[0:1] if fixed_locality else azi_rot[0:1]).inverse()
        img0 = F.interpolate(x[0:1], size=input_size, mode="bilinear", align_corners=False)
        fmap = encoder_forward(enc_p, img0, bottom)
        feat = fmap.flatten(2).permute(0, 2, 1)
        slots, attn = slot_attention_forward(sa_p, feat, num_slots, attn_iter, noise_fg, noise_bg)
        slots, attn = slots[0], attn[0]
        K = num_slots
        N = cam2world.shape[0]
        W, H, D = spec.W, spec.H, spec.D
        scale = H // spec.render_size
        Hs, Ws = H // scale, W // scale
        coords, zv, rd = sampling_coords(spec, cam2world, partitioned=True)
        x_recon = torch.zeros(N, 3, H, W)
        depth = torch.zeros(N, H, W)
        masked_all = torch.zeros(K, N, D, H, W, 4) if with_slot_outputs else None
        unmasked_all = torch.zeros(K, N, D, H, W, 4) if with_slot_outputs else None
        for j in range(scale * scale):
            h0, w0 = divmod(j, scale)
            c = coords[j]
            raws, masked, unmasked, _ = decoder_forward(
                dec_p, c, c[None].expand(K - 1, -1, -1), slots, nss2cam0, n_freq=n_freq, n_layers=n_layers,
                locality=False, locality_ratio=obj_scale / spec.nss_scale, fixed_locality=fixed_locality,
                noise=torch.zeros(K, c.shape[0], 1))
            raws = raws.view(N, D, Hs, Ws, 4).permute(0, 2, 3, 1, 4).flatten(0, 2)
            rgb, dep, _ = composite(raws, zv[j], rd[j])
            x_recon[..., h0::scale, w0::scale] = rgb.view(N, Hs, Ws, 3).permute(0, 3, 1, 2) * 2 - 1
            depth[:, h0::scale, w0::scale] = dep.view(N, Hs, Ws)
            if with_slot_outputs:
                masked_all[..., h0::scale, w0::scale, :] = masked.view(K, N, D, Hs, Ws, 4)
                unmasked_all[..., h0::scale, w0::scale, :] = unmasked.view(K, N, D, Hs, Ws, 4)
        return dict(x_recon=x_recon, depth=depth, z_slots=slots, attn=attn,
                    masked_raws=masked_all, unmasked_raws=unmasked_all)


def slot_mask_maps(masked_raws: Tensor, z_vals: Tensor, ray_dir: Tensor):
    """uorf_eval_model.py:181-193: per-slot composited rgb + mask_map and the argmax segmentation.
    masked_raws K,N,D,H,W,4 ; z_vals (N*H*W)xD ; ray_dir (N*H*W)x3."""
    K, N, D, H, W, _ = masked_raws.shape
    rgbs, maps = [], []
    for k in range(K):
        r = masked_raws[k].permute(0, 2, 3, 1, 4).flatten(0, 2)
        rgb, _, _, m = composite(r, z_vals, ray_dir, render_mask=True)
        rgbs.append(rgb.view(N, H, W, 3).permute(0, 3, 1, 2) * 2 - 1)
        maps.append(m.view(N, H, W))
    maps = torch.stack(maps)
    return torch.stack(rgbs), maps, maps.argmax(dim=0)


# --------------------------------------------------------------------------------------
# training forward (no perceptual term)            models/uorf_nogan_model.py:128-183
# --------------------------------------------------------------------------------------
def train_forward(enc_p: Params, sa_p: Params, dec_p: Params, x: Tensor, cam2world: Tensor, azi_rot: Tensor,
                  spec: FrustumSpec, *, num_slots: int, stage: str = "coarse", input_size: int = 64,
                  supervision_size: int = 64, bottom: bool = False, n_freq: int = 5, n_layers: int = 3,
                  attn_iter: int = 3, obj_scale: float = 4.5, fixed_locality: bool = False,
                  locality: bool = True, dens_noise: float = 0.0, noise_fg=None, noise_bg=None,
                  dec_noise=None, crop: Tuple[int, int] = (0, 0)):
    """Differentiable forward of one training scene; returns (loss_recon, x_recon, rendered, x_target).
    The VGG perceptual term (uorf_nogan_model.py:181-183) needs downloaded weights and is
    outside the hot path; the driver adds it on top of ``rendered``."""
    nss2cam0 = (cam2world[0:1] if fixed_locality else azi_rot[0:1]).inverse()
    img0 = F.interpolate(x[0:1], size=input_size, mode="bilinear", align_corners=False)
    feat = encoder_forward(enc_p, img0, bottom).flatten(2).permute(0, 2, 1)
    slots, attn = slot_attention_forward(sa_p, feat, num_slots, attn_iter, noise_fg, noise_bg)
    slots = slots[0]
    K, N = num_slots, cam2world.shape[0]
    D = spec.D
    coords, zv, rd = sampling_coords(spec, cam2world)
    if stage == "coarse":
        tgt = F.interpolate(x, size=supervision_size, mode="bilinear", align_corners=False)
        H = W = supervision_size
    else:
        rs = spec.render_size
        h0, w0 = crop
        coords = coords.view(N, D, spec.H, spec.W, 3)[:, :, h0:h0 + rs, w0:w0 + rs].flatten(0, 3)
        zv = zv.view(N, spec.H, spec.W, D)[:, h0:h0 + rs, w0:w0 + rs].flatten(0, 2)
        rd = rd.view(N, spec.H, spec.W, 3)[:, h0:h0 + rs, w0:w0 + rs].flatten(0, 2)
        tgt = x[:, :, h0:h0 + rs, w0:w0 + rs]
        H = W = rs
    P = coords.shape[0]
    if dec_noise is None:
        dec_noise = torch.zeros(K, P, 1)
    raws, masked, unmasked, masks = decoder_forward(
        dec_p, coords, coords[None].expand(K - 1, -1, -1), slots, nss2cam0, n_freq=n_freq, n_layers=n_layers,
        locality=locality, locality_ratio=obj_scale / spec.nss_scale, fixed_locality=fixed_locality,
        dens_noise=dens_noise, noise=dec_noise)
    raws = raws.view(N, D, H, W, 4).permute(0, 2, 3, 1, 4).flatten(0, 2)
    rgb, _, _ = composite(raws, zv, rd)
    rendered = rgb.view(N, H, W, 3).permute(0, 3, 1, 2)
    x_recon = rendered * 2 - 1
    loss = F.mse_loss(x_recon, tgt)
    return loss, x_recon, rendered, tgt, dict(z_slots=slots, attn=attn[0], masked_raws=masked, unmasked_raws=unmasked)


# --------------------------------------------------------------------------------------
# synthetic scene + random-init weights (SURVEY.md section 8(c),(d))
# --------------------------------------------------------------------------------------
def lookat_cameras(n_views: int, radius: float = 12.0, elevation: float = 0.6, az0: float = 0.3):
    """cam2world Nx4x4 with columns [right, down, forward | position] and azi_rot Nx3x3
    (rotation about world z by the camera azimuth), cameras on a sphere looking at the origin."""
    c2w, azi = [], []
    for i in range(n_views):
        az = 2 * math.pi * i / n_views + az0
        pos = torch.tensor([radius * math.cos(elevation) * math.cos(az),
                            radius * math.cos(elevation) * math.sin(az),
                            radius * math.sin(elevation)], dtype=torch.float64)
        fwd = -pos / pos.norm()
        up = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64)
        right = torch.linalg.cross(fwd, up)
        right = right / right.norm()
        down = torch.linalg.cross(fwd, right)
        m = torch.eye(4, dtype=torch.float64)
        m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, down, fwd, pos
        c2w.append(m)
        azi.append(torch.tensor([[math.cos(az), -math.sin(az), 0], [math.sin(az), math.cos(az), 0], [0, 0, 1]],
                                dtype=torch.float64))
    return torch.stack(c2w).float(), torch.stack(azi).float()


def _xavier_normal(gen, out_f, in_f):
    std = math.sqrt(2.0 / (in_f + out_f))
    return torch.randn(out_f, in_f, generator=gen) * std


def init_decoder_params(z_dim: int = 64, n_freq: int = 5, n_layers: int = 3, seed: int = 0) -> Params:
    """Random decoder weights with the reference's init RULE (xavier_normal weights, zero bias:
    networks.py:86-95 via uorf_nogan_model.py:94-95).  Not the same RNG stream as the reference."""
    g = torch.Generator().manual_seed(seed)
    I = 6 * n_freq + 3 + z_dim
    p: Params = {}

    def add(name, o, i):
        p[name + ".weight"] = _xavier_normal(g, o, i)
        p[name + ".bias"] = torch.zeros(o)

    for pre in ("f", "b"):
        add(f"{pre}_before.0", z_dim, I)
        add(f"{pre}_after.0", z_dim, I + z_dim)
        for i in range(1, n_layers):
            add(f"{pre}_before.{2 * i}", z_dim, z_dim)
            add(f"{pre}_after.{2 * i}", z_dim, z_dim)
    add(f"b_after.{2 * n_layers}", 4, z_dim)
    add("f_after_latent", z_dim, z_dim)
    add("f_after_shape", 1, z_dim)
    add("f_color.0", z_dim // 4, z_dim)
    add("f_color.2", 3, z_dim // 4)
    return p


def init_encoder_params(z_dim: int = 64, bottom: bool = False, seed: int = 1) -> Params:
    """N(0, 0.02) conv weights, zero bias (networks.py:84-85)."""
    g = torch.Generator().manual_seed(seed)
    p: Params = {}

    def add(name, o, i):
        p[name + ".0.weight"] = torch.randn(o, i, 3, 3, generator=g) * 0.02
        p[name + ".0.bias"] = torch.zeros(o)

    if bottom:
        add("enc_down_0", z_dim, 7)
        add("enc_down_1", z_dim, z_dim)
    else:
        add("enc_down_1", z_dim, 7)
    add("enc_down_2", z_dim, z_dim)
    add("enc_down_3", z_dim, z_dim)
    add("enc_up_3", z_dim, z_dim)
    add("enc_up_2", z_dim, 2 * z_dim)
    add("enc_up_1", z_dim, 2 * z_dim)
    return p


def init_slot_attention_params(z_dim: int = 64, hidden: int = 128, seed: int = 2) -> Params:
    """Linear weights N(0,0.02)/zero bias; GRUCell U(-1/sqrt(C),1/sqrt(C)); LayerNorm (1,0);
    slots_mu randn, slots_logsigma xavier_uniform (model.py:172-177, networks.py:84-85)."""
    g = torch.Generator().manual_seed(seed)
    C, Hd = z_dim, max(z_dim, hidden)
    p: Params = {}
    # torch computes the fans of a (1,1,C) tensor as fan_in = fan_out = C  -> bound sqrt(6/(2C))
    bound = math.sqrt(6.0 / (2 * C))
    for suf in ("", "_bg"):
        p["slots_mu" + suf] = torch.randn(1, 1, C, generator=g)
        p["slots_logsigma" + suf] = (torch.rand(1, 1, C, generator=g) * 2 - 1) * bound
    for name in ("to_k", "to_v"):
        p[name + ".weight"] = torch.randn(C, C, generator=g) * 0.02
    for name in ("to_q", "to_q_bg"):
        p[name + ".0.weight"], p[name + ".0.bias"] = torch.ones(C), torch.zeros(C)
        p[name + ".1.weight"] = torch.randn(C, C, generator=g) * 0.02
    k = 1.0 / math.sqrt(C)
    for name in ("gru", "gru_bg"):
        p[name + ".weight_ih"] = (torch.rand(3 * C, C, generator=g) * 2 - 1) * k
        p[name + ".weight_hh"] = (torch.rand(3 * C, C, generator=g) * 2 - 1) * k
        p[name + ".bias_ih"] = (torch.rand(3 * C, generator=g) * 2 - 1) * k
        p[name + ".bias_hh"] = (torch.rand(3 * C, generator=g) * 2 - 1) * k
    for name in ("to_res", "to_res_bg"):
        p[name + ".0.weight"], p[name + ".0.bias"] = torch.ones(C), torch.zeros(C)
        p[name + ".1.weight"] = torch.randn(Hd, C, generator=g) * 0.02
        p[name + ".1.bias"] = torch.zeros(Hd)
        p[name + ".3.weight"] = torch.randn(C, Hd, generator=g) * 0.02
        p[name + ".3.bias"] = torch.zeros(C)
    p["norm_feat.weight"], p["norm_feat.bias"] = torch.ones(C), torch.zeros(C)
    return p


def synthetic_scene(n_views: int, load_size: int, seed: int = 2021):
    """img_data in [-1,1] (N,3,L,L), cam2world, azi_rot - the recipe of SURVEY.md 8(d)."""
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(n_views, 3, load_size, load_size, generator=g) * 2 - 1
    c2w, azi = lookat_cameras(n_views)
    return img, c2w, azi
