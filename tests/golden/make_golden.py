"""Generate golden vectors by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/*.npz (inputs + reference outputs, fp32).  The oracle
(oracle/uorf_oracle.py) and the CUDA path are both checked against these files.
The reference has no tests/fixtures of its own - these vectors are the pin.
"""
import argparse
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("UORF_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))


def install_stubs():
    """The four stubs listed in SURVEY.md 8(c): vgg16 download, lpips, piq, util.util."""
    import torchvision
    lp = types.ModuleType("lpips")

    class LPIPS(torch.nn.Module):
        def forward(self, a, b):
            return ((a - b) ** 2).mean(dim=(1, 2, 3))
    lp.LPIPS = LPIPS
    sys.modules["lpips"] = lp
    pq = types.ModuleType("piq")
    pq.ssim = lambda a, b, data_range=1.: torch.zeros(())
    pq.psnr = lambda a, b, data_range=1.: torch.zeros(())
    sys.modules["piq"] = pq
    uu = types.ModuleType("util.util")

    class AverageMeter:
        def __init__(self):
            self.val = 0
            self.n = 0

        def update(self, v, n=1):
            self.val = (self.val * self.n + v * n) / (self.n + n)
            self.n += n
    uu.AverageMeter = AverageMeter
    up = types.ModuleType("util")
    up.util = uu
    up.__path__ = []
    sys.modules["util"] = up
    sys.modules["util.util"] = uu
    import models.model as mm
    mm.vgg16 = lambda pretrained=True: torchvision.models.vgg16(weights=None)


def sd_np(sd, prefix):
    return {f"{prefix}/{k}": v.detach().cpu().numpy() for k, v in sd.items()}


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(OUT, name + ".npz")
    np.savez(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1e3:.1f} KB")


def make_opt(model_name, is_train, **over):
    from models import get_option_setter
    p = argparse.ArgumentParser()
    # base options the model code reads
    p.add_argument("--batch_size", type=int, default=1)
    p.add_argument("--lr", type=float, default=3e-4)
    p.add_argument("--niter", type=int, default=100)
    p.add_argument("--niter_decay", type=int, default=0)
    p.add_argument("--dataset_mode", default="multiscenes")
    p.add_argument("--custom_lr", action="store_true")
    p.add_argument("--lr_policy", default="warmup")
    p.add_argument("--exp_id", default="golden")
    p = get_option_setter(model_name)(p, is_train)
    opt = p.parse_args([])
    opt.gpu_ids, opt.isTrain, opt.model = [], is_train, model_name
    opt.checkpoints_dir, opt.name = "/tmp/uorf_golden_ckpt", "golden"
    opt.verbose, opt.continue_train, opt.load_iter, opt.epoch, opt.epoch_count = False, False, 0, "latest", 1
    for k, v in over.items():
        setattr(opt, k, v)
    return opt


def main():
    install_stubs()
    from models.model import Encoder, Decoder, SlotAttention, raw2outputs, sin_emb
    from models.projection import Projection
    from models import networks, create_model
    from oracle import uorf_oracle as O

    # ---- sin_emb ------------------------------------------------------------------
    torch.manual_seed(10)
    x = (torch.rand(257, 3) * 2 - 1) * 2.0
    save("sin_emb", x=x, out=sin_emb(x, n_freq=5))

    # ---- projection ---------------------------------------------------------------
    c2w, azi = O.lookat_cameras(2)
    for tag, fs, rs in (("proj_8x8x6_rs4", [8, 8, 6], 4), ("proj_16x16x8_rs16", [16, 16, 8], 16)):
        pr = Projection(frustum_size=fs, near=6, far=20, nss_scale=7, render_size=(rs, rs), device="cpu")
        c, z, r = pr.construct_sampling_coor(c2w)
        cp, zp, rp = pr.construct_sampling_coor(c2w, partitioned=True)
        save(tag, cam2world=c2w, frustum=np.array(fs), render_size=rs, coords=c, z_vals=z.contiguous(),
             ray_dir=r.contiguous(), coords_p=cp, z_vals_p=zp.contiguous(), ray_dir_p=rp.contiguous())

    # ---- raw2outputs --------------------------------------------------------------
    torch.manual_seed(11)
    R, D = 37, 16
    raw = torch.rand(R, D, 4)
    raw[..., 3] = torch.relu(torch.randn(R, D)) * 3
    zv = torch.linspace(0, 1, D)[None].expand(R, D).contiguous()
    rd = torch.randn(R, 3)
    rgb, dep, wn, mm_ = raw2outputs(raw, zv, rd, render_mask=True)
    save("raw2outputs", raw=raw, z_vals=zv, rays_d=rd, rgb_map=rgb, depth_map=dep, weights_norm=wn, mask_map=mm_)

    # ---- decoder ------------------------------------------------------------------
    for Z, K, P, seed in ((64, 3, 203, 12), (40, 4, 130, 13)):
        for fixed in (False, True):
            torch.manual_seed(seed)
            dec = networks.init_net(Decoder(n_freq=5, input_dim=33 + Z, z_dim=Z, n_layers=3, locality=True,
                                            locality_ratio=4.5 / 7, fixed_locality=fixed), init_type="xavier")
            # non-zero biases make the test stricter than the reference's zero-bias init
            with torch.no_grad():
                for n_, p_ in dec.named_parameters():
                    if n_.endswith("bias"):
                        p_.copy_(torch.randn_like(p_) * 0.1)
            cb = (torch.rand(P, 3) * 2 - 1) * 1.2
            cf = cb[None].expand(K - 1, -1, -1).clone()
            cf[0] += 0.05  # one shifted slot, as uorf_manip_model.py:202-203 does
            zs = torch.randn(K, Z)
            T = O.lookat_cameras(1)[0][0:1].inverse() if fixed else O.lookat_cameras(1)[1][0:1].inverse()
            outs = {}
            for loc in (True, False):
                dec.locality = loc
                torch.manual_seed(seed + 100)
                noise = torch.randn(K, P, 1)
                torch.manual_seed(seed + 100)
                with torch.no_grad():
                    raws, masked, unmasked, masks = dec(cb, cf, zs, T, dens_noise=0.3)
                outs[f"raws_loc{int(loc)}"] = raws
                outs[f"masked_loc{int(loc)}"] = masked
                outs[f"unmasked_loc{int(loc)}"] = unmasked
                outs[f"masks_loc{int(loc)}"] = masks
                outs["noise"] = noise
            # gradients (locality on, no noise) of a scalar probe loss
            dec.locality = True
            zs_g = zs.clone().requires_grad_(True)
            torch.manual_seed(0)
            raws, masked, unmasked, masks = dec(cb, cf, zs_g, T, dens_noise=0.)
            probe = torch.randn(P, 4)
            (raws * probe).sum().backward()
            grads = {f"grad/{n_}": p_.grad for n_, p_ in dec.named_parameters()}
            save(f"decoder_z{Z}_fixed{int(fixed)}", coor_bg=cb, coor_fg=cf, z_slots=zs, fg_transform=T,
                 dens_noise=0.3, locality_ratio=4.5 / 7, probe=probe, grad_z_slots=zs_g.grad, raws_nonoise=raws,
                 **sd_np(dec.state_dict(), "dec"), **grads, **outs)

    # ---- slot attention -------------------------------------------------------------
    for Z, K, Ntok, seed in ((64, 5, 256, 14), (40, 8, 100, 15)):
        torch.manual_seed(seed)
        sa = networks.init_net(SlotAttention(num_slots=K, in_dim=Z, slot_dim=Z, iters=3), init_type="normal")
        feat = torch.randn(1, Ntok, Z, requires_grad=True)
        torch.manual_seed(seed + 100)
        nfg = torch.randn(1, K - 1, Z)
        nbg = torch.randn(1, 1, Z)
        torch.manual_seed(seed + 100)
        slots, attn = sa(feat)
        ps, pa = torch.randn_like(slots), torch.randn_like(attn)
        ((slots * ps).sum() + (attn * pa).sum()).backward()
        grads = {f"grad/{n_}": p_.grad for n_, p_ in sa.named_parameters()}
        save(f"slot_attention_z{Z}", feat=feat, noise_fg=nfg, noise_bg=nbg, slots=slots, attn=attn, K=K,
             probe_slots=ps, probe_attn=pa, grad_feat=feat.grad, **sd_np(sa.state_dict(), "sa"), **grads)

    # ---- encoder ------------------------------------------------------------------
    for bottom, S in ((False, 32), (True, 64)):
        torch.manual_seed(16)
        enc = networks.init_net(Encoder(3, z_dim=8, bottom=bottom), init_type="normal")
        xi = torch.rand(1, 3, S, S) * 2 - 1
        with torch.no_grad():
            fm = enc(xi)
        save(f"encoder_bottom{int(bottom)}", x=xi, out=fm, **sd_np(enc.state_dict(), "enc"))

    # ---- eval driver (uorf_eval_model.py:116-157 + compute_visuals:181-193) -----------
    torch.manual_seed(17)
    opt = make_opt("uorf_eval", False, num_slots=4, z_dim=40, n_samp=8, frustum_size=16, render_size=8,
                   input_size=32, n_img_each_scene=2)
    model = create_model(opt)
    img, c2w, azi = O.synthetic_scene(2, 16, seed=2021)
    model.set_input({"img_data": img, "cam2world": c2w, "azi_rot": azi, "paths": []})
    torch.manual_seed(117)
    nfg, nbg = torch.randn(1, 3, 40), torch.randn(1, 1, 40)
    torch.manual_seed(117)
    with torch.no_grad():
        model.forward()
    x_recon = torch.stack([getattr(model, f"x_rec{i}") for i in range(2)])
    # per-slot mask maps exactly as compute_visuals does (lines 181-193), without the GT-mask part
    _, zv, rd = model.projection.construct_sampling_coor(c2w)
    maps = []
    for k in range(4):
        r = model.masked_raws[k].permute([0, 2, 3, 1, 4]).flatten(0, 2)
        _, _, _, m = raw2outputs(r, zv, rd, render_mask=True)
        maps.append(m.view(2, 16, 16))
    maps = torch.stack(maps)
    save("eval_driver", img=img, cam2world=c2w, azi_rot=azi, noise_fg=nfg, noise_bg=nbg, x_recon=x_recon,
         masked_raws=model.masked_raws, unmasked_raws=model.unmasked_raws, mask_maps=maps,
         mask_idx=maps.argmax(0), **sd_np(model.netEncoder.state_dict(), "enc"),
         **sd_np(model.netSlotAttention.state_dict(), "sa"), **sd_np(model.netDecoder.state_dict(), "dec"))

    # ---- train driver (uorf_nogan_model.py:128-183, 223-227), coarse and fine ---------
    for stage in ("coarse", "fine"):
        torch.manual_seed(18)
        opt = make_opt("uorf_nogan", True, num_slots=3, z_dim=40, n_samp=8, frustum_size=8, frustum_size_fine=16,
                       render_size=8, supervision_size=8, input_size=32, n_img_each_scene=2, stage=stage,
                       bottom=(stage == "fine"))
        model = create_model(opt)
        L = 16
        img, c2w, azi = O.synthetic_scene(2, L, seed=2022)
        model.set_input({"img_data": img, "cam2world": c2w, "azi_rot": azi, "paths": []})
        torch.manual_seed(118)
        nfg, nbg = torch.randn(1, 2, 40), torch.randn(1, 1, 40)
        crop = np.array([0, 0])
        if stage == "fine":
            hh = torch.randint(low=0, high=8, size=(1,))
            ww = torch.randint(low=0, high=8, size=(1,))
            crop = np.array([int(hh), int(ww)])
        torch.manual_seed(118)
        model.forward(epoch=0)
        for o in model.optimizers:
            o.zero_grad()
        model.backward()
        grads = {}
        for pre, net in (("enc", model.netEncoder), ("sa", model.netSlotAttention), ("dec", model.netDecoder)):
            for n_, p_ in net.named_parameters():
                grads[f"grad/{pre}/{n_}"] = p_.grad if p_.grad is not None else torch.zeros_like(p_)
        x_recon = torch.stack([getattr(model, f"x_rec{i}") for i in range(2)])
        save(f"train_driver_{stage}", img=img, cam2world=c2w, azi_rot=azi, noise_fg=nfg, noise_bg=nbg, crop=crop,
             loss_recon=model.loss_recon, x_recon=x_recon,
             **sd_np(model.netEncoder.state_dict(), "enc"), **sd_np(model.netSlotAttention.state_dict(), "sa"),
             **sd_np(model.netDecoder.state_dict(), "dec"), **grads)


if __name__ == "__main__":
    main()
