"""Golden G-step / D-step of the UNMODIFIED reference adversarial driver (models/uorf_gan_model.py) on CPU.

    python tests/golden/make_golden_gan.py        (build container only: needs /root/reference)

One small synthetic scene, coarse stage, epoch = gan_train_epoch + gan_in so the adversarial weight is active
(uorf_gan_model.py:140); percept_in is set beyond the epoch, so the random-init VGG stub has no influence.  Sequence, as the
reference's train_with_gan.py:39-57 alternates them:
  1. generator step:   set_input -> forward(epoch) -> zero_grad -> backward()          (optimize_parameters without the step)
     recorded: loss_recon, loss_gan (unweighted, as backward() leaves it), x_recon, every Encoder / SlotAttention / Decoder gradient
  2. discriminator step: d_scheduler.step() twice (the warm-up schedule starts at lr = 0) -> forward_disc(it = 32)
     (it % 32 == 0: logistic step + R1 step); recorded: loss_d_real, loss_d_fake, the discriminator weights after the step
The slot-attention noise (models/model.py:216,219) is the only random draw of a coarse step; it is pinned by seeding the global
generator before each forward and recorded by drawing the same tensors from the same seed.
Written to tests/golden/gan_driver.npz.
"""
import os
import sys

import numpy as np
import torch

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
from oracle import ref_harness as R            # noqa: E402
from oracle import uorf_oracle as O            # noqa: E402


def sd_np(sd, prefix):
    return {f"{prefix}/{k}": v.detach().cpu().numpy().copy() for k, v in sd.items()}


def main():
    R.install()
    torch.manual_seed(47)
    # supervision 16: the reference discriminator only closes on a 4x4 map for sizes 16, 32, 64, 128
    cfg = dict(num_slots=3, z_dim=40, n_samp=8, frustum_size=16, frustum_size_fine=16, render_size=16, supervision_size=16,
               input_size=32, n_img_each_scene=2, stage="coarse", bottom=False, warmup_steps=5, ndf=8, percept_in=100,
               gan_train_epoch=2, gan_in=1)
    model, opt = R.create_model("uorf_gan", True, **cfg)
    model.d_scheduler = __import__("models.networks", fromlist=["get_scheduler"]).get_scheduler(model.d_optimizer, opt) \
        if not hasattr(model, "d_scheduler") else model.d_scheduler
    epoch = opt.gan_train_epoch + opt.gan_in
    img, c2w, azi = O.synthetic_scene(2, 16, seed=99)
    init = {}
    for pre, net in (("enc", model.netEncoder), ("sa", model.netSlotAttention), ("dec", model.netDecoder), ("disc", model.netDisc)):
        init.update(sd_np(net.state_dict(), "init/" + pre))
    K, Z = cfg["num_slots"], cfg["z_dim"]
    noise = {}
    for tag, seed in (("g", 501), ("d", 502)):
        torch.manual_seed(seed)
        noise[f"noise_fg_{tag}"] = torch.randn(1, K - 1, Z).numpy()
        noise[f"noise_bg_{tag}"] = torch.randn(1, 1, Z).numpy()
    # ---- generator step (no optimizer step: the gradients are the golden) ----
    model.set_input({"img_data": img, "cam2world": c2w, "azi_rot": azi, "paths": []})
    torch.manual_seed(501)
    model.forward(epoch)
    x_recon = torch.stack([getattr(model, f"x_rec{i}") for i in range(2)]).detach().clone()
    for o in model.optimizers:
        o.zero_grad()
    model.backward()
    grads = {}
    for pre, net in (("enc", model.netEncoder), ("sa", model.netSlotAttention), ("dec", model.netDecoder)):
        for n, p in net.named_parameters():
            grads[f"grad/{pre}/{n}"] = (p.grad if p.grad is not None else torch.zeros_like(p)).detach().numpy().copy()
    g_losses = dict(loss_recon=float(model.loss_recon), loss_gan=float(model.loss_gan), weight_gan=float(model.weight_gan))
    # ---- discriminator step ----
    model.d_scheduler.step()
    model.d_scheduler.step()
    d_lr = model.d_optimizer.param_groups[0]["lr"]
    torch.manual_seed(502)
    model.forward_disc(32)
    d_losses = dict(loss_d_real=float(model.loss_d_real), loss_d_fake=float(model.loss_d_fake), d_lr=d_lr)
    final = sd_np(model.netDisc.state_dict(), "final/disc")
    path = os.path.join(OUT, "gan_driver.npz")
    np.savez(path, img=img.numpy(), cam2world=c2w.numpy(), azi_rot=azi.numpy(), epoch=epoch, x_recon=x_recon.numpy(),
             **{k: np.float64(v) for k, v in {**g_losses, **d_losses}.items()}, **noise, **init, **grads, **final)
    print("gan_driver:", os.path.getsize(path) / 1e3, "KB", g_losses, d_losses)


if __name__ == "__main__":
    main()
