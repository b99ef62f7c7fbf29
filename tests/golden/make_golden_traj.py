"""Golden TRAINING TRAJECTORY: 20 optimize_parameters() steps of the UNMODIFIED reference driver (uorf_nogan, coarse stage) on CPU.

    python tests/golden/make_golden_traj.py        (build container only: needs /root/reference)

Loop of the reference's train_without_gan.py:31-41 on one small synthetic scene: set_input -> optimize_parameters(epoch) ->
update_learning_rate, lr_policy 'warmup' (5 warm-up steps, so the learning rate is at 3e-4 from step 5 on).  The slot-attention
noise is the only random draw of a coarse step (models/model.py:216,219); it is pinned by re-seeding the global generator
before every step and recorded by drawing the same two tensors from the same seed.  Saved: initial weights, per-step noise,
per-step loss_recon, final weights.  Perceptual weight is 0 at epoch 0 (uorf_nogan_model.py:130), so the random-init VGG stub
has no influence on the trajectory.  Written to tests/golden/train_trajectory.npz.
"""
import os
import sys

import numpy as np
import torch

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
from oracle import ref_harness as R            # noqa: E402
from oracle import uorf_oracle as O            # noqa: E402

STEPS = 20


def sd_np(sd, prefix):
    return {f"{prefix}/{k}": v.detach().cpu().numpy().copy() for k, v in sd.items()}


def main():
    R.install()
    torch.manual_seed(31)
    model, opt = R.create_model("uorf_nogan", True, num_slots=3, z_dim=40, n_samp=8, frustum_size=8, frustum_size_fine=16,
                                render_size=8, supervision_size=8, input_size=32, n_img_each_scene=2, stage="coarse",
                                bottom=False, warmup_steps=5)
    img, c2w, azi = O.synthetic_scene(2, 16, seed=2023)
    init = {}
    for pre, net in (("enc", model.netEncoder), ("sa", model.netSlotAttention), ("dec", model.netDecoder)):
        init.update(sd_np(net.state_dict(), "init/" + pre))
    nfg, nbg, losses, lrs = [], [], [], []
    for i in range(STEPS):
        torch.manual_seed(1000 + i)
        nfg.append(torch.randn(1, 2, 40))
        nbg.append(torch.randn(1, 1, 40))
        torch.manual_seed(1000 + i)
        model.set_input({"img_data": img, "cam2world": c2w, "azi_rot": azi, "paths": []})
        lrs.append(model.optimizers[0].param_groups[0]["lr"])
        model.optimize_parameters(epoch=0)
        model.update_learning_rate()
        losses.append(float(model.loss_recon))
    final = {}
    for pre, net in (("enc", model.netEncoder), ("sa", model.netSlotAttention), ("dec", model.netDecoder)):
        final.update(sd_np(net.state_dict(), "final/" + pre))
    path = os.path.join(OUT, "train_trajectory.npz")
    np.savez(path, img=img.numpy(), cam2world=c2w.numpy(), azi_rot=azi.numpy(), noise_fg=torch.stack(nfg).numpy(),
             noise_bg=torch.stack(nbg).numpy(), loss_recon=np.array(losses, np.float64), lr=np.array(lrs, np.float64), **init, **final)
    print("train_trajectory:", os.path.getsize(path) / 1e3, "KB; loss", losses[0], "->", losses[-1], "lr", lrs[:7])


if __name__ == "__main__":
    main()
