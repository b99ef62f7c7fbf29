"""Golden vectors for the adversarial-training pieces (reference models/model.py:327-349, 494-549), produced by running the
UNMODIFIED reference Discriminator and loss functions on CPU.  Build container only:

    python tests/golden/make_golden_disc.py      # writes tests/golden/discriminator.npz
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as G  # noqa: E402  (puts /root/reference on sys.path)


def main():
    G.install_stubs()
    from models.model import Discriminator, d_logistic_loss, d_r1_loss, g_nonsaturating_loss
    torch.manual_seed(77)
    size, ndf = 16, 8
    disc = Discriminator(size=size, ndf=ndf)
    real = torch.rand(4, 3, size, size) * 2 - 1
    fake = torch.rand(4, 3, size, size) * 2 - 1
    real_pred, fake_pred = disc(real), disc(fake)
    l_real, l_fake = d_logistic_loss(real_pred, fake_pred)
    disc.zero_grad()
    (l_real + l_fake).backward()
    grads = {f"grad/{n}": (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for n, p in disc.named_parameters()}
    g_loss = g_nonsaturating_loss(disc(fake))
    real_r = real.clone().requires_grad_(True)
    r1 = d_r1_loss(disc(real_r), real_r)
    disc.zero_grad()
    r1.backward()
    r1_grads = {f"r1grad/{n}": (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for n, p in disc.named_parameters()}
    G.save("discriminator", size=size, ndf=ndf, real=real, fake=fake, real_pred=real_pred, fake_pred=fake_pred, loss_real=l_real,
           loss_fake=l_fake, g_loss=g_loss, r1=r1, **G.sd_np(disc.state_dict(), "disc"), **grads, **r1_grads)


if __name__ == "__main__":
    main()
