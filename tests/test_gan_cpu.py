"""Adversarial-training pieces (discriminator + losses, reference models/model.py:327-349, 494-549) against golden vectors
produced by the unmodified reference (tests/golden/make_golden_disc.py).  These parts are plain torch, so they are checked
on the CPU; the driver itself needs the CUDA kernels (tests/test_gpu_kernels.py::test_gan_driver_*)."""
import os

import numpy as np
import torch

from uorf_b200.gan_model import Discriminator, d_logistic_loss, d_r1_loss, g_nonsaturating_loss

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "discriminator.npz"))


def _disc():
    d = Discriminator(size=int(G["size"]), ndf=int(G["ndf"]))
    sd = {k[len("disc/"):]: torch.from_numpy(G[k]) for k in G.files if k.startswith("disc/")}
    assert set(sd) == set(d.state_dict()), "state-dict keys differ from the reference Discriminator"
    d.load_state_dict(sd)
    return d


def test_discriminator_forward_and_losses_match_reference():
    d = _disc()
    real, fake = torch.from_numpy(G["real"]), torch.from_numpy(G["fake"])
    rp, fp = d(real), d(fake)
    assert torch.allclose(rp, torch.from_numpy(G["real_pred"]), atol=1e-5, rtol=1e-5)
    assert torch.allclose(fp, torch.from_numpy(G["fake_pred"]), atol=1e-5, rtol=1e-5)
    lr_, lf = d_logistic_loss(rp, fp)
    assert abs(float(lr_) - float(G["loss_real"])) < 1e-5 and abs(float(lf) - float(G["loss_fake"])) < 1e-5
    assert abs(float(g_nonsaturating_loss(fp)) - float(G["g_loss"])) < 1e-5
    d.zero_grad()
    (lr_ + lf).backward()
    for n, p in d.named_parameters():
        want = torch.from_numpy(G["grad/" + n])
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        assert torch.allclose(got, want, atol=1e-5 + 1e-4 * want.abs().max().item()), n


def test_r1_penalty_matches_reference():
    d = _disc()
    real = torch.from_numpy(G["real"]).clone().requires_grad_(True)
    r1 = d_r1_loss(d(real), real)
    assert abs(float(r1) - float(G["r1"])) < 1e-5 * max(1.0, abs(float(G["r1"])))
    d.zero_grad()
    r1.backward()
    for n, p in d.named_parameters():
        want = torch.from_numpy(G["r1grad/" + n])
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        assert torch.allclose(got, want, atol=1e-5 + 1e-4 * want.abs().max().item()), n
