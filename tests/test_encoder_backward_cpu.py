"""The training-path encoder runs its forward through the direct-convolution kernels and its backward through
`modules._encoder_backward` (cuDNN convolution_backward per layer on the kept activations).  The backward chain is plain torch,
so it is checked here on the CPU against autograd through the reference layer stack (models/model.py:36-61); the kernels'
forward is checked on the GPU (tests/test_gpu_kernels.py::test_encoder_*)."""
import pytest
import torch
import torch.nn.functional as F

from uorf_b200.modules import Encoder, _encoder_backward


@pytest.mark.parametrize("Z,S,bottom", [(16, 16, False), (8, 32, True)])
def test_encoder_backward_chain_matches_autograd(Z, S, bottom):
    torch.manual_seed(11)
    enc = Encoder(3, z_dim=Z, bottom=bottom).double()
    for p in enc.parameters():
        torch.nn.init.normal_(p, 0.0, 0.2)
    x = (torch.rand(1, 3, S, S, dtype=torch.float64) * 2 - 1).requires_grad_(True)
    planes = enc._pixel_planes(S, S, x.device).double()
    # reference layer stack, keeping the activations the kernels' forward keeps (u3 / u2 before their upsample)
    x_ = torch.cat([x, planes], dim=1)
    d0 = enc.enc_down_0(x_) if bottom else None
    d1 = enc.enc_down_1(d0 if bottom else x_)
    d2 = enc.enc_down_2(d1)
    d3 = enc.enc_down_3(d2)
    u3 = enc.enc_up_3[1](enc.enc_up_3[0](d3))
    u2 = enc.enc_up_2[1](enc.enc_up_2[0](torch.cat([enc.enc_up_3[2](u3), d2], dim=1)))
    out = enc.enc_up_1(torch.cat([enc.enc_up_2[2](u2), d1], dim=1))
    probe = torch.randn_like(out)
    (out * probe).sum().backward()
    want = {n: p.grad.clone() for n, p in enc.named_parameters()}
    want_x = x.grad.clone()
    a = dict(img=x.detach()[0], planes=planes[0], d1=d1.detach()[0], d2=d2.detach()[0], d3=d3.detach()[0], u3=u3.detach()[0],
             u2=u2.detach()[0], out=out.detach()[0])
    if bottom:
        a["d0"] = d0.detach()[0]
    gx, grads = _encoder_backward(enc, a, probe.double(), True)
    assert set(grads) == set(want)
    for n in want:
        assert torch.allclose(grads[n], want[n], rtol=1e-9, atol=1e-9), n
    assert torch.allclose(gx, want_x, rtol=1e-9, atol=1e-9)
