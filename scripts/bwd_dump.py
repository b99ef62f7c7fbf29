"""decoder backward at the training size with a dense random d raws: gradients saved to argv[1] (variant comparison), run twice to check reproducibility"""
import sys
import torch
from uorf_b200.modules import Decoder

dev = torch.device("cuda:0")
torch.manual_seed(7)
K, Z, P = 8, 64, 4 * 64 * 64 * 64
dec = Decoder(locality=True, locality_ratio=4.5 / 7).to(dev)
for p_ in dec.parameters():
    torch.nn.init.normal_(p_, 0.0, 0.08)
coords = (torch.rand(P, 3, device=dev) * 2 - 1) * 1.2
zs = torch.randn(K, Z, device=dev)
T = torch.eye(3, device=dev)[None]
g = torch.randn(P, 4, device=dev) * 1e-3
out = {}
for rep in range(3):
    dec.zero_grad()
    z_d = zs.clone().requires_grad_(True)
    raws, *_ = dec(coords, coords[None].expand(K - 1, -1, -1), z_d, T)
    raws.backward(g)
    cur = {n: p_.grad.detach().cpu().clone() for n, p_ in dec.named_parameters()}
    cur["z_slots"] = z_d.grad.detach().cpu().clone()
    if out:
        same = all(torch.equal(cur[k], out[k]) for k in cur)
        print("rep", rep, "bit-equal to rep 0:", same)
    else:
        out = cur
torch.save(out, sys.argv[1])
if len(sys.argv) > 2:
    other = torch.load(sys.argv[2])
    worst = max(((out[k] - other[k]).abs().max().item() / max(other[k].abs().max().item(), 1e-30), k) for k in out)
    print("max relative difference against", sys.argv[2], worst, "bit-equal:", all(torch.equal(out[k], other[k]) for k in out))
