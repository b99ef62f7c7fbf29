import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import uorf_oracle as O
from uorf_b200.modules import Decoder
dev = torch.device("cuda:0")

def run(K, P, Z, bias_shift, tag, seed=0, bg_sigma_bias=None, coherent=False):
    gen = torch.Generator().manual_seed(seed)
    params = O.init_decoder_params(Z, seed=K + 1)
    for k in params:
        if k.endswith("bias"):
            params[k] = torch.randn(params[k].shape, generator=gen) * 0.1
            if bias_shift and ("before" in k or ("after." in k and not k.startswith("b_after.6"))):
                params[k] = params[k] + bias_shift
    if bg_sigma_bias is not None:
        params["b_after.6.bias"][3] = bg_sigma_bias
    cb = (torch.rand(P, 3, generator=gen) * 2 - 1) * 1.5
    zs = torch.randn(K, Z, generator=gen)
    probe = torch.randn(P, 4, generator=gen)
    if coherent:
        probe = torch.tensor([0.3, -0.2, 0.5, 0.1]).expand(P, 4).contiguous()
    c2w, azi = O.lookat_cameras(1)
    T = azi[0:1].inverse()
    ratio = 4.5 / 7
    p_ref = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    zs_ref = zs.clone().requires_grad_(True)
    r_o, *_ = O.decoder_forward(p_ref, cb, cb[None].expand(K - 1, -1, -1), zs_ref, T, locality=True, locality_ratio=ratio,
                                noise=torch.zeros(K, P, 1))
    (r_o * probe).sum().backward()
    dec = Decoder(n_freq=5, input_dim=33 + Z, z_dim=Z, locality=True, locality_ratio=ratio).to(dev)
    dec.load_state_dict(params)
    zs_g = zs.to(dev).requires_grad_(True)
    cbd = cb.to(dev)
    raws, *_ = dec(cbd, cbd[None].expand(K - 1, -1, -1), zs_g, T.to(dev))
    (raws * probe.to(dev)).sum().backward()
    out = {}
    for k, p in dec.named_parameters():
        ref = p_ref[k].grad
        out[k] = (p.grad.cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-20)
    out["z_slots"] = (zs_g.grad.cpu() - zs_ref.grad).abs().max().item() / zs_ref.grad.abs().max().item()
    print(tag, "fwd err", (raws.detach().cpu() - r_o.detach()).abs().max().item())
    for name in ():
        gg, rr = dict(dec.named_parameters())[name].grad.cpu(), p_ref[name].grad
        sc = rr.abs().max().item()
        rowerr = ((gg - rr).abs().max(dim=1).values / sc)
        colerr = ((gg - rr).abs().max(dim=0).values / sc)
        print("   ", name, "row err x100:", [int(100 * v) for v in rowerr.tolist()])
        print("   ", name, "col err x100:", [int(100 * v) for v in colerr.tolist()])
        print("    ratio got/ref on largest entries:", [(round((gg.flatten()[i] / rr.flatten()[i]).item(), 3)) for i in rr.abs().flatten().topk(8).indices.tolist()])
    print("   ", {k.replace(".weight", ".w").replace(".bias", ".b"): f"{v:.1e}" for k, v in out.items()})

run(5, 128 * 148 + 55, 64, 0.0, "K5 big coherent", coherent=True)
run(5, 128 * 148 + 55, 64, 0.0, "K5 big coherent bg+1", coherent=True, bg_sigma_bias=1.0)
run(8, 9000, 40, 0.0, "K8 Z40 coherent", coherent=True)
run(16, 3000, 64, 0.0, "K16 coherent", coherent=True)
