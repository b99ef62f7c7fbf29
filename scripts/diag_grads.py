"""Diagnostic (GPU): where does the training gradient of the B200 path differ from fp32 autograd?  One coarse training step of the
tests/golden/train_trajectory.npz scene at its initial weights: the oracle (CPU, fp32 autograd) supplies every intermediate
gradient; each B200 module is then driven with the oracle's inputs and upstream gradient in isolation and its parameter
gradients are compared tensor by tensor (max error relative to the tensor's max, relative L2 error, fraction of elements whose
SIGN differs - what Adam's normalised step is sensitive to)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
from conftest import load_golden
from oracle import uorf_oracle as O
from uorf_b200.modules import Decoder, Encoder, SlotAttention, Projection, raw2outputs

dev = torch.device("cuda:0")
g = load_golden("train_trajectory")
step = int(sys.argv[1]) if len(sys.argv) > 1 else 0
K, Z, Dn, Fr, N = 3, 40, 8, 8, 2
spec = O.FrustumSpec([Fr, Fr, Dn], render_size=Fr)
P_ = {pre: {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in g["init"][pre].items()} for pre in ("enc", "sa", "dec")}
x, c2w, azi = g["img"], g["cam2world"], g["azi_rot"]
nfg, nbg = g["noise_fg"][step], g["noise_bg"][step]

# ---- oracle forward with retained intermediates (oracle.train_forward restated so the intermediates are reachable)
T = azi[0:1].inverse()
img0 = F.interpolate(x[0:1], size=32, mode="bilinear", align_corners=False)
fmap = O.encoder_forward(P_["enc"], img0, False); fmap.retain_grad()
feat = fmap.flatten(2).permute(0, 2, 1)
slots, attn = O.slot_attention_forward(P_["sa"], feat, K, 3, nfg, nbg); slots.retain_grad()
coords, zv, rd = O.sampling_coords(spec, c2w)
raws, masked, unmasked, masks = O.decoder_forward(P_["dec"], coords, coords[None].expand(K - 1, -1, -1), slots[0], T, locality=True,
                                                 locality_ratio=4.5 / 7, noise=torch.zeros(K, coords.shape[0], 1))
raws.retain_grad()
rr = raws.view(N, Dn, Fr, Fr, 4).permute(0, 2, 3, 1, 4).flatten(0, 2); rr.retain_grad()
rgb, _, _ = O.composite(rr, zv, rd); rgb.retain_grad()
rendered = rgb.view(N, Fr, Fr, 3).permute(0, 3, 1, 2)
tgt = F.interpolate(x, size=Fr, mode="bilinear", align_corners=False)
loss = F.mse_loss(rendered * 2 - 1, tgt)
loss.backward()
print(f"oracle loss {loss.item():.6f} (golden step-{step} loss {g['loss_recon'][step].item():.6f})")


def cmp(tag, got, ref):
    got, ref = got.detach().cpu().double().flatten(), ref.detach().double().flatten()
    mx = ref.abs().max().item()
    e = (got - ref).abs()
    big = ref.abs() > 1e-3 * mx
    sign_bad = ((torch.sign(got) != torch.sign(ref)) & (ref != 0)).double().mean().item()
    sign_bad_big = ((torch.sign(got) != torch.sign(ref)) & big).double().sum().item() / max(big.sum().item(), 1)
    print(f"  {tag:34s} max|ref| {mx:9.3e}  max err/max {e.max().item() / max(mx, 1e-30):8.2e}  rel L2 {(e.norm() / ref.norm().clamp(min=1e-30)).item():8.2e}"
          f"  sign flips {sign_bad:6.2%} (among |g|>1e-3 max: {sign_bad_big:6.2%})")


# ---- composite in isolation
print("composite backward (d rgb -> d raw):")
raw_d = rr.detach().to(dev).requires_grad_(True)
rgb_d, _, _ = raw2outputs(raw_d, zv.to(dev), rd.to(dev))
cmp("rgb_map forward", rgb_d, rgb)
rgb_d.backward(rgb.grad.to(dev))
cmp("d raw", raw_d.grad, rr.grad)

# ---- decoder in isolation (oracle coords are in the reference order)
print("decoder backward (d raws -> weights, z_slots):")
dec = Decoder(n_freq=5, input_dim=33 + Z, z_dim=Z, locality=True, locality_ratio=4.5 / 7).to(dev)
dec.load_state_dict({k: v.detach() for k, v in P_["dec"].items()})
zs = slots[0].detach().to(dev).requires_grad_(True)
cd = coords.to(dev)
raws_d, *_ = dec(cd, cd[None].expand(K - 1, -1, -1), zs, T.to(dev))
cmp("raws forward", raws_d, raws)
raws_d.backward(raws.grad.to(dev))
for n, p in dec.named_parameters():
    cmp(n, p.grad, P_["dec"][n].grad)
cmp("z_slots", zs.grad, slots.grad[0])

# ---- slot attention in isolation
print("slot attention backward (d slots -> weights, feat):")
sa = SlotAttention(num_slots=K, in_dim=Z, slot_dim=Z, iters=3).to(dev)
sa.load_state_dict({k: v.detach() for k, v in P_["sa"].items()})
feat_d = feat.detach().to(dev).requires_grad_(True)
s_d, a_d = sa(feat_d, noise=(nfg.to(dev), nbg.to(dev)))
cmp("slots forward", s_d, slots)
s_d.backward(slots.grad.to(dev))
for n, p in sa.named_parameters():
    cmp(n, p.grad, P_["sa"][n].grad)
cmp("feat", feat_d.grad, (fmap.grad.flatten(2).permute(0, 2, 1)))

# ---- encoder in isolation
print("encoder backward (d feature map -> weights):")
enc = Encoder(3, z_dim=Z, bottom=False).to(dev)
enc.load_state_dict({k: v.detach() for k, v in P_["enc"].items()})
out = enc(img0.to(dev))
cmp("feature map forward", out, fmap)
out.backward(fmap.grad.to(dev))
for n, p in enc.named_parameters():
    cmp(n, p.grad, P_["enc"][n].grad)
