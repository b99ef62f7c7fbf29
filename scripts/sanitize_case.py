"""Small-shape pass over every kernel of libuorf_b200.so for compute-sanitizer (scripts/gpu_sanitize.sh): frustum sampling (all three
kernels), encoder convolutions (cluster / DSMEM split included), slot attention fwd + bwd, fused decoder fwd (3-pass four-channel
and 1-pass variants, ragged last tile) + bwd, compositing fwd + bwd.  Shapes are tiny: the tools slow kernels down 10-100x."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uorf_b200.modules import Decoder, Encoder, Projection, SlotAttention, raw2outputs
from uorf_b200.eval_model import init_weights_like_reference

dev = torch.device("cuda:0")
torch.manual_seed(0)
K, Z, N, F_, D = 3, 64, 1, 8, 20
c2w = torch.eye(4)[None].repeat(N, 1, 1); c2w[:, 2, 3] = -12.0
proj = Projection(frustum_size=[F_, F_, D], near=6, far=20, nss_scale=7, render_size=(4, 4), device=dev)
for kw in (dict(ray_major=True), dict(), dict(partitioned=True), dict(ray_major=True, crop=(2, 3)),
           dict(ray_major=True, crop=torch.tensor([1, 2], dtype=torch.int32, device=dev))):
    proj.construct_sampling_coor(c2w.to(dev), **kw)
coords, z_vals, ray_dir = proj.construct_sampling_coor(c2w.to(dev), ray_major=True)          # P = 1280 points = 10 tiles
enc = init_weights_like_reference(Encoder(3, z_dim=Z), "normal").to(dev)
sa = init_weights_like_reference(SlotAttention(num_slots=K, in_dim=Z, slot_dim=Z), "normal").to(dev)
dec = init_weights_like_reference(Decoder(n_freq=5, input_dim=33 + Z, z_dim=Z, locality=True, locality_ratio=4.5 / 7), "xavier").to(dev)
T = torch.eye(3, device=dev)[None]
x = torch.rand(1, 3, 32, 32, device=dev) * 2 - 1
with torch.no_grad():                                                                       # inference kernels
    fm = enc(x)
    slots, attn = sa(fm.flatten(2).permute(0, 2, 1))
    for prec in ("fp16x3", "bf16x3", "bf16", "fp16"):
        dec.precision = prec
        P_ = coords.shape[0] - 37                                                           # ragged last tile
        raws, masked, unmasked, masks = dec(coords[:P_], coords[:P_][None].expand(K - 1, -1, -1), slots[0], T)
    dec.precision = "fp16x3"
    dec.fg_slot_offset = torch.zeros(K - 1, 3, device=dev)
    raws, masked, unmasked, masks = dec(coords, coords[None].expand(K - 1, -1, -1), slots[0], T)
    dec.fg_slot_offset = None
    raw2outputs(masked.reshape(-1, D, 4), z_vals, ray_dir, render_mask=True)                # K stacked ray sets, one geometry
# training kernels: one forward + backward through every module
fm = enc(x)
slots, attn = sa(fm.flatten(2).permute(0, 2, 1))
raws, *_ = dec(coords, coords[None].expand(K - 1, -1, -1), slots[0], T)
rgb, depth, wn = raw2outputs(raws.view(-1, D, 4), z_vals, ray_dir)
rgb.square().mean().backward()
torch.cuda.synchronize()
assert all(p.grad is not None and torch.isfinite(p.grad).all() for m in (enc, sa, dec) for p in m.parameters())
print("sanitize_case: ok")
