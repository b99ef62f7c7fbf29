"""Hand-off timeline of the decoder BACKWARD kernel (library built with -DUORF_TIMELINE):
    python -m uorf_b200.build -DUORF_TIMELINE=1 --tag=timeline
    UORF_B200_LIB=$PWD/uorf_b200/_build/variants/libuorf_b200_timeline.so python scripts/timeline_bwd.py
Per hand-off h (0 = positional encoding ... 6 = head input, 7 = head gradient ... 13 = layer-0 gradient) of the fg sweep, mean
cycles of: publish -> issuer wake | issuer wake -> step issued (incl. the trailing dW group) | issuer wake -> epilogue wake
(MMAs of the critical group + completion) | epilogue wake -> publish (the epilogue work)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from uorf_b200 import _lib
from uorf_b200.modules import Decoder

K, Z, P = 8, 64, 4 * 64 * 64 * 64
dev = torch.device("cuda:0")
torch.manual_seed(0)
dec = Decoder(locality=True, locality_ratio=4.5 / 7).to(dev)
coords = (torch.rand(P, 3, device=dev) * 2 - 1) * 1.5
z = torch.randn(K, Z, device=dev, requires_grad=True)
T = torch.eye(3, device=dev)[None]
for _ in range(2):
    raws, _, _, _ = dec(coords, coords[None].expand(K - 1, -1, -1), z, T)
    raws.square().mean().backward()
torch.cuda.synchronize()
lib = _lib.lib()
n = 8192
epi, iss = (C.c_longlong * n)(), (C.c_longlong * n)()
lib.uorf_debug_timeline_bwd.restype = C.c_int
assert lib.uorf_debug_timeline_bwd(epi, iss, n) == 0
def split(a):
    a = np.array(a); a = a[a != 0]
    return a >> 8, a & 0xFF
et, ei = split(epi)
it_, ii = split(iss)
pub, prep, wake = et[ei == 0x40], et[ei == 0x50], et[ei == 0x30]
iw, idone = it_[ii == 0x10], it_[ii == 0x20]
# per slot: 14 publishes; epilogue wakes: 14 per slot for slots after the first (wait_d(300) is skipped on the very first)
nslot = min(len(pub) // 14, len(iw) // 14) - 2
pub = pub[14:14 * (nslot + 1)].reshape(nslot, 14); prep = prep[14:14 * (nslot + 1)].reshape(nslot, 14)
iw = iw[14:14 * (nslot + 1)].reshape(nslot, 14)
idone = idone[15:15 + 14 * nslot].reshape(nslot, 14)          # 0x20 stamps precede the NEXT wait
wake = wake[13:13 + 14 * nslot].reshape(nslot, 14)            # first slot has 13 wakes; wake[h] = the wait that precedes publish h
print(f"{nslot} fg slots of CTA 0 (cycles, mean)")
print("  h | publish->issuer wake | issuer busy | issuer wake->next epilogue wake | epilogue wake->publish (store+fence part)")
tot = 0
for h in range(14):
    a = (iw[:, h] - pub[:, h]).mean()
    b = (idone[:, h] - iw[:, h]).mean()
    nxt = wake[:, h + 1] if h < 13 else np.roll(wake[:, 0], -1)
    c = (nxt - iw[:, h])[:-1].mean()
    e = (pub[:, h] - wake[:, h]).mean()
    f = (pub[:, h] - prep[:, h]).mean()
    tot += a + c + e
    print(f" {h:2d} | {a:8.0f} | {b:8.0f} | {c:8.0f} | {e:8.0f} ({f:.0f})")
print(f"slot period: {np.diff(pub[:, 0]).mean():.0f} cycles; sum of the chain segments {tot:.0f}")
