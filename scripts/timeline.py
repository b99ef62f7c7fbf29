"""Hand-off timeline of the decoder forward kernel (tuning variant built with -DUORF_TIMELINE, see decoder_fwd.cu):
    python -m uorf_b200.build -DUORF_TIMELINE=1 --tag=timeline
    UORF_B200_LIB=$PWD/uorf_b200/_build/variants/libuorf_b200_timeline.so python scripts/timeline.py [precision]
CTA 0 / channel 0 stamps clock64 at: epilogue wake (d_full), epilogue arrive (a_full), issuer wake, issuer commit.
Prints the mean cycles of the four segments of a stage, per stage index, for the fg sweep of the bench workload."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from uorf_b200 import _lib
from uorf_b200.modules import Decoder

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16x3"
K, Z, P = 5, 64, 4 * 128 * 128 * 64
dev = torch.device("cuda:0")
torch.manual_seed(0)
dec = Decoder(locality=False, locality_ratio=4.5 / 7).to(dev)
dec.precision = prec
dec.emit = ("raws", "masked_raws", "unmasked_raws")
coords = (torch.rand(P, 3, device=dev) * 2 - 1) * 1.5
z = torch.randn(K, Z, device=dev)
T = torch.eye(3, device=dev)[None]
with torch.no_grad():
    for _ in range(2):
        dec(coords, coords[None].expand(K - 1, -1, -1), z, T)
torch.cuda.synchronize()
lib = _lib.lib()
n = 8192
epi, iss = (C.c_longlong * n)(), (C.c_longlong * n)()
lib.uorf_debug_timeline.restype = C.c_int
assert lib.uorf_debug_timeline(epi, iss, n) == 0
epi, iss = np.array(epi), np.array(iss)
def split(a):
    a = a[a != 0]
    return a >> 8, a & 0xFF
et, ei = split(epi)
it_, ii = split(iss)
# epilogue stream per slot: arrive(0x40), [wake, arrive(0x41..0x46)] x6, wake(head)
ev = {}
def add(k, v):
    ev.setdefault(k, []).append(v)
# index issuer events: wake 0x10+st, commit 0x20+st, in order
iw = {st: it_[ii == 0x10 + st] for st in range(7)}
ic = {st: it_[ii == 0x20 + st] for st in range(7)}
ea = {st: et[ei == 0x40 + st] for st in range(7)}
wake_mask = (ei >= 0x30) & (ei < 0x40)
ew_all = et[wake_mask]
nslot = min(len(ea[0]), len(ic[6]), len(ew_all) // 7) - 1
ew = ew_all[: nslot * 7].reshape(nslot, 7)      # wakes for stages 1..6 and head
print(f"{prec}: {nslot} slots of CTA 0 / channel 0 (cycles, mean over slots)")
print(" st | arrive->issuer wake | issue (wake->commit) | commit->epilogue wake | epilogue (wake->arrive)")
tot = 0.0
for st in range(7):
    a = ea[st][:nslot]
    w = iw[st][:nslot]
    c = ic[st][:nslot]
    nxt = ew[:, st]
    t1, t2, t3 = (w - a).mean(), (c - w).mean(), (nxt - c).mean()
    t4 = (ea[st + 1][:nslot] - nxt).mean() if st < 6 else float("nan")
    tot += t1 + t2 + t3 + (0 if st == 6 else t4)
    print(f" {st}  | {t1:8.0f}            | {t2:8.0f}             | {t3:8.0f}              | {t4:8.0f}")
if (ei == 0x51).any():
    print(" epilogue split (stages 1..6): wake -> ld+math+st issued -> st complete -> arrive")
    for st in range(1, 7):
        w = ew[:, st - 1]
        b = et[ei == 0x50 + st][:nslot]; c2 = et[ei == 0x60 + st][:nslot]; a = ea[st][:nslot]
        print(f"   st {st}: {(b - w).mean():7.0f} {(c2 - b).mean():7.0f} {(a - c2).mean():7.0f}")
print(f"sum over the 7 stages: {tot:.0f} cycles per slot (excluding the head epilogue / composition)")
span = (ea[0][1:nslot] - ea[0][: nslot - 1])
print(f"slot period (arrive0 -> arrive0): mean {span.mean():.0f}, median {np.median(span):.0f}")
