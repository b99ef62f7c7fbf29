"""Hand-off timeline of the half-tile pipelined 3-pass decoder forward (decoder_fwd_x3.cu built with -DUORF_TIMELINE_X3):
    python -m uorf_b200.build -DUORF_TIMELINE_X3=1 --tag=timeline_x3
    UORF_B200_LIB=$PWD/uorf_b200/_build/variants/libuorf_b200_timeline_x3.so python scripts/timeline_x3.py [precision]
CTA 0 / channel 0 stamps clock64 (fg sweep of the bench workload):
  epilogue (warp 0):  0x01 stage-0 arrive | 0x10+st wake on d[0] | 0x20+8n+st converted | 0x30+8n+st gate passed |
                      0x40+8n+st a[n] arrived | 0x17 heads ready | 0x02 composition written
  issuer:             0x60+st wake on a[0] | 0x70+st wake on a[1] | 0x80+st d[0] committed | 0x90+st d[1] committed
Prints mean cycles of the segments per stage."""
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
n = 16384
epi, iss = (C.c_longlong * n)(), (C.c_longlong * n)()
lib.uorf_debug_timeline_x3.restype = C.c_int
assert lib.uorf_debug_timeline_x3(epi, iss, n) == 0
epi, iss = np.array(epi), np.array(iss)


def split(a):
    a = a[a != 0]
    return a >> 8, a & 0xFF


et, ei = split(epi)
it_, ii = split(iss)


def ev(t, i, code):
    return t[i == code]


nslot = min(len(ev(et, ei, 0x17)), len(ev(it_, ii, 0x96))) - 2
print(f"{prec}: {nslot} slots of CTA 0 / channel 0 (cycles, mean over slots)")
e01 = ev(et, ei, 0x01)[:nslot]
print(" MMA stage | a0 arrive->iss wake | a1 arrive->iss wake | a1 wake->d0 commit | d0 commit->d1 commit | d0 commit->epi wake")
for st in range(7):
    a0 = e01 if st == 0 else ev(et, ei, 0x40 + st)[:nslot]
    a1 = e01 if st == 0 else ev(et, ei, 0x48 + st)[:nslot]
    w0 = ev(it_, ii, 0x60 + st)[:nslot]
    c1 = ev(it_, ii, 0x90 + st)[:nslot]
    if 0 < st < 6:
        w1 = ev(it_, ii, 0x70 + st)[:nslot]
        c0 = ev(it_, ii, 0x80 + st)[:nslot]
    else:
        w1 = c0 = None
    nxt = ev(et, ei, 0x10 + st + 1)[:nslot] if st < 6 else ev(et, ei, 0x17)[:nslot]
    s = f"    {st}     | {(w0 - a0).mean():8.0f}           "
    if w1 is not None:
        s += f"| {(w1 - a1).mean():8.0f}            | {(c0 - w1).mean():8.0f}           | {(c1 - c0).mean():8.0f}             | {(nxt - c0).mean():8.0f}"
    else:
        s += f"|      -              | (a0 wake->commit) {(c1 - w0).mean():6.0f} |                      | {(nxt - c1).mean():8.0f}"
    print(s)
print(" epilogue stage | wake d0->E(0) converted | gate wait | store+arrive a0 | E(1) to converted | store+arrive a1 | total")
tot = 0.0
for st in range(1, 7):
    w = ev(et, ei, 0x10 + st)[:nslot]
    c0 = ev(et, ei, 0x20 + st)[:nslot]
    g0 = ev(et, ei, 0x30 + st)[:nslot]
    a0 = ev(et, ei, 0x40 + st)[:nslot]
    c1 = ev(et, ei, 0x28 + st)[:nslot]
    a1 = ev(et, ei, 0x48 + st)[:nslot]
    tot += (a1 - w).mean()
    print(f"      {st}        | {(c0 - w).mean():8.0f}               | {(g0 - c0).mean():6.0f}    | {(a0 - g0).mean():8.0f}        | {(c1 - a0).mean():8.0f}          "
          f"| {(a1 - c1).mean():8.0f}        | {(a1 - w).mean():6.0f}")
hd = ev(et, ei, 0x17)[:nslot]
span = e01[1:nslot] - e01[: nslot - 1]
print(f"sum of the epilogue stages 1..6: {tot:.0f}")
print(f"slot period (stage-0 arrive -> stage-0 arrive): mean {span.mean():.0f}, median {np.median(span):.0f}")
a16 = ev(et, ei, 0x4e)[:nslot]
print(f"last a1 arrive -> heads ready: {(hd - a16).mean():.0f}; heads ready -> next stage-0 arrive (head epilogue, [composition,] stage 0): "
      f"mean {(e01[1:nslot] - hd[: nslot - 1]).mean():.0f}, median {np.median(e01[1:nslot] - hd[: nslot - 1]):.0f}")
