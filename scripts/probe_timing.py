import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uorf_b200 import _lib
out = torch.zeros(256, dtype=torch.int64, device="cuda")
_lib.check(_lib.probe_lib().uorf_probe_timing(out.data_ptr(), 50, None), "probe_timing")
torch.cuda.synchronize()
o = out.cpu().tolist()
names = ["N=64 TS same D", "N=64 TS 2 rotating D", "N=64 TS 4 rotating D", "N=128 TS same D", "N=256 TS same D",
         "N=64 SS same D", "N=64 SS 4 rotating D", "N=32 TS same D"]
for v, nm in enumerate(names):
    print(f"A[{nm}] 48 MMAs: issue {o[32+2*v+1]} cyc, complete {o[32+2*v]} cyc -> {o[32+2*v]/48:.1f} cyc/MMA")
print(f"B: ld32+wait {o[16]}  compute(split x16 pairs) {o[17]}  st16x2+wait {o[18]}  fence+arrive {o[19]}  total {o[20]}")
print(f"C: round trip arrive->issuer wake->4 MMAs->commit->wake: {o[24]} cyc")
for m, nm in enumerate(["LDTM (2 x ld32 + wait)", "STTM (4 x st16 + wait)", "ALU (16 hi/lo splits)"]):
    print(f"D[{nm}] 64 iterations on 4 warps: alone {o[64+4*m]} cyc ({o[64+4*m]/64:.1f}/iter), with an MMA stream {o[65+4*m]} cyc "
          f"({o[65+4*m]/64:.1f}/iter); 192 MMAs alone {o[80]} cyc, against it {o[66+4*m]} cyc")
