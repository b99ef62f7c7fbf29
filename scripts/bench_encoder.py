"""Encoder forward (batch 1, 64x64 input, z_dim 64): direct-conv inference kernels vs the torch / cuDNN layers, both replayed as
CUDA graphs (so launch overhead is excluded, as in bench.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uorf_b200.modules import Encoder
torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
enc = Encoder(3, 64).to(dev).eval()
x = torch.rand(1, 3, 64, 64, device=dev)
def timed(flag):
    enc.use_kernels = flag
    with torch.no_grad():
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3): enc(x)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            y = enc(x)
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 200 * 1000
print("cudnn/torch layers: %.1f us   direct-conv kernels: %.1f us" % (timed(False), timed(True)))
from torch.profiler import profile, ProfilerActivity
for flag in (False, True):
    enc.use_kernels = flag
    with torch.no_grad():
        for _ in range(3): enc(x)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(10): enc(x)
            torch.cuda.synchronize()
    print("use_kernels =", flag)
    for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
        print(f"  {e.device_time_total / 10:8.1f} us/fwd  {e.count / 10:4.1f}x  {e.key[:80]}")
if True:
    # per-launch times of the direct-conv path, in call order
    enc.use_kernels = True
    with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA]) as prof:
        enc(x); torch.cuda.synchronize()
    print([round(e.device_time_total, 1) for e in prof.events() if "encoder_conv" in e.name])
