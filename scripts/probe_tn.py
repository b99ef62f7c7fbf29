import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uorf_b200 import _lib
g = torch.Generator().manual_seed(0)
a, b1, b2 = [torch.randn(128, 64, generator=g) for _ in range(3)]
d1 = torch.full((64, 64), float("nan"), device="cuda"); d2 = d1.clone()
A, B1, B2 = a.cuda(), b1.cuda(), b2.cuda()
_lib.check(_lib.probe_lib().uorf_probe_gemm_tn(A.data_ptr(), B1.data_ptr(), B2.data_ptr(), d1.data_ptr(), d2.data_ptr(), None), "tn")
torch.cuda.synchronize()
r1 = a.bfloat16().double().t() @ b1.bfloat16().double()
r2 = a.bfloat16().double().t() @ b2.bfloat16().double()
print("d1 err", (d1.cpu().double() - r1).abs().max().item(), "d2 err", (d2.cpu().double() - r2).abs().max().item(), "scale", r1.abs().max().item())
print("d1 vs r2 (should be large)", (d1.cpu().double() - r2).abs().max().item())
print("d1^T err (transposed?)", (d1.cpu().double().t() - r1).abs().max().item())
