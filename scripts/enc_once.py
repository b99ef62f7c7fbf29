import os, sys
sys.path.insert(0, os.getcwd())
import torch
from uorf_b200.modules import Encoder
enc = Encoder(3, 64).cuda().eval()
x = torch.rand(1, 3, 64, 64, device="cuda")
with torch.no_grad():
    for _ in range(3): enc(x)
torch.cuda.synchronize()
