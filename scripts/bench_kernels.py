"""Achieved HBM bandwidth of the two HBM-bound kernels (frustum sampling, compositing): CUDA events over many launches,
buffers rotated through >> 126 MB so L2 does not serve the reads.  Shapes: the bench workload's 128x128x64 frustum with
N_VIEWS views (default 32 = 8x the bench workload, so a ~100 us kernel is timed rather than the ~10 us launch overhead;
pass 4 for the bench workload itself)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uorf_b200.modules import Projection, raw2outputs
import bench
dev = torch.device("cuda:0")
pk = bench.peaks()
N, F_, D = (int(sys.argv[1]) if len(sys.argv) > 1 else 32), 128, 64
R, P = N * F_ * F_, N * F_ * F_ * D
c2w, _ = bench.lookat_cameras(N)
c2w = c2w.to(dev)
pr = Projection(frustum_size=[F_, F_, D], near=6, far=20, nss_scale=7, render_size=(64, 64), device=dev)
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3
out = {}
import ctypes as C
from uorf_b200 import _lib
co = torch.empty(P, 3, device=dev); zo = torch.empty(R, D, device=dev); ro = torch.empty(R, 3, device=dev)
def coords_kernel_only():      # the C-ABI call alone, outputs preallocated (no allocator / wrapper work in the timed loop)
    _lib.check(_lib.lib().uorf_sample_coords(C.byref(pr._frustum), pr._z_cam.data_ptr(), c2w.data_ptr(), N, 1, -1, -1, F_, F_, 1,
                                             co.data_ptr(), zo.data_ptr(), ro.data_ptr(), None), "coords")
t = timeit(coords_kernel_only)
bytes_ = P * 16 + R * 12          # coords 12 B + z_vals 4 B per sample, ray_dir 12 B per ray
out["sample_coords"] = dict(us=t * 1e6, algorithmic_bytes=bytes_, GBps=bytes_ / t / 1e9, frac_of_measured_hbm=bytes_ / t / 1e9 / pk["hbm"])
nbuf = 3                           # rotating buffers: each launch reads data that left L2 long ago
raws = [torch.rand(R, D, 4, device=dev) for _ in range(nbuf)]
_, zv, rd = pr.construct_sampling_coor(c2w, ray_major=True)
i = [0]
def comp():
    raw2outputs(raws[i[0] % nbuf], zv, rd); i[0] += 1
t = timeit(comp, 60)
bytes_ = R * D * 24 + R * 28       # raw 16 B + z 4 B in, weights_norm 4 B out per sample; 28 B per ray
out["composite_fwd"] = dict(us=t * 1e6, algorithmic_bytes=bytes_, GBps=bytes_ / t / 1e9, frac_of_measured_hbm=bytes_ / t / 1e9 / pk["hbm"])
g = torch.rand(R, 3, device=dev)
rq = [r.clone().requires_grad_(True) for r in raws]
def comp_bwd():
    r = rq[i[0] % nbuf]; i[0] += 1
    rgb, _, _ = raw2outputs(r, zv, rd)
    r.grad = None
    rgb.backward(g)
t_fb = timeit(comp_bwd, 60)
bytes_b = R * D * (16 + 4 + 16) + R * 24
out["composite_fwd_plus_bwd"] = dict(us=t_fb * 1e6, note="forward + backward launch pair", bwd_algorithmic_bytes=bytes_b,
                                     bwd_GBps_estimate=bytes_b / max(t_fb - t, 1e-9) / 1e9)
big = torch.empty(256 * 1024 * 1024, device=dev)       # 1 GiB fp32
t_w = timeit(lambda: big.fill_(1.0), 20)
out["context_write_only_fill_GBps"] = big.numel() * 4 / t_w / 1e9
src = torch.empty_like(big)
t_c = timeit(lambda: big.copy_(src), 20)
out["context_copy_GBps"] = 2 * big.numel() * 4 / t_c / 1e9
out["sample_coords"]["frac_of_write_only_fill"] = out["sample_coords"]["GBps"] / out["context_write_only_fill_GBps"]
out["peak_hbm_GBps"] = pk["hbm"]; out["peak_source"] = pk["src"]; out["n_views"] = N
print(json.dumps(out, indent=1))
