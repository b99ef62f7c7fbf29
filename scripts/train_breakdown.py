"""Per-kernel device-time breakdown of the training step (torch.profiler, CUDA activities)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from uorf_b200.nogan_model import uorfNoGanModel, default_train_options
c = bench.TRAIN_CFG
torch.manual_seed(2021)
opt = default_train_options(num_slots=c["K"], z_dim=c["z_dim"], n_samp=c["n_samp"], frustum_size=c["frustum"], render_size=64,
                            supervision_size=64, input_size=64, n_img_each_scene=4, stage="coarse", gpu_ids=[0], warmup_steps=10, vgg_random_init=True)
m = uorfNoGanModel(opt, precision=sys.argv[1] if len(sys.argv) > 1 else "fp16x3")
img, c2w, azi = bench.synthetic_scene(4, 128)
inp = {k: v.cuda() for k, v in {"img_data": img, "cam2world": c2w, "azi_rot": azi}.items()}
for _ in range(4):
    m.set_input(inp); m.optimize_parameters(epoch=m.opt.percept_in)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5):
        m.set_input(inp); m.optimize_parameters(epoch=m.opt.percept_in)
    torch.cuda.synchronize()
from torch.autograd import DeviceType
kern = [e for e in prof.events() if e.device_type == DeviceType.CUDA]
busy = sum(e.device_time_total for e in kern)
t_lo, t_hi = min(e.time_range.start for e in kern), max(e.time_range.end for e in kern)
print(f"kernel events: {len(kern) / 5:.0f} per step, span {(t_hi - t_lo) / 5 / 1000:.3f} ms per step")
print(f"sum of kernel durations per step: {busy / 5 / 1000:.3f} ms (compare with bench.py --mode train ms_per_step: the rest is idle gaps)")
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.device_time_total for e in ev)
print(f"total device time per step: {tot / 5 / 1000:.3f} ms")
for e in sorted(ev, key=lambda e: -e.device_time_total)[:22]:
    print(f"{e.device_time_total / 5:9.1f} us/step  {e.count / 5:5.1f}x  {100 * e.device_time_total / tot:5.1f}%  {e.key[:90]}")
