"""Row-sharded moved-object render at the manipulation configuration (F=256, D=128, one view), one process
per GPU:  torchrun --nproc-per-node G scripts/manip_sharded.py [K].  Rank 0 encodes and broadcasts the slot latents, every
rank renders its block of image rows with one object shifted, the blocks are all-gathered, and the result is compared bit
for bit with a single-GPU render of the whole image on rank 0.  Prints one JSON line (validation + timing, not a bench arm)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synthetic_scene  # noqa: E402
from uorf_b200 import sharded  # noqa: E402
from uorf_b200.eval_model import default_eval_options  # noqa: E402
from uorf_b200.manip_model import uorfManipModel  # noqa: E402


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    F, D, Z = 256, 128, 64
    world, rank, local_rank = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(2021)
    opt = default_eval_options(num_slots=K, z_dim=Z, n_samp=D, frustum_size=F, render_size=F, input_size=64, n_img_each_scene=1,
                               gpu_ids=[local_rank])
    m = uorfManipModel(opt, layout="native", precision=os.environ.get("PRECISION", "fp16x3")).eval()
    m.netDecoder.emit = ("raws",)
    img, c2w, azi = synthetic_scene(1, 128)
    m.set_input({"img_data": img, "cam2world": c2w, "azi_rot": azi, "paths": [], "movement": torch.tensor([[1.0, 0.5, 0.0]])})
    g = torch.Generator().manual_seed(7)      # fixed slot-initialisation noise: every encode() yields the same latents
    m.slot_noise = (torch.randn(1, K - 1, Z, generator=g).to(dev), torch.randn(1, 1, Z, generator=g).to(dev))
    offsets = torch.zeros(K - 1, 3, device=dev)
    offsets[1] = m.movement[0] / opt.nss_scale

    def encode():
        with torch.no_grad():
            return m.encode()[0]

    def rows(z, h0, h1):
        with torch.no_grad():
            return m.render(z, m.cam2world, m.nss2cam0, fg_slot_offset=offsets, rows=(h0, h1))["rendered"]

    def step():
        return sharded.sharded_render_rows(encode, rows, F, K, Z, dev, gather=True)

    for _ in range(3):
        local, (h0, h1), full = step()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    e0.record()
    for _ in range(steps):
        local, (h0, h1), full = step()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    same = torch.tensor([1], device=dev)
    if rank == 0:
        z = encode()
        whole = rows(z, 0, F)
        same[0] = int(torch.equal(whole, full))
    dist.broadcast(same, 0)
    if rank == 0:
        units = F * F * D * K
        print(json.dumps({"workload": f"manip_render_K{K}_F{F}_D{D}_N1", "n_gpus": world, "ms_per_step": ms.item(),
                          "sample_slots_per_s": units / (ms.item() * 1e-3), "rows_per_rank": h1 - h0,
                          "gathered_equals_single_gpu_render": bool(same.item())}))
    dist.destroy_process_group()
    if not same.item():
        sys.exit(1)


if __name__ == "__main__":
    main()
