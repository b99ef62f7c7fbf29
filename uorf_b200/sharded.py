"""Ray-sharded multi-GPU rendering (one process per GPU, torch.distributed).

Rays are independent once the slot latents exist, so the only exchange on the path is:
  1. rank 0 runs Encoder -> SlotAttention on view 0 and BROADCASTS z_slots (K x C fp32, < 5 KB) and the fg transform;
  2. every rank renders its own contiguous block of (view, row) rays with the single-GPU kernels;
  3. (optional) the rendered image tiles are ALL-GATHERED.
No collective sits inside the per-point work.  The backend is NCCL on GPUs; the same host logic runs over gloo in
the CPU tests, where `render_fn` is a stand-in supplied by the test.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """contiguous [lo, hi) block of `n_items` for `rank`; the first n_items % world ranks get one extra item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_views(n_views: int, world: int, rank: int) -> Tuple[int, int]:
    return shard_range(n_views, world, rank)


def broadcast_slots(z_slots: Optional[torch.Tensor], K: int, C: int, device, src: int = 0, group=None) -> torch.Tensor:
    """rank `src` passes its z_slots; the others pass None and receive them."""
    if z_slots is None:
        z_slots = torch.empty(K, C, device=device, dtype=torch.float32)
    z_slots = z_slots.contiguous()
    dist.broadcast(z_slots, src=src, group=group)
    return z_slots


def sharded_render(encode_fn: Callable[[], torch.Tensor], render_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor],
                   cam2world: torch.Tensor, K: int, C: int, gather: bool = True, group=None):
    """One scene, views sharded over the ranks of `group`.

    encode_fn()                        -> z_slots KxC         (called on rank 0 only)
    render_fn(z_slots, cam2world[lo:hi]) -> images (hi-lo)x3xHxW
    Returns (local images, (lo, hi), all images Nx3xHxW on every rank or None).
    Views must divide evenly when gather=True (all_gather_into_tensor needs equal shards)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    device = cam2world.device
    z_slots = encode_fn() if rank == 0 else None
    z_slots = broadcast_slots(z_slots, K, C, device, 0, group)
    lo, hi = shard_views(cam2world.shape[0], world, rank)
    local = render_fn(z_slots, cam2world[lo:hi]).contiguous()
    full = None
    if gather:
        if cam2world.shape[0] % world != 0:
            raise ValueError("gather=True needs n_views divisible by the world size")
        full = torch.empty((cam2world.shape[0],) + tuple(local.shape[1:]), device=device, dtype=local.dtype)
        dist.all_gather_into_tensor(full, local, group=group)
    return local, (lo, hi), full


def sharded_render_rows(encode_fn: Callable[[], torch.Tensor], render_rows_fn: Callable[[torch.Tensor, int, int], torch.Tensor],
                        height: int, K: int, C: int, device, gather: bool = True, group=None):
    """One scene, the image ROWS of every view sharded over the ranks of `group` (few views, large frustum: the
    manipulation render at 256x256x128).

    encode_fn()                      -> z_slots KxC          (rank 0 only)
    render_rows_fn(z_slots, h0, h1)  -> images Nx3x(h1-h0)xW for rows h0..h1-1
    Returns (local rows, (h0, h1), full images Nx3xHxW on every rank or None)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    z_slots = encode_fn() if rank == 0 else None
    z_slots = broadcast_slots(z_slots, K, C, device, 0, group)
    h0, h1 = shard_range(height, world, rank)
    local = render_rows_fn(z_slots, h0, h1).contiguous()
    full = None
    if gather:
        if height % world != 0:
            raise ValueError("gather=True needs the image height divisible by the world size")
        n = local.shape[0]
        stacked = torch.empty((world * n,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
        dist.all_gather_into_tensor(stacked, local, group=group)
        # [world*N, 3, h, W] -> N x 3 x (world*h) x W
        full = stacked.view((world, n) + tuple(local.shape[1:])).permute(1, 2, 0, 3, 4).reshape(n, local.shape[1], height, local.shape[3])
    return local, (h0, h1), full


class GraphedShardedRender:
    """encode (rank 0) -> NCCL broadcast of the slot latents -> rank-local render [-> all-gather of the image tiles], captured as
    ONE CUDA graph per rank.

    The eager form of this step costs rank 0 about 60 small launches (encoder, slot attention) and every rank a separate
    broadcast launch per scene; replaying it as a graph (NCCL collectives are capturable) removes that host time from the
    multi-GPU step.  `model` is a uorfEvalModel whose set_input() has been called with THIS rank's views (rank 0 must hold
    view 0's image); shapes must stay fixed after the first call.

      rows=(h0, h1)     this rank renders image rows h0..h1-1 of every view (rays of ONE scene sharded over ranks: the 256^2 x 128
                        manipulation render, reference models/uorf_manip_model.py:196-232) instead of whole views;
      fg_slot_offset    (K-1)x3 per-object translation of the moved-object render (static tensor, read at replay time);
      gather=True       the rendered tiles of all ranks are all-gathered inside the graph: every rank ends the step holding
                        the full image set (`out['full']`: views sharded -> (world*N)x3xHxW, rows sharded -> Nx3xHxW).
    `release()` drops the graph (and with it its references to the NCCL communicator) before the process group is destroyed."""

    def __init__(self, model, K: int, C: int, group=None, rows=None, fg_slot_offset=None, gather: bool = False, height=None):
        self.model, self.K, self.C, self.group = model, K, C, group
        self.rows, self.fg_slot_offset, self.gather, self.height = rows, fg_slot_offset, gather, height
        self.graph, self.static, self.out = None, None, None

    def _body(self):
        m, st = self.model, self.static
        rank, world = (dist.get_rank(self.group), dist.get_world_size(self.group)) if dist.is_initialized() else (0, 1)
        with torch.no_grad():
            m.x = st["x"]
            z = m.encode()[0] if rank == 0 else torch.empty(self.K, self.C, device=st["x"].device)
            if world > 1:
                z = broadcast_slots(z, self.K, self.C, z.device, 0, self.group)
            out = m.render(z, st["c2w"], st["T"], fg_slot_offset=self.fg_slot_offset, rows=self.rows)
            if self.gather:
                local = out["rendered"].contiguous()
                if world == 1:
                    out["full"] = local
                else:
                    n = local.shape[0]
                    stacked = torch.empty((world * n,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
                    dist.all_gather_into_tensor(stacked, local, group=self.group)
                    if self.rows is None:
                        out["full"] = stacked
                    else:   # [world*N, 3, h, W] -> N x 3 x (world*h) x W
                        out["full"] = stacked.view((world, n) + tuple(local.shape[1:])).permute(1, 2, 0, 3, 4).reshape(
                            n, local.shape[1], world * local.shape[2], local.shape[3])
            return out

    def __call__(self):
        m = self.model
        dev = m.x.device
        if self.graph is None:
            self.static = {"x": m.x.clone(), "c2w": m.cam2world.clone(), "T": m.nss2cam0.clone()}
            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(3):          # NCCL communicator, cuDNN autotuning and allocator pools exist before the capture
                    self._body()
            cur.wait_stream(side)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self.out = self._body()
            self.graph = g
        else:
            st = self.static
            st["x"].copy_(m.x, non_blocking=True)
            st["c2w"].copy_(m.cam2world, non_blocking=True)
            st["T"].copy_(m.nss2cam0, non_blocking=True)
        self.graph.replay()
        return self.out

    def run_eager(self):
        """the same step launched eagerly on the model's current inputs (no graph)"""
        m = self.model
        self.static = {"x": m.x, "c2w": m.cam2world, "T": m.nss2cam0}
        return self._body()

    def release(self):
        self.graph, self.out, self.static = None, None, None


def allreduce_gradients(params, group=None, average: bool = True) -> None:
    """Data-parallel training: scenes are sharded over ranks, weights replicated; after backward
    every gradient is summed over ranks with ONE flat-bucket all-reduce (0.7 - 1.9 MB fp32: latency-bound on NVLink, so
    one message beats per-tensor calls) and divided by the world size (the step then averages the per-scene losses;
    world size 1 reproduces the reference's one-scene step exactly)."""
    world = dist.get_world_size(group)
    if world == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= world
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def flatten_gradients(params) -> torch.Tensor:
    """One flat fp32 buffer for all gradients; every `p.grad` becomes a view of its slice (16-byte aligned slices, so the
    backward kernels' vector stores and the fused optimizer's multi-tensor loads stay aligned).  Returns the buffer."""
    params = [p for p in params if p.requires_grad]
    offs, total = [], 0
    for p in params:
        offs.append(total)
        total += (p.numel() + 3) // 4 * 4
    flat = torch.zeros(total, device=params[0].device, dtype=torch.float32)
    for p, o in zip(params, offs):
        p.grad = flat[o:o + p.numel()].view(p.shape)
    return flat


def allreduce_flat(flat: torch.Tensor, group=None, average: bool = True) -> None:
    """the data-parallel exchange: ONE all-reduce over the flat gradient buffer (NCCL: the average is taken inside the
    collective; gloo has no AVG, so the CPU test path divides afterwards)"""
    world = dist.get_world_size(group)
    if world == 1:
        return
    if average and dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
        return
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= world
