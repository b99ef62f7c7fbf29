"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: view sharding + slot broadcast + image gather, and the
flat-bucket gradient all-reduce.  The per-rank compute is a stand-in (the CUDA kernels need a GPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from uorf_b200 import sharded


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    K, C, N = 5, 8, 4
    cam2world = torch.arange(N * 16, dtype=torch.float32).view(N, 4, 4)
    slots_rank0 = torch.randn(K, C)

    def encode():
        return slots_rank0.clone()

    def render(z, c2w):          # stand-in render: depends on the slots and on the view's camera
        return (z.sum() + c2w.sum(dim=(1, 2))).view(-1, 1, 1, 1).expand(-1, 3, 2, 2).contiguous()

    local, (lo, hi), full = sharded.sharded_render(encode, render, cam2world, K, C, gather=True)
    expect = render(slots_rank0, cam2world)
    ok_render = torch.equal(full, expect) and torch.equal(local, expect[lo:hi]) and (hi - lo) == N // world
    # gradient all-reduce: rank r holds grads r+1 -> average (1+2)/2
    params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2))]
    params[0].grad = torch.full((3, 4), float(rank + 1))
    params[1].grad = torch.arange(7, dtype=torch.float32) * (rank + 1)
    sharded.allreduce_gradients(params)
    ok_grad = torch.allclose(params[0].grad, torch.full((3, 4), 1.5)) and torch.allclose(params[1].grad, torch.arange(7.) * 1.5) \
        and params[2].grad is None
    # the training driver's form: every .grad a view of ONE flat bucket, one all-reduce (average) over the bucket
    ps = [torch.nn.Parameter(torch.zeros(3, 5)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2, 2))]
    flat = sharded.flatten_gradients(ps)
    views_ok = all(p_.grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr() and p_.grad.shape == p_.shape and
                   p_.grad.data_ptr() % 16 == 0 for p_ in ps)
    for i, p_ in enumerate(ps):
        p_.grad.copy_(torch.full(p_.shape, float((rank + 1) * (i + 1))))       # written in place, as the backward kernels do
    sharded.allreduce_flat(flat)
    ok_flat = views_ok and all(torch.allclose(p_.grad, torch.full(p_.shape, 1.5 * (i + 1))) for i, p_ in enumerate(ps))
    # row sharding (one view, large frustum): rank r renders rows [r*H/world, (r+1)*H/world)
    Hh, Ww = 6, 5
    image = torch.arange(2 * 3 * Hh * Ww, dtype=torch.float32).view(2, 3, Hh, Ww)

    def render_rows(z, h0, h1):
        return image[:, :, h0:h1] + z.sum()
    loc, (h0, h1), full_rows = sharded.sharded_render_rows(encode, render_rows, Hh, K, C, torch.device("cpu"), gather=True)
    ok_rows = torch.equal(full_rows, image + slots_rank0.sum()) and (h1 - h0) == Hh // world and torch.equal(loc, full_rows[:, :, h0:h1])
    out[rank] = bool(ok_render and ok_grad and ok_rows and ok_flat)
    dist.destroy_process_group()


def test_sharded_render_and_grad_allreduce_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_shard_ranges_cover_everything():
    for n in (1, 4, 7, 32):
        for w in (1, 2, 3, 8):
            got = []
            for r in range(w):
                lo, hi = sharded.shard_range(n, w, r)
                got += list(range(lo, hi))
            assert got == list(range(n))
