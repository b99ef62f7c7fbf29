#!/usr/bin/env python
"""bench.py - headline benchmark of the uORF per-scene render hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--precision fp16x3|bf16x3|bf16|fp16]
                    [--mode render|train] [--workload clevr567|room_diverse_fine] [--no-graph] [--no-cpu-baseline]

Metric (BASELINE.json): composited ray-samples/s x K slots = P_total * K / t, where one STEP is one scene rendered
through Encoder -> SlotAttention -> Projection -> Decoder -> raw2outputs (the eval driver's forward,
reference models/uorf_eval_model.py:116-157) with per-slot masked/unmasked raws materialised as the reference does.
Workload (BASELINE.json configs[1]): Room-Chair novel-view render, K=5 slots, z_dim 64, 128x128 frustum x 64 samples,
4 views per scene, synthetic images/cameras, random-init weights (seed 2021).
N > 1 GPUs: rays shard by view - every rank renders 4 views of a 4N-view scene, rank 0 encodes and broadcasts the
slot latents over NCCL (weak scaling, no collective inside the per-point work).

The step is replayed as ONE CUDA graph per rank (N > 1: [rank 0: encoder + slot attention] -> NCCL broadcast -> render, the
collective captured in the graph); `--no-graph` launches eagerly.  `--mode train` measures BASELINE configs[2] (CLEVR-567
coarse-stage training step, K=8, scene-sharded data parallel, whole step incl. the gradient all-reduce as one graph).
The JSON line carries `roofline` (dominant kernel: CUDA-event time over eagerly launched steps, algorithmic flops counted
once), `cpu_baseline` (oracle on the host cores + the as-written algorithm as torch eager ops on the same GPU), `e2e`
(host buffers, H2D + D2H inside the timed region), `clocks` (NVML samples under load) and `gpu_launches`.
`UORF_PROFILE_RANGE=1` brackets the timed steps with cudaProfilerStart/Stop (ncu --profile-from-start off).

`--impl reference` times the CPU restatement of the reference algorithm (oracle/, the reference itself cannot travel
to the GPU box) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(workload="room_chair_eval_K5_F128_D64_N4", K=5, z_dim=64, n_views=4, frustum=128, n_samp=64, render_size=64,
           input_size=64, load_size=128)
# the other shapes BASELINE.json names (parity-test cases; `--workload` runs them through the same arms for the record)
WORKLOADS = {
    "clevr567": dict(workload="clevr567_eval_K8_F64_D64_N4", K=8, z_dim=64, n_views=4, frustum=64, n_samp=64, render_size=64,
                     input_size=64, load_size=128),
}
TRAIN_WORKLOADS = {
    "room_diverse_fine": dict(workload="room_diverse_train_fine_K5_F128crop64_D64_N4", K=5, z_dim=64, n_views=4, frustum=64, n_samp=64,
                              supervision_size=64, input_size=128, load_size=128, stage="fine", bottom=True, frustum_fine=128),
}
METRIC = "composited_ray_samples_x_slots_per_sec"
UNIT = "sample-slots/s"
HBM_FALLBACK_GBS, BF16_FALLBACK_TFLOPS = 6650.0, 1590.0


def f_min_per_point(K: int, Z: int) -> float:
    """ALGORITHMIC forward flops per point (all K slots), SURVEY section 8(d): slot-latent columns hoisted into a
    per-slot bias, f_after_latent folded into f_color.0, no padding."""
    fg = 2 * (66 * Z + 5.25 * Z * Z + 1.75 * Z)
    bg = 2 * (66 * Z + 5 * Z * Z + 4 * Z)
    return (K - 1) * fg + bg


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=HBM_FALLBACK_GBS, tf_burst=BF16_FALLBACK_TFLOPS, tf_sust=1400.0, src="fallback")


# ----------------------------------------------------------------------------------------------------
# clocks sampler (NVML), runs during the timed region
# ----------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake",
               0x100: "display_clock_setting"}

    def __init__(self, index: int, period=0.02):
        super().__init__(daemon=True)
        self.period, self.index = period, index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    self.nv, "nvmlDeviceGetCurrentClocksEventReasons") else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------------------------------
# synthetic scene (SURVEY section 8(d) recipe; same generator as oracle.synthetic_scene, restated here because the
# product path may not import oracle/)
# ----------------------------------------------------------------------------------------------------
def lookat_cameras(n_views, radius=12.0, elevation=0.6, az0=0.3):
    import math
    c2w, azi = [], []
    for i in range(n_views):
        az = 2 * math.pi * i / n_views + az0
        pos = torch.tensor([radius * math.cos(elevation) * math.cos(az), radius * math.cos(elevation) * math.sin(az),
                            radius * math.sin(elevation)], dtype=torch.float64)
        fwd = -pos / pos.norm()
        up = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64)
        right = torch.linalg.cross(fwd, up)
        right = right / right.norm()
        down = torch.linalg.cross(fwd, right)
        m = torch.eye(4, dtype=torch.float64)
        m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, down, fwd, pos
        c2w.append(m)
        azi.append(torch.tensor([[math.cos(az), -math.sin(az), 0], [math.sin(az), math.cos(az), 0], [0, 0, 1]],
                                dtype=torch.float64))
    return torch.stack(c2w).float(), torch.stack(azi).float()


def synthetic_scene(n_views, load_size, seed=2021):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(n_views, 3, load_size, load_size, generator=g) * 2 - 1
    c2w, azi = lookat_cameras(n_views)
    return img, c2w, azi


# ----------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (oracle port) on the host cores
# ----------------------------------------------------------------------------------------------------
def cpu_reference_run(steps: int, warmup: int, views_sample: int = 1):
    """Times oracle.eval_render (the CPU restatement of uorf_eval_model.forward) on `views_sample` of the 4 views."""
    from oracle import uorf_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    c = CFG
    img, c2w, azi = synthetic_scene(c["n_views"], c["load_size"])
    spec = O.FrustumSpec([c["frustum"], c["frustum"], c["n_samp"]], near=6.0, far=20.0, nss_scale=7.0, render_size=c["render_size"])
    enc_p, sa_p, dec_p = O.init_encoder_params(c["z_dim"]), O.init_slot_attention_params(c["z_dim"]), O.init_decoder_params(c["z_dim"])
    g = torch.Generator().manual_seed(7)
    nfg, nbg = torch.randn(1, c["K"] - 1, c["z_dim"], generator=g), torch.randn(1, 1, c["z_dim"], generator=g)

    def one():
        return O.eval_render(enc_p, sa_p, dec_p, img[:views_sample], c2w[:views_sample], azi[:views_sample], spec,
                             num_slots=c["K"], input_size=c["input_size"], noise_fg=nfg, noise_bg=nbg)
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    units = views_sample * c["frustum"] ** 2 * c["n_samp"] * c["K"]
    return units / dt, dt, cores, f"{views_sample} of {c['n_views']} views per step ({units} sample-slots), oracle eval_render, {cores} threads"


def eager_gpu_run(dev, reps: int = 3):
    """The as-written algorithm (oracle decoder_forward + the permute copy + composite: ~93 % of the reference's render path,
    SURVEY 8(a)) executed as stock torch eager ops on the GPU, TF32 off like the reference: what running the reference's own
    modules on this B200 would cost for one of the 4 chunks of the bench workload.  A baseline leg like cpu_baseline."""
    from oracle import uorf_oracle as O
    c = CFG
    K, Z, N, D, rs = c["K"], c["z_dim"], c["n_views"], c["n_samp"], c["render_size"]
    _, c2w, azi = synthetic_scene(N, c["load_size"])
    spec = O.FrustumSpec([c["frustum"], c["frustum"], D], near=6.0, far=20.0, nss_scale=7.0, render_size=rs)
    coords, zv, rd = O.sampling_coords(spec, c2w, partitioned=True)
    coor, zv0, rd0 = coords[0].to(dev), zv[0].to(dev), rd[0].to(dev)
    dec_p = {k: v.to(dev) for k, v in O.init_decoder_params(Z).items()}
    g = torch.Generator().manual_seed(11)
    slots = torch.randn(K, Z, generator=g).to(dev)
    T = azi[0:1].inverse().to(dev)
    noise = torch.zeros(K, coor.shape[0], 1, device=dev)

    def one():
        with torch.no_grad():
            raws, masked, unmasked, _ = O.decoder_forward(dec_p, coor, coor[None].expand(K - 1, -1, -1), slots, T, locality=False,
                                                          locality_ratio=4.5 / 7.0, noise=noise)
            raws = raws.view(N, D, rs, rs, 4).permute(0, 2, 3, 1, 4).flatten(0, 2)
            return O.composite(raws, zv0, rd0)
    one()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        one()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    units = coor.shape[0] * K
    torch.cuda.empty_cache()
    return {"value": units / (ms * 1e-3), "unit": UNIT, "ms": ms,
            "sample": f"1 of {(c['frustum'] // rs) ** 2} chunks ({units} sample-slots): oracle decoder_forward + permute + composite as torch "
                      "eager fp32 (TF32 off) on cuda, encoder/slot attention/sampling excluded"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    value, dt, cores, sample = cpu_reference_run(steps, warm)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference", "config": dict(CFG, note="CPU reference arm; bounded sample, steps capped at 3"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def _shutdown(dist):
    """End of a multi-rank run.  The NCCL collectives of the step live inside CUDA graphs, and destroy_process_group() can block
    on a communicator that graphs still reference; everything has been printed and synchronised, so leave without teardown."""
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


# ----------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    from uorf_b200.eval_model import uorfEvalModel, default_eval_options
    from uorf_b200 import sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    c = CFG
    torch.manual_seed(2021)
    opt = default_eval_options(num_slots=c["K"], z_dim=c["z_dim"], n_samp=c["n_samp"], frustum_size=c["frustum"],
                               render_size=c["render_size"], input_size=c["input_size"], n_img_each_scene=c["n_views"],
                               gpu_ids=[local_rank])
    model = uorfEvalModel(opt, layout="native", precision=args.precision).eval()
    use_graph = world == 1 and not args.no_graph
    n_total = c["n_views"] * world
    img, c2w, azi = synthetic_scene(n_total, c["load_size"])
    lo, hi = sharded.shard_views(n_total, world, rank)
    # host (pinned) copies for the end-to-end arm; device-resident copies for the kernel arm.
    # every rank needs view 0's azimuth for the fg transform; rank 0 (lo == 0) holds image 0 for the encoder.
    host = {"img_data": img[lo:hi].clone().pin_memory(), "cam2world": c2w[lo:hi].clone().pin_memory(),
            "azi_rot": azi[0:1].clone().pin_memory()}
    resident = {k: v.to(dev) for k, v in host.items()}
    K, Z = c["K"], c["z_dim"]
    out_host = torch.empty(hi - lo, 3, c["frustum"], c["frustum"]).pin_memory()

    graph_on = [True]

    graphed_sharded = sharded.GraphedShardedRender(model, K, Z) if world > 1 else None

    def step(inputs):
        model.set_input(inputs)
        if world == 1:
            model.forward()
            return model.x_recon
        with torch.no_grad():
            if not (args.no_graph or not graph_on[0]):      # encode + broadcast + render replayed as one graph per rank
                return graphed_sharded()["rendered"] * 2 - 1
            z = model.encode()[0] if rank == 0 else None
            z = sharded.broadcast_slots(z, K, Z, dev)
            return model.render(z, model.cam2world, model.nss2cam0)["rendered"] * 2 - 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    sampler = ClockSampler(local_rank)
    # ---- kernel arm: inputs resident in HBM ----
    step(resident)
    if use_graph:
        model.enable_cuda_graph()
    for _ in range(max(args.warmup, 3)):
        step(resident)
    sampler.start()
    prof_range = bool(os.environ.get("UORF_PROFILE_RANGE"))    # ncu --profile-from-start off: list exactly the timed launches
    if prof_range:
        torch.cuda.profiler.start()
    ms_total = timed(lambda: step(resident), args.steps)
    if prof_range:
        torch.cuda.profiler.stop()
    # duration of the dominant kernel: CUDA events around the decoder launch, over the same steps launched eagerly
    # (events cannot be recorded inside a replayed graph); same stream, same inputs, same launch configuration
    if use_graph:
        model.enable_cuda_graph(False)
    graph_on[0] = False
    model.netDecoder.events = []
    for _ in range(args.steps):
        step(resident)
    torch.cuda.synchronize()
    ev = model.netDecoder.events
    model.netDecoder.events = None
    graph_on[0] = True
    dec_ms = sum(a.elapsed_time(b) for a, b in ev) / max(len(ev), 1)
    # keep the device busy for ~1 s so the sampler has readings under load (not part of any reported number); the number of
    # extra steps derives from the all-reduced time, so every rank runs the same count (the steps contain collectives)
    for _ in range(int(min(400, 1000.0 / max(ms_total / args.steps, 1e-3))) + 1):
        step(resident)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    if use_graph:
        model.enable_cuda_graph()
        step(resident)          # re-capture

    # ---- informational: the render alone (sampling -> decoder -> compositing), without the Encoder / SlotAttention prefix ----
    render_only_ms = None
    if world == 1:
        with torch.no_grad():
            z_fixed = model.z_slots.detach().clone()
            ro = (model.render if args.no_graph else model.render_graphed)
            for _ in range(3):
                ro(z_fixed, model.cam2world, model.nss2cam0)
            render_only_ms = timed(lambda: ro(z_fixed, model.cam2world, model.nss2cam0), args.steps) / args.steps

    # ---- end-to-end arm: host buffers, H2D + D2H inside the timed region ----
    def e2e_step():
        x_recon = step(host)
        out_host.copy_(x_recon, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    for _ in range(3):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)

    pts_rank = (hi - lo) * c["frustum"] ** 2 * c["n_samp"]
    units_total = n_total * c["frustum"] ** 2 * c["n_samp"] * K
    value = units_total / (ms_total / args.steps * 1e-3)
    e2e_value = units_total / (ms_e2e / args.steps * 1e-3)
    pk = peaks()
    ach_tflops = pts_rank * f_min_per_point(K, Z) / (dec_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "decoder_traffic.json")
    if os.path.exists(tpath) and args.workload is None:     # the ncu capture is of the headline workload
        traffic = json.load(open(tpath)).get(args.precision)
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = out_host.numel() * out_host.element_size()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.precision + " (tensor-core operands; fp32 accumulate, fp32 elsewhere)", "data": "synthetic",
                "config": dict(c, n_views_total=n_total, sharding=f"views over {world} rank(s), slot latents broadcast from rank 0",
                               layout="native ray-major, raws+masked+unmasked materialised",
                               launch=("eager launches" if args.no_graph else "forward replayed as one CUDA graph" if world == 1
                                       else "per rank one CUDA graph: [rank 0: encoder + slot attention] -> NCCL broadcast of the slot latents -> render"),
                               l2="no flush: each step streams ~0.8 GB of outputs per GPU (>> 126 MB L2)",
                               weights="random init, seed 2021",
                               render_only=(None if render_only_ms is None else
                                            {"ms_per_step": render_only_ms, "value": units_total / (render_only_ms * 1e-3),
                                             "note": "same step without the Encoder + SlotAttention prefix (SURVEY 8(d))"})),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                # launches of this library's own kernels per step on rank 0 (profiles/*_launches.md): encoder 6 (7 with
                # `bottom`), slot attention 8, weight image 1, sampling 1, decoder 1, compositing 1
                "gpu_launches": (18 + (1 if c.get("bottom") else 0)) * args.steps,
                "roofline": {"kernel": "decoder_fwd_kernel", "bound": "tensor", "achieved": ach_tflops, "peak": pk["tf_burst"],
                             "unit": "TFLOP/s", "frac": ach_tflops / pk["tf_burst"], "traffic": traffic,
                             "peak_source": pk["src"] + " bf16 burst", "kernel_ms": dec_ms,
                             "flops_per_point_algorithmic": f_min_per_point(K, Z), "points_per_launch": pts_rank,
                             "note": "achieved counts ALGORITHMIC flops once; the x3 modes issue 3 MMA passes for them"},
                "impl": "b200"}
        if not args.no_cpu_baseline and world == 1:
            v, dt, cores, sample = cpu_reference_run(1, 0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
            try:
                line["cpu_baseline"]["eager_gpu"] = eager_gpu_run(dev)
            except Exception as e:   # a baseline leg must never take the product line down with it
                line["cpu_baseline"]["eager_gpu"] = {"error": repr(e)[:200]}
        print(json.dumps(line))
    if world > 1:
        _shutdown(dist)


# ----------------------------------------------------------------------------------------------------
# training arm (BASELINE.json configs[2]): CLEVR-567 coarse-stage step, fwd + bwd + Adam, scene-sharded data parallel
# ----------------------------------------------------------------------------------------------------
TRAIN_CFG = dict(workload="clevr567_train_coarse_K8_F64_D64_N4", K=8, z_dim=64, n_views=4, frustum=64, n_samp=64,
                 supervision_size=64, input_size=64, load_size=128)


def run_train(args):
    import torch.distributed as dist
    from uorf_b200.nogan_model import uorfNoGanModel, default_train_options
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --mode train needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    c = TRAIN_CFG
    torch.manual_seed(2021)   # identical replicas
    opt = default_train_options(num_slots=c["K"], z_dim=c["z_dim"], n_samp=c["n_samp"], frustum_size=c["frustum"],
                                render_size=c["supervision_size"], supervision_size=c["supervision_size"], input_size=c["input_size"],
                                n_img_each_scene=c["n_views"], stage=c.get("stage", "coarse"), bottom=c.get("bottom", False),
                                frustum_size_fine=c.get("frustum_fine", 128), gpu_ids=[local_rank], warmup_steps=10)
    use_graph = not args.no_graph and opt.stage == "coarse"     # the fine stage draws its crop window on the host every step
    opt.cuda_graph = use_graph
    model = uorfNoGanModel(opt, precision=args.precision)
    if world > 1:
        model.process_group = dist.group.WORLD
    if use_graph:
        model.enable_cuda_graph()
    img, c2w, azi = synthetic_scene(c["n_views"], c["load_size"], seed=2021 + rank)   # one scene per rank
    host = {"img_data": img.pin_memory(), "cam2world": c2w.pin_memory(), "azi_rot": azi.pin_memory()}
    resident = {k: v.to(dev) for k, v in host.items()}
    loss_host = torch.empty(1).pin_memory()

    def step(inputs):
        model.set_input(inputs)
        model.optimize_parameters(epoch=opt.percept_in)    # perceptual loss active: the steady-state step of the schedule
        model.update_learning_rate()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    sampler = ClockSampler(local_rank)
    for _ in range(5 if use_graph else 0):      # three eager steps + the capture happen before the warm-up
        step(resident)
    for _ in range(max(args.warmup, 3)):
        step(resident)
    sampler.start()
    prof_range = bool(os.environ.get("UORF_PROFILE_RANGE"))
    if prof_range:
        torch.cuda.profiler.start()
    ms_total = timed(lambda: step(resident), args.steps)
    if prof_range:
        torch.cuda.profiler.stop()
    for _ in range(int(min(400, 1000.0 / max(ms_total / args.steps, 1e-3))) + 1):   # same count on every rank (collectives inside)
        step(resident)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    # duration of the two decoder calls: CUDA events around them over a few eagerly launched steps (events cannot be recorded
    # inside a replayed graph); same stream, same inputs, same launch configuration
    fwd_ms = bwd_ms = None
    if world == 1:
        model.enable_cuda_graph(False) if use_graph else None
        model.netDecoder.events, model.netDecoder.events_bwd = [], []
        for _ in range(min(args.steps, 5)):
            step(resident)
        torch.cuda.synchronize()
        ef, eb = model.netDecoder.events, model.netDecoder.events_bwd
        model.netDecoder.events = model.netDecoder.events_bwd = None
        fwd_ms = sum(a.elapsed_time(b) for a, b in ef) / max(len(ef), 1)
        bwd_ms = sum(a.elapsed_time(b) for a, b in eb) / max(len(eb), 1)
        if use_graph:
            model.enable_cuda_graph()
            for _ in range(5):
                step(resident)

    def e2e_step():
        step(host)
        loss_host.copy_(model.loss_recon.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    for _ in range(3):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    if rank == 0:
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        line = {"metric": "train_scene_steps_per_sec", "value": world * args.steps / (ms_total * 1e-3), "unit": "scene-steps/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.precision + " forward; fp16 tensor-core backward with fp32 accumulation; fp32 elsewhere", "data": "synthetic",
                "config": dict(c, step="Encoder+SlotAttention+sampling+decoder+compositing fwd+bwd, MSE + VGG16-relu4_3 perceptual "
                               "(random-init VGG, weight_percept 0.006 active: epoch = percept_in), Adam",
                               parallelism=f"scene-sharded data parallel x{world}, one flat gradient all-reduce (NCCL)",
                               launch=("whole step (forward + backward" + (" + gradient all-reduce" if world > 1 else "") + " + Adam) replayed as one CUDA graph")
                               if use_graph else "eager launches",
                               loss_last=float(model.loss_recon.detach())),
                "clocks": clocks,
                "e2e": {"value": world * args.steps / (ms_e2e * 1e-3), "unit": "scene-steps/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
                # launches of this library's own kernels per step and rank (profiles/*_launches_train.md): slot attention 8 + 10,
                # sampling 1, weight image 2, decoder fwd 1 + bwd 4 (compose, main, reduce, unfold), compositing 1 + 1
                "gpu_launches": 27 * args.steps,
                "impl": "b200"}
        if bwd_ms:
            pk = peaks()
            pts = c["n_views"] * c["frustum"] ** 2 * c["n_samp"]
            fl = f_min_per_point(c["K"], c["z_dim"])
            ach = 4.0 * fl * pts / (bwd_ms * 1e-3) / 1e12
            line["roofline"] = {"kernel": "decoder_bwd_kernel (+ composition-backward, reduce, unfold)", "bound": "tensor", "achieved": ach,
                                "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": ach / pk["tf_burst"], "traffic": None,
                                "peak_source": pk["src"] + " bf16 burst", "kernel_ms": bwd_ms, "decoder_fwd_kernel_ms": fwd_ms,
                                "flops_per_point_algorithmic": 4.0 * fl, "points_per_launch": pts,
                                "note": "algorithmic flops = 4 x forward F_min (recompute + dX + dW at M=64 half rate)"}
        print(json.dumps(line))
    if world > 1:
        _shutdown(dist)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fp16x3", choices=["fp16x3", "bf16x3", "bf16", "fp16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default=None, choices=[None] + sorted(WORKLOADS) + sorted(TRAIN_WORKLOADS),
                    help="another BASELINE.json shape instead of the headline one (clevr567: render; room_diverse_fine: --mode train)")
    ap.add_argument("--no-graph", action="store_true", help="launch the render step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--mode", default="render", choices=["render", "train"],
                    help="render: the headline eval-render metric (default); train: CLEVR-567 coarse training step")
    args = ap.parse_args()
    global CFG, TRAIN_CFG
    if args.workload in WORKLOADS:
        CFG = WORKLOADS[args.workload]
    elif args.workload in TRAIN_WORKLOADS:
        TRAIN_CFG = TRAIN_WORKLOADS[args.workload]
        args.mode = "train"
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "train":
        run_train(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
