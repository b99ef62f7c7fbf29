#!/usr/bin/env python
"""bench.py - the uORF per-scene render hot path on B200: BASELINE.json's whole metric in one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--precision fp16x3|bf16x3|bf16|fp16]
                    [--mode all|render|train|strong] [--workload clevr567|room_diverse_fine] [--no-graph] [--no-cpu-baseline]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...       (N > 1, one rank per GPU)

BASELINE.json's metric is "composited ray-samples/s x K slots AND train steps/s at 1/2/4/8 B200".  The headline fields of the
line are the render half on the configuration the metric is quoted on; the sub-records carry the rest, all from this one run:

  headline      configs[1]  Room-Chair novel-view render, K=5, z_dim 64, 128x128 frustum x 64 samples, 4 views per rank
                (weak scaling: rays shard by view, rank 0 encodes and broadcasts the slot latents, the rendered tiles are
                all-gathered inside the timed step when N > 1).  One STEP = one scene through Encoder -> SlotAttention ->
                Projection -> Decoder -> raw2outputs (reference models/uorf_eval_model.py:116-157), per-slot masked /
                unmasked raws materialised as the reference does.  value = P_total * K / t.
  "train"       configs[2]  CLEVR-567 coarse-stage training step (K=8, 64^3 frustum, 4 views): forward + backward + Adam
                (reference models/uorf_nogan_model.py:229-245), one scene per rank, ONE flat gradient all-reduce (NCCL) when
                N > 1; scene-steps/s summed over ranks.
  "train_fine"  configs[3]  Room-Diverse fine stage (K=5, --bottom, 128 frustum, 64x64 crop drawn on the device).
  "strong"      configs[4]  scene-manipulation render, ONE view of 256x256 rays x 128 samples, K in {5, 8, 16}: image rows
                sharded over the N ranks, slot latents broadcast, tiles all-gathered inside the timed step (reference
                models/uorf_manip_model.py:196-232).  Total work is fixed: ms per scene should fall ~1/N (strong scaling).

Every step is replayed as ONE CUDA graph per rank (NCCL collectives captured inside); `--no-graph` launches eagerly.
The line carries `roofline` (dominant kernel: CUDA-event time over eagerly launched steps, algorithmic flops counted once),
`cpu_baseline` (the unmodified reference, oracle/_ref, on the host cores and - `eager_gpu` - on the same GPU), `e2e` (host
buffers, H2D + D2H inside the timed region), `clocks` (NVML samples under load) and `gpu_launches` (counted by the library:
uorf_launch_count).  `UORF_PROFILE_RANGE=1` brackets the headline's timed steps with cudaProfilerStart/Stop.

`--impl reference` times the reference's own CPU implementation of the headline step - the unmodified reference's
uorfEvalModel.forward from oracle/_ref (scripts/install_ref.sh), or the oracle port when that copy is missing - on the host
cores, all 4 views per step, a bounded number of steps.
"""
from __future__ import annotations

import argparse
import gc
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
# the other render shape BASELINE.json names (`--workload clevr567` runs it through the same arms for the record)
WORKLOADS = {
    "clevr567": dict(workload="clevr567_eval_K8_F64_D64_N4", K=8, z_dim=64, n_views=4, frustum=64, n_samp=64, render_size=64,
                     input_size=64, load_size=128),
}
TRAIN_CFG = dict(workload="clevr567_train_coarse_K8_F64_D64_N4", K=8, z_dim=64, n_views=4, frustum=64, n_samp=64,
                 supervision_size=64, input_size=64, load_size=128)
TRAIN_FINE_CFG = dict(workload="room_diverse_train_fine_K5_F128crop64_D64_N4", K=5, z_dim=64, n_views=4, frustum=64, n_samp=64,
                      supervision_size=64, input_size=128, load_size=128, stage="fine", bottom=True, frustum_fine=128)
STRONG_CFG = dict(frustum=256, n_samp=128, z_dim=64, n_views=1, input_size=64, load_size=128, slots=(5, 8, 16))
METRIC = "composited_ray_samples_x_slots_per_sec"
UNIT = "sample-slots/s"
HBM_FALLBACK_GBS, BF16_FALLBACK_TFLOPS = 6650.0, 1590.0
L2_NOTE = "no flush: each step streams ~0.8 GB of outputs per GPU (>> 126 MB L2)"


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


def headline_config():
    """`config` of the JSON line - identical for the b200 and the reference arm (the driver compares them)"""
    return dict(CFG, l2=L2_NOTE, weights="random init, seed 2021")


# ----------------------------------------------------------------------------------------------------
# clocks sampler (NVML), runs during the timed regions
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
# reference legs (the only places that execute oracle/): the unmodified reference from oracle/_ref, else the oracle port
# ----------------------------------------------------------------------------------------------------
def _ref_eval_model(device_ids):
    """the reference's uorfEvalModel (models/uorf_eval_model.py) built by its own factory on the headline configuration"""
    from oracle import ref_harness as R
    c = CFG
    torch.manual_seed(2021)
    model, _ = R.create_model("uorf_eval", False, num_slots=c["K"], z_dim=c["z_dim"], n_samp=c["n_samp"], frustum_size=c["frustum"],
                              render_size=c["render_size"], input_size=c["input_size"], n_img_each_scene=c["n_views"],
                              gpu_ids=list(device_ids))
    return model


def cpu_reference_run(steps: int, warmup: int):
    """-> (value, seconds per step, cores, kind, sample).  kind 'reference': uorfEvalModel.forward of the unmodified reference
    (oracle/_ref), all views; kind 'port': oracle.eval_render when that copy is not installed."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    c = CFG
    img, c2w, azi = synthetic_scene(c["n_views"], c["load_size"])
    units = c["n_views"] * c["frustum"] ** 2 * c["n_samp"] * c["K"]
    from oracle import ref_harness as R
    if R.available():
        model = _ref_eval_model([])
        model.set_input({"img_data": img, "cam2world": c2w, "azi_rot": azi, "paths": []})

        def one():
            with torch.no_grad():
                model.forward()
        kind = "reference"
        what = "unmodified reference uorfEvalModel.forward (oracle/_ref, models/uorf_eval_model.py:116-157)"
    else:
        from oracle import uorf_oracle as O
        spec = O.FrustumSpec([c["frustum"], c["frustum"], c["n_samp"]], near=6.0, far=20.0, nss_scale=7.0, render_size=c["render_size"])
        enc_p, sa_p, dec_p = O.init_encoder_params(c["z_dim"]), O.init_slot_attention_params(c["z_dim"]), O.init_decoder_params(c["z_dim"])
        g = torch.Generator().manual_seed(7)
        nfg, nbg = torch.randn(1, c["K"] - 1, c["z_dim"], generator=g), torch.randn(1, 1, c["z_dim"], generator=g)

        def one():
            return O.eval_render(enc_p, sa_p, dec_p, img, c2w, azi, spec, num_slots=c["K"], input_size=c["input_size"],
                                 noise_fg=nfg, noise_bg=nbg)
        kind = "port"
        what = "oracle eval_render (oracle/_ref not installed)"
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return units / dt, dt, cores, kind, f"all {c['n_views']} views per step ({units} sample-slots), {what}, {cores} threads, {steps} step(s)"


def eager_gpu_reference(dev, reps: int = 3):
    """The real bar: the reference's OWN eval driver, unmodified, as stock torch eager ops on this GPU (TF32 off as the
    reference's set_seed does) - the whole step Encoder -> ... -> raw2outputs, all views.  A baseline leg like cpu_baseline."""
    from oracle import ref_harness as R
    if not R.available():
        return {"unavailable": "oracle/_ref not installed (scripts/install_ref.sh)"}
    c = CFG
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = _ref_eval_model([dev.index])
    img, c2w, azi = synthetic_scene(c["n_views"], c["load_size"])
    model.set_input({"img_data": img, "cam2world": c2w, "azi_rot": azi, "paths": []})
    with torch.no_grad():
        model.forward()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            model.forward()
        e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    units = c["n_views"] * c["frustum"] ** 2 * c["n_samp"] * c["K"]
    del model
    gc.collect()
    torch.cuda.empty_cache()
    return {"value": units / (ms * 1e-3), "unit": UNIT, "ms": ms,
            "sample": f"all {c['n_views']} views ({units} sample-slots): unmodified reference uorfEvalModel.forward (oracle/_ref) as torch "
                      "eager fp32 (TF32 off) on cuda, whole step"}


def eager_gpu_reference_train(dev, c, reps: int = 3):
    """the reference's own training step (uorfNoGanModel.optimize_parameters, unmodified, torch eager + autograd) on this GPU"""
    from oracle import ref_harness as R
    if not R.available():
        return {"unavailable": "oracle/_ref not installed (scripts/install_ref.sh)"}
    torch.manual_seed(2021)
    model, opt = R.create_model("uorf_nogan", True, num_slots=c["K"], z_dim=c["z_dim"], n_samp=c["n_samp"], frustum_size=c["frustum"],
                                frustum_size_fine=c.get("frustum_fine", 128), render_size=c["supervision_size"],
                                supervision_size=c["supervision_size"], input_size=c["input_size"], n_img_each_scene=c["n_views"],
                                stage=c.get("stage", "coarse"), bottom=c.get("bottom", False), gpu_ids=[dev.index], warmup_steps=10)
    img, c2w, azi = synthetic_scene(c["n_views"], c["load_size"])
    model.set_input({"img_data": img, "cam2world": c2w, "azi_rot": azi, "paths": []})
    model.optimize_parameters(epoch=opt.percept_in)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        model.optimize_parameters(epoch=opt.percept_in)
        model.update_learning_rate()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    peak_gb = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    del model
    gc.collect()
    torch.cuda.empty_cache()
    return {"value": 1e3 / ms, "unit": "scene-steps/s", "ms": ms, "peak_mem_gib": round(peak_gb, 1),
            "sample": "unmodified reference uorfNoGanModel.optimize_parameters (oracle/_ref) as torch eager + autograd on cuda, same shapes"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 2))            # one step is the whole 4-view scene: ~10-20 s on the host cores
    warm = 1 if args.warmup > 0 else 0
    value, dt, cores, kind, sample = cpu_reference_run(steps, warm)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference", "config": headline_config(),
            "detail": {"note": "CPU reference arm: whole scene per step, steps capped at 2 and warm-up at 1 to bound the run"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
# process context
# ----------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device(f"cuda:{self.local_rank}")
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.graph_owners = []       # objects holding CUDA graphs with captured NCCL work: released before the group is destroyed

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps, finish=None):
        """K steps bracketed by barrier + synchronize on both sides, device-timed, MAX over ranks -> total ms
        (finish: called after the last step and before the closing event, e.g. to make the timed stream wait for side streams)"""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return ms.item()

    def shutdown(self):
        """Ordered teardown: the step graphs hold captured NCCL kernels, so they are destroyed FIRST (their owners release them),
        then the communicator.  A watchdog ends the process if NCCL teardown still blocks: every line has been printed by then."""
        if self.world == 1:
            return
        torch.cuda.synchronize()
        self.dist.barrier()
        for o in self.graph_owners:
            o.release()
        self.graph_owners = []
        gc.collect()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        t = threading.Timer(20.0, lambda: os._exit(0))
        t.daemon = True
        t.start()
        self.dist.destroy_process_group()
        t.cancel()


def keep_busy(step, ms_per_step):
    """~0.5 s more of the same steps so the clock sampler has readings under load (not part of any reported number); the count
    derives from an all-reduced time, so every rank runs the same number (the steps contain collectives)"""
    for _ in range(int(min(300, 500.0 / max(ms_per_step, 1e-3))) + 1):
        step()
    torch.cuda.synchronize()


# ----------------------------------------------------------------------------------------------------
# headline: eval render (BASELINE.json configs[1])
# ----------------------------------------------------------------------------------------------------
def bench_render(ctx, args):
    from uorf_b200 import _lib, sharded
    from uorf_b200.eval_model import uorfEvalModel, default_eval_options
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    c = CFG
    torch.manual_seed(2021)
    opt = default_eval_options(num_slots=c["K"], z_dim=c["z_dim"], n_samp=c["n_samp"], frustum_size=c["frustum"],
                               render_size=c["render_size"], input_size=c["input_size"], n_img_each_scene=c["n_views"],
                               gpu_ids=[ctx.local_rank])
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
    graphed_sharded = None
    if world > 1:
        graphed_sharded = sharded.GraphedShardedRender(model, K, Z, gather=True)
        ctx.graph_owners.append(graphed_sharded)

    def step(inputs):
        model.set_input(inputs)
        if world == 1:
            model.forward()
            return model.x_recon
        return step_render_only()

    def step_render_only():
        with torch.no_grad():
            if not (args.no_graph or not graph_on[0]):      # encode + broadcast + render + all-gather replayed as one graph per rank
                return graphed_sharded()["rendered"] * 2 - 1
            z = model.encode()[0] if rank == 0 else None
            z = sharded.broadcast_slots(z, K, Z, dev)
            return model.render(z, model.cam2world, model.nss2cam0)["rendered"] * 2 - 1

    sampler = ClockSampler(ctx.local_rank)
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
    ms_total = ctx.timed(lambda: step(resident), args.steps)
    if prof_range:
        torch.cuda.profiler.stop()
    # duration of the dominant kernel: CUDA events around the decoder launch, over the same steps launched eagerly
    # (events cannot be recorded inside a replayed graph); same stream, same inputs, same launch configuration.
    # The library's launch counter over these eager steps is the number of its kernels one step runs.
    if use_graph:
        model.enable_cuda_graph(False)
    graph_on[0] = False
    model.netDecoder.events = []
    n0 = _lib.launch_count()
    for _ in range(args.steps):
        step(resident)
    torch.cuda.synchronize()
    launches_per_step = (_lib.launch_count() - n0) / args.steps
    ev = model.netDecoder.events
    model.netDecoder.events = None
    graph_on[0] = True
    dec_ms = sum(a.elapsed_time(b) for a, b in ev) / max(len(ev), 1)
    if use_graph:
        model.enable_cuda_graph()
        step(resident)          # re-capture
    keep_busy(lambda: step(resident), ms_total / args.steps)
    clocks = sampler.stop()

    # ---- informational: the render alone (sampling -> decoder -> compositing), without the Encoder / SlotAttention prefix ----
    render_only_ms = None
    if world == 1:
        with torch.no_grad():
            z_fixed = model.z_slots.detach().clone()
            ro = (model.render if args.no_graph else model.render_graphed)
            for _ in range(3):
                ro(z_fixed, model.cam2world, model.nss2cam0)
            render_only_ms = ctx.timed(lambda: ro(z_fixed, model.cam2world, model.nss2cam0), args.steps) / args.steps

    # ---- end-to-end arm: host buffers, H2D + D2H inside the timed region ----
    # Every step copies its inputs from pinned host memory and its image back to pinned host memory.  The caller keeps two steps in
    # flight, as a render server would: the H2D copy of step i's inputs runs on a copy stream under the render of step i-1, the
    # image of step i goes through a device staging buffer and back to the host on a second copy stream under the render of step
    # i+1, and the host waits for step i-1's image only after step i has been enqueued.  The closing event of the timed region is
    # recorded after the main stream has waited for both copy streams.
    main = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    out_ring = [out_host, torch.empty_like(out_host).pin_memory()]
    dev_ring, ev_out, ev_host = [None, None], [None, None], [None]
    e2e_count = [0]

    def e2e_step():
        i = e2e_count[0] & 1
        e2e_count[0] += 1
        s_in.wait_stream(main) if e2e_count[0] == 1 else None
        with torch.cuda.stream(s_in):                # H2D of this step's inputs (fresh device tensors), off the render stream
            model.set_input(host)
            for t_ in (model.x, model.cam2world, getattr(model, "cam2world_azi", None), model.nss2cam0):
                if t_ is not None:
                    t_.record_stream(main)
        main.wait_stream(s_in)
        if world == 1:
            model.forward()
            x_recon = model.x_recon
        else:
            x_recon = step_render_only()
        if dev_ring[i] is None:
            dev_ring[i] = torch.empty_like(x_recon)
        if ev_out[i] is not None:
            main.wait_event(ev_out[i])               # the staging buffer's previous image has left for the host
        dev_ring[i].copy_(x_recon)
        s_out.wait_stream(main)
        with torch.cuda.stream(s_out):
            out_ring[i].copy_(dev_ring[i], non_blocking=True)
            ev_out[i] = torch.cuda.Event()
            ev_out[i].record()
        if ev_host[0] is not None:
            ev_host[0].synchronize()                 # the image of the previous step is on the host now
        ev_host[0] = ev_out[i]

    def e2e_finish():
        main.wait_stream(s_out)
        main.wait_stream(s_in)
    for _ in range(3):
        e2e_step()
    ms_e2e = ctx.timed(e2e_step, args.steps, finish=e2e_finish)
    ev_host[0].synchronize()

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
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision + " (tensor-core operands; fp32 accumulate, fp32 elsewhere)", "data": "synthetic",
            "config": headline_config(),
            "detail": {"n_views_total": n_total,
                       "sharding": f"views over {world} rank(s), slot latents broadcast from rank 0" + (", rendered tiles all-gathered in the timed step" if world > 1 else ""),
                       "layout": "native ray-major, raws+masked+unmasked materialised",
                       "launch": ("eager launches" if args.no_graph else "forward replayed as one CUDA graph" if world == 1
                                  else "per rank one CUDA graph: [rank 0: encoder + slot attention] -> NCCL broadcast of the slot latents -> render -> NCCL all-gather of the tiles"),
                       "render_only": (None if render_only_ms is None else
                                       {"ms_per_step": render_only_ms, "value": units_total / (render_only_ms * 1e-3),
                                        "note": "same step without the Encoder + SlotAttention prefix (SURVEY 8(d))"})},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps,
                    "pipeline": "two steps in flight: the H2D copy of step i's inputs (pinned host memory, copy stream) and the D2H copy of "
                                "its image (second copy stream, device staging buffer) overlap the renders of the neighbouring steps; "
                                "the host waits for step i-1's image after step i has been enqueued; the closing event follows both copy streams"},
            # kernels of libuorf_b200.so launched inside the timed region on rank 0: uorf_launch_count() over one eagerly launched
            # step x the timed steps (the timed steps replay exactly those launches from a graph)
            "gpu_launches": int(round(launches_per_step * args.steps)),
            "gpu_launches_per_step": launches_per_step,
            "roofline": {"kernel": "decoder_fwd_kernel", "bound": "tensor", "achieved": ach_tflops, "peak": pk["tf_burst"],
                         "unit": "TFLOP/s", "frac": ach_tflops / pk["tf_burst"], "traffic": traffic,
                         "peak_source": pk["src"] + " bf16 burst", "kernel_ms": dec_ms,
                         "flops_per_point_algorithmic": f_min_per_point(K, Z), "points_per_launch": pts_rank,
                         "note": "achieved counts ALGORITHMIC flops once; the x3 modes issue 3 MMA passes for them"},
            "impl": "b200"}
    if use_graph:
        model.enable_cuda_graph(False)
    if graphed_sharded is not None:
        graphed_sharded.release()
    del model
    gc.collect()
    torch.cuda.empty_cache()
    return line


# ----------------------------------------------------------------------------------------------------
# training step (BASELINE.json configs[2] coarse / configs[3] fine): fwd + bwd + Adam, scene-sharded data parallel
# ----------------------------------------------------------------------------------------------------
def bench_train(ctx, args, c, with_reference_gpu=True):
    from uorf_b200 import _lib
    from uorf_b200.nogan_model import uorfNoGanModel, default_train_options
    world, rank, dev, dist = ctx.world, ctx.rank, ctx.dev, ctx.dist
    torch.manual_seed(2021)   # identical replicas
    opt = default_train_options(num_slots=c["K"], z_dim=c["z_dim"], n_samp=c["n_samp"], frustum_size=c["frustum"],
                                render_size=c["supervision_size"], supervision_size=c["supervision_size"], input_size=c["input_size"],
                                n_img_each_scene=c["n_views"], stage=c.get("stage", "coarse"), bottom=c.get("bottom", False),
                                frustum_size_fine=c.get("frustum_fine", 128), gpu_ids=[ctx.local_rank], warmup_steps=10,
                                vgg_random_init=True)       # no network for the ImageNet weights: arithmetic-identical random VGG
    use_graph = not args.no_graph
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

    sampler = ClockSampler(ctx.local_rank)
    # eager steps first (cuDNN autotuning, optimizer state); the library's launch counter over one of them = its kernels per step
    step(resident)
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    step(resident)
    launches_per_step = _lib.launch_count() - n0
    for _ in range(3 if use_graph else 0):      # the remaining eager step(s) + the capture happen before the warm-up
        step(resident)
    for _ in range(max(args.warmup, 3)):
        step(resident)
    sampler.start()
    prof_range = bool(os.environ.get("UORF_PROFILE_RANGE"))    # ncu --profile-from-start off: list exactly the timed launches
    if prof_range:
        torch.cuda.profiler.start()
    ms_total = ctx.timed(lambda: step(resident), args.steps)
    if prof_range:
        torch.cuda.profiler.stop()
    keep_busy(lambda: step(resident), ms_total / args.steps)
    clocks = sampler.stop()

    def e2e_step():
        step(host)
        loss_host.copy_(model.loss_recon.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    for _ in range(3):
        e2e_step()
    ms_e2e = ctx.timed(e2e_step, args.steps)
    loss_last = float(loss_host[0])
    # duration of the two decoder calls: CUDA events around them over a few eagerly launched steps (events cannot be recorded
    # inside a replayed graph); same stream, same inputs, same launch configuration.  Every rank runs them (collectives inside).
    if use_graph:
        model.enable_cuda_graph(False)
    model.netDecoder.events, model.netDecoder.events_bwd = [], []
    for _ in range(min(args.steps, 5)):
        step(resident)
    torch.cuda.synchronize()
    ef, eb = model.netDecoder.events, model.netDecoder.events_bwd
    model.netDecoder.events = model.netDecoder.events_bwd = None
    fwd_ms = sum(a.elapsed_time(b) for a, b in ef) / max(len(ef), 1)
    bwd_ms = sum(a.elapsed_time(b) for a, b in eb) / max(len(eb), 1)
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    pk = peaks()
    pts = c["n_views"] * c["supervision_size"] ** 2 * c["n_samp"]
    fl = f_min_per_point(c["K"], c["z_dim"])
    ach = 4.0 * fl * pts / (bwd_ms * 1e-3) / 1e12 if bwd_ms else None
    rec = {"metric": "train_scene_steps_per_sec", "workload": c["workload"], "value": world * args.steps / (ms_total * 1e-3),
           "unit": "scene-steps/s", "n_gpus": world, "steps": args.steps, "ms_per_step": ms_total / args.steps, "scaling": "weak",
           "dtype": args.precision + " decoder forward; fp16 tensor-core decoder backward with fp32 accumulation; VGG convolutions fp16x3 forward / "
                    "bf16x3 data gradient on the tensor cores (fp32 parity); fp32 elsewhere",
           "config": dict(c, step="Encoder+SlotAttention+sampling+decoder+compositing fwd+bwd, MSE + VGG16-relu4_3 perceptual "
                          "(random-init VGG, weight_percept 0.006 active: epoch = percept_in), Adam (reference uorf_nogan_model.py:229-245)",
                          parallelism=f"scene-sharded data parallel x{world}: one scene per rank, ONE flat-bucket gradient all-reduce (NCCL, average)"),
           "launch": ("whole step (forward + backward" + (" + gradient all-reduce" if world > 1 else "") + " + Adam) replayed as one CUDA graph")
           if use_graph else "eager launches",
           "loss_last": loss_last, "clocks": clocks,
           "e2e": {"value": world * args.steps / (ms_e2e * 1e-3), "unit": "scene-steps/s", "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
           "gpu_launches": int(launches_per_step) * args.steps, "gpu_launches_per_step": int(launches_per_step),
           "roofline": {"kernel": "decoder_bwd_kernel (+ composition-backward, reduce, unfold)", "bound": "tensor", "achieved": ach,
                        "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": (ach / pk["tf_burst"]) if ach else None, "traffic": None,
                        "peak_source": pk["src"] + " bf16 burst", "kernel_ms": bwd_ms, "decoder_fwd_kernel_ms": fwd_ms,
                        "flops_per_point_algorithmic": 4.0 * fl, "points_per_launch": pts,
                        "note": "algorithmic flops = 4 x forward F_min (recompute + dX + dW at M=64 half rate)"}}
    model._graph = None
    del model
    gc.collect()
    torch.cuda.empty_cache()
    if with_reference_gpu and world == 1 and not args.no_cpu_baseline:
        try:
            rec["reference_eager_gpu"] = eager_gpu_reference_train(dev, c)
        except Exception as e:   # a baseline leg must never take the product line down with it
            rec["reference_eager_gpu"] = {"error": repr(e)[:300]}
    return rec


# ----------------------------------------------------------------------------------------------------
# strong scaling (BASELINE.json configs[4]): one 256x256x128 view, rows sharded over the ranks, tiles gathered in the step
# ----------------------------------------------------------------------------------------------------
def bench_strong(ctx, args):
    from uorf_b200 import sharded
    from uorf_b200.eval_model import default_eval_options
    from uorf_b200.manip_model import uorfManipModel
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    s = STRONG_CFG
    F_, D, Z = s["frustum"], s["n_samp"], s["z_dim"]
    if F_ % world != 0:
        return {"error": f"{F_} image rows do not divide over {world} ranks"}
    h0, h1 = sharded.shard_range(F_, world, rank)
    pk = peaks()
    runs = []
    for K in s["slots"]:
        torch.manual_seed(2021)
        opt = default_eval_options(num_slots=K, z_dim=Z, n_samp=D, frustum_size=F_, render_size=F_, input_size=s["input_size"],
                                   n_img_each_scene=1, gpu_ids=[ctx.local_rank])
        m = uorfManipModel(opt, layout="native", precision=args.precision).eval()
        m.netDecoder.emit = ("raws",)       # the reference stores per-slot raws of the moved render but never reads them (:205-216)
        img, c2w, azi = synthetic_scene(1, s["load_size"])
        m.set_input({"img_data": img, "cam2world": c2w, "azi_rot": azi, "paths": [], "movement": torch.tensor([1.0, 0.5, 0.0])})
        g = torch.Generator().manual_seed(7)
        m.slot_noise = (torch.randn(1, K - 1, Z, generator=g).to(dev), torch.randn(1, 1, Z, generator=g).to(dev))
        offsets = torch.zeros(K - 1, 3, device=dev)
        offsets[1] = m.movement[0] / opt.nss_scale           # object 1 moved (uorf_manip_model.py:203)
        gr = sharded.GraphedShardedRender(m, K, Z, rows=(h0, h1), fg_slot_offset=offsets, gather=True)
        ctx.graph_owners.append(gr)

        step = gr.run_eager if args.no_graph else gr
        for _ in range(max(args.warmup, 3)):
            out = step()
        ms = ctx.timed(step, args.steps) / args.steps
        full = out["full"]
        ok = bool(torch.isfinite(full).all().item()) and tuple(full.shape) == (1, 3, F_, F_)
        # decoder kernel time on this rank's rows (eager launches, CUDA events)
        with torch.no_grad():
            z_fixed = m.encode()[0].clone()
            m.netDecoder.events = []
            for _ in range(3):
                m.render(z_fixed, m.cam2world, m.nss2cam0, fg_slot_offset=offsets, rows=(h0, h1))
        torch.cuda.synchronize()
        ev = m.netDecoder.events
        m.netDecoder.events = None
        dec_ms = sum(a.elapsed_time(b) for a, b in ev) / max(len(ev), 1)
        pts_rank = (h1 - h0) * F_ * D
        ach = pts_rank * f_min_per_point(K, Z) / (dec_ms * 1e-3) / 1e12
        units = F_ * F_ * D * K
        runs.append({"workload": f"manip_K{K}_F{F_}_D{D}", "K": K, "ms_per_step": ms, "value": units / (ms * 1e-3), "unit": UNIT,
                     "rows_per_rank": h1 - h0, "points_per_rank": pts_rank, "decoder_kernel_ms": dec_ms,
                     "decoder_roofline_frac": ach / pk["tf_burst"], "gathered_image_ok": ok})
        gr.release()
        del m, gr
        gc.collect()
        torch.cuda.empty_cache()
    return {"metric": METRIC, "scaling": "strong", "n_gpus": world, "steps": args.steps, "gather_in_timed_region": True,
            "dtype": args.precision,
            "step": "one view, 256x256 rays x 128 samples, one object moved: [rank 0: encoder + slot attention] -> NCCL broadcast of the "
                    "slot latents -> each rank renders its block of image rows -> NCCL all-gather of the tiles; one CUDA graph per rank "
                    "(reference models/uorf_manip_model.py:196-232; only `raws` is materialised)",
            "limiter": "the encoder + slot-attention prefix (~0.2 ms, rank 0 only) and the two collectives do not shrink with N; the "
                       "render itself is tensor-bound and divides by N",
            "runs": runs}


# ----------------------------------------------------------------------------------------------------
def run_b200(args):
    ctx = Ctx()
    rank = ctx.rank
    line = None
    try:
        if args.mode in ("all", "render"):
            line = bench_render(ctx, args)
            if not args.no_cpu_baseline and ctx.world == 1 and rank == 0:
                v, dt, cores, kind, sample = cpu_reference_run(1, 0)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
                try:
                    line["cpu_baseline"]["eager_gpu"] = eager_gpu_reference(ctx.dev)
                except Exception as e:   # a baseline leg must never take the product line down with it
                    line["cpu_baseline"]["eager_gpu"] = {"error": repr(e)[:300]}
        subs = {}
        if args.mode in ("all", "train"):
            subs["train"] = bench_train(ctx, args, TRAIN_CFG)
        if args.mode == "all":
            subs["train_fine"] = bench_train(ctx, args, TRAIN_FINE_CFG)
        if args.mode in ("all", "strong"):
            subs["strong"] = bench_strong(ctx, args)
        if line is None:          # a single sub-record requested: it is the line
            name, rec = next(iter(subs.items()))
            line = dict(rec, impl="b200", higher_is_better=True, warmup=max(args.warmup, 3), vs_baseline=None, data="synthetic")
            subs.pop(name)
        line.update(subs)
        if rank == 0:
            print(json.dumps(line))
            sys.stdout.flush()
    finally:
        ctx.shutdown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fp16x3", choices=["fp16x3", "bf16x3", "bf16", "fp16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default=None, choices=[None] + sorted(WORKLOADS) + ["room_diverse_fine"],
                    help="another BASELINE.json shape instead of the default one (clevr567: render; room_diverse_fine: --mode train)")
    ap.add_argument("--no-graph", action="store_true", help="launch the steps eagerly instead of replaying CUDA graphs")
    ap.add_argument("--mode", default="all", choices=["all", "render", "train", "strong"],
                    help="all (default): the headline render line with the train / train_fine / strong sub-records; "
                         "render / train / strong: only that part")
    args = ap.parse_args()
    global CFG, TRAIN_CFG
    if args.workload in WORKLOADS:
        CFG = WORKLOADS[args.workload]
        if args.mode == "all":
            args.mode = "render"
    elif args.workload == "room_diverse_fine":
        TRAIN_CFG = TRAIN_FINE_CFG
        args.mode = "train"
    if args.impl == "reference":
        run_reference(args)
    else:
        try:
            run_b200(args)
        except Exception:
            # a kernel's bounded wait records its site in pinned host memory before it traps (csrc/common.cuh mbar_wait*): readable
            # after the launch failure that follows
            try:
                from uorf_b200._lib import lib
                print(f"[bench] uorf_debug_flag (failing wait site, 0 = none) = {lib().uorf_debug_flag()}", file=sys.stderr, flush=True)
            except Exception:
                pass
            raise


if __name__ == "__main__":
    main()
