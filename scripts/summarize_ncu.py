"""Turns the ncu outputs a GPU visit leaves in gpurun_out/ into the committed summaries under profiles/.

    python scripts/summarize_ncu.py <tag>      # e.g. r1_v1

Reads gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum launch list of `bench.py --steps 2 --warmup 1`)
and gpurun_out/prof_decoder.ncu-rep (ncu --set full capture of one decoder launch); writes
profiles/<tag>_launches.md, profiles/<tag>_decoder_ncu.md and updates profiles/decoder_traffic.json."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def launches(tag, fname="launches.csv", suffix="launches", cmd="python bench.py --mode render --steps 2 --warmup 1 --no-cpu-baseline"):
    if not os.path.exists(os.path.join(G, fname)):
        return
    rows = list(csv.reader(open(os.path.join(G, fname))))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, mi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, tot, n = collections.OrderedDict(), 0.0, 0
    for r in data:
        if len(r) <= mi:
            continue
        v = float(r[mi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
        a = agg.setdefault(r[ki][:90], [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        n += 1
    out = [f"# {tag}: ncu launch list of `{cmd}`",
           "", "`UORF_PROFILE_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none`: exactly the launches of the "
           "timed steps (cold-cache, serialised: compare SHARES).",
           f"{n} launches captured, {tot:.1f} us in total.", "", "| us | launches | share | kernel |", "|---:|---:|---:|---|"]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        out.append(f"| {t:.1f} | {c} | {100 * t / tot:.1f}% | `{k}` |")
    open(os.path.join(P, f"{tag}_{suffix}.md"), "w").write("\n".join(out) + "\n")


def decoder(tag, precision, repname="prof_decoder.ncu-rep", title=None, outname=None):
    rep = os.path.join(G, repname)
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
    out = [f"# {tag}: `ncu --set full --clock-control none` of one " + (title or f"decoder_fwd_kernel launch ({precision}, bench workload)"), "",
           "| metric | value | unit |", "|---|---:|---|"]
    for k in keys:
        if k in d:
            out.append(f"| {k} | {d[k][1]} | {d[k][0]} |")
    stalls = [(h, float(d[h][1])) for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and d[h][1]]
    tot = sum(v for _, v in stalls) or 1.0
    out += ["", "Warp stall reasons (PC samples, share of all samples):", "", "| reason | samples | share |", "|---|---:|---:|"]
    for h, v in sorted(stalls, key=lambda x: -x[1])[:10]:
        out.append(f"| {h.replace('smsp__pcsamp_warps_issue_stalled_', '')} | {v:.0f} | {100 * v / tot:.1f}% |")
    open(os.path.join(P, outname or f"{tag}_decoder_ncu_{precision}.md"), "w").write("\n".join(out) + "\n")
    if outname:
        return
    try:
        rd, wr = float(d["dram__bytes_read.sum"][1]), float(d["dram__bytes_write.sum"][1])
        mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        traffic = rd * mult[d["dram__bytes_read.sum"][0]] + wr * mult[d["dram__bytes_write.sum"][0]]
        tp = os.path.join(P, "decoder_traffic.json")
        t = json.load(open(tp)) if os.path.exists(tp) else {}
        t[precision] = traffic
        t["_source"] = f"{tag}: dram__bytes_read.sum + dram__bytes_write.sum of one decoder launch (bench workload), bytes"
        json.dump(t, open(tp, "w"), indent=1)
    except Exception as e:
        print("traffic not updated:", e)


def conv(tag, repname="prof_conv.ncu-rep"):
    """one row per captured launch of the tensor-core convolution kernel"""
    rep = os.path.join(G, repname)
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, data = rows[0], rows[2:]
    cols = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct"]
    cols = [c for c in cols if c in hdr]
    out = [f"# {tag}: `ncu --set full --clock-control none` of conv3x3_x3_kernel launches inside one training step (VGG16 perceptual loss, "
           "8 images of 64 x 64; launches 15-17 of the step: block-3 / block-4 layers)", "",
           "| " + " | ".join(cols) + " |", "|" + "---|" * len(cols)]
    for r in data:
        out.append("| " + " | ".join(r[hdr.index(c)][:60] for c in cols) + " |")
    open(os.path.join(P, f"{tag}_conv_ncu.md"), "w").write("\n".join(out) + "\n")


def encoder(tag, repname="prof_encoder.ncu-rep"):
    """one row per launch of the six (seven with `bottom`) encoder convolution kernels of one Encoder.forward"""
    rep = os.path.join(G, repname)
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, data = rows[0], rows[2:]
    cols = ["Kernel Name", "Grid Size", "launch__cluster_dim_z", "gpu__time_duration.sum", "launch__registers_per_thread",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum"]
    cols = [c for c in cols if c in hdr]
    out = [f"# {tag}: `ncu --set full --clock-control none` of the encoder convolution launches of one Encoder.forward (64x64 input, z_dim 64; "
           "scripts/enc_once.py)", "", "Layers in launch order: down1 (7->64 @64^2), down2 (s2 -> 32^2), down3 (s2 -> 16^2), up3 (16^2), "
           "up2 (upsample + cat, 128->64 @32^2), up1 (upsample + cat, 128->64 @64^2).", "",
           "| " + " | ".join(cols) + " |", "|" + "---|" * len(cols)]
    for r in data:
        out.append("| " + " | ".join(r[hdr.index(c)][:60] for c in cols) + " |")
    open(os.path.join(P, f"{tag}_encoder_ncu.md"), "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    tag = sys.argv[1]
    prec = sys.argv[2] if len(sys.argv) > 2 else "fp16x3"
    os.makedirs(P, exist_ok=True)
    launches(tag)
    launches(tag, "launches_train.csv", "launches_train", "python bench.py --mode train --steps 2 --warmup 1 --no-graph --no-cpu-baseline")
    decoder(tag, prec)
    decoder(tag, "bf16", "prof_decoder_bf16.ncu-rep")
    decoder(tag, "", "prof_composite.ncu-rep", "composite_fwd_kernel launch (scripts/bench_kernels.py 32: 524,288 rays x 64 samples)",
            f"{tag}_composite_ncu.md")
    decoder(tag, "", "prof_sample_coords.ncu-rep", "sample_coords launch (scripts/bench_kernels.py 32: 32 views, 128x128x64 frustum, ray-major)",
            f"{tag}_sample_coords_ncu.md")
    decoder(tag, "", "prof_decoder_bwd.ncu-rep", "decoder_bwd_kernel launch (bench.py --mode train: K=8, 64^3 frustum, 4 views)",
            f"{tag}_decoder_bwd_ncu.md")
    conv(tag)
    encoder(tag)
    print("wrote profiles/%s_*.md" % tag)
