"""Builds libuorf_b200.so (sm_100a only) in-tree with plain nvcc: no torch headers, no JIT cache.

    python -m uorf_b200.build [--force] [--verbose]

The shared object lands in uorf_b200/_build/ (git-ignored, travels to the GPU box with the snapshot).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libuorf_b200.so")
SOURCES = ["abi.cu", "encoder.cu", "sample_coords.cu", "decoder_prep.cu", "decoder_fwd.cu", "decoder_fwd_x3.cu", "decoder_bwd.cu", "composite.cu", "conv3x3.cu", "adam.cu", "losses.cu",
           "slot_attention.cu", "slot_attention_bwd.cu"]
# test / diagnostic hooks (tcgen05 layout probes, hand-off timing): their own library, never linked into the product .so
PROBE_LIB = os.path.join(OUT, "libuorf_b200_probe.so")
PROBE_SOURCES = ["probe.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libuorf_b200 cannot be built (there is no non-CUDA fallback)")


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _fingerprint() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join("..", "..", "include", "uorf_b200.h"),
                                        os.path.join("..", "..", "include", "uorf_b200_probe.h")]
    for f in files:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            with open(p, "rb") as fh:
                h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False, defines=(), tag: str = "") -> str:
    """`defines`/`tag`: tuning variants (e.g. defines=["-DUORF_WAIT_HINT_NS=1000"], tag="hint1000") are built next to the
    product library as _build/variants/libuorf_b200_<tag>.so; UORF_B200_LIB=<path> makes the loader pick one up."""
    out, lib, probe_lib = OUT, LIB, PROBE_LIB
    if tag:
        out = os.path.join(OUT, "variants", tag)
        lib = os.path.join(OUT, "variants", f"libuorf_b200_{tag}.so")
        probe_lib = None
    os.makedirs(out, exist_ok=True)
    stamp = os.path.join(out, "fingerprint.txt")
    fp = _fingerprint() + " " + " ".join(defines)
    if (not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read().strip() == fp
            and (probe_lib is None or os.path.exists(probe_lib))):
        return lib
    nvcc = _nvcc()
    flags = list(NVCC_FLAGS) + list(defines) + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):
        obj = os.path.join(out, src.replace(".cu", ".o"))
        cmd = [nvcc] + flags + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    probe_srcs = PROBE_SOURCES if probe_lib else []
    with ThreadPoolExecutor(max_workers=8) as ex:
        all_objs = list(ex.map(compile_one, _sources() + probe_srcs))
    objs, probe_objs = all_objs[:len(all_objs) - len(probe_srcs)], all_objs[len(all_objs) - len(probe_srcs):]
    for target, o in ((lib, objs), (probe_lib, probe_objs)):
        if not target:
            continue
        cmd = [nvcc, "-shared", "-o", target] + o + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(fp)
    return lib


if __name__ == "__main__":
    _defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    _tag = next((a[len("--tag="):] for a in sys.argv[1:] if a.startswith("--tag=")), "")
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=_defs, tag=_tag))
