"""TEST INFRASTRUCTURE - installs and drives the UNMODIFIED reference (KovenYu/uORF) as `oracle/_ref`.

The reference is pure Python; nothing has to be compiled.  `install()` copies the reference's
`models/`, `util/`, `options/`, `data/` packages from /root/reference into `oracle/_ref/` (git-ignored, NOT gpurun-ignored:
it travels to the GPU box like the built .so) and writes a sha256 manifest, so a run on the GPU box can show that the
files it executed are byte-identical to the reference's.  No reference source is committed to this repository.

Only `tests/`, `__graft_entry__.py` and `bench.py`'s baseline legs (`cpu_baseline`, `--impl reference`) may import this
module; the product package `uorf_b200/` never does (tests/test_abi_cpu.py::test_product_never_imports_oracle).

    python -m oracle.ref_harness install       # scripts/install_ref.sh does exactly this

What needs stubbing to import the reference's drivers offline: the VGG16 weight download
(models/model.py:320), `lpips`, `piq` (uorf_eval_model.py:13-15) and `util.util` (imports matplotlib / skimage,
util/util.py:3-19; only AverageMeter and set_seed are used by the drivers).  The stubs live HERE, not in the copy.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("UORF_REFERENCE", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
PACKAGES = ("models", "util", "options", "data")
MANIFEST = "MANIFEST.json"


def _sha(path: str) -> str:
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def _walk(root: str):
    for pkg in PACKAGES:
        for d, _, files in os.walk(os.path.join(root, pkg)):
            if "__pycache__" in d:
                continue
            for f in sorted(files):
                if f.endswith(".pyc"):
                    continue
                yield os.path.relpath(os.path.join(d, f), root)


def install(force: bool = False):
    """Copy the reference packages into oracle/_ref (skipped when the source tree is absent, e.g. on the GPU box, where the
    copy made in the build container has travelled with the snapshot).  Returns the destination or None."""
    if not os.path.isdir(REF_SRC):
        return REF_DST if available() else None
    man_path = os.path.join(REF_DST, MANIFEST)
    want = {rel: _sha(os.path.join(REF_SRC, rel)) for rel in _walk(REF_SRC)}
    if not force and os.path.exists(man_path):
        try:
            if json.load(open(man_path))["files"] == want and verify():
                return REF_DST
        except Exception:
            pass
    if os.path.isdir(REF_DST):
        shutil.rmtree(REF_DST)
    os.makedirs(REF_DST)
    for rel in want:
        dst = os.path.join(REF_DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF_SRC, rel), dst)
    with open(man_path, "w") as fh:
        json.dump({"source": "KovenYu/uORF (unmodified copy of " + REF_SRC + ")", "files": want}, fh, indent=1, sort_keys=True)
    return REF_DST


def available() -> bool:
    return os.path.exists(os.path.join(REF_DST, MANIFEST)) and os.path.exists(os.path.join(REF_DST, "models", "model.py"))


def verify() -> bool:
    """every file of oracle/_ref still has the sha256 recorded at install time (= the reference's bytes)"""
    man = json.load(open(os.path.join(REF_DST, MANIFEST)))["files"]
    return all(os.path.exists(os.path.join(REF_DST, rel)) and _sha(os.path.join(REF_DST, rel)) == h for rel, h in man.items())


_active = False


def activate():
    """Put oracle/_ref first on sys.path and install the four stubs (idempotent).  Raises when oracle/_ref is missing."""
    global _active
    if _active:
        return
    if not available():
        raise RuntimeError("oracle/_ref is not installed: run scripts/install_ref.sh in the build container")
    import torch
    import torchvision
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    lp = types.ModuleType("lpips")

    class LPIPS(torch.nn.Module):                       # stand-in metric: the path under test never reads it
        def forward(self, a, b):
            return ((a - b) ** 2).mean(dim=(1, 2, 3))
    lp.LPIPS = LPIPS
    sys.modules["lpips"] = lp
    pq = types.ModuleType("piq")
    pq.ssim = lambda a, b, data_range=1.: torch.zeros((), device=a.device)
    pq.psnr = lambda a, b, data_range=1.: -10.0 * torch.log10(((a.clamp(0, 1) - b.clamp(0, 1)) ** 2).mean())
    sys.modules["piq"] = pq
    uu = types.ModuleType("util.util")

    class AverageMeter:                                 # util/util.py:AverageMeter (the real module needs matplotlib / skimage)
        def __init__(self):
            self.val, self.n = 0, 0

        def update(self, v, n=1):
            self.val = (self.val * self.n + v * n) / (self.n + n)
            self.n += n
    uu.AverageMeter = AverageMeter
    up = types.ModuleType("util")
    up.util, up.__path__ = uu, []
    sys.modules["util"], sys.modules["util.util"] = up, uu
    import models.model as mm                            # noqa: E402  (the reference's module, from oracle/_ref)
    mm.vgg16 = lambda pretrained=True: torchvision.models.vgg16(weights=None)   # no network: random-init VGG (seeded by the caller)
    _active = True


def make_opt(model_name: str, is_train: bool, **over):
    """The argparse Namespace the reference's option classes would hand to create_model (options/base_options.py),
    built from the model's own modify_commandline_options plus the base options the model code reads."""
    activate()
    from models import get_option_setter
    p = argparse.ArgumentParser()
    p.add_argument("--batch_size", type=int, default=1)
    p.add_argument("--lr", type=float, default=3e-4)
    p.add_argument("--niter", type=int, default=100)
    p.add_argument("--niter_decay", type=int, default=0)
    p.add_argument("--dataset_mode", default="multiscenes")
    p.add_argument("--custom_lr", action="store_true")
    p.add_argument("--lr_policy", default="warmup")
    p.add_argument("--exp_id", default="ref")
    p = get_option_setter(model_name)(p, is_train)
    opt = p.parse_args([])
    opt.gpu_ids, opt.isTrain, opt.model = [], is_train, model_name
    opt.checkpoints_dir, opt.name = "/tmp/uorf_ref_ckpt", "ref"
    opt.verbose, opt.continue_train, opt.load_iter, opt.epoch, opt.epoch_count = False, False, 0, "latest", 1
    opt.n_img_each_scene = 4
    for k, v in over.items():
        setattr(opt, k, v)
    return opt


def create_model(model_name: str, is_train: bool, quiet: bool = True, **over):
    """models.create_model(opt) of the reference (models/__init__.py:54-67): the name-based plug-in factory."""
    import contextlib
    import io
    opt = make_opt(model_name, is_train, **over)
    from models import create_model as ref_create
    with contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext():
        model = ref_create(opt)
        if is_train:                                   # BaseModel.setup without the checkpoint load: schedulers only
            from models import networks
            model.schedulers = [networks.get_scheduler(o, opt) for o in model.optimizers]
    return model, opt


class swapped_imports:
    """INTEGRATION.md Option A, executed: inside the reference's driver modules the names bound by
        from .model import Encoder, Decoder, SlotAttention, raw2outputs        (uorf_eval_model.py:10, uorf_nogan_model.py:12)
        from .projection import Projection                                     (:9 / :10)
    are re-bound to uorf_b200.modules' classes - the same effect as editing those two import lines.  Context manager:
    the reference's own bindings are restored on exit, so the same process can build both variants."""
    NAMES = ("Encoder", "Decoder", "SlotAttention", "raw2outputs", "Projection")

    def __init__(self, *driver_modules):
        self.driver_modules = driver_modules or ("models.uorf_eval_model", "models.uorf_nogan_model")
        self.saved = []

    def __enter__(self):
        activate()
        import importlib
        from uorf_b200 import modules as B
        for name in self.driver_modules:
            mod = importlib.import_module(name)
            for n in self.NAMES:
                if hasattr(mod, n):
                    self.saved.append((mod, n, getattr(mod, n)))
                    setattr(mod, n, getattr(B, n))
        return self

    def __exit__(self, *exc):
        for mod, n, v in self.saved:
            setattr(mod, n, v)
        self.saved = []
        return False


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "install":
        print(install(force="--force" in sys.argv))
    elif len(sys.argv) > 1 and sys.argv[1] == "verify":
        print("oracle/_ref", "matches the manifest" if available() and verify() else "MISSING or modified")
    else:
        print(__doc__)
