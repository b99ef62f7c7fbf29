"""oracle/_ref - the unmodified reference copied by oracle/ref_harness.py - is what it claims to be and can be driven here:
its files match the reference byte for byte, its own factory builds its own drivers, they reproduce the committed golden
vectors on the CPU, and the repo's on-disk readers agree with the reference's dataset class on the same directory."""
import os

import numpy as np
import pytest
import torch
from PIL import Image

from conftest import load_golden
from oracle import ref_harness as R


@pytest.fixture(scope="module")
def ref():
    if R.install() is None:
        pytest.skip("no reference tree and no installed copy (oracle/_ref)")
    R.activate()
    return R


def test_copy_matches_manifest_and_source(ref):
    assert ref.verify()
    if os.path.isdir(ref.REF_SRC):
        for rel in ref._walk(ref.REF_SRC):
            assert open(os.path.join(ref.REF_SRC, rel), "rb").read() == open(os.path.join(ref.REF_DST, rel), "rb").read(), rel


def test_reference_eval_driver_reproduces_golden(ref):
    """models.create_model -> uorfEvalModel.forward (uorf_eval_model.py:116-157) of the copy == tests/golden/eval_driver.npz"""
    g = load_golden("eval_driver")
    torch.manual_seed(17)
    model, opt = ref.create_model("uorf_eval", False, num_slots=4, z_dim=40, n_samp=8, frustum_size=16, render_size=8,
                                  input_size=32, n_img_each_scene=2)
    assert type(model).__module__ == "models.uorf_eval_model" and model.device.type == "cpu"
    model.netEncoder.load_state_dict(g["enc"])
    model.netSlotAttention.load_state_dict(g["sa"])
    model.netDecoder.load_state_dict(g["dec"])
    model.set_input({"img_data": g["img"], "cam2world": g["cam2world"], "azi_rot": g["azi_rot"], "paths": []})
    torch.manual_seed(117)                 # the slot-attention draws of the golden (make_golden.py)
    with torch.no_grad():
        model.forward()
    x = torch.stack([model.x_rec0, model.x_rec1])
    assert (x - g["x_recon"]).abs().max().item() == 0.0
    assert torch.equal(model.masked_raws, g["masked_raws"])


def test_reference_gan_driver_reproduces_golden(ref):
    """models.create_model -> uorfGanModel.forward / backward (uorf_gan_model.py:137-245) of the copy == tests/golden/gan_driver.npz
    (the generator step of make_golden_gan.py, re-run from the recorded initial weights)"""
    g = load_golden("gan_driver")
    torch.manual_seed(1)
    model, opt = ref.create_model("uorf_gan", True, num_slots=3, z_dim=40, n_samp=8, frustum_size=16, frustum_size_fine=16,
                                  render_size=16, supervision_size=16, input_size=32, n_img_each_scene=2, stage="coarse",
                                  bottom=False, warmup_steps=5, ndf=8, percept_in=100, gan_train_epoch=2, gan_in=1)
    assert type(model).__module__ == "models.uorf_gan_model"
    for net, key in ((model.netEncoder, "enc"), (model.netSlotAttention, "sa"), (model.netDecoder, "dec"), (model.netDisc, "disc")):
        net.load_state_dict(g["init"][key])
    model.set_input({"img_data": g["img"], "cam2world": g["cam2world"], "azi_rot": g["azi_rot"], "paths": []})
    torch.manual_seed(501)
    model.forward(int(g["epoch"]))
    for o in model.optimizers:
        o.zero_grad()
    model.backward()
    assert abs(float(model.loss_recon) - float(g["loss_recon"])) < 1e-6 and abs(float(model.loss_gan) - float(g["loss_gan"])) < 1e-6
    w = model.netDecoder.f_after[0].weight
    assert torch.allclose(w.grad, g["grad"]["dec"]["f_after.0.weight"], atol=1e-7, rtol=1e-4)


def test_swapped_imports_rebinds_and_restores(ref):
    import importlib
    import uorf_b200.modules as B
    mod = importlib.import_module("models.uorf_nogan_model")
    before = {n: getattr(mod, n) for n in R.swapped_imports.NAMES}
    with R.swapped_imports("models.uorf_nogan_model"):
        assert all(getattr(mod, n) is getattr(B, n) for n in R.swapped_imports.NAMES)
    assert all(getattr(mod, n) is before[n] for n in R.swapped_imports.NAMES)
    assert before["Decoder"].__module__ == "models.model"


def _write_view(root, idx, scene, with_mask):
    rng = np.random.default_rng(100 + idx)
    stem = os.path.join(root, f"{idx:05d}_sc{scene:04d}_az{idx % 4:02d}")
    Image.fromarray(rng.integers(0, 256, (24, 24, 3), dtype=np.uint8)).save(stem + ".png")
    pose = np.eye(4)
    pose[:3, :3] = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    pose[:3, 3] = rng.normal(size=3)
    np.savetxt(stem + "_RT.txt", pose)
    np.savetxt(stem + "_azi_rot.txt", np.linalg.qr(rng.normal(size=(3, 3)))[0])
    if with_mask:
        m = np.full((24, 24), 64, np.uint8)
        m[:6] = 0
        m[14:, 10:] = 200
        m[3:9, 15:20] = 130
        Image.fromarray(np.stack([m] * 3, -1)).save(stem + "_mask.png")
        Image.fromarray(np.zeros((24, 24, 3), np.uint8)).save(stem + "_moved.png")


@pytest.mark.parametrize("with_mask", [False, True])
def test_dataset_matches_reference_dataset(ref, tmp_path, with_mask):
    """uorf_b200.dataset.MultiscenesDataset + collate_fn against the reference's data/multiscenes_dataset.py (:12-157) reading
    the same directory: same scene grouping, same tensors (images to float rounding, everything else exactly)."""
    import argparse
    from data.multiscenes_dataset import MultiscenesDataset as RefDataset, collate_fn as ref_collate
    from uorf_b200.dataset import MultiscenesDataset, collate_fn
    root = str(tmp_path)
    for i in range(8):
        _write_view(root, i, i // 4, with_mask)
    opt = argparse.Namespace(dataroot=root, n_scenes=2, n_img_each_scene=3, load_size=16, mask_size=12, isTrain=False, no_shuffle=True,
                             fixed_locality=False)
    rds = RefDataset(opt)
    ds = MultiscenesDataset(root, n_scenes=2, n_img_each_scene=3, load_size=16, mask_size=12, is_train=False)
    assert len(rds) == len(ds) and rds.scenes == ds.scenes
    for s in range(2):
        want, got = ref_collate([rds[s]]), collate_fn([ds[s]])
        assert set(want) == set(got) and want["paths"] == got["paths"]
        assert (want["img_data"] - got["img_data"]).abs().max().item() < 1e-6
        for k in ("cam2world", "azi_rot"):
            assert torch.equal(want[k], got[k]), k
        assert want["depths"] is None and got["depths"] is None
        if with_mask:
            assert (want["masks"] - got["masks"]).abs().max().item() < 1e-6
            for k in ("mask_idx", "fg_idx", "obj_idxs"):
                assert want[k].shape == got[k].shape and torch.equal(want[k].to(got[k].dtype), got[k]), k
