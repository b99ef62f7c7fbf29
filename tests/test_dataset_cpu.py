"""On-disk scene format (reference data/multiscenes_dataset.py) -> driver input dict, on a generated miniature dataset."""
import os

import numpy as np
import pytest
import torch
from PIL import Image

from uorf_b200.dataset import MultiscenesDataset, collate_fn, make_loader


def _write_scene(root, idx, scene, with_mask):
    rng = np.random.default_rng(idx)
    stem = os.path.join(root, f"{idx:05d}_sc{scene:04d}_az{idx % 4:02d}")
    Image.fromarray(rng.integers(0, 256, (20, 20, 3), dtype=np.uint8)).save(stem + ".png")
    pose = np.eye(4); pose[:3, 3] = rng.normal(size=3); np.savetxt(stem + "_RT.txt", pose)
    np.savetxt(stem + "_azi_rot.txt", np.eye(3))
    if with_mask:
        m = np.full((20, 20), 64, np.uint8); m[:5] = 0; m[12:, 12:] = 200           # levels 0, 64 (= background), 200
        Image.fromarray(np.stack([m] * 3, -1)).save(stem + "_mask.png")
        Image.fromarray(np.zeros((20, 20, 3), np.uint8)).save(stem + "_moved.png")  # auxiliary image: not a view


@pytest.mark.parametrize("with_mask", [False, True])
def test_dataset_to_driver_dict(tmp_path, with_mask):
    root = str(tmp_path)
    for i in range(6):
        _write_scene(root, i, i // 3, with_mask)
    ds = MultiscenesDataset(root, n_scenes=2, n_img_each_scene=2, load_size=16, mask_size=8, is_train=False)
    assert len(ds) == 2 and [len(s) for s in ds.scenes] == [3, 3]            # *_mask / *_moved are not views
    d = collate_fn([ds[1]])
    assert d["img_data"].shape == (2, 3, 16, 16) and d["img_data"].min() >= -1 and d["img_data"].max() <= 1
    assert d["cam2world"].shape == (2, 4, 4) and d["azi_rot"].shape == (2, 3, 3) and d["depths"] is None
    assert all("sc0001" in p for p in d["paths"])
    want = np.loadtxt(d["paths"][0][:-4] + "_RT.txt")
    assert np.allclose(d["cam2world"][0].numpy(), want)
    if with_mask:
        assert d["masks"].shape == (2, 3, 8, 8) and d["mask_idx"].shape == (2, 64) and d["fg_idx"].shape == (2, 64)
        assert d["obj_idxs"].shape == (3, 1, 8, 8)
        assert set(d["mask_idx"].unique().tolist()) == {0, 1, 2}
        assert torch.equal(d["fg_idx"], d["mask_idx"] != 1)                  # the second grey level is the background
    # training mode samples views at random; the loader applies the same collate
    ds_train = MultiscenesDataset(root, n_scenes=2, n_img_each_scene=2, load_size=16, mask_size=8, is_train=True)
    batch = next(iter(make_loader(ds_train, batch_size=1, num_workers=0)))
    assert batch["img_data"].shape == (2, 3, 16, 16)
    with pytest.raises(FileNotFoundError):
        os.remove(ds.scenes[0][0][:-4] + "_RT.txt")
        ds[0]
