"""On-disk scene format of the reference (data/multiscenes_dataset.py:12-157, data/README.md:10-43) -> driver input dicts.

A dataset directory holds, per view,
    <a>sc<scene:04d><b>.png              RGB image
    <a>sc<scene:04d><b>_RT.txt           4x4 camera-to-world matrix (text)
    <a>sc<scene:04d><b>_azi_rot.txt      3x3 azimuth rotation (not needed with fixed_locality)
    <a>sc<scene:04d><b>_mask.png         optional instance mask (grey levels = objects; the SECOND level is the background)
    <a>sc<scene:04d><b>_depth.npy        optional HxWx1 depth
plus the auxiliary images of the manipulation sets (*_moved.png, *_changed.png, *_providing_bg.png and their masks), which are
not views of their own.  One item = one scene = `n_img_each_scene` views; `collate_fn` turns a list of items into the dict
`set_input()` of the drivers takes (img_data Nx3xLxL in [-1,1], cam2world Nx4x4, azi_rot Nx3x3, paths, depths,
masks / mask_idx / fg_idx / obj_idxs when masks exist).  Pure host code: PIL + numpy + torch, no GPU work.
"""
from __future__ import annotations

import glob
import os
import random
from typing import Dict, List, Optional

import numpy as np
import torch
from PIL import Image

_AUX_SUFFIXES = ("_mask.png", "_mask_for_moving.png", "_moved.png", "_mask_for_bg.png", "_mask_for_providing_bg.png",
                 "_changed.png", "_providing_bg.png")


def _to_unit_range(img: Image.Image, size: int, nearest: bool = False) -> torch.Tensor:
    """resize to size x size, HWC uint8 -> CHW float in [-1, 1] (reference _transform / _transform_mask, :50-61)"""
    img = img.resize((size, size), Image.NEAREST if nearest else Image.BILINEAR)
    a = np.asarray(img, dtype=np.float32) / 255.0
    if a.ndim == 2:
        a = a[:, :, None]
    return torch.from_numpy(a).permute(2, 0, 1).contiguous() * 2.0 - 1.0


class MultiscenesDataset(torch.utils.data.Dataset):
    def __init__(self, dataroot: str, n_scenes: int = 1000, n_img_each_scene: int = 10, load_size: int = 128,
                 mask_size: int = 128, is_train: bool = True, no_shuffle: bool = False, fixed_locality: bool = False):
        self.dataroot, self.n_scenes, self.n_img_each_scene = dataroot, int(n_scenes), int(n_img_each_scene)
        self.load_size, self.mask_size = int(load_size), int(mask_size)
        self.is_train, self.no_shuffle, self.fixed_locality = bool(is_train), bool(no_shuffle), bool(fixed_locality)
        views = [f for f in sorted(glob.glob(os.path.join(dataroot, "*.png"))) if not f.endswith(_AUX_SUFFIXES)]
        # a file belongs to scene i when its name contains 'sc{i:04d}' (data/README.md:14-22)
        self.scenes: List[List[str]] = [[f for f in views if f"sc{i:04d}" in os.path.basename(f)] for i in range(self.n_scenes)]

    def __len__(self) -> int:
        return self.n_scenes

    def _mask_fields(self, mask_path: str) -> Dict[str, torch.Tensor]:
        rgb = Image.open(mask_path).convert("RGB")
        grey = _to_unit_range(rgb.convert("L"), self.mask_size, nearest=True)         # 1xHxW
        flat = grey.flatten()
        levels = flat.unique(sorted=True)
        mask_idx = (flat[:, None] == levels).to(torch.uint8).argmax(dim=1)             # index of each pixel's grey level
        bg_level = levels[1] if levels.numel() > 1 else levels[0]                      # reference: second level is the background
        return {"mask": _to_unit_range(rgb, self.mask_size, nearest=True), "mask_idx": mask_idx, "fg_idx": flat != bg_level,
                "obj_idxs": torch.stack([grey == lv for lv in levels])}                # Kx1xHxW

    def __getitem__(self, index: int) -> List[Dict]:
        files = self.scenes[index]
        if self.is_train and not self.no_shuffle:
            files = random.sample(files, self.n_img_each_scene)
        else:
            files = files[:self.n_img_each_scene]
        views = []
        for path in files:
            stem = path[:-len(".png")]
            if not os.path.isfile(stem + "_RT.txt"):
                raise FileNotFoundError(f"camera-to-world matrix missing for {path}: {stem}_RT.txt")
            item = {"img_data": _to_unit_range(Image.open(path).convert("RGB"), self.load_size), "path": path,
                    "cam2world": torch.tensor(np.loadtxt(stem + "_RT.txt"), dtype=torch.float32),
                    "azi_rot": torch.eye(3) if self.fixed_locality
                    else torch.tensor(np.loadtxt(stem + "_azi_rot.txt"), dtype=torch.float32)}
            if os.path.isfile(stem + "_depth.npy"):
                item["depth"] = torch.from_numpy(np.load(stem + "_depth.npy")).permute(2, 0, 1)
            if os.path.isfile(stem + "_mask.png"):
                item.update(self._mask_fields(stem + "_mask.png"))
            views.append(item)
        return views


def collate_fn(batch: List[List[Dict]]) -> Dict:
    """list (batch) of lists (views) of dicts -> one driver input dict (reference :130-157); obj_idxs are view 0's."""
    flat = [v for scene in batch for v in scene]
    out = {"img_data": torch.stack([v["img_data"] for v in flat]), "paths": [v["path"] for v in flat],
           "cam2world": torch.stack([v["cam2world"] for v in flat]), "azi_rot": torch.stack([v["azi_rot"] for v in flat]),
           "depths": torch.stack([v["depth"] for v in flat]) if "depth" in flat[0] else None}
    if "mask" in flat[0]:
        out["masks"] = torch.stack([v["mask"] for v in flat])
        out["mask_idx"] = torch.stack([v["mask_idx"] for v in flat])
        out["fg_idx"] = torch.stack([v["fg_idx"] for v in flat])
        out["obj_idxs"] = flat[0]["obj_idxs"]
    return out


def make_loader(dataset: MultiscenesDataset, batch_size: int = 1, shuffle: Optional[bool] = None, num_workers: int = 0,
                sampler=None) -> torch.utils.data.DataLoader:
    """DataLoader with the reference's collate (data/__init__.py:60-75); pass a DistributedSampler to shard scenes over ranks."""
    if shuffle is None:
        shuffle = dataset.is_train and sampler is None
    return torch.utils.data.DataLoader(dataset, batch_size=batch_size, shuffle=shuffle and sampler is None, sampler=sampler,
                                       num_workers=num_workers, collate_fn=collate_fn, pin_memory=True)
