"""Host-side data layer (no GPU): image-folder listing, cache layout, PNG naming, launcher config
(reference: data/datasets/base_dataset.py, cache_manager.py, dataloader_utils.py, engine/utils/save_image.py,
scripts/eval.py)."""
import json
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch
from PIL import Image

from ucod_dpl_b200.data.datasets import MultiCacheManager, USCODDataset, collate_fn, list_dir_image
from ucod_dpl_b200.engine.utils.save_image import save_tensor_binary_mask_as_image


def _make_set(root, name, stems, with_gt=True, ext="png"):
    for sub in ("im", "gt") if with_gt else ("im",):
        os.makedirs(root / name / sub, exist_ok=True)
    for i, s in enumerate(stems):
        Image.fromarray(np.full((8 + i, 9 + i, 3), i, np.uint8)).save(root / name / "im" / f"{s}.{ext}")
        if with_gt:
            Image.fromarray(np.full((8 + i, 9 + i), 255, np.uint8), mode="L").save(root / name / "gt" / f"{s}.png")


def test_listing_is_sorted_filtered_and_concatenates_datasets(tmp_path):
    _make_set(tmp_path, "B", ["z", "a"])
    _make_set(tmp_path, "A", ["m"])
    (tmp_path / "B" / "im" / "notes.txt").write_text("x")
    (tmp_path / "B" / "im" / "upper.PNG").write_bytes(b"")      # suffix match is case sensitive, like the reference
    assert [p.name for p in list_dir_image(tmp_path / "B" / "im")] == ["a.png", "z.png"]
    cfg = SimpleNamespace(DATASET="B+A", image_size=(32, 32), require_label=True)
    ds = USCODDataset(cfg, SimpleNamespace(type="dinov2"), "test", str(tmp_path), None)
    assert [os.path.basename(os.path.dirname(os.path.dirname(str(p)))) + "/" + p.name for p in ds.image_paths] == \
        ["A/m.png", "B/a.png", "B/z.png"]
    assert len(ds) == 3 and len(ds.label_paths) == 3
    assert ds.load_all and ds.transform_label.size is None      # test mode keeps the label size


def test_label_mapping_is_checked(tmp_path):
    _make_set(tmp_path, "A", ["x", "y"])
    os.remove(tmp_path / "A" / "gt" / "y.png")
    cfg = SimpleNamespace(DATASET="A", image_size=(32, 32), require_label=True)
    with pytest.raises(AssertionError):
        USCODDataset(cfg, SimpleNamespace(type="dinov2"), "test", str(tmp_path), None)


def test_cache_layout_and_item_dict(tmp_path):
    _make_set(tmp_path / "data", "TR", ["a", "b"], with_gt=False)
    mgr = MultiCacheManager(str(tmp_path / "cache"), "dinov2", "train", "TR")
    assert mgr.cache_path("features").endswith(os.path.join("features_cache", "dinov2", "train", "TR"))
    assert mgr.cache_path("patch").endswith(os.path.join("patch_cache", "dinov2", "train", "TR"))
    assert mgr.cache_path("pseudo_label").endswith(os.path.join("pseudo_label_cache", "TR"))
    assert MultiCacheManager(str(tmp_path), "dinov2", "test", "TR").get_pseudo_label_cache() is None
    feats = [torch.full((768, 2, 2), float(i)) for i in range(2)]
    mgr.get_features_cache().dump_list(feats)
    mgr.get_pseudo_label_cache().dump_list([torch.ones(1, 4, 4), torch.zeros(1, 4, 4)])
    with open(os.path.join(mgr.cache_path("features"), "index.json")) as f:
        assert json.load(f) == {"0": "data_0.pkl", "1": "data_1.pkl"}
    cfg = SimpleNamespace(DATASET="TR", image_size=(32, 32), require_label=False)
    ds = USCODDataset(cfg, SimpleNamespace(type="dinov2"), "train", str(tmp_path / "data"), str(tmp_path / "cache"))
    item = ds[1]
    assert list(item.keys()) == ["pseudo_label", "label_tensor", "features", "img_path"]
    assert item["label_tensor"] is None and torch.equal(item["features"], feats[1])
    assert item["pseudo_label"].sum() == 0 and item["img_path"].endswith("b.png")
    batch = collate_fn([ds[0], ds[1]])
    assert batch["features"].shape == (2, 768, 2, 2) and batch["label_tensor"] == [None, None]
    assert isinstance(batch["img_path"], list)


def test_png_naming_rules(tmp_path):
    m = torch.zeros(1, 6, 5)
    m[0, 2:4, 1:3] = 1
    save_tensor_binary_mask_as_image(m > 0.5, str(tmp_path / "preds" / "SET" / "camo_1.jpg"))
    arr = np.asarray(Image.open(tmp_path / "preds" / "SET" / "camo_1.png"))
    assert arr.dtype == np.uint8 and arr.shape == (6, 5) and arr.max() == 255 and int((arr > 0).sum()) == 4
    save_tensor_binary_mask_as_image(torch.ones(3, 1, 4, 4), str(tmp_path / "batch.png"))
    assert sorted(os.listdir(tmp_path / "batch")) == ["0.png", "1.png", "2.png"]


def test_launcher_config_and_work_dir(tmp_path, monkeypatch):
    from ucod_dpl_b200.scripts import eval as ev
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    monkeypatch.chdir(root)
    args = ev.parse_train_args(["--config", "configs/uscod/CORAL_dinov2.py", "--work_dir", str(tmp_path / "w"),
                                "--load_from", "a.safetensors", "--refiner_path", "r.safetensors"])
    cfg = ev.init_cfg(args)
    assert cfg.work_dir == os.path.join(str(tmp_path / "w"), "uscod", "CORAL_dinov2") and os.path.isdir(cfg.work_dir)
    assert cfg.mode == "eval" and cfg.dataset_cfg.valset_cfg.keep_size is True
    assert cfg.train_cfg.checkpoint == "a.safetensors" and cfg.train_cfg.refiner_path == "r.safetensors"
    assert ev.DATASET == ["CHAMELEON", "TE-CAMO", "TE-COD10K", "NC4K"]


def test_async_mask_writer(tmp_path):
    from ucod_dpl_b200.engine.utils.save_image import AsyncMaskWriter
    w = AsyncMaskWriter(workers=2)
    for i in range(5):
        m = torch.zeros(7, 9, dtype=torch.uint8)
        m[i, :] = 1
        w.submit(m, str(tmp_path / "p" / f"im{i}.jpg"))
    w.close()
    for i in range(5):
        arr = np.asarray(Image.open(tmp_path / "p" / f"im{i}.png"))
        assert arr.shape == (7, 9) and (arr[i] == 255).all() and int((arr > 0).sum()) == 9


def test_dataloader_factory_on_cached_features(tmp_path, monkeypatch):
    """`DataLoaderFactory.create_train_loader` over existing caches (no GPU work needed: nothing to extract)."""
    from ucod_dpl_b200.data.datasets import DataLoaderFactory, init_trainloader
    from ucod_dpl_b200.engine.config import CfgNode
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    monkeypatch.chdir(root)
    cfg = CfgNode(CfgNode.load_with_base("configs/uscod/UCOD-DPL_dinov2.py")).dataset_cfg
    _make_set(tmp_path / "data", "TR-F", [f"i{k}" for k in range(5)], with_gt=False)
    cfg.trainset_cfg.DATASET = "TR-F"
    cfg.dataset_dir, cfg.cache_dir = str(tmp_path / "data"), str(tmp_path / "cache")
    cfg.trainloader_cfg.batch_size, cfg.trainloader_cfg.shuffle = 2, False
    mgr = MultiCacheManager(cfg.cache_dir, "dinov2", "train", "TR-F")
    mgr.get_features_cache().dump_list([torch.full((768, 2, 2), float(k)) for k in range(5)])
    mgr.get_pseudo_label_cache().dump_list([torch.full((1, 16, 16), float(k % 2)) for k in range(5)])
    loader = DataLoaderFactory.create_train_loader(cfg)
    batches = list(loader)
    assert len(loader) == 3 and [b["features"].shape[0] for b in batches] == [2, 2, 1]
    assert batches[1]["features"][0, 0, 0, 0].item() == 2.0 and batches[1]["pseudo_label"].shape == (2, 1, 16, 16)
    assert batches[0]["label_tensor"] == [None, None] and len(batches[2]["img_path"]) == 1
    assert len(init_trainloader(cfg)) == 3


def test_checkpoint_discovery(tmp_path):
    """`find_latest_checkpoint` / `resolve_checkpoint`: newest `*.pth|*.pt|*.safetensors` entry of any run directory's
    `ckp` (or `refiner_ckp`) folder; a directory entry (what `accelerator.save_model` writes) resolves to its
    `model.safetensors` (reference: runner.py:165-241)."""
    import time
    from types import SimpleNamespace as NS

    from ucod_dpl_b200.scripts.eval import find_latest_checkpoint, resolve_checkpoint
    work = tmp_path / "work" / "uscod" / "cfg"
    cfg = NS(work_dir=str(work), log_cfg=NS(log_path=str(work / "eval_run")))
    os.makedirs(work / "eval_run")
    assert find_latest_checkpoint(cfg, "ckp") is None
    for k, name in enumerate(["epoch5.pth", "epoch10.pth"]):
        d = work / "train_run" / "ckp" / name
        os.makedirs(d)
        (d / "model.safetensors").write_bytes(b"x")
        t = time.time() - 100 + 10 * k
        os.utime(d, (t, t))
    (work / "train_run" / "ckp" / "notes.txt").write_text("ignored")
    os.makedirs(work / "train_run" / "refiner_ckp")
    (work / "train_run" / "refiner_ckp" / "epoch8.pth").write_bytes(b"y")
    found = find_latest_checkpoint(cfg, "ckp")
    assert found.endswith(os.path.join("ckp", "epoch10.pth"))
    assert resolve_checkpoint(found).endswith(os.path.join("epoch10.pth", "model.safetensors"))
    r = find_latest_checkpoint(cfg, "refiner_ckp")
    assert r.endswith("epoch8.pth") and resolve_checkpoint(r) == r
