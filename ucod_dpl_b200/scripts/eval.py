"""Test-set sweep of both stages (reference: scripts/eval.py:8-37 + `Runner.launch_val_look_twice`,
runner.py:390-398; scripts/LTeval.py:8-35 + `Runner_local_refine.launch_val`, runner.py:583-590; the loops
loop_UCOD_DPL.py:297-324 and loop_CORAL.py:247-341).

    python -m ucod_dpl_b200.scripts.eval   --config configs/uscod/UCOD-DPL_dinov2.py --load_from weights/UCOD_DPL_dinov2.safetensors
    python -m ucod_dpl_b200.scripts.LTeval --config configs/uscod/CORAL_dinov2.py --load_from ... --refiner_path ...
    python -m torch.distributed.run --nproc-per-node 8 -m ucod_dpl_b200.scripts.eval ...      (images sharded by rank)

Same work_dir layout (`{work_dir}/{config dir relative to ./configs}/{config name}/{exp}/preds/{DATASET}/x.png`,
`config.yaml`, `eval{rank}.log`) and the same result table per test set.  The reference streams one image at a time
through cached features; here a rank decodes a batch on host threads, and resize, both looks, the final resize,
binarisation and the fp64 metric suite all run on the GPU; only the PNG encode goes back to the host.
"""
from __future__ import annotations

import logging
import os
from datetime import datetime
from types import SimpleNamespace

import numpy as np
import torch

from .. import dist as udist
from .. import ops
from ..data.datasets import USCODDataset, pack_padded
from ..data.utils.feature_extractor import backbone
from ..engine.config import CfgNode
from ..engine.runner.loop_CORAL import CoralEvaluator
from ..engine.runner.loop_UCOD_DPL import LookTwiceEvaluator
from ..engine.utils.metrics.metric import statistics
from ..engine.utils.save_image import AsyncMaskWriter
from ..models.UDLR import SparseRefiner
from ..models.uscod import baseline
from .args import parse_train_args

DATASET = ["CHAMELEON", "TE-CAMO", "TE-COD10K", "NC4K"]


def init_cfg(args) -> CfgNode:
    cfg = CfgNode(CfgNode.load_with_base(args.config))
    cfg.dataset_cfg.valset_cfg.keep_size = True
    cfg.train_cfg.checkpoint = args.load_from
    cfg.train_cfg.refiner_path = args.refiner_path
    cfg.mode = "eval"
    cfg.work_dir = os.path.join(args.work_dir, os.path.relpath(os.path.dirname(args.config), "./configs"),
                                os.path.splitext(os.path.basename(args.config))[0])
    os.makedirs(cfg.work_dir, exist_ok=True)
    cfg.launcher = args.launcher
    if args.dataset_dir:
        cfg.dataset_cfg.dataset_dir = args.dataset_dir
    if args.exp_name:
        cfg.exp_name = args.exp_name
    return cfg


def _setup_run_dir(cfg, rank: int) -> logging.Logger:
    """runner.py:125-163: `{work_dir}/{exp_name | exp_<timestamp>}`, the config dump and a per-rank log file."""
    exp = cfg.get("exp_name", None)
    log_path = os.path.join(cfg.work_dir, str(exp) if exp is not None else datetime.now().strftime("exp_%Y%m%d_%H%M%S"))
    if udist.world()[1] > 1:  # every rank must agree on the timestamped directory
        box = [log_path]
        torch.distributed.broadcast_object_list(box, src=0)
        log_path = box[0]
    os.makedirs(log_path, exist_ok=True)
    cfg.log_cfg.log_path = log_path
    logger = logging.getLogger(f"ucod_dpl_b200.eval{rank}")
    logger.setLevel(logging.INFO)
    logger.handlers.clear()
    logger.addHandler(logging.FileHandler(os.path.join(log_path, f"{cfg.mode}{rank}.log")))
    if rank == 0:
        logger.addHandler(logging.StreamHandler())
        with open(os.path.join(log_path, "config.yaml"), "w") as f:
            f.write(cfg.dump())
    return logger


def resolve_checkpoint(path) -> str:
    """a safetensors file, or the directory `accelerator.save_model` writes (`.../epoch{n}.pth/model.safetensors`)."""
    path = str(path)
    return os.path.join(path, "model.safetensors") if os.path.isdir(path) else path


def find_latest_checkpoint(cfg, ckp_type: str = "ckp"):
    """`BaseRunner._find_latest_checkpoint` (runner.py:209-241): newest `*.pth` / `*.pt` entry of
    `{dirname(log_path)}/{ckp_type}`; additionally every run directory `{work_dir}/*/{ckp_type}` is searched, since
    that is where `save_checkpoint` (runner.py:165-185) actually puts them."""
    roots = [os.path.join(os.path.dirname(cfg.log_cfg.log_path), ckp_type)]
    if os.path.isdir(cfg.work_dir):
        roots += [os.path.join(cfg.work_dir, d, ckp_type) for d in sorted(os.listdir(cfg.work_dir))]
    found = [os.path.join(r, f) for r in roots if os.path.isdir(r) for f in os.listdir(r)
             if f.endswith((".pth", ".pt", ".safetensors"))]
    return max(found, key=os.path.getmtime) if found else None


def load_first_stage(cfg, device) -> baseline:
    from safetensors.torch import load_file

    model = baseline(SimpleNamespace(dim=cfg.model_cfg.dim))
    ckpt = cfg.train_cfg.get("checkpoint", None) or find_latest_checkpoint(cfg, "ckp")
    if ckpt is None:
        raise FileNotFoundError("no --load_from and no checkpoint under the work_dir")
    model.load_state_dict(load_file(resolve_checkpoint(ckpt)), strict=True)
    return model.to(device).eval()


def evaluate_dataset(cfg, name: str, extractor, model, refiner, args, logger) -> dict:
    """one test set on this rank's shard; returns the reduced result dict (identical on every rank)."""
    device = extractor.feature_extractor.device
    vcfg = cfg.dataset_cfg.valset_cfg
    vcfg.DATASET = name
    ds = USCODDataset(vcfg, cfg.dataset_cfg.feature_extractor_cfg, "test", cfg.dataset_cfg.dataset_dir, None,
                      feature_extractor=extractor)
    size = tuple(vcfg.image_size)
    if refiner is None:
        looker = LookTwiceEvaluator(extractor.feature_extractor, model, size, cfg.model_cfg.feature_size,
                                    look_twice_th=float(cfg.val_cfg.look_twice_th),
                                    expand_type=cfg.val_cfg.expand_type, look_twice=bool(cfg.val_cfg.look_twice))
    else:
        looker = CoralEvaluator(extractor.feature_extractor, model, refiner, size,
                                window_size=int(cfg.model_cfg.get("window_size", 3)),
                                window_length=int(cfg.model_cfg.get("window_length", 56)),
                                require_m_patches=bool(cfg.dataset_cfg.valset_cfg.get("require_m_patches", False)))
    stats = statistics(device=device)
    writer = None if args.no_save else AsyncMaskWriter()
    out_dir = os.path.join(cfg.log_cfg.log_path, "preds", name)
    shard = udist.shard_indices(len(ds))
    for batch in ds.iter_image_batches(args.batch_size, indices=shard, with_labels=True):
        label_sizes = [lab.shape for lab in batch["labels"]]
        if refiner is None:
            canvas, sizes = pack_padded(batch["originals"], device)
            final, _ = looker(batch["images"], originals=canvas, layout="HWC", orig_sizes=sizes)
            # loop_UCOD_DPL.py:315-317: `final` is a [0,1] mask (pasted second looks / 255); bilinear to the label
            # size, then `> 0.5` on the interpolated value itself (binarize = 3; mode 1 would threshold sigmoid(.))
            masks = [ops.upsample_bilinear(final[i], label_sizes[i], binarize=3) for i in range(len(label_sizes))]
        else:
            groups: dict = {}
            for i, im in enumerate(batch["originals"]):  # CORAL windows need equal-size originals per launch
                groups.setdefault(im.shape, []).append(i)
            res = {}
            for _, idx in groups.items():
                originals = torch.from_numpy(np.stack([batch["originals"][i] for i in idx])).to(device)
                for i, m in zip(idx, looker(originals, label_sizes=[label_sizes[i] for i in idx], layout="HWC")):
                    res[i] = m
            masks = [res[i] for i in range(len(label_sizes))]
        for i, m in enumerate(masks):
            gt = ds.transform_label(batch["labels"][i])  # ToTensor only (keep_size): [1,h,w] = label / 255
            stats.step(gt, m[None])
            if writer is not None:
                writer.submit(m, os.path.join(out_dir, os.path.basename(batch["img_path"][i])))
    if writer is not None:
        writer.close()
    result = stats.get_result()
    table = {k: [round(float(v), 4)] for k, v in result.items()}
    logger.info("%s (%d images): %s", name, len(ds), table)
    return result


def main(argv=None, second_stage: bool = False) -> dict:
    args = parse_train_args(argv)
    rank, world = udist.world()
    if "RANK" in os.environ and not torch.distributed.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        torch.distributed.init_process_group("nccl")
        rank, world = udist.world()
    device = torch.device("cuda", torch.cuda.current_device())
    cfg = init_cfg(args)
    logger = _setup_run_dir(cfg, rank)
    extractor = backbone(cfg.dataset_cfg.feature_extractor_cfg, device=device)
    model = load_first_stage(cfg, device)
    refiner = None
    if second_stage:
        from safetensors.torch import load_file

        refiner = SparseRefiner.from_config(cfg.model_cfg).to(device).eval()
        rpath = cfg.train_cfg.get("refiner_path", None) or find_latest_checkpoint(cfg, "refiner_ckp")
        if rpath is None:
            raise FileNotFoundError("no --refiner_path and no refiner checkpoint under the work_dir")
        refiner.load_state_dict(load_file(resolve_checkpoint(rpath)), strict=True)
    results = {}
    for name in (args.datasets.split(",") if args.datasets else DATASET):
        if rank == 0:
            print("running {}".format(name))
        results[name] = evaluate_dataset(cfg, name, extractor, model, refiner, args, logger)
    return results


if __name__ == "__main__":
    main()
