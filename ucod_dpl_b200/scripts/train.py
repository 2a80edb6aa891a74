"""First-stage training launcher (reference: scripts/train.py:9-32 + `StandardRunner`, runner.py:253-398 +
`TrainLoop`, loop_UCOD_DPL.py:35-255).

    python -m ucod_dpl_b200.scripts.train --config configs/uscod/UCOD-DPL_dinov2.py [--dataset_dir ...]
    python -m torch.distributed.run --nproc-per-node 8 -m ucod_dpl_b200.scripts.train --config ...

Same caches (features / pseudo labels, `MetaListPickleIO`), schedule, work_dir layout and checkpoint names
(`{run}/ckp/epoch{n}.pth/model.safetensors`, what `accelerator.save_model` writes).  Missing caches are produced
first (batched on the GPU); then the whole training set lives in HBM as bf16 key tokens and every step is the fused
forward / backward / AdamW+EMA kernel sequence of `FirstStageTrainer`.
"""
from __future__ import annotations

import os
import random
from types import SimpleNamespace

import numpy as np
import torch

from .. import dist as udist
from ..data.datasets import USCODDataset
from ..data.utils.feature_extractor import backbone, load_vit_state_dict
from ..engine.config import CfgNode
from ..engine.runner.loop_UCOD_DPL import TrainLoop
from ..generate_pseudo_label import PseudoLabelGenerator, generate_from_folders
from ..models.discriminator import Discriminator
from ..models.uscod import baseline
from ..train import DiscriminatorTrainer, FirstStageTrainer
from . import eval as ev
from .args import parse_train_args


def set_random_seed(seed: int) -> None:
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)


def init_cfg(args) -> CfgNode:
    cfg = CfgNode(CfgNode.load_with_base(args.config))
    cfg.dataset_cfg.valset_cfg.keep_size = False
    cfg.mode = "train"
    cfg.work_dir = os.path.join(args.work_dir, os.path.relpath(os.path.dirname(args.config), "./configs"),
                                os.path.splitext(os.path.basename(args.config))[0])
    os.makedirs(cfg.work_dir, exist_ok=True)
    cfg.launcher = args.launcher
    cfg.train_cfg.checkpoint = args.load_from
    if args.dataset_dir:
        cfg.dataset_cfg.dataset_dir = args.dataset_dir
    if args.cache_dir:
        cfg.dataset_cfg.cache_dir = args.cache_dir
    if args.exp_name:
        cfg.exp_name = args.exp_name
    if args.max_epoch is not None:
        cfg.train_cfg.max_epoch = args.max_epoch
    return cfg


def load_training_set(cfg, extractor, device, logger):
    """-> (keys bf16 [N, gh*gw, dim] on the device, pseudo labels fp32 [N,1,g,g], (gh, gw)); fills missing caches."""
    dcfg = cfg.dataset_cfg
    rank, world = udist.world()
    pl_dir = os.path.join(dcfg.cache_dir, "pseudo_label_cache", dcfg.trainset_cfg.DATASET)
    ds = USCODDataset(dcfg.trainset_cfg, dcfg.feature_extractor_cfg, "train", dcfg.dataset_dir, dcfg.cache_dir,
                      feature_extractor=extractor, prepare_cache=(rank == 0))
    if not os.path.exists(os.path.join(pl_dir, "index.json")):
        logger.info("pseudo-label cache %s is missing: generating it", pl_dir)
        fe = dcfg.feature_extractor_cfg  # pseudo labels always come from DINOv2-B/14 (generate_pseudo_label.py:110)
        sd = load_vit_state_dict(SimpleNamespace(type="dinov2", backbone="facebook/dinov2-base",
                                                 backbone_weights=fe.get("backbone_weights", None),
                                                 backbone_weight_base=fe.get("backbone_weight_base", None)))
        generate_from_folders(PseudoLabelGenerator(sd, "dinov2", th_bkg=float(dcfg.trainset_cfg.get("bkg_th", 0.6)),
                                                   device=device), [str(p) for p in ds.image_paths], pl_dir)
    if world > 1:
        torch.distributed.barrier()
    ds.cache_manager._caches.clear()  # re-open: the caches may just have been written (by rank 0)
    fc, pc = ds.cache_manager.get_features_cache(), ds.cache_manager.get_pseudo_label_cache()
    if fc.mode != "r" or pc.mode != "r" or fc.length() != len(ds) or pc.length() != len(ds):
        raise RuntimeError(f"feature / pseudo-label caches under {dcfg.cache_dir} do not match the {len(ds)} images")
    f0 = fc.read_file(0)
    C, gh, gw = f0.shape
    keys = torch.empty(len(ds), gh * gw, C, dtype=torch.bfloat16, device=device)
    pls = []
    for i in range(len(ds)):
        f = fc.read_file(i)
        keys[i] = f.reshape(C, gh * gw).t().to(device=device, dtype=torch.bfloat16)
        pl = pc.read_file(i).float()
        pls.append(pl.reshape(1, *pl.shape[-2:]))
    return keys, torch.stack(pls).to(device), (gh, gw), ds


def save_model(model, path: str) -> None:
    """`accelerator.save_model(model, path)`: a directory holding `model.safetensors`."""
    from safetensors.torch import save_file
    os.makedirs(path, exist_ok=True)
    save_file({k: v.detach().contiguous().cpu() for k, v in model.state_dict().items()},
              os.path.join(path, "model.safetensors"))


def main(argv=None):
    set_random_seed(42)
    args = parse_train_args(argv)
    rank, world = udist.world()
    if "RANK" in os.environ and not torch.distributed.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        torch.distributed.init_process_group("nccl")
        rank, world = udist.world()
    device = torch.device("cuda", torch.cuda.current_device())
    cfg = init_cfg(args)
    logger = ev._setup_run_dir(cfg, rank)
    extractor = backbone(cfg.dataset_cfg.feature_extractor_cfg, device=device)
    keys, pls, grid_in, _ = load_training_set(cfg, extractor, device, logger)
    logger.info("training set resident in HBM: %d images, keys %s (%.2f GB)", keys.shape[0], tuple(keys.shape),
                keys.numel() * 2 / 1e9)

    model = baseline(cfg.model_cfg).to(device)
    disc = Discriminator(cfg.model_cfg).to(device)
    if cfg.train_cfg.get("checkpoint", None):
        from safetensors.torch import load_file
        model.load_state_dict(load_file(ev.resolve_checkpoint(cfg.train_cfg.checkpoint)), strict=True)
    if world > 1:  # same initial weights on every rank (DDP broadcasts them in the reference)
        for t in list(model.state_dict().values()) + list(disc.state_dict().values()):
            torch.distributed.broadcast(t, src=0)
    tc, mc = cfg.train_cfg, cfg.model_cfg
    trainer = FirstStageTrainer(model, disc, lr0=tc.lr0, step_lr_size=tc.step_lr_size, step_lr_gamma=tc.step_lr_gamma,
                                feature_size=mc.feature_size, ema_weight=mc.ema_weight, max_epoch=tc.max_epoch,
                                start_finetune=tc.start_finetune)
    dis_trainer = DiscriminatorTrainer(disc, lr0=tc.dis_lr0, step_lr_size=tc.dis_step_lr_size,
                                       step_lr_gamma=tc.dis_step_lr_gamma)

    def save_checkpoint(epoch: int) -> None:
        if rank == 0:
            save_model(model, os.path.join(cfg.log_cfg.log_path, "ckp", f"epoch{epoch}.pth"))
        if world > 1:
            torch.distributed.barrier()

    def validate():
        model.eval()
        res = ev.evaluate_dataset(cfg, cfg.dataset_cfg.valset_cfg.DATASET, extractor, model, None, args, logger)
        model.train()
        return res

    loop = TrainLoop(cfg, trainer, dis_trainer, keys, pls, grid_in, validate=validate if not args.no_val else None,
                     save_checkpoint=save_checkpoint, logger=logger, seed=42, rank=rank, world_size=world)
    best = loop.run()
    # per-rank fingerprint of the trained weights (data-parallel ranks must agree bit for bit)
    digest = float(sum(v.double().abs().sum() for v in model.state_dict().values()))
    return {"best": best, "losses": loop.losses, "log_path": cfg.log_cfg.log_path, "weights_l1": digest}


if __name__ == "__main__":
    main()
