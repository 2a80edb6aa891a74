"""`TrainLoop` schedule and sharding (host logic, no GPU): the decisions of engine/runner/loop_UCOD_DPL.py:93-215
replayed with recording stubs for the two trainers."""
from types import SimpleNamespace

import torch

from ucod_dpl_b200.engine.config import CfgNode
from ucod_dpl_b200.engine.runner.loop_UCOD_DPL import TrainLoop


class _Trainer:
    def __init__(self):
        self.model, self.fs, self.cur_epoch, self.finetune = object(), 68, 0, False
        self.calls, self.resets = [], 0

    def process_batch(self, keys, grid, pl):
        self.calls.append((self.cur_epoch, self.finetune, keys[:, 0, 0].tolist()))
        return torch.tensor(1.0)

    def start_finetune_phase(self):
        self.finetune = True
        self.resets += 1


class _Dis:
    def __init__(self):
        self.epochs, self.resets = [], 0

    def epoch_step(self, model, keys, grid, pl, feature_size=68):
        self.epochs.append(keys.shape[0])
        return torch.tensor(0.5)

    def reset_optimizer(self):
        self.resets += 1


def _cfg(**train):
    cfg = CfgNode(CfgNode.load_with_base("configs/uscod/UCOD-DPL_dinov2.py"))
    for k, v in train.items():
        cfg.train_cfg[k] = v
    return cfg


def _data(n):
    keys = torch.arange(n, dtype=torch.float32).reshape(n, 1, 1).expand(n, 4, 8).contiguous()
    return keys, torch.zeros(n, 1, 16, 16)


def test_schedule_matches_reference_decisions(monkeypatch):
    import os
    monkeypatch.chdir(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    cfg = _cfg()
    assert (cfg.train_cfg.max_epoch, cfg.train_cfg.start_finetune, cfg.train_cfg.dis_intertrain) == (25, -5, 2)
    keys, pl = _data(40)                       # 40 images, batch 16 -> 3 batches (16, 16, 8) per epoch
    tr, dis = _Trainer(), _Dis()
    saved, vals = [], []
    events = []

    def validate():
        vals.append(loop._cur_epoch)
        events.append(("val", loop._cur_epoch))
        return {"MAE": 0.5 - 0.01 * loop._cur_epoch if loop._cur_epoch < 20 else 0.9}

    def save(epoch):
        saved.append(epoch)
        events.append(("save", epoch))

    loop = TrainLoop(cfg, tr, dis, keys, pl, (2, 2), validate=validate, save_checkpoint=save)
    best = loop.run()
    assert len(tr.calls) == 25 * 3
    assert [c[0] for c in tr.calls[::3]] == list(range(25))
    # finetune from epoch max_epoch + start_finetune = 20 on: optimiser rebuilt once, discriminator frozen
    assert [c[1] for c in tr.calls[::3]] == [False] * 20 + [True] * 5 and tr.resets == 1 and dis.resets == 1
    assert len(dis.epochs) == 10 * 3           # epochs 0, 2, ..., 18, dis_epoch = 1
    # start_save / start_val = -50 -> from the beginning, every 5 epochs, checked after the epoch counter moved
    assert saved == [5, 10, 15, 20, 25] and vals == [5, 10, 15, 20, 25]
    assert events[:2] == [("save", 5), ("val", 5)]
    assert best["MAE"] == 0.5 - 0.15 and loop.best_mae == best["MAE"]   # epoch 15 is the best (later ones are worse)
    # every epoch is a permutation of the data set
    for e in range(25):
        seen = sorted(i for c in tr.calls[3 * e:3 * e + 3] for i in c[2])
        assert seen == [float(i) for i in range(40)]
    assert tr.calls[0][2] != tr.calls[3][2]    # reshuffled between epochs


def test_rank_sharding_and_ragged_tail(monkeypatch):
    import os
    monkeypatch.chdir(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    cfg = _cfg(max_epoch=1)
    keys, pl = _data(40)
    per_rank = []
    for rank in range(2):
        tr = _Trainer()
        TrainLoop(cfg, tr, _Dis(), keys, pl, (2, 2), seed=7, rank=rank, world_size=2).run()
        per_rank.append([c[2] for c in tr.calls])
    # 40 images, 2 ranks x 16: step 0 uses 32 images; the last global batch (8 left) is completed by wrapping around to
    # the head of the permutation like accelerate's BatchSamplerShard(even_batches=True): every rank gets a full batch
    # in every step (so every rank joins every all-reduce and no sample is weighted more than another within a step)
    assert [len(b) for b in per_rank[0]] == [16, 16] and [len(b) for b in per_rank[1]] == [16, 16]
    epoch = per_rank[0][0] + per_rank[1][0] + per_rank[0][1] + per_rank[1][1]
    assert sorted(epoch[:40]) == [float(i) for i in range(40)]
    assert epoch[40:] == epoch[:24]            # the 24 fill-in samples are the head of the same permutation


def test_no_discriminator_when_merge_method_differs(monkeypatch):
    import os
    monkeypatch.chdir(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    cfg = _cfg(max_epoch=2, merge_method="none")
    keys, pl = _data(16)
    dis = _Dis()
    TrainLoop(cfg, _Trainer(), dis, keys, pl, (2, 2)).run()
    assert dis.epochs == []
