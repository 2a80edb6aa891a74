"""Discriminator forward (BN train/eval) and APM fusion (CUDA) vs. the oracle."""
import copy
from types import SimpleNamespace

import pytest
import torch

from oracle import decoder as odec
from ucod_dpl_b200.models.discriminator import Discriminator, merge_pseudo_label

pytestmark = pytest.mark.gpu


def _disc(fs=68, seed=0):
    sd = odec.random_discriminator_state_dict(fs, seed)
    d = Discriminator(SimpleNamespace(dis_use_features=False, dim=768, feature_size=fs))
    d.load_state_dict(sd, strict=True)
    return d.cuda(), sd


def _masks(B, fs, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(B, 1, fs, fs, generator=g) < torch.rand(B, 1, 1, 1, generator=g)).float()


@pytest.mark.parametrize("B,fs", [(16, 68), (1, 68), (5, 16)])
def test_discriminator_train_bn(B, fs):
    d, sd = _disc(fs, 1)
    m = _masks(B, fs, 2)
    d.train()
    ref = odec.discriminator_forward(sd, m, train_bn=True)
    got = d(m.cuda(), None).cpu()
    assert got.shape == (B, 1)
    assert (got - ref).abs().max().item() < 2e-4
    # running statistics follow nn.BatchNorm2d (momentum 0.1, unbiased variance)
    import torch.nn.functional as F
    y = F.conv2d(m, sd["maskConv.layers.0.weight"], None, padding=1)
    want_mean = 0.9 * sd["maskConv.layers.1.running_mean"] + 0.1 * y.mean(dim=(0, 2, 3))
    want_var = 0.9 * sd["maskConv.layers.1.running_var"] + 0.1 * y.var(dim=(0, 2, 3), unbiased=True)
    assert (d.maskConv.layers[1].running_mean.cpu() - want_mean).abs().max().item() < 1e-4
    assert (d.maskConv.layers[1].running_var.cpu() - want_var).abs().max().item() < 1e-4
    assert int(d.maskConv.layers[1].num_batches_tracked) == 1


def test_discriminator_eval_bn():
    d, sd = _disc(68, 3)
    m = _masks(4, 68, 4)
    d.eval()
    ref = odec.discriminator_forward(sd, m, train_bn=False)
    got = d(m.cuda(), None).cpu()
    assert (got - ref).abs().max().item() < 2e-4


@pytest.mark.parametrize("train", [True, False])
@pytest.mark.parametrize("B,fs", [(16, 68), (3, 16)])
def test_two_calls_in_one_launch_equal_two_forwards(train, B, fs):
    """`forward_calls(masks, 2)` (what the APM uses) == forward(a) then forward(b): probabilities, every BatchNorm
    running buffer (updated in call order) and the batch counters."""
    d1, _ = _disc(fs, 7)
    d2 = copy.deepcopy(d1)
    d1.train(train), d2.train(train)
    a, b = _masks(B, fs, 8).cuda(), _masks(B, fs, 9).cuda()
    want = torch.cat([d1(a, None), d1(b, None)])
    got = d2.forward_calls(torch.cat([a, b]), 2)
    assert got.shape == (2 * B, 1)
    assert (got - want).abs().max().item() < 1e-6
    sd1, sd2 = d1.state_dict(), d2.state_dict()
    for k in sd1:
        if "running" in k:
            assert (sd1[k] - sd2[k]).abs().max().item() < 1e-6, k
        elif "num_batches" in k:
            assert int(sd1[k]) == int(sd2[k]) == (2 if train else 0), k


@pytest.mark.parametrize("epoch", [0, 3, 19])
def test_apm_merge(epoch):
    B, fs = 16, 68
    d, sd = _disc(fs, 5)
    d.train()
    g = torch.Generator().manual_seed(6)
    pl = torch.rand(B, 1, fs, fs, generator=g)
    teacher = torch.randn(B, 1, fs, fs, generator=g)
    student = torch.randn(B, 1, fs, fs, generator=g) + 0.3
    merged_r, loss_r, w_r, ps_r, pp_r = odec.apm_merge(sd, pl, teacher, student, epoch)
    merged, loss = merge_pseudo_label(d, pl.cuda(), teacher.cuda(), student.cuda(), None, cur_epoch=epoch)
    extra = merge_pseudo_label.last
    assert (extra["p_s"].cpu() - ps_r).abs().max().item() < 2e-4
    assert (extra["p_p"].cpu() - pp_r).abs().max().item() < 2e-4
    assert (extra["weight"].cpu() - w_r).abs().max().item() < 1e-3
    assert (merged.cpu() - merged_r).abs().max().item() < 1e-3
    assert abs(loss.item() - loss_r.item()) < 1e-3
