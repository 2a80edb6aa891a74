"""First-stage training step (CUDA) vs. the CPU oracle / the reference-generated golden vectors."""
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch
from safetensors.torch import load_file

from oracle import decoder as odec
from oracle import train as otr
from ucod_dpl_b200 import ops
from ucod_dpl_b200.models.discriminator import Discriminator
from ucod_dpl_b200.models.uscod import baseline
from ucod_dpl_b200.train import FirstStageTrainer

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def train_inputs(seed: int, B: int = 2):
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(B, 768, 37, 37, generator=g)
    pl = (torch.rand(B, 1, 16, 16, generator=g) < 0.35).float()
    return feats, pl


def _models():
    sd = load_file(str(ROOT / "weights" / "UCOD_DPL_dinov2.safetensors"))
    model = baseline(SimpleNamespace(dim=768))
    model.load_state_dict(sd, strict=True)
    D = Discriminator(SimpleNamespace(dis_use_features=False, dim=768, feature_size=68))
    D.load_state_dict(odec.random_discriminator_state_dict(68, seed=31), strict=True)
    return sd, model.cuda().train(), D.cuda().train()


def test_autograd_gradients_match_oracle():
    """model(features) + loss.backward() through the custom autograd Function == torch autograd on the fp32 oracle
    (the 1x1 conv input is bf16 on the GPU: tolerance 2 % of the largest gradient entry)."""
    sd, model, _ = _models()
    feats, _ = train_inputs(100)
    x = torch.nn.functional.interpolate(feats, size=(68, 68), mode="bilinear")
    g = torch.Generator().manual_seed(5)
    tgt = torch.rand(2, 1, 68, 68, generator=g)
    p = {k: sd["decoder." + k].clone().float().requires_grad_(True) for k in otr.PARAM_ORDER}
    fg, bg, ortho = otr.student_forward(p, x)
    bce = torch.nn.functional.binary_cross_entropy_with_logits
    (bce(fg, tgt) + bce(bg, 1 - tgt) + ortho).backward()
    fg2, bg2, ortho2 = model(x.cuda())
    loss = bce(fg2, tgt.cuda()) + bce(bg2, 1 - tgt.cuda()) + ortho2
    loss.backward()
    assert abs(float(ortho2) - float(ortho)) < 2e-2 * abs(float(ortho))
    for k in otr.PARAM_ORDER:
        ref = p[k].grad
        got = dict(model.decoder.named_parameters())[k].grad.cpu()
        tol = 2e-2 * ref.abs().max().item() + 1e-7
        assert (got - ref).abs().max().item() < tol, (k, (got - ref).abs().max().item(), tol)


def test_three_fused_steps_match_oracle():
    """Three fused steps (teacher fwd, student fwd, APM, BCE + ortho, backward, AdamW, StepLR, EMA) vs the oracle,
    which is itself pinned to the reference's TrainLoop._process_batch (tests/test_oracle_train.py).
    Stage isolation: with random features the binarised student / teacher masks sit on the sigmoid = 0.5 boundary,
    and a BatchNorm-in-train-mode discriminator at batch 2 amplifies single-pixel flips, so the APM outputs of the
    GPU run (parity-tested on their own in test_discriminator_apm_gpu.py) are injected into the oracle step."""
    gold = np.load(ROOT / "tests" / "golden" / "train.npz")
    sd0, model, D = _models()
    dis_sd = odec.random_discriminator_state_dict(68, seed=31)
    tr = FirstStageTrainer(model, D, lr0=2e-4)
    tr.cur_epoch = 3
    sd = {k: v.clone() for k, v in sd0.items()}
    state = otr.new_state(sd)
    for step in range(3):
        feats, pl = train_inputs(100 + step)
        tok = ops.features_to_tokens_bf16(feats.cuda())
        loss = tr.process_batch(tok, (37, 37), pl.cuda())
        out = otr.train_step(sd, dis_sd, state, feats, pl, cur_epoch=3, global_step=2 * step,
                             lr=otr.step_lr(2e-4, step), merged_override=tr.last["merged"].cpu(),
                             dis_loss_override=tr.last["dis_loss"].cpu())
        assert abs(float(loss) - float(out["loss"])) < 2e-3 * abs(float(out["loss"])) + 2e-4
        assert abs(float(loss) - float(gold[f"loss_{step}"])) < 0.15      # same ballpark as the un-isolated reference
        for name in otr.PARAM_ORDER:
            ref = out["grads"][name].reshape(-1).numpy()
            got = tr.views[name].cpu().numpy()
            tol = 2e-2 * np.abs(ref).max() + 1e-7
            assert np.abs(got - ref).max() < tol, (step, name, np.abs(got - ref).max(), tol)
    # AdamW moves every element by ~lr per step; compare the movement of the parameters after the three steps
    for k, v in model.state_dict().items():
        ref, init = sd[k].numpy(), sd0[k].numpy()
        moved_ref, moved = ref - init, v.cpu().numpy() - init
        assert np.abs(moved - moved_ref).mean() < 0.25 * np.abs(moved_ref).mean() + 1e-7, k
        np.testing.assert_allclose(v.cpu().numpy(), ref, atol=1.3e-3)
    assert tr.global_step == 6 and tr.opt_steps == 3


def test_teacher_follows_the_ema_weights():
    """The teacher forward must see the EMA weights of the CURRENT step (they are updated in place by the fused
    AdamW / EMA kernel): after a few steps its logits equal a fresh EMA decoder loaded from the state_dict."""
    _, model, D = _models()
    tr = FirstStageTrainer(model, D, lr0=2e-2)          # large lr: the EMA weights move visibly
    tr.cur_epoch = 3
    feats, pl = train_inputs(900, B=2)
    tok = ops.features_to_tokens_bf16(feats.cuda())
    for _ in range(4):
        tr.process_batch(tok, (37, 37), pl.cuda())
    fresh = baseline(SimpleNamespace(dim=768))
    fresh.load_state_dict({k: v.detach().cpu().clone() for k, v in model.state_dict().items()}, strict=True)
    fresh = fresh.cuda().eval()
    want, _, _ = fresh.decoder_ema.forward_tokens(tok, (37, 37), (68, 68), want_bg=False)
    model.decoder_ema._packed = None
    now, _, _ = model.decoder_ema.forward_tokens(tok, (37, 37), (68, 68), want_bg=False)
    assert torch.allclose(now, want, atol=1e-5)
    # what the trainer's own next teacher pass computes (cache rebuilt inside the step)
    tr._forward_backward(tok, (37, 37), pl.cuda())
    cached = model.decoder_ema._w_dec_bf16()
    assert torch.equal(cached, model.decoder_ema.decoupling.weight.detach().reshape(128, 768).to(torch.bfloat16))


@pytest.mark.parametrize("cur_epoch,B,tol", [(3, 16, 0.05), (22, 2, 0.03)])
def test_fused_step_matches_oracle_without_injection(cur_epoch, B, tol):
    """One fused step against the oracle with NOTHING injected: the oracle computes its own teacher / student masks,
    discriminator probabilities, APM weights and merged targets.
    * epoch 22 of 25 (finetune schedule term 22/20 >= 1): the APM weight saturates at 1, the target is the binarised
      teacher mask; a few threshold flips between bf16 and fp32 logits move a few of 9 248 target pixels.
    * epoch 3, batch 16: the discriminator's train-mode BatchNorm sees 16 masks, so single-pixel flips no longer swing
      its statistics; the APM weights follow the oracle's to a few 1e-3.
    Gradients are compared at `tol` of the largest entry, the loss and the per-image APM weights directly."""
    sd0, model, D = _models()
    dis_sd = odec.random_discriminator_state_dict(68, seed=31)
    tr = FirstStageTrainer(model, D, lr0=2e-4)
    tr.cur_epoch = cur_epoch
    feats, pl = train_inputs(300 + cur_epoch, B)
    tok = ops.features_to_tokens_bf16(feats.cuda())
    loss = tr.process_batch(tok, (37, 37), pl.cuda())
    sd = {k: v.clone() for k, v in sd0.items()}
    out = otr.train_step(sd, dis_sd, otr.new_state(sd), feats, pl, cur_epoch=cur_epoch, global_step=0, lr=2e-4)
    flips = (tr.last["merged"].cpu() - out["merged"]).abs()
    print(f"epoch {cur_epoch} B {B}: loss {float(loss):.5f} vs {float(out['loss']):.5f}; merged target mean |diff| "
          f"{flips.mean().item():.2e}, APM weight oracle {out['weight'].flatten()[:4].tolist()}")
    # loss = BCE(fg) + BCE(bg) + ortho - dis_loss: the decoder part is compared tightly; dis_loss = BCE(D(student mask), 0)
    # goes through the discriminator's train-mode BatchNorm (at batch 2 a single flipped pixel moves it by 1e-2)
    dis, dis_o = float(tr.last["dis_loss"]), float(out["dis_loss"])
    dec, dec_o = float(loss) + dis, float(out["loss"]) + dis_o
    print(f"   decoder part of the loss {dec:.5f} vs {dec_o:.5f}; dis_loss {dis:.5f} vs {dis_o:.5f}")
    assert abs(dec - dec_o) < 1e-2 * abs(dec_o) + 1e-3
    assert abs(dis - dis_o) < (0.05 if B < 8 else 5e-3)
    assert flips.mean().item() < 5e-3
    for name in otr.PARAM_ORDER:
        ref = out["grads"][name].reshape(-1).numpy()
        got = tr.views[name].cpu().numpy()
        lim = tol * np.abs(ref).max() + 1e-7
        print(f"   {name}: max |grad diff| / max |grad| = {np.abs(got - ref).max() / (np.abs(ref).max() + 1e-12):.4f}")
        assert np.abs(got - ref).max() < lim, (name, np.abs(got - ref).max(), lim)


def test_adamw_ema_kernel_matches_torch():
    from ucod_dpl_b200 import _lib
    import ctypes
    g = torch.Generator().manual_seed(0)
    n = 98690
    p = torch.randn(n, generator=g)
    ema = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([ref], lr=2e-4)
    dp, dm, dv, de = p.cuda(), torch.zeros(n).cuda(), torch.zeros(n).cuda(), ema.cuda()
    ema_ref = ema.clone()
    for t in range(1, 4):
        grad = torch.randn(n, generator=g) * 0.01
        ref.grad = grad.clone()
        opt.step()
        alpha = min(1 - 1 / (2 * (t - 1) + 1), 0.99)
        ema_ref.mul_(alpha).add_(ref.data, alpha=1 - alpha)
        _lib.call("ucod_adamw_ema_step", _lib.ptr(dp), _lib.ptr((grad * 2).cuda()), _lib.ptr(dm), _lib.ptr(dv),
                  _lib.ptr(de), ctypes.c_uint64(n), _lib.c_float(2e-4), _lib.c_float(0.9), _lib.c_float(0.999),
                  _lib.c_float(1e-8), _lib.c_float(0.01), t, _lib.c_float(0.5), _lib.c_float(alpha), _lib.stream_ptr())
    np.testing.assert_allclose(dp.cpu().numpy(), ref.data.numpy(), atol=2e-6)
    np.testing.assert_allclose(de.cpu().numpy(), ema_ref.numpy(), atol=2e-6)


def dis_inputs(seed: int, B: int = 4):
    g = torch.Generator().manual_seed(seed)
    pseudo = (torch.rand(B, 1, 68, 68, generator=g) < 0.4).float()
    student = (torch.rand(B, 1, 68, 68, generator=g) < torch.rand(B, 1, 1, 1, generator=g)).float()
    return pseudo, student


def test_discriminator_epoch_steps_match_reference_golden():
    """Three Discriminator_epoch iterations (fwd + hand-written conv/BN/LeakyReLU/linear backward, AdamW, StepLR,
    running statistics) vs the reference module + torch autograd (tests/golden/train.npz).  All fp32."""
    from ucod_dpl_b200.train import DiscriminatorTrainer
    gold = np.load(ROOT / "tests" / "golden" / "train.npz")
    D = Discriminator(SimpleNamespace(dis_use_features=False, dim=768, feature_size=68))
    D.load_state_dict(odec.random_discriminator_state_dict(68, seed=31), strict=True)
    D = D.cuda().train()
    tr = DiscriminatorTrainer(D, lr0=1e-3)
    names = [n for n, _ in D.named_parameters()]
    for step in range(3):
        pseudo, student = dis_inputs(200 + step)
        loss = tr.step(pseudo.cuda(), student.cuda())
        assert abs(float(loss) - float(gold[f"dis_loss_{step}"])) < 2e-4 * abs(float(gold[f"dis_loss_{step}"])) + 1e-5
        if step == 0:
            for n, view in zip(names, tr.grad_views):
                ref = gold["dis_grad0_" + n].reshape(-1)
                tol = 2e-3 * np.abs(ref).max() + 1e-7
                assert np.abs(view.cpu().numpy() - ref).max() < tol, (n, np.abs(view.cpu().numpy() - ref).max(), tol)
    for k, v in D.state_dict().items():
        if "num_batches_tracked" in k:
            assert int(v) == int(gold["dis_final_" + k])
            continue
        np.testing.assert_allclose(v.cpu().numpy(), gold["dis_final_" + k], atol=3e-5, rtol=2e-3, err_msg=k)


def test_graph_replay_matches_eager_steps():
    """CUDA-graph replay of the forward / APM / backward sequence: same losses and parameters as the eager path over
    six steps with changing inputs."""
    outs = []
    for use_graph in (False, True):
        _, model, D = _models()
        tr = FirstStageTrainer(model, D, lr0=2e-4, use_graph=use_graph)
        tr.cur_epoch = 3
        losses = []
        for step in range(6):
            feats, pl = train_inputs(500 + step, B=4)
            tok = ops.features_to_tokens_bf16(feats.cuda())
            losses.append(float(tr.process_batch(tok, (37, 37), pl.cuda())))
        assert tr.global_step == 12 and tr.opt_steps == 6
        assert (tr._graph is not None) == use_graph
        outs.append((losses, {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}))
    (l0, p0), (l1, p1) = outs
    assert np.allclose(l0, l1, rtol=2e-4, atol=2e-5), (l0, l1)
    for k in p0:
        assert torch.allclose(p0[k], p1[k], rtol=1e-4, atol=5e-5), k


def test_training_steps_are_bit_reproducible():
    """No float atomics on the step's path (fixed-order partial sums in the decoder forward / backward, split-K weight
    gradient reduced in order): two runs from the same state give bit-identical gradients, losses and parameters."""
    outs = []
    for _ in range(2):
        _, model, D = _models()
        tr = FirstStageTrainer(model, D, lr0=2e-4)
        tr.cur_epoch = 3
        grads, losses = [], []
        for step in range(4):
            feats, pl = train_inputs(900 + step, B=4)
            tok = ops.features_to_tokens_bf16(feats.cuda())
            losses.append(tr.process_batch(tok, (37, 37), pl.cuda()).clone())
            grads.append(tr.flat_g.clone())
        outs.append((torch.stack(losses), torch.stack(grads), tr.flat_p.clone(), tr.flat_ema.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    assert outs[0][1].abs().max().item() > 0


def _ddp_worker(rank, world, port, out):
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", rank))
    try:
        _, model, D = _models()
        tr = FirstStageTrainer(model, D, lr0=2e-4)
        tr.cur_epoch = 3
        for step in range(3):
            feats, pl = train_inputs(700 + 10 * step + rank, B=4)      # every rank its own shard of the global batch
            tok = ops.features_to_tokens_bf16(feats.cuda())
            _, _ = tr._forward_backward(tok, (37, 37), pl.cuda())
            local = tr.flat_g.clone()
            gathered = [torch.empty_like(local) for _ in range(world)]
            torch.distributed.all_gather(gathered, local)
            tr._optimizer_step()                                        # NCCL all-reduce + AdamW(1/world) + EMA
            mean = torch.stack(gathered).mean(0)
            assert torch.allclose(tr.flat_g / world, mean, rtol=1e-5, atol=1e-8), step
        params = [torch.empty_like(tr.flat_p) for _ in range(world)]
        torch.distributed.all_gather(params, tr.flat_p)
        assert all(torch.equal(params[0], p) for p in params[1:])      # replicas stay bit-identical
        if rank == 0:
            out.put("ok")
    finally:
        torch.distributed.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_rank_nccl_gradient_allreduce():
    """The path's only data-path collective on real hardware: two ranks, different shards, one flat NCCL all-reduce of
    the decoder gradients per step; the applied gradient is the mean over ranks and the replicas stay identical."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    mp.spawn(_ddp_worker, args=(2, 29591, out), nprocs=2, join=True)
    assert out.get(timeout=10) == "ok"
