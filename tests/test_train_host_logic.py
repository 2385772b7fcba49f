"""Stage-3 fine-tune step (clap2diffusion_b200/train.py): the hand-written reverse pass through the frozen UNet, the
processor adjoint, clipping, AdamW and the cosine schedule, checked on CPU against torch AUTOGRAD through the oracle
(oracle/sd15.py + oracle/audio.py) with the libc2d ops swapped for their torch doubles (tests/torch_ops.py).  The CUDA
kernels themselves are checked by tests/test_gpu_train.py."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import torch_ops
from oracle import audio as A
from oracle import pipeline as PL
from oracle import sd15

from clap2diffusion_b200.models.hierarchical_audio_v4 import ImprovedHierarchicalAudioEncoder
from clap2diffusion_b200.train import LEVELS, Stage3Trainer


def _t(a):
    return torch.from_numpy(np.asarray(a))


def make_batch(B=2, h=16, w=16, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {"audio_embedding": torch.stack([_t(PL.clap_embedding(10 + i)) for i in range(B)]),
            "image_latents": torch.randn(B, 4, h, w, generator=g),
            "text_embedding": torch.stack([_t(PL.text_states(p)) for p in ("a beach", "a forest", "a train", "a cafe")[:B]]),
            "noise": torch.randn(B, 4, h, w, generator=g),
            "timesteps": torch.tensor([731, 12, 405, 998][:B])}


def oracle_grads(W, batch, weight=2.0):
    """d [weight * mse(UNet(noisy, t, ctx + audio), noise)] / d processor parameters by autograd through the oracle."""
    psd = {lvl: {k: v.clone().requires_grad_(True) for k, v in W[f"proc_{lvl}"].items()} for lvl in LEVELS}
    Wg = dict(W)
    for lvl in LEVELS:
        Wg[f"proc_{lvl}"] = psd[lvl]
    with torch.no_grad():
        routed = A.improved_hier_forward(W["hier"], batch["audio_embedding"])["routed"]
    t = batch["timesteps"].float()
    a = (1.0 - t / 1000.0).view(-1, 1, 1, 1)
    noisy = a * batch["image_latents"] + (1.0 - a) * batch["noise"]                 # train_stage3.py:203-204
    eps = sd15.unet_forward(W["unet"], noisy, t, batch["text_embedding"], PL.make_attn2_hook(Wg, routed, "add"))
    loss = weight * F.mse_loss(eps, batch["noise"])
    loss.backward()
    return float(loss.detach()), {lvl: {k: v.grad.clone() for k, v in psd[lvl].items()} for lvl in LEVELS}


@pytest.fixture(scope="module")
def W():
    return PL.build_weights(seed=0, with_vae=False)


def _trainer(W, **kw):
    hier = ImprovedHierarchicalAudioEncoder().eval()
    hier.load_state_dict(W["hier"])
    return Stage3Trainer(W["unet"], hier, {lvl: W[f"proc_{lvl}"] for lvl in LEVELS}, device="cpu", dtype=torch.float32, **kw)


def test_frozen_unet_backward_matches_autograd(W):
    batch = make_batch()
    ref_loss, ref = oracle_grads(W, batch)
    with torch_ops.installed(), torch.no_grad():
        tr = _trainer(W)
        assert tr.num_params == 3 * 99_137
        loss = tr.forward_backward(batch["audio_embedding"], batch["image_latents"], batch["text_embedding"], batch["noise"],
                                   batch["timesteps"])
        got = tr.named_grads()
    assert abs(float(loss) - ref_loss) < 1e-5 * max(1.0, abs(ref_loss))
    for lvl in LEVELS:
        for k, g in ref[lvl].items():
            e = float((got[lvl][k].double() - g.double()).norm() / (g.double().norm() + 1e-30))
            assert float(g.norm()) > 0 and e < 2e-4, (lvl, k, e)


def test_train_step_matches_torch_adamw_and_schedule(W):
    """One full step (clip 0.5 -> AdamW(lr, wd 0.01) -> cosine schedule) equals torch.optim.AdamW + clip_grad_norm_ +
    CosineAnnealingLR applied to the oracle's autograd gradients (train_stage3.py:33-46, :182-189)."""
    batch = make_batch(seed=1)
    _, ref = oracle_grads(W, batch)
    params = [W[f"proc_{lvl}"][k].clone().requires_grad_(True) for lvl in LEVELS for k in sorted(W[f"proc_{lvl}"])]
    for p, (lvl, k) in zip(params, [(lvl, k) for lvl in LEVELS for k in sorted(W[f"proc_{lvl}"])]):
        p.grad = ref[lvl][k].clone()
    opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=0.01)
    sch = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=10, eta_min=1e-6)
    norm = torch.nn.utils.clip_grad_norm_(params, 0.5)
    opt.step(); sch.step()
    with torch_ops.installed(), torch.no_grad():
        tr = _trainer(W, learning_rate=1e-3, num_steps=10)
        out = tr.train_step(batch)
    assert abs(float(out["grad_norm"]) - float(norm)) < 1e-4 * float(norm)
    new = {lvl: tr.procs[lvl].state_dict() for lvl in LEVELS}
    for p, (lvl, k) in zip(params, [(lvl, k) for lvl in LEVELS for k in sorted(W[f"proc_{lvl}"])]):
        step = (p.detach() - W[f"proc_{lvl}"][k]).norm()
        assert float(step) > 0                                                     # the step moved the parameter
        # Adam's first step is ~ lr * sign(g): elements whose gradient is at rounding level may flip, so compare the UPDATE
        e = float((new[lvl][k] - p.detach()).norm() / step)
        assert e < 2e-2, (lvl, k, e)
    assert abs(tr.lr(1) - sch.get_last_lr()[0]) < 1e-12 and abs(tr.lr(0) - 1e-3) < 1e-15
    assert abs(tr.lr(10) - 1e-6) < 1e-15 and tr.lr(5) == pytest.approx(1e-6 + (1e-3 - 1e-6) * (1 + math.cos(math.pi / 2)) / 2)
    sd = tr.state_dict()
    assert set(sd) >= {"processor_early", "processor_mid", "processor_late", "mode", "optimizer_state_dict"}
