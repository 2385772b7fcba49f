"""Stage-3 fine-tune step on the GPU (-m gpu): every backward / optimiser kernel of csrc/train.cu against torch autograd
of the same op, and the whole step -- frozen-UNet reverse pass through libc2d kernels -- against autograd through the
oracle (oracle/sd15.py on the GPU as the checker).  Gates: fp32 mode <= 1e-3 relative L2 per gradient tensor, bf16
mode cosine >= 0.99 and relative L2 <= 1e-1 per tensor, 1.5e-1 for the three gate scalars (bf16 activations AND bf16
activation gradients through ~600 layers)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import torch_ops as T
from oracle import pipeline as PL
from test_train_host_logic import make_batch, oracle_grads

from clap2diffusion_b200 import ops
from clap2diffusion_b200.models.hierarchical_audio_v4 import ImprovedHierarchicalAudioEncoder
from clap2diffusion_b200.train import LEVELS, Stage3Trainer

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda"
F32, BF16 = torch.float32, torch.bfloat16


def rnd(*shape, dtype=F32, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def tol(dt):
    return 2e-5 if dt == F32 else 2e-2


@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("B,HW,C,silu", [(2, 256, 320, True), (3, 64, 1280, True), (2, 1024, 640, False), (1, 4096, 960, True)])
def test_group_norm_bwd(B, HW, C, silu, dtype):
    x, dy = rnd(B, HW, C, dtype=dtype) * 2 + 0.5, rnd(B, HW, C, dtype=dtype, seed=1)
    g, b = 1 + 0.1 * rnd(C, seed=2), 0.1 * rnd(C, seed=3)
    add = rnd(B, HW, C, dtype=dtype, seed=4)
    for a in (None, add):
        got = ops.group_norm_bwd(x, dy, g, b, 32, 1e-5, silu, add=a)
        ref = T.group_norm_bwd(x.cpu(), dy.cpu(), g.cpu(), b.cpu(), 32, 1e-5, silu, add=None if a is None else a.cpu())
        assert rel(got, ref) < tol(dtype)
    if dtype == BF16:       # the forward pass's channel statistics handed to the adjoint: same result, no statistics pass
        st = ops.channel_stats(x, torch.zeros(B * C * 2, device=DEV, dtype=torch.int64))
        n0 = ops._lib.launch_count()
        got2 = ops.group_norm_bwd(x, dy, g, b, 32, 1e-5, silu, add=add, stats=st)
        n1 = ops._lib.launch_count()
        ops.group_norm_bwd(x, dy, g, b, 32, 1e-5, silu, add=add)
        print("launches with / without forward statistics:", n1 - n0, ops._lib.launch_count() - n1)
        assert rel(got2, got) < 1e-3 and n1 - n0 < ops._lib.launch_count() - n1


@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("M,C", [(512, 320), (100, 1280), (64, 640)])
def test_layer_norm_bwd(M, C, dtype):
    x, dy = rnd(M, C, dtype=dtype) * 1.5 - 0.3, rnd(M, C, dtype=dtype, seed=1)
    g = 1 + 0.1 * rnd(C, seed=2)
    add = rnd(M, C, dtype=dtype, seed=4)
    assert rel(ops.layer_norm_bwd(x, dy, g), T.layer_norm_bwd(x.cpu(), dy.cpu(), g.cpu())) < tol(dtype)
    assert rel(ops.layer_norm_bwd(x, dy, g, add=add), T.layer_norm_bwd(x.cpu(), dy.cpu(), g.cpu(), add=add.cpu())) < tol(dtype)


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_geglu_bwd_and_layout_adjoints(dtype):
    ag, dy = rnd(300, 2 * 640, dtype=dtype), rnd(300, 640, dtype=dtype, seed=1)
    assert rel(ops.geglu_bwd(ag, dy), T.geglu_bwd(ag.cpu(), dy.cpu())) < tol(dtype)
    x = rnd(2, 6, 4, 24, dtype=dtype)
    assert torch.equal(ops.zero_insert2x(x).cpu(), T.zero_insert2x(x.cpu()))
    x2 = rnd(2, 8, 12, 40, dtype=dtype)
    assert rel(ops.sumpool2x2(x2), T.sumpool2x2(x2.cpu())) < tol(dtype)
    add = rnd(2, 8, 12, 16, dtype=dtype, seed=3)
    assert torch.equal(ops.slice_channels(x2, 8, 16).cpu(), x2[..., 8:24].contiguous().cpu())
    assert rel(ops.slice_channels(x2, 24, 16, add=add), T.slice_channels(x2.cpu(), 24, 16, add=add.cpu())) < tol(dtype)
    # adjoint identities: <upsample2x(x), y> == <x, sumpool2x2(y)>
    xs, ys = rnd(1, 4, 4, 8), rnd(1, 8, 8, 8, seed=5)
    assert abs(float((ops.upsample2x(xs) * ys).sum()) - float((xs * ops.sumpool2x2(ys)).sum())) < 1e-3


@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("stride,up", [(1, False), (2, False), (1, True)])
def test_conv3x3_data_gradient_through_forward_kernel(stride, up, dtype):
    """dgrad of a pad-1 3x3 convolution = the forward kernel on the flipped / transposed weight (after zero insertion for
    stride 2, before the 2x2 sum for the fused upsample) -- against autograd of conv2d."""
    B, H, Cin, Cout = 2, 16, 64, 96
    x = rnd(B, H, H, Cin, dtype=dtype)
    w32 = rnd(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=1)
    w = ops.pack_conv3x3(w32, dtype)
    y = ops.conv3x3(ops.upsample2x(x) if up else x, w, None, stride=stride)
    gy = rnd(*y.shape, dtype=dtype, seed=2)
    wd = w.flip(1, 2).permute(3, 1, 2, 0).contiguous()
    gx = ops.conv3x3(ops.zero_insert2x(gy) if stride == 2 else gy, wd, None)
    if up:
        gx = ops.sumpool2x2(gx)
    xr = x.float().cpu().permute(0, 3, 1, 2).requires_grad_(True)
    with torch.enable_grad():
        xin = F.interpolate(xr, scale_factor=2.0, mode="nearest") if up else xr
        yr = F.conv2d(xin, w.float().cpu().permute(0, 3, 1, 2), None, stride=stride, padding=1)
        ref = torch.autograd.grad(yr, xr, gy.float().cpu().permute(0, 3, 1, 2))[0].permute(0, 2, 3, 1)
    assert tuple(gx.shape) == tuple(x.shape) and rel(gx, ref) < (1e-4 if dtype == F32 else 2e-2)


@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("B,heads,Nq,Nkv,d,packed", [(2, 8, 256, 256, 40, True), (1, 8, 1024, 1024, 40, True), (2, 8, 64, 64, 160, True),
                                                      (2, 8, 256, 77, 40, False), (2, 8, 100, 77, 80, False), (1, 4, 64, 10, 160, False),
                                                      (2, 8, 1024, 1024, 80, True), (1, 8, 4096, 4096, 40, True), (2, 4, 200, 333, 64, False),
                                                      (1, 2, 384, 130, 128, False), (2, 8, 256, 256, 160, True), (1, 4, 300, 77, 160, False)])
def test_attention_bwd(B, heads, Nq, Nkv, d, packed, dtype):
    C = heads * d
    if packed:                       # self-attention: q, k, v are column slices of one [B, N, 3C] buffer
        qkv = rnd(B, Nq, 3 * C, dtype=dtype, scale=0.7)
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        dqkv = torch.zeros_like(qkv)
        dq, dk, dv = dqkv[..., :C], dqkv[..., C:2 * C], dqkv[..., 2 * C:]
    else:                            # cross-attention: k, v are the halves of the cached [B, T, 2C] K/V
        q = rnd(B, Nq, C, dtype=dtype, scale=0.7)
        kv = rnd(B, Nkv, 2 * C, dtype=dtype, seed=1, scale=0.7)
        k, v = kv[..., :C], kv[..., C:]
        dq, dkv = torch.zeros_like(q), torch.zeros_like(kv)
        dk, dv = dkv[..., :C], dkv[..., C:]
    o = ops.attention(q, k, v, heads)
    do = rnd(B, Nq, C, dtype=dtype, seed=2)
    ops.attention_bwd(q, k, v, o, do, heads, dq, dk, dv)
    rq, rk, rv = (torch.zeros(t.shape) for t in (q, k, v))
    T.attention_bwd(q.cpu(), k.cpu(), v.cpu(), o.cpu(), do.cpu(), heads, rq, rk, rv)
    # the forward kernel's log-sum-exp by-product (long sequences, head_dim <= 64, bf16) lets the adjoint skip a sweep
    lse = torch.empty(B, heads, Nq, device=DEV, dtype=torch.float32)
    o2, written = ops.attention(q, k, v, heads, lse=lse)
    assert torch.equal(o2, o) and written == (dtype == BF16 and d <= 64 and Nkv > 128)
    if written:
        ref_lse = torch.logsumexp((q.float().view(B, Nq, heads, d).transpose(1, 2) @ k.float().view(B, Nkv, heads, d).permute(0, 2, 3, 1))
                                  * d ** -0.5, -1) * 1.4426950408889634
        assert float((lse - ref_lse).abs().max()) < 2e-2
        g2 = [torch.zeros_like(t) for t in (dq, dk, dv)]
        ops.attention_bwd(q, k, v, o, do, heads, g2[0], g2[1], g2[2], lse=lse)
        for a_, b_ in zip(g2, (dq, dk, dv)):
            assert rel(a_, b_) < 5e-3
    t_ = 1e-4 if dtype == F32 else 3e-2
    assert rel(dq, rq) < t_ and rel(dk, rk) < t_ and rel(dv, rv) < t_


def test_loss_reductions_and_optimizer_kernels():
    pred, tgt = rnd(2, 8, 8, 4, dtype=BF16), rnd(2, 4, 8, 8, seed=1)
    acc = torch.zeros(1, device=DEV, dtype=torch.float64)
    g = ops.mse_loss_grad(pred, tgt, 2.0, acc)
    racc = torch.zeros(1, dtype=torch.float64)
    rg = T.mse_loss_grad(pred.cpu(), tgt.cpu(), 2.0, racc)
    assert rel(g, rg) < 1e-2 and abs(float(acc) - float(racc)) < 1e-5 * float(racc)
    x = rnd(3, 77, 768, dtype=BF16)
    assert rel(ops.colsum(x), x.float().sum(1)) < 1e-5
    out = torch.ones(3, 768, device=DEV)
    ops.colsum(x, out=out, accumulate=True)
    assert rel(out, 1 + x.float().sum(1)) < 1e-5
    s, af, alpha = rnd(4, 768), rnd(4, 768, seed=1), torch.tensor([0.3], device=DEV)
    da, rda = torch.zeros(1, device=DEV), torch.zeros(1)
    daf = ops.gate_bwd(s, af, alpha, da)
    rdaf = T.gate_bwd(s.cpu(), af.cpu(), alpha.cpu(), rda)
    assert rel(daf, rdaf) < 1e-6 and abs(float(da) - float(rda)) < 1e-4 * abs(float(rda))
    z, dh = rnd(40, 64), rnd(4, 64, seed=1)
    assert rel(ops.gelu_bwd_bcast(z, dh, 10), T.gelu_bwd_bcast(z.cpu(), dh.cpu(), 10)) < 1e-5
    # clip + AdamW against torch.optim.AdamW / clip_grad_norm_ over three steps
    p0, gr = rnd(1000), [rnd(1000, seed=10 + i) * 3 for i in range(3)]
    p = p0.clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    rp = p0.clone().cpu().requires_grad_(True)
    opt = torch.optim.AdamW([rp], lr=1e-2, weight_decay=0.01)
    sc, nrm, ss = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV), torch.zeros(1, device=DEV, dtype=torch.float64)
    for i in range(3):
        ss.zero_()
        ops.sumsq(gr[i], ss)
        ops.clip_scale(ss, 0.5, sc, nrm)
        ops.adamw_step(p, gr[i], m, v, 1e-2, 0.9, 0.999, 1e-8, 0.01, i + 1, sc)
        rp.grad = gr[i].cpu().clone()
        n = torch.nn.utils.clip_grad_norm_([rp], 0.5)
        opt.step()
        assert abs(float(nrm) - float(n)) < 1e-4 * float(n)
    assert rel(p, rp.detach()) < 1e-5


def _trainer(W, dtype, **kw):
    hier = ImprovedHierarchicalAudioEncoder().to(DEV).eval()
    hier.load_state_dict({k: v.to(DEV) for k, v in W["hier"].items()})
    return Stage3Trainer(W["unet"], hier, {lvl: W[f"proc_{lvl}"] for lvl in LEVELS}, device=DEV, dtype=dtype, **kw)


@pytest.fixture(scope="module")
def W():
    return PL.build_weights(seed=0, with_vae=False)


@pytest.fixture(scope="module")
def ref16(W):
    """Autograd through the oracle on the GPU (fp32, TF32 off) for a 2-sample batch of 16 x 16 latents."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    batch = make_batch(B=2, h=16, w=16)
    Wg = {k: {n: t.to(DEV) for n, t in v.items()} for k, v in W.items()}
    loss, grads = oracle_grads(Wg, {k: v.to(DEV) for k, v in batch.items()})
    return batch, loss, grads


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_stage3_gradients_vs_oracle_autograd(W, ref16, dtype):
    batch, ref_loss, ref = ref16
    tr = _trainer(W, dtype)
    n0 = ops._lib.launch_count()
    loss = tr.forward_backward(batch["audio_embedding"], batch["image_latents"], batch["text_embedding"], batch["noise"],
                               batch["timesteps"])
    assert ops._lib.launch_count() - n0 > 800                        # the reverse pass runs on libc2d kernels
    got = tr.named_grads()
    assert abs(float(loss) - ref_loss) < (1e-4 if dtype == F32 else 3e-2) * abs(ref_loss)
    worst = 0.0
    for lvl in LEVELS:
        for k, g in ref[lvl].items():
            a, b = got[lvl][k].double().reshape(-1).cpu(), g.double().reshape(-1).cpu()
            e = float((a - b).norm() / (b.norm() + 1e-30))
            cos = float((a * b).sum() / (a.norm() * b.norm() + 1e-30)) if a.numel() > 1 else 1.0
            worst = max(worst, e)
            print(f"  {dtype} {lvl:5s} {k:22s} rel-L2 {e:.2e} cos {cos:.5f}")
            if dtype == F32:
                assert e < 1e-3, (lvl, k, e)
            elif a.numel() > 1:
                assert e < 1e-1 and cos > 0.99, (lvl, k, e, cos)
            else:
                assert e < 1.5e-1, (lvl, k, e)              # the gate scalar: one number summed over every site of the level
    print(f"{dtype}: loss {float(loss):.6f} (oracle {ref_loss:.6f}); worst gradient rel-L2 {worst:.2e}")


def test_stage3_train_steps_reduce_the_loss(W):
    """A few optimiser steps on one fixed batch (bf16 product mode, raised learning rate): the loss goes down, the
    parameters move, and the saved state has the layout scripts/inference.py loads."""
    batch = make_batch(B=2, h=16, w=16, seed=3)
    tr = _trainer(W, BF16, learning_rate=3e-3, num_steps=20)
    before = tr.flat.clone()
    losses = [float(tr.train_step(batch)["diffusion"]) for _ in range(6)]
    print("stage-3 losses:", ["%.5f" % v for v in losses])
    assert losses[-1] < losses[0] and float((tr.flat - before).abs().max()) > 0
    sd = tr.state_dict()
    assert tuple(sd["processor_mid"]["audio_proj.0.weight"].shape) == (64, 768) and sd["step"] == 6


def test_stage3_step_as_cuda_graph_matches_eager(W):
    """Three optimiser steps replayed from ONE captured CUDA graph (device-resident schedule and step counter) land on the
    parameters of three eager steps, report the same losses, and launch nothing from the host."""
    batches = [make_batch(B=2, h=16, w=16, seed=20 + i) for i in range(3)]
    eager = _trainer(W, BF16, learning_rate=1e-3, num_steps=10)
    le = [float(eager.train_step(b)["diffusion"]) for b in batches]
    gr = _trainer(W, BF16, learning_rate=1e-3, num_steps=10)
    gr.capture(batches[0])
    assert gr.step_count == 0 and int(gr._step_dev) == 0 and torch.equal(gr.flat, _trainer(W, BF16).flat)
    n0 = ops._lib.launch_count()
    lg = [float(gr.train_step(b)["diffusion"]) for b in batches]
    assert ops._lib.launch_count() == n0 and gr.graph_launches > 800
    assert gr.step_count == 3 and int(gr._step_dev) == 3
    print("losses eager", le, "graph", lg)
    for a, b in zip(le, lg):
        assert abs(a - b) < 1e-4 * abs(a)
    moved = float((eager.flat - _trainer(W, BF16).flat).norm())
    assert float((gr.flat - eager.flat).norm()) < 1e-3 * moved          # double atomics in the GroupNorm adjoint: order-dependent rounding only
    with pytest.raises(ValueError):
        gr.train_step(make_batch(B=2, h=8, w=8))


def test_train_stage3_cli_writes_reference_layout(tmp_path):
    """scripts/train_stage3.py (the reference's configuration keys and file names): a few graph-replayed steps on batches
    in the reference's format, then unet_adapter_final.pth in the layout scripts/inference.py loads."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("train_stage3_cli", os.path.join(ROOT, "scripts", "train_stage3.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    data = tmp_path / "data"
    data.mkdir()
    for i in range(2):
        b = make_batch(B=2, h=16, w=16, seed=40 + i)
        torch.save({k: b[k] for k in ("audio_embedding", "image_latents", "text_embedding")}, data / f"batch{i}.pt")
    ck = tmp_path / "ck"
    tr = cli.main(["--checkpoint-dir", str(ck), "--data", str(data), "--num-steps", "4", "--batch-size", "2", "--learning-rate", "1e-3",
                   "--log-interval", "2", "--save-interval", "4"])
    assert tr.step_count == 4 and tr._graph is not None
    saved = torch.load(ck / "unet_adapter_final.pth", weights_only=True)
    assert saved["mode"] == "add" and saved["step"] == 4
    for lvl in LEVELS:
        for k, v in tr.procs[lvl].state_dict().items():
            assert torch.equal(saved[f"processor_{lvl}"][k], v.detach().cpu())
    full = torch.load(ck / "audio_projector_stage3_finetuned.pth", weights_only=True)
    assert {"step", "hierarchical_state_dict", "optimizer_state_dict", "config"} <= set(full) and full["config"]["gradient_clipping"] == 0.5
    # resume: the processors just written are picked up
    unet_sd, hier, procs = cli.load_models(ck, torch.device(DEV), 0)
    assert procs is not None and torch.equal(procs["mid"]["alpha"], saved["processor_mid"]["alpha"])
