"""CLAP HTSAT (unfused) audio tower on libc2d: waveform -> log-mel -> Swin tower -> 512-d unit-norm embedding.

Follows what the reference calls in ``models/audio_encoder.py:164-174`` -- Hugging Face ``ClapFeatureExtractor``
(rand_trunc path: slaney filter bank, dB) and ``ClapModel.get_audio_features`` (``ClapAudioEncoder.forward``,
transformers/models/clap/modeling_clap.py:814-918, + ``ClapProjectionLayer`` + ``F.normalize``) -- with the
host-side numpy STFT and the host->device hop removed: the waveform goes to the GPU once and everything,
including framing / DFT / mel projection, runs there.

Design notes (B200-first, not a port of the HF module tree):
  * the DFT is an fp32 GEMM of the windowed frames against a constant [1026, 1024] cos|sin matrix and the mel
    projection an fp32 GEMM against the [64, 513] filter bank (c2d_linear); framing, |.|^2 and dB + the folded
    eval-mode BatchNorm are three small kernels; clips are processed in chunks so the frame matrix stays small;
  * bicubic 1001 -> 1024 resampling, the 4-chunk fold to a 256 x 256 "image" and the 4 x 4 patch gather are one
    kernel, the patch embedding a K = 16 GEMM;
  * tokens stay in image order for the whole tower: window partition, cyclic shift, their inverses and the shift
    mask are index arithmetic inside the window-attention kernel, so there is not a single permute / roll copy;
  * Q, K, V are one fused GEMM with bias; GELU is fused into the MLP's first GEMM, residual adds into the
    epilogues of the projection / second MLP GEMM.
State-dict keys are HF's (``audio_model.audio_encoder.*``, ``audio_projection.*``).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from . import ops

SR, N_FFT, HOP, N_MEL, N_FRAMES, N_SAMPLES = 48000, 1024, 480, 64, 1001, 480000
DFT_IM_OFF, DFT_N_PAD = 520, 1040        # tensor-core DFT output row: re at 0..512, im at 520..1032, zero padded to 1040
SPEC, PATCH, EMBED, WINDOW = 256, 4, 96, 8
DEPTHS, HEADS = (2, 2, 6, 2), (4, 8, 16, 32)
HIDDEN, PROJ = 768, 512


def _slaney_mel_filters(frequency_min: float = 0.0, frequency_max: float = 14000.0) -> np.ndarray:
    """[513, 64] triangular slaney-scale, area-normalised filters between frequency_min and frequency_max at 48 kHz
    (HF audio_utils.mel_filter_bank(norm="slaney", mel_scale="slaney"), what ClapFeatureExtractor builds as
    ``mel_filters_slaney`` from its ``frequency_min`` / ``frequency_max``).  The class defaults are 0 / 14000; the
    ``laion/clap-htsat-unfused`` checkpoint's preprocessor_config.json sets frequency_min = 50."""
    nb = N_FFT // 2 + 1
    fft_freqs = np.linspace(0, SR // 2, nb)

    def hz2mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-12) / 1000.0) * (27.0 / np.log(6.4)), 3.0 * f / 200.0)

    def mel2hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), 200.0 * m / 3.0)

    f = mel2hz(np.linspace(hz2mel(float(frequency_min)), hz2mel(float(frequency_max)), N_MEL + 2))
    slopes = f[None, :] - fft_freqs[:, None]
    fb = np.maximum(0.0, np.minimum(-slopes[:, :-2] / np.diff(f)[:-1], slopes[:, 2:] / np.diff(f)[1:]))
    return fb * (2.0 / (f[2:] - f[:-2]))[None, :]


def param_shapes() -> Dict[str, tuple]:
    """HF state-dict names -> shapes of the audio tower + projection (28,190,872 parameters + BatchNorm running stats)."""
    e = "audio_model.audio_encoder"
    out: Dict[str, tuple] = {}

    def lin(n, cin, cout, bias=True):
        out[f"{n}.weight"] = (cout, cin)
        if bias:
            out[f"{n}.bias"] = (cout,)

    def norm(n, c):
        out[f"{n}.weight"] = (c,)
        out[f"{n}.bias"] = (c,)

    norm(f"{e}.batch_norm", N_MEL)
    out[f"{e}.batch_norm.running_mean"] = (N_MEL,)
    out[f"{e}.batch_norm.running_var"] = (N_MEL,)
    out[f"{e}.patch_embed.proj.weight"] = (EMBED, 1, PATCH, PATCH)
    out[f"{e}.patch_embed.proj.bias"] = (EMBED,)
    norm(f"{e}.patch_embed.norm", EMBED)
    for i, (depth, heads) in enumerate(zip(DEPTHS, HEADS)):
        c = EMBED * 2 ** i
        for j in range(depth):
            b = f"{e}.layers.{i}.blocks.{j}"
            norm(f"{b}.layernorm_before", c)
            out[f"{b}.attention.self.relative_position_bias_table"] = ((2 * WINDOW - 1) ** 2, heads)
            for n in ("query", "key", "value"):
                lin(f"{b}.attention.self.{n}", c, c)
            lin(f"{b}.attention.output.dense", c, c)
            norm(f"{b}.layernorm_after", c)
            lin(f"{b}.intermediate.dense", c, 4 * c)
            lin(f"{b}.output.dense", 4 * c, c)
        if i < len(DEPTHS) - 1:
            lin(f"{e}.layers.{i}.downsample.reduction", 4 * c, 2 * c, bias=False)
            norm(f"{e}.layers.{i}.downsample.norm", 4 * c)
    norm(f"{e}.norm", HIDDEN)
    lin("audio_projection.linear1", HIDDEN, PROJ)
    lin("audio_projection.linear2", PROJ, PROJ)
    return out


class ClapAudioTower:
    """``ClapAudioTower(state_dict, device, dtype)``: ``encode(waves fp32 [B, 480000]) -> fp32 [B, 512]`` (unit norm)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", dtype=torch.bfloat16, clip_chunk: int = 16,
                 frequency_min: float = 0.0, frequency_max: float = 14000.0):
        """frequency_min / frequency_max: mel filter-bank range of the checkpoint's feature extractor
        (preprocessor_config.json; ClapFeatureExtractor defaults 0 / 14000)."""
        self.device, self.dtype, self.clip_chunk = torch.device(device), dtype, int(clip_chunk)
        self.frequency_min, self.frequency_max = float(frequency_min), float(frequency_max)
        self._sd = state_dict
        self.w: Dict[str, torch.Tensor] = {}
        self._pack()
        self._sd = None

    # ------------------------------------------------------------------ weights / constants
    def _f32(self, name):
        return self._sd[name].detach().to(self.device, torch.float32).contiguous()

    def _mat(self, w):
        w = w.contiguous()
        return w if self.dtype == torch.float32 else ops.cast(w, self.dtype)

    def _pack(self):
        e = "audio_model.audio_encoder"
        w = self.w
        dev = self.device
        # front-end constants (fp32): periodic Hann window, DFT matrix [cos | sin] of the 513 non-negative bins, mel filters
        n = np.arange(N_FFT, dtype=np.float64)
        k = np.arange(N_FFT // 2 + 1, dtype=np.float64)
        ang = 2.0 * np.pi * np.outer(k, n) / N_FFT
        w["window"] = torch.from_numpy(np.hanning(N_FFT + 1)[:-1].astype(np.float32)).to(dev)
        w["dft"] = torch.from_numpy(np.concatenate([np.cos(ang), np.sin(ang)], 0).astype(np.float32)).to(dev).contiguous()
        if self.dtype == torch.bfloat16:
            # tensor-core DFT: constant rows [HI | HI | LO] (bf16 split of the fp32 matrix) against frames [hi | lo | hi];
            # cos rows at 0..512, sin rows at DFT_IM_OFF..+512, zero rows pad N to a multiple of 16
            d3 = torch.zeros(DFT_N_PAD, N_FFT, device=dev, dtype=torch.float32)
            d3[:N_FFT // 2 + 1] = w["dft"][:N_FFT // 2 + 1]
            d3[DFT_IM_OFF:DFT_IM_OFF + N_FFT // 2 + 1] = w["dft"][N_FFT // 2 + 1:]
            hi = d3.to(torch.bfloat16)
            lo = (d3 - hi.float()).to(torch.bfloat16)
            w["dft3"] = torch.cat([hi, hi, lo], dim=1).contiguous()                  # [1040, 3072]
        w["mel_fb"] = torch.from_numpy(_slaney_mel_filters(self.frequency_min, self.frequency_max).T.astype(np.float32).copy()).to(dev).contiguous()    # [64, 513]
        # eval-mode BatchNorm2d over mel bins folded into the dB kernel: y = dB * a + b
        g, b = self._f32(f"{e}.batch_norm.weight"), self._f32(f"{e}.batch_norm.bias")
        rm, rv = self._f32(f"{e}.batch_norm.running_mean"), self._f32(f"{e}.batch_norm.running_var")
        a = g / torch.sqrt(rv + 1e-5)
        w["bn_a"], w["bn_b"] = a.contiguous(), (b - rm * a).contiguous()
        w["pe.weight"] = self._mat(self._f32(f"{e}.patch_embed.proj.weight").reshape(EMBED, PATCH * PATCH))
        w["pe.bias"] = self._f32(f"{e}.patch_embed.proj.bias")
        w["pe.norm.weight"], w["pe.norm.bias"] = self._f32(f"{e}.patch_embed.norm.weight"), self._f32(f"{e}.patch_embed.norm.bias")
        # relative-position index of an 8 x 8 window (ClapAudioSelfAttention.create_relative_position_index)
        c = np.stack(np.meshgrid(np.arange(WINDOW), np.arange(WINDOW), indexing="ij")).reshape(2, -1)
        rel = (c[:, :, None] - c[:, None, :]).transpose(1, 2, 0) + (WINDOW - 1)
        rpi = torch.from_numpy((rel[:, :, 0] * (2 * WINDOW - 1) + rel[:, :, 1]).reshape(-1)).to(dev)
        for i, (depth, heads) in enumerate(zip(DEPTHS, HEADS)):
            for j in range(depth):
                p = f"{e}.layers.{i}.blocks.{j}"
                q = f"l{i}.{j}"
                for nm in ("layernorm_before", "layernorm_after"):
                    w[f"{q}.{nm}.weight"], w[f"{q}.{nm}.bias"] = self._f32(f"{p}.{nm}.weight"), self._f32(f"{p}.{nm}.bias")
                w[f"{q}.qkv.weight"] = self._mat(torch.cat([self._f32(f"{p}.attention.self.{n}.weight") for n in ("query", "key", "value")], 0))
                w[f"{q}.qkv.bias"] = torch.cat([self._f32(f"{p}.attention.self.{n}.bias") for n in ("query", "key", "value")], 0).contiguous()
                tbl = self._f32(f"{p}.attention.self.relative_position_bias_table")              # [225, heads]
                w[f"{q}.rel_bias"] = tbl[rpi].view(64, 64, heads).permute(2, 0, 1).contiguous()  # [heads, 64, 64]
                w[f"{q}.proj.weight"] = self._mat(self._f32(f"{p}.attention.output.dense.weight"))
                w[f"{q}.proj.bias"] = self._f32(f"{p}.attention.output.dense.bias")
                w[f"{q}.fc1.weight"] = self._mat(self._f32(f"{p}.intermediate.dense.weight"))
                w[f"{q}.fc1.bias"] = self._f32(f"{p}.intermediate.dense.bias")
                w[f"{q}.fc2.weight"] = self._mat(self._f32(f"{p}.output.dense.weight"))
                w[f"{q}.fc2.bias"] = self._f32(f"{p}.output.dense.bias")
            if i < len(DEPTHS) - 1:
                p = f"{e}.layers.{i}.downsample"
                w[f"ds{i}.norm.weight"], w[f"ds{i}.norm.bias"] = self._f32(f"{p}.norm.weight"), self._f32(f"{p}.norm.bias")
                w[f"ds{i}.reduction.weight"] = self._mat(self._f32(f"{p}.reduction.weight"))
        w["norm.weight"], w["norm.bias"] = self._f32(f"{e}.norm.weight"), self._f32(f"{e}.norm.bias")
        for n in ("linear1", "linear2"):
            w[f"proj.{n}.weight"] = self._f32(f"audio_projection.{n}.weight")       # head in fp32 (tiny)
            w[f"proj.{n}.bias"] = self._f32(f"audio_projection.{n}.bias")

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def log_mel(self, waves: torch.Tensor) -> torch.Tensor:
        """waves fp32 [B, 480000] (device) -> BatchNorm-ed log-mel fp32 [B, 1001, 64] (``ClapFeatureExtractor`` + batch_norm)."""
        w = self.w
        B = waves.shape[0]
        out = torch.empty(B, N_FRAMES, N_MEL, device=self.device, dtype=torch.float32)
        for b0 in range(0, B, self.clip_chunk):
            wv = waves[b0:b0 + self.clip_chunk].contiguous()
            if "dft3" in w:
                frames3 = ops.stft_frames_split(wv, w["window"], HOP, N_FRAMES)     # [b*1001, 3072] bf16 = [hi | lo | hi]
                dft = ops.linear(frames3, w["dft3"])                                # [b*1001, 1040] tcgen05, fp32 accumulation
                power = ops.power_spectrum(dft, nb=N_FFT // 2 + 1, im_off=DFT_IM_OFF)
            else:
                frames = ops.stft_frames(wv, w["window"], HOP, N_FRAMES)            # [b*1001, 1024]
                dft = ops.linear(frames, w["dft"])                                  # [b*1001, 1026]  fp32 GEMM (parity mode)
                power = ops.power_spectrum(dft)                                     # [b*1001, 513]
            mel = ops.linear(power, w["mel_fb"])                                    # [b*1001, 64]
            ops.log_mel_affine(mel, w["bn_a"], w["bn_b"], 1e-10, out=out[b0:b0 + wv.shape[0]].view(-1, N_MEL))
        return out

    @torch.no_grad()
    def tower(self, mel_bn: torch.Tensor, taps: dict = None) -> torch.Tensor:
        """BatchNorm-ed log-mel fp32 [B, 1001, 64] -> fp32 [B, 512] unit-norm embedding."""
        w = self.w
        B = mel_bn.shape[0]
        patches = ops.clap_patches(mel_bn.contiguous(), self.dtype)                 # [B*4096, 16]
        h = ops.linear(patches, w["pe.weight"], w["pe.bias"]).view(B, 4096, EMBED)
        h = ops.layer_norm(h, w["pe.norm.weight"], w["pe.norm.bias"])
        if taps is not None:
            taps["patch_embed"] = h
        Hc = Wc = SPEC // PATCH
        for i, (depth, heads) in enumerate(zip(DEPTHS, HEADS)):
            for j in range(depth):
                q = f"l{i}.{j}"
                shift = WINDOW // 2 if (j % 2 == 1 and min(Hc, Wc) > WINDOW) else 0
                t = ops.layer_norm(h, w[f"{q}.layernorm_before.weight"], w[f"{q}.layernorm_before.bias"])
                qkv = ops.linear(t, w[f"{q}.qkv.weight"], w[f"{q}.qkv.bias"])
                a = ops.window_attention(qkv, w[f"{q}.rel_bias"], Hc, Wc, heads, shift)
                h = ops.linear(a, w[f"{q}.proj.weight"], w[f"{q}.proj.bias"], residual=h)
                t = ops.layer_norm(h, w[f"{q}.layernorm_after.weight"], w[f"{q}.layernorm_after.bias"])
                t = ops.linear(t, w[f"{q}.fc1.weight"], w[f"{q}.fc1.bias"], act=ops.ACT_GELU)
                h = ops.linear(t, w[f"{q}.fc2.weight"], w[f"{q}.fc2.bias"], residual=h)
            if taps is not None:
                taps[f"stage{i}"] = h
            if i < len(DEPTHS) - 1:
                t = ops.patch_merge(h, Hc, Wc)
                t = ops.layer_norm(t, w[f"ds{i}.norm.weight"], w[f"ds{i}.norm.bias"])
                h = ops.linear(t, w[f"ds{i}.reduction.weight"])
                Hc, Wc = Hc // 2, Wc // 2
        h = ops.layer_norm(h, w["norm.weight"], w["norm.bias"])
        pooled = ops.token_mean(h)                                                   # fp32 [B, 768]
        if taps is not None:
            taps["pooled"] = pooled
        z = ops.linear(pooled, w["proj.linear1.weight"], w["proj.linear1.bias"], act=ops.ACT_RELU)
        z = ops.linear(z, w["proj.linear2.weight"], w["proj.linear2.bias"])
        return ops.l2_normalize(z)

    @torch.no_grad()
    def encode(self, waves: torch.Tensor, taps: dict = None) -> torch.Tensor:
        mel = self.log_mel(waves.to(self.device, torch.float32).contiguous())
        if taps is not None:
            taps["mel_bn"] = mel
        return self.tower(mel, taps)


def flops_per_clip() -> float:
    """Algorithmic FLOPs of one clip: DFT + mel GEMMs and the tower's dense layers / window attention."""
    f = 2.0 * N_FRAMES * N_FFT * (N_FFT + 2) + 2.0 * N_FRAMES * 513 * N_MEL + 2.0 * 4096 * 16 * EMBED
    n = 4096
    for i, depth in enumerate(DEPTHS):
        c = EMBED * 2 ** i
        f += depth * (2.0 * n * c * (3 * c + c + 8 * c) + 4.0 * n * 64 * c)
        if i < len(DEPTHS) - 1:
            f += 2.0 * (n // 4) * 4 * c * 2 * c
            n //= 4
    return f + 2.0 * HIDDEN * PROJ + 2.0 * PROJ * PROJ
