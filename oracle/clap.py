"""CPU restatement of the CLAP audio path the reference calls: ``CLAPAudioEncoder.preprocess_audio`` /
``encode_audio`` (/root/reference/models/audio_encoder.py:87-176) = HF ``ClapFeatureExtractor`` (rand_trunc,
slaney filter bank) + ``ClapModel.get_audio_features`` + L2 normalisation.  TEST INFRASTRUCTURE (oracle/__init__.py).

The arithmetic lives in a third-party dependency that is not under /root/reference:
``transformers==4.35.2`` (requirements.txt:8); the installed 5.5.0 is what this restatement follows and is
pinned against (``oracle/make_golden_clap.py`` runs the UNMODIFIED ``ClapAudioModel`` / ``ClapProjectionLayer`` /
``ClapFeatureExtractor`` on portable synthetic weights and stores their outputs under tests/golden/clap_*.npz):
  * log-mel: transformers/audio_utils.py ``spectrogram`` (center/reflect pad, periodic Hann 1024, hop 480,
    |rfft|^2 through complex64, slaney mel filters, floor 1e-10, 10 log10) -- feature_extraction_clap.py:154-174
  * tower:   transformers/models/clap/modeling_clap.py ``ClapAudioEncoder.forward`` :814-918
             (``reshape_mel2img`` :777-811, ``ClapAudioPatchEmbed`` :246-339, ``ClapAudioLayer`` :505-622,
             ``ClapAudioSelfAttention`` :360-437, ``ClapAudioPatchMerging`` :680-731)
  * head:    ``ClapProjectionLayer`` :921-936, ``F.normalize`` :1555-1556
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

from .weights import P, conv, linear, norm

SR, N_FFT, HOP, N_MEL, N_FRAMES, N_SAMPLES = 48000, 1024, 480, 64, 1001, 480000
FMIN, FMAX = 0.0, 14000.0
SPEC, PATCH, EMBED, WINDOW = 256, 4, 96, 8
DEPTHS, HEADS = (2, 2, 6, 2), (4, 8, 16, 32)
HIDDEN, PROJ = 768, 512
LN_EPS, BN_EPS = 1e-5, 1e-5


# ---------------------------------------------------------------------------------------------------------
# spec of the audio tower + projection (HF state-dict names under ``audio_model.`` / ``audio_projection.``)
# ---------------------------------------------------------------------------------------------------------
def clap_audio_spec() -> List[P]:
    e = "audio_model.audio_encoder"
    s: List[P] = []
    s += [P(f"{e}.batch_norm.weight", (N_MEL,), "gain"), P(f"{e}.batch_norm.bias", (N_MEL,), "beta"),
          P(f"{e}.batch_norm.running_mean", (N_MEL,), "scalar", shift=-12.0),
          P(f"{e}.batch_norm.running_var", (N_MEL,), "var", scale=60.0)]
    s += conv(f"{e}.patch_embed.proj", 1, EMBED, PATCH) + norm(f"{e}.patch_embed.norm", EMBED)
    for i, (depth, heads) in enumerate(zip(DEPTHS, HEADS)):
        c = EMBED * 2 ** i
        for j in range(depth):
            b = f"{e}.layers.{i}.blocks.{j}"
            s += norm(f"{b}.layernorm_before", c)
            s += [P(f"{b}.attention.self.relative_position_bias_table", ((2 * WINDOW - 1) ** 2, heads), "emb", scale=0.3)]
            s += linear(f"{b}.attention.self.query", c, c) + linear(f"{b}.attention.self.key", c, c)
            s += linear(f"{b}.attention.self.value", c, c) + linear(f"{b}.attention.output.dense", c, c)
            s += norm(f"{b}.layernorm_after", c)
            s += linear(f"{b}.intermediate.dense", c, 4 * c) + linear(f"{b}.output.dense", 4 * c, c)
        if i < len(DEPTHS) - 1:
            s += linear(f"{e}.layers.{i}.downsample.reduction", 4 * c, 2 * c, bias=False)
            s += norm(f"{e}.layers.{i}.downsample.norm", 4 * c)
    s += norm(f"{e}.norm", HIDDEN)
    s += linear("audio_projection.linear1", HIDDEN, PROJ) + linear("audio_projection.linear2", PROJ, PROJ)
    return s


# ---------------------------------------------------------------------------------------------------------
# log-mel front end
# ---------------------------------------------------------------------------------------------------------
def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    mels = 3.0 * f / 200.0
    logstep = 27.0 / np.log(6.4)
    hi = f >= 1000.0
    return np.where(hi, 15.0 + np.log(np.maximum(f, 1e-12) / 1000.0) * logstep, mels)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    logstep = np.log(6.4) / 27.0
    return np.where(m >= 15.0, 1000.0 * np.exp(logstep * (m - 15.0)), 200.0 * m / 3.0)


def mel_filters_slaney() -> np.ndarray:
    """[513, 64] triangular filters, slaney scale + slaney (area) normalisation (audio_utils.mel_filter_bank)."""
    nb = N_FFT // 2 + 1
    fft_freqs = np.linspace(0, SR // 2, nb)
    mel_pts = np.linspace(_hz_to_mel_slaney(FMIN), _hz_to_mel_slaney(FMAX), N_MEL + 2)
    f = _mel_to_hz_slaney(mel_pts)
    fdiff = np.diff(f)
    slopes = f[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / fdiff[:-1]
    up = slopes[:, 2:] / fdiff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    fb *= (2.0 / (f[2:N_MEL + 2] - f[:N_MEL]))[None, :]
    return fb


def hann_window() -> np.ndarray:
    return np.hanning(N_FFT + 1)[:-1]


def preprocess_audio(audio: np.ndarray, sample_rate: int = SR) -> np.ndarray:
    """reference models/audio_encoder.py:87-131 (resampling needs librosa and is out of the synthetic path)."""
    a = np.asarray(audio, dtype=np.float32)
    if a.ndim > 1:
        a = a.mean(axis=-1)
    if sample_rate != SR:
        raise NotImplementedError("oracle: synthetic audio is generated at 48 kHz")
    if len(a) < N_SAMPLES:
        a = np.pad(a, (0, N_SAMPLES - len(a)))
    return a[:N_SAMPLES]


def log_mel(wave: np.ndarray) -> np.ndarray:
    """wave [480000] float32 -> [1001, 64] float32 log-mel in dB."""
    w = np.pad(wave.astype(np.float32), (N_FFT // 2, N_FFT // 2), mode="reflect").astype(np.float64)
    idx = np.arange(N_FRAMES)[:, None] * HOP + np.arange(N_FFT)[None, :]
    frames = w[idx] * hann_window()[None, :]
    spec = np.fft.rfft(frames, axis=1).astype(np.complex64)
    power = np.abs(spec).astype(np.float64) ** 2
    mel = np.maximum(1e-10, power @ mel_filters_slaney())
    return (10.0 * np.log10(mel)).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------
# HTSAT tower
# ---------------------------------------------------------------------------------------------------------
def _rel_pos_index() -> torch.Tensor:
    c = torch.stack(torch.meshgrid(torch.arange(WINDOW), torch.arange(WINDOW), indexing="ij")).flatten(1)
    r = (c[:, :, None] - c[:, None, :]).permute(1, 2, 0).contiguous()
    r[:, :, 0] += WINDOW - 1
    r[:, :, 1] += WINDOW - 1
    r[:, :, 0] *= 2 * WINDOW - 1
    return r.sum(-1)


def _shift_mask(H: int, W: int, shift: int) -> torch.Tensor:
    img = torch.zeros(1, H, W, 1)
    cnt = 0
    for hs in (slice(0, -WINDOW), slice(-WINDOW, -shift), slice(-shift, None)):
        for ws in (slice(0, -WINDOW), slice(-WINDOW, -shift), slice(-shift, None)):
            img[:, hs, ws, :] = cnt
            cnt += 1
    m = img.view(1, H // WINDOW, WINDOW, W // WINDOW, WINDOW, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1, WINDOW * WINDOW)
    d = m[:, None, :] - m[:, :, None]
    return torch.where(d != 0, torch.full_like(d, -100.0), torch.zeros_like(d))


def mel_to_image(mel_bn: torch.Tensor) -> torch.Tensor:
    """[B,1,1001,64] -> [B,1,256,256]: bicubic (align_corners) 1001 -> 1024 frames, 4 time chunks stacked on frequency."""
    x = F.interpolate(mel_bn, (SPEC * 4, N_MEL), mode="bicubic", align_corners=True)
    B = x.shape[0]
    x = x.reshape(B, 4, SPEC, N_MEL).permute(0, 1, 3, 2).contiguous()
    return x.reshape(B, 1, SPEC, SPEC)


def tower_forward(W: Dict[str, torch.Tensor], mel: torch.Tensor, taps: dict = None) -> torch.Tensor:
    """mel [B,1,1001,64] fp32 -> unit-norm CLAP embedding [B,512]."""
    e = "audio_model.audio_encoder"
    x = mel.transpose(1, 3)
    x = F.batch_norm(x, W[f"{e}.batch_norm.running_mean"], W[f"{e}.batch_norm.running_var"], W[f"{e}.batch_norm.weight"],
                     W[f"{e}.batch_norm.bias"], False, 0.0, BN_EPS).transpose(1, 3)
    img = mel_to_image(x)
    h = F.conv2d(img, W[f"{e}.patch_embed.proj.weight"], W[f"{e}.patch_embed.proj.bias"], stride=PATCH)
    h = h.flatten(2).transpose(1, 2)
    h = F.layer_norm(h, (EMBED,), W[f"{e}.patch_embed.norm.weight"], W[f"{e}.patch_embed.norm.bias"], LN_EPS)
    if taps is not None:
        taps["patch_embed"] = h
    rpi = _rel_pos_index().view(-1)
    B = h.shape[0]
    Hc = Wc = SPEC // PATCH
    for i, (depth, heads) in enumerate(zip(DEPTHS, HEADS)):
        C = EMBED * 2 ** i
        d = C // heads
        for j in range(depth):
            b = f"{e}.layers.{i}.blocks.{j}"
            shift = WINDOW // 2 if (j % 2 == 1 and min(Hc, Wc) > WINDOW) else 0
            sc = h
            t = F.layer_norm(h, (C,), W[f"{b}.layernorm_before.weight"], W[f"{b}.layernorm_before.bias"], LN_EPS)
            t = t.view(B, Hc, Wc, C)
            if shift:
                t = torch.roll(t, (-shift, -shift), (1, 2))
            nh, nw = Hc // WINDOW, Wc // WINDOW
            win = t.view(B, nh, WINDOW, nw, WINDOW, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, WINDOW * WINDOW, C)
            q = F.linear(win, W[f"{b}.attention.self.query.weight"], W[f"{b}.attention.self.query.bias"])
            k = F.linear(win, W[f"{b}.attention.self.key.weight"], W[f"{b}.attention.self.key.bias"])
            v = F.linear(win, W[f"{b}.attention.self.value.weight"], W[f"{b}.attention.self.value.bias"])
            nWB = win.shape[0]
            q, k, v = (z.view(nWB, 64, heads, d).transpose(1, 2) for z in (q, k, v))
            s = q @ k.transpose(-1, -2) / math.sqrt(d)
            bias = W[f"{b}.attention.self.relative_position_bias_table"][rpi].view(64, 64, heads).permute(2, 0, 1)
            s = s + bias[None]
            if shift:
                m = _shift_mask(Hc, Wc, shift)
                s = (s.view(B, nh * nw, heads, 64, 64) + m[None, :, None]).view(-1, heads, 64, 64)
            o = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(nWB, 64, C)
            o = F.linear(o, W[f"{b}.attention.output.dense.weight"], W[f"{b}.attention.output.dense.bias"])
            o = o.view(B, nh, nw, WINDOW, WINDOW, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hc, Wc, C)
            if shift:
                o = torch.roll(o, (shift, shift), (1, 2))
            h = sc + o.reshape(B, Hc * Wc, C)
            t = F.layer_norm(h, (C,), W[f"{b}.layernorm_after.weight"], W[f"{b}.layernorm_after.bias"], LN_EPS)
            t = F.gelu(F.linear(t, W[f"{b}.intermediate.dense.weight"], W[f"{b}.intermediate.dense.bias"]))
            h = h + F.linear(t, W[f"{b}.output.dense.weight"], W[f"{b}.output.dense.bias"])
        if taps is not None:
            taps[f"stage{i}"] = h
        if i < len(DEPTHS) - 1:
            p = f"{e}.layers.{i}.downsample"
            t = h.view(B, Hc, Wc, C)
            t = torch.cat([t[:, 0::2, 0::2], t[:, 1::2, 0::2], t[:, 0::2, 1::2], t[:, 1::2, 1::2]], -1)
            t = t.view(B, -1, 4 * C)
            t = F.layer_norm(t, (4 * C,), W[f"{p}.norm.weight"], W[f"{p}.norm.bias"], LN_EPS)
            h = F.linear(t, W[f"{p}.reduction.weight"])
            Hc, Wc = Hc // 2, Wc // 2
    h = F.layer_norm(h, (HIDDEN,), W[f"{e}.norm.weight"], W[f"{e}.norm.bias"], LN_EPS)
    pooled = h.mean(1)                      # the frequency regrouping of :896-907 is a permutation before a global mean
    z = F.linear(F.relu(F.linear(pooled, W["audio_projection.linear1.weight"], W["audio_projection.linear1.bias"])),
                 W["audio_projection.linear2.weight"], W["audio_projection.linear2.bias"])
    return F.normalize(z, dim=-1)


def encode_audio(W: Dict[str, torch.Tensor], waves: np.ndarray) -> torch.Tensor:
    """reference audio_encoder.py:133-176: waves [B, n] -> [B,512] (re-normalised as :174 does)."""
    mel = np.stack([log_mel(preprocess_audio(w)) for w in waves])[:, None]
    z = tower_forward(W, torch.from_numpy(mel))
    return z / z.norm(p=2, dim=-1, keepdim=True)
