"""Drop-in for the reference's ``models/audio_encoder.py`` (``CLAPAudioEncoder`` :15-213, ``CLAPTextEncoder`` :216-283,
``compute_audio_text_similarity`` :286-): same class names, method names, argument meaning and return shapes, with the
audio path -- log-mel front end + HTSAT tower + projection -- on libc2d (``clap2diffusion_b200/clap.py``).

What differs from the reference (by design):
  * weights come from a state dict in Hugging Face's key layout (``audio_model.audio_encoder.*``,
    ``audio_projection.*``).  ``CLAPAudioEncoder(model_name=...)`` tries ``ClapModel.from_pretrained`` exactly like
    the reference and takes the audio tower's tensors from it; on a box without the checkpoint (no network) it raises
    -- pass ``state_dict=`` or use ``CLAPAudioEncoder.random_init`` there;
  * ``encode_audio`` never round-trips through numpy / the HF feature extractor: the 10 s clip goes to the GPU once
    and the STFT, mel projection and dB conversion run there (reference: host numpy STFT, :164-168);
  * CUDA only -- there is no CPU path; resampling (``librosa``) is only attempted when the input rate differs;
  * the text tower is not on the hot path: ``CLAPTextEncoder`` raises on construction.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Union

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .._lib import C2DError
from ..clap import ClapAudioTower, param_shapes


class CLAPAudioEncoder(nn.Module):
    """Audio encoder producing 512-d unit-norm CLAP embeddings (reference :15-213)."""

    def __init__(self, model_name: str = "laion/clap-htsat-unfused", sample_rate: int = 48000, target_length: float = 10.0,
                 device: str = "cuda", freeze: bool = False, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 dtype: torch.dtype = torch.bfloat16, feature_config: Optional[Dict[str, float]] = None):
        """feature_config: the checkpoint's feature-extractor settings (``frequency_min``, ``frequency_max``,
        ``truncation``, ``padding``) -- the reference gets them from ``ClapProcessor.from_pretrained(model_name)``
        (:47), i.e. from preprocessor_config.json.  Default: that file when it can be found (model directory or the
        local HF cache), else the values ``laion/clap-htsat-*`` checkpoints publish (frequency_min 50, frequency_max
        14000, rand_trunc / repeatpad), else the ClapFeatureExtractor class defaults."""
        super().__init__()
        self.model_name = model_name
        self.sample_rate = sample_rate
        self.target_length = target_length
        self.device = device
        if state_dict is None:
            state_dict = self._load_pretrained(model_name)
        missing = set(param_shapes()) - set(state_dict)
        if missing:
            raise KeyError(f"CLAP state dict lacks {len(missing)} tensors, e.g. {sorted(missing)[0]}")
        self.feature_config = self._feature_config(model_name, feature_config)
        if self.feature_config.get("truncation", "rand_trunc") != "rand_trunc" or self.feature_config.get("padding", "repeatpad") != "repeatpad":
            raise C2DError(f"feature extractor mode {self.feature_config} is not the unfused rand_trunc / repeatpad path")
        self.tower = ClapAudioTower(state_dict, device=device, dtype=dtype,
                                    frequency_min=self.feature_config["frequency_min"],
                                    frequency_max=self.feature_config["frequency_max"])
        self.embedding_dim = 512
        self._frozen = True          # inference engine: the tower holds packed, non-trainable weights
        if freeze:
            self.freeze_encoder()

    @staticmethod
    def _feature_config(model_name: str, given: Optional[Dict[str, float]]) -> Dict[str, float]:
        import json
        import os
        cfg = {"frequency_min": 0.0, "frequency_max": 14000.0, "truncation": "rand_trunc", "padding": "repeatpad", "source": "ClapFeatureExtractor defaults"}
        if "clap-htsat" in str(model_name):
            cfg.update(frequency_min=50.0, source="published preprocessor_config.json of laion/clap-htsat-* (frequency_min 50)")
        path = os.path.join(str(model_name), "preprocessor_config.json") if os.path.isdir(str(model_name)) else None
        if path is None:
            try:
                from transformers.utils import cached_file
                path = cached_file(model_name, "preprocessor_config.json", local_files_only=True)
            except Exception:
                path = None
        if path and os.path.exists(path):
            with open(path) as f:
                js = json.load(f)
            for k in ("frequency_min", "frequency_max", "truncation", "padding"):
                if k in js:
                    cfg[k] = js[k]
            cfg["source"] = path
        if given:
            cfg.update(given)
            cfg["source"] = "caller"
        cfg["frequency_min"], cfg["frequency_max"] = float(cfg["frequency_min"]), float(cfg["frequency_max"])
        return cfg

    @staticmethod
    def _load_pretrained(model_name: str) -> Dict[str, torch.Tensor]:
        try:
            from transformers import ClapModel
            model = ClapModel.from_pretrained(model_name)
        except Exception as e:          # no network / no cached checkpoint
            raise RuntimeError(f"cannot load {model_name!r} ({type(e).__name__}: {e}); pass state_dict= (HF key layout) "
                               "or use CLAPAudioEncoder.random_init()") from e
        sd = model.state_dict()
        return {k: v for k, v in sd.items() if k.startswith(("audio_model.", "audio_projection."))}

    @classmethod
    def random_init(cls, seed: int = 0, device: str = "cuda", dtype: torch.dtype = torch.bfloat16, **kw) -> "CLAPAudioEncoder":
        """Random-init weights of the ``laion/clap-htsat-unfused`` architecture (benchmarks, tests)."""
        from .. import synthetic
        sd = synthetic.random_state_dict(param_shapes(), seed, device)
        sd["audio_model.audio_encoder.batch_norm.running_mean"] = sd["audio_model.audio_encoder.batch_norm.running_mean"] * 40.0 - 12.0
        sd["audio_model.audio_encoder.batch_norm.running_var"] = sd["audio_model.audio_encoder.batch_norm.running_var"].abs() * 400.0 + 40.0
        kw.setdefault("feature_config", {"frequency_min": 0.0, "frequency_max": 14000.0})   # the goldens' extractor (class defaults)
        return cls(device=device, state_dict=sd, dtype=dtype, **kw)

    def freeze_encoder(self):
        self._frozen = True

    def unfreeze_encoder(self):
        raise C2DError("the libc2d CLAP tower is an inference engine (packed weights); fine-tuning it is not on the hot path")

    # ------------------------------------------------------------------ reference :87-131
    def preprocess_audio(self, audio: Union[np.ndarray, torch.Tensor], sample_rate: int) -> np.ndarray:
        if isinstance(audio, torch.Tensor):
            if audio.dtype == torch.bfloat16:
                audio = audio.float()
            audio = audio.cpu().numpy()
        audio = np.asarray(audio)
        if len(audio.shape) > 1:
            audio = audio.mean(axis=-1)
        if sample_rate != self.sample_rate:
            try:
                import librosa
            except ImportError as e:
                raise RuntimeError(f"resampling {sample_rate} -> {self.sample_rate} Hz needs librosa (not installed)") from e
            audio = librosa.resample(audio, orig_sr=sample_rate, target_sr=self.sample_rate)
        target_samples = int(self.sample_rate * self.target_length)
        if len(audio) < target_samples:
            audio = np.pad(audio, (0, target_samples - len(audio)), mode="constant")
        else:
            audio = audio[:target_samples]
        return audio

    # ------------------------------------------------------------------ reference :133-176
    @torch.no_grad()
    def encode_audio(self, audio: Union[np.ndarray, torch.Tensor, List], sample_rate: int = None) -> torch.Tensor:
        """Audio (one clip, a batch [B, n] or a list of clips) -> embeddings [B, 512] (fp32, unit norm)."""
        if sample_rate is None:
            sample_rate = self.sample_rate
        target = int(self.sample_rate * self.target_length)
        if (isinstance(audio, torch.Tensor) and audio.is_cuda and audio.dim() == 2 and audio.shape[1] == target
                and sample_rate == self.sample_rate):
            waves = audio.float()                         # already on the device in the tower's format: no host hop
        else:
            if isinstance(audio, list):
                arr = np.stack([self.preprocess_audio(a, sample_rate) for a in audio])
            else:
                a = audio
                if isinstance(a, torch.Tensor):
                    a = a.float().cpu().numpy()
                a = np.asarray(a)
                if a.ndim == 2 and a.shape[1] == target and sample_rate == self.sample_rate:
                    arr = a                               # a batch of ready clips
                else:
                    arr = self.preprocess_audio(a, sample_rate)
                    if arr.ndim == 1:
                        arr = arr[np.newaxis, :]
            waves = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).to(self.device)
        feats = self.tower.encode(waves)
        # the reference re-normalises the (already unit-norm) features (:174); kept for exact drop-in semantics
        return ops.l2_normalize(feats, eps=0.0)

    def forward(self, audio, sample_rate: int = None) -> torch.Tensor:
        return self.encode_audio(audio, sample_rate)

    def get_audio_embeds_from_file(self, audio_path: str) -> torch.Tensor:
        try:
            import librosa
        except ImportError as e:
            raise RuntimeError("loading audio files needs librosa (not installed)") from e
        audio, sr = librosa.load(audio_path, sr=None)
        return self.encode_audio(audio, sr)


class CLAPTextEncoder(nn.Module):
    """Placeholder with the reference's name (:216-283): the CLAP text tower is not on the inference hot path."""

    def __init__(self, model_name: str = "laion/clap-htsat-unfused", device: str = "cuda", freeze: bool = True):
        super().__init__()
        raise C2DError("CLAPTextEncoder is outside the B200 hot path (SURVEY §6); use transformers.ClapModel for text features")


def compute_audio_text_similarity(audio_embeds: torch.Tensor, text_embeds: torch.Tensor, temperature: float = 0.07) -> torch.Tensor:
    """Cosine-similarity logits between unit-norm audio and text embeddings, scaled by 1 / temperature (reference :286-)."""
    # CUDA only, like every libc2d-backed entry point: a CPU tensor raises C2DError inside ops (no fallback)
    a = ops.l2_normalize(audio_embeds.float().contiguous(), eps=0.0)
    t = ops.l2_normalize(text_embeds.float().contiguous(), eps=0.0)
    return ops.linear(a, (t / temperature).contiguous())
