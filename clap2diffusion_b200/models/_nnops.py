"""Glue between ``nn.Module`` parameter containers and the libc2d ops (host side only)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .. import ops
from .audio_attention_processor import _CastCache


class Native(nn.Module):
    """Base for the drop-in modules: parameters live in ordinary nn containers (so ``state_dict`` keys match
    the reference), compute goes through libc2d in the dtype of the incoming activation."""

    def __init__(self):
        super().__init__()
        object.__setattr__(self, "_cc", _CastCache())

    # weights are used in the activation dtype; biases / norm affine parameters always in fp32
    def _w(self, p: torch.Tensor, dt: torch.dtype) -> torch.Tensor:
        return self._cc.get(p, dt, ("w", p.data_ptr(), tuple(p.shape)))

    def _f(self, p: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        return None if p is None else self._cc.get(p, torch.float32, ("f", p.data_ptr(), tuple(p.shape)))

    def lin(self, layer: nn.Linear, x: torch.Tensor, act: int = ops.ACT_NONE, residual=None, out=None) -> torch.Tensor:
        return ops.linear(x, self._w(layer.weight, x.dtype), self._f(layer.bias), act=act, residual=residual, out=out)

    def lin_w(self, weight, bias, x, act: int = ops.ACT_NONE, residual=None) -> torch.Tensor:
        return ops.linear(x, self._w(weight, x.dtype), self._f(bias), act=act, residual=residual)

    def ln(self, layer: nn.LayerNorm, x: torch.Tensor) -> torch.Tensor:
        return ops.layer_norm(x.contiguous(), self._f(layer.weight), self._f(layer.bias), layer.eps)


def require_cuda(x: torch.Tensor, who: str) -> None:
    ops.require_cuda(x, who)
