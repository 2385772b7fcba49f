"""Portable synthetic ("random-init") weights.  TEST INFRASTRUCTURE (see oracle/__init__.py).

There is no network for checkpoints, so every weight on the hot path is synthetic.
To let the build container (where the unmodified reference can be imported) and the
GPU box (where it cannot) agree on the *same* weights without shipping them, each
tensor is a pure function of (name, shape, kind, seed) drawn from numpy's PCG64 --
bit-stable across machines, unlike torch's default module init order.

kinds
  'w'     weight of a Linear/Conv: U(-b, b), b = 1/sqrt(fan_in), fan_in = prod(shape[1:])
          (the bound torch's default kaiming_uniform(a=sqrt(5)) produces)
  'b'     bias: U(-b, b) with b = 1/sqrt(fan_in) passed via ``fan_in``
  'gain'  norm weight: 1 + 0.1*U(-1,1)
  'beta'  norm bias:  0.05*U(-1,1)
  'emb'   free parameter (queries, offsets, positional tables): ``scale``*U(-sqrt3, sqrt3)
          (unit-variance uniform times scale)
  'scalar' gates/alpha: U(-0.5, 0.5) + ``shift``
  'var'   positive running variance: ``scale`` * (1 + 0.3*U(-1,1))
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Iterable, NamedTuple, Optional, Tuple

import numpy as np


class P(NamedTuple):
    name: str
    shape: Tuple[int, ...]
    kind: str
    fan_in: Optional[int] = None   # for 'b'
    scale: float = 1.0             # for 'emb'
    shift: float = 0.0             # for 'scalar'


def _rng(name: str, seed: int) -> np.random.Generator:
    key = (zlib.crc32(name.encode()) << 32) | (seed & 0xFFFFFFFF)
    return np.random.Generator(np.random.PCG64(key))


def synth(p: P, seed: int) -> np.ndarray:
    shape = tuple(int(s) for s in p.shape)
    n = int(np.prod(shape)) if len(shape) else 1
    u = _rng(p.name, seed).random(n, dtype=np.float32) * 2.0 - 1.0   # U(-1,1)
    if p.kind == "w":
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
        u *= np.float32(1.0 / math.sqrt(fan_in))
    elif p.kind == "b":
        fan_in = p.fan_in if p.fan_in else n
        u *= np.float32(1.0 / math.sqrt(fan_in))
    elif p.kind == "gain":
        u = np.float32(1.0) + np.float32(0.1) * u
    elif p.kind == "beta":
        u *= np.float32(0.05)
    elif p.kind == "emb":
        u *= np.float32(p.scale * math.sqrt(3.0))
    elif p.kind == "scalar":
        u = np.float32(0.5) * u + np.float32(p.shift)
    elif p.kind == "var":
        u = np.float32(p.scale) * (np.float32(1.0) + np.float32(0.3) * u)
    else:
        raise ValueError(f"unknown kind {p.kind!r} for {p.name}")
    return np.asarray(u, dtype=np.float32).reshape(shape)   # (ascontiguousarray would turn 0-d into 1-d)


def synth_state_dict(spec: Iterable[P], seed: int, prefix: str = "") -> Dict[str, np.ndarray]:
    return {prefix + p.name: synth(p, seed) for p in spec}


def count(spec: Iterable[P]) -> int:
    return sum(int(np.prod(p.shape)) if len(p.shape) else 1 for p in spec)


# ---------------------------------------------------------------------------------------
# small spec helpers shared by sd15.py and audio.py
# ---------------------------------------------------------------------------------------
def linear(name: str, cin: int, cout: int, bias: bool = True):
    out = [P(f"{name}.weight", (cout, cin), "w")]
    if bias:
        out.append(P(f"{name}.bias", (cout,), "b", fan_in=cin))
    return out


def conv(name: str, cin: int, cout: int, k: int):
    return [P(f"{name}.weight", (cout, cin, k, k), "w"),
            P(f"{name}.bias", (cout,), "b", fan_in=cin * k * k)]


def norm(name: str, c: int):
    return [P(f"{name}.weight", (c,), "gain"), P(f"{name}.bias", (c,), "beta")]
