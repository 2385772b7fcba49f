"""clap2diffusion_b200 -- B200 (sm_100a) implementation of the CLAP2Diffusion inference hot path.

Layout
  csrc/ + libc2d.so   hand-written CUDA kernels behind the C ABI in include/c2d.h
  _lib.py, ops.py     ctypes binding and the torch-tensor front end (device memory + streams only)
  unet.py, vae.py     SD-1.5 UNet / VAE-decoder launch sequences (host logic only)
  sampler.py          DDIM / Euler + CFG loop, CUDA-graph capture, data-parallel sharding
  models/             drop-in modules with the reference's names, signatures and state-dict keys
"""
from . import _lib  # noqa: F401  (fails loudly when libc2d.so has not been built)

__all__ = ["ops", "unet", "vae", "sampler", "models"]
