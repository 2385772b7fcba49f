"""ctypes binding of libc2d.so (the C ABI declared in include/c2d.h).

The library is the product: importing this module FAILS LOUDLY when it has not been built
(``python -c "import __graft_entry__ as g; g.build()"``) and ``ensure_init`` fails loudly without a
CUDA device -- there is no CPU or PyTorch fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("C2D_LIB") or os.path.join(_HERE, "libc2d.so")

OK, ERR_ARG, ERR_CUDA, ERR_UNSUPPORTED = 0, 1, 2, 3
F32, BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_SILU, ACT_RELU = 0, 1, 2, 3
IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05 = 0, 1, 2
AUDIO_NONE, AUDIO_ADD, AUDIO_CONCAT = 0, 1, 2


class C2DError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: the CUDA library has not been built. Run "
        "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). There is no fallback path.")

lib = C.CDLL(LIB_PATH)

_p, _i, _f, _ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong

# name -> argtypes (restype is int unless listed in _RESTYPES).  Mirrors include/c2d.h one to one;
# tests/test_abi.py checks every declaration in the header is bound here and exported by the .so.
SIGNATURES = {
    "c2d_abi_version": [],
    "c2d_last_error": [],
    "c2d_init": [_i],
    "c2d_destroy": [_i],
    "c2d_splitk_workspace_bytes": [],
    "c2d_set_workspace": [_i, _p, _ll],
    "c2d_launch_count": [],
    "c2d_last_kernel": [],
    "c2d_linear": [_p, _p, _p, _p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "c2d_geglu_linear": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "c2d_conv3x3": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "c2d_group_norm": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _i, _i, _p],
    "c2d_linear_ex": [_p, _p, _i, _i, _p, _p, _p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _i, _p, _p, _p, _f, _i, _p],
    "c2d_geglu_linear_ex": [_p, _p, _p, _p, _p, _f, _p, _i, _i, _i, _i, _p],
    "c2d_pack_lnfold": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p],
    "c2d_conv3x3_ex": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _i, _p],
    "c2d_conv3x3_down": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _i, _i, _p],
    "c2d_channel_stats": [_p, _p, _i, _i, _i, _i, _p],
    "c2d_group_norm_apply": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _i, _i, _p],
    "c2d_layer_norm": [_p, _p, _p, _p, _i, _i, _f, _i, _p],
    "c2d_attention": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _ll, _ll, _ll, _ll, _f, _p, _i, _i, _p],
    "c2d_attention_lse": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _ll, _ll, _ll, _ll, _f, _p, _p, _p, _i, _i, _p],
    "c2d_xattn_supported": [_i, _i, _i, _i, _i, _i],
    "c2d_xattn_packed_bytes": [_i, _i, _i, _i],
    "c2d_xattn_pack_kv": [_p, _p, _ll, _ll, _i, _p, _p, _ll, _ll, _i, _p, _i, _i, _i, _i, _p],
    "c2d_xattn_fwd": [_p, _ll, _p, _p, _p, _p, _f, _p, _i, _i, _f, _p, _ll, _i, _i, _i, _i, _f, _i, _p],
    "c2d_audio_context": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "c2d_timestep_embedding": [_p, _p, _i, _i, _p],
    "c2d_unary": [_p, _p, _ll, _i, _i, _i, _p],
    "c2d_add": [_p, _p, _p, _ll, _i, _p],
    "c2d_geglu": [_p, _p, _i, _i, _i, _p],
    "c2d_upsample2x": [_p, _p, _i, _i, _i, _i, _i, _p],
    "c2d_concat": [_p, _p, _p, _ll, _i, _i, _i, _p],
    "c2d_nchw_to_nhwc": [_p, _p, _i, _i, _i, _i, _p],
    "c2d_nhwc_to_nchw": [_p, _p, _i, _i, _i, _i, _p],
    "c2d_cfg_sched_step": [_p, _p, _p, _p, _i, _i, _f, _p, _i, _i, _p],
    "c2d_softmax_rows": [_p, _p, _i, _i, _f, _i, _p],
    "c2d_transpose": [_p, _p, _i, _i, _i, _i, _p],
    "c2d_bcast_add": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "c2d_hier_assign": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "c2d_hier_route": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "c2d_norm_scale": [_p, _p, _i, _i, _i, _f, _i, _i, _p],
    "c2d_legacy_combine": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "c2d_stft_frames": [_p, _p, _p, _i, _i, _i, _i, _i, _p],
    "c2d_stft_frames_split": [_p, _p, _p, _i, _i, _i, _i, _i, _p],
    "c2d_power_spectrum": [_p, _p, _ll, _i, _i, _i, _i, _p],
    "c2d_log_mel_affine": [_p, _p, _p, _p, _ll, _i, _f, _p],
    "c2d_clap_patches": [_p, _p, _i, _i, _i, _i, _p],
    "c2d_window_attention": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _i, _p],
    "c2d_patch_merge": [_p, _p, _i, _i, _i, _i, _i, _p],
    "c2d_token_mean": [_p, _p, _i, _i, _i, _i, _p],
    "c2d_l2_normalize": [_p, _p, _i, _i, _f, _p],
    "c2d_group_norm_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _i, _i, _p],
    "c2d_layer_norm_bwd": [_p, _p, _p, _p, _p, _i, _i, _f, _i, _p],
    "c2d_geglu_bwd": [_p, _p, _p, _i, _i, _i, _p],
    "c2d_attention_bwd": [_p] * 10 + [_i] * 5 + [_ll] * 16 + [_f, _i, _i, _p],
    "c2d_zero_insert2x": [_p, _p, _i, _i, _i, _i, _i, _p],
    "c2d_sumpool2x2": [_p, _p, _i, _i, _i, _i, _i, _p],
    "c2d_slice_channels": [_p, _p, _p, _ll, _i, _i, _i, _i, _p],
    "c2d_mse_loss_grad": [_p, _p, _p, _p, _i, _i, _i, _f, _i, _p],
    "c2d_colsum": [_p, _p, _i, _i, _i, _i, _i, _p],
    "c2d_gate_bwd": [_p, _p, _p, _p, _p, _i, _p],
    "c2d_gelu_bwd_bcast": [_p, _p, _p, _i, _i, _i, _p],
    "c2d_sumsq": [_p, _ll, _p, _p],
    "c2d_clip_scale": [_p, _f, _p, _p, _p],
    "c2d_adamw_step": [_p, _p, _p, _p, _ll, _f, _f, _f, _f, _f, _i, _p, _p],
    "c2d_adamw_step_sched": [_p, _p, _p, _p, _ll, _p, _i, _p, _f, _f, _f, _f, _p, _p],
    "c2d_pack_conv3x3": [_p, _p, _i, _i, _i, _p],
    "c2d_pack_geglu": [_p, _p, _p, _p, _i, _i, _i, _p],
}
_RESTYPES = {"c2d_xattn_packed_bytes": C.c_longlong, "c2d_splitk_workspace_bytes": C.c_longlong, "c2d_last_error": C.c_char_p, "c2d_last_kernel": C.c_char_p, "c2d_launch_count": C.c_ulonglong}

for _name, _args in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here = symbol missing from the .so
    _fn.argtypes = _args
    _fn.restype = _RESTYPES.get(_name, C.c_int)


def last_error() -> str:
    return (lib.c2d_last_error() or b"").decode(errors="replace")


def check(rc: int, what: str = "") -> None:
    if rc != OK:
        kind = {ERR_ARG: "bad argument", ERR_CUDA: "CUDA error", ERR_UNSUPPORTED: "unsupported"}.get(rc, f"rc={rc}")
        raise C2DError(f"libc2d {what}: {kind}: {last_error()}")


_inited = set()
_workspaces = {}          # device -> torch uint8 tensor registered with c2d_set_workspace (the CALLER owns the memory)


def ensure_init(device: int) -> None:
    """c2d_init once per device; raises C2DError when no sm_100 CUDA device is present.  Also allocates (through
    torch's allocator) and registers the split-K workspace: the library itself allocates nothing."""
    if device not in _inited:
        check(lib.c2d_init(int(device)), "c2d_init")
        _inited.add(device)
        import torch
        nbytes = int(lib.c2d_splitk_workspace_bytes())
        set_workspace(device, torch.empty(nbytes, device=f"cuda:{int(device)}", dtype=torch.uint8))


def set_workspace(device: int, ws) -> None:
    """Register a caller-owned uint8 CUDA tensor as the device's split-K workspace (None removes it)."""
    if ws is None:
        check(lib.c2d_set_workspace(int(device), None, 0), "c2d_set_workspace")
        _workspaces.pop(device, None)
        return
    assert ws.is_cuda and ws.is_contiguous() and ws.element_size() == 1
    check(lib.c2d_set_workspace(int(device), ws.data_ptr(), ws.numel()), "c2d_set_workspace")
    _workspaces[device] = ws


def destroy(device: int) -> None:
    """c2d_destroy + release of the workspace tensor; the next op on the device re-initialises."""
    check(lib.c2d_destroy(int(device)), "c2d_destroy")
    _workspaces.pop(device, None)
    _inited.discard(device)


def launch_count() -> int:
    return int(lib.c2d_launch_count())
