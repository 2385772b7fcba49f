"""The C-ABI library loads, exports every symbol include/c2d.h declares, and refuses to run without CUDA."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "c2d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(c2d_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from clap2diffusion_b200 import _lib
    names = _header_functions()
    assert len(names) >= 25
    so = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(so, n), f"{n} declared in c2d.h but not exported by libc2d.so"
        assert n in _lib.SIGNATURES, f"{n} declared in c2d.h but not bound in _lib.SIGNATURES"
    assert set(_lib.SIGNATURES) == set(names), set(_lib.SIGNATURES) ^ set(names)


def test_abi_version_and_error_string():
    from clap2diffusion_b200 import _lib
    assert _lib.lib.c2d_abi_version() == 7
    assert isinstance(_lib.last_error(), str)


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly (no silent torch / CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from clap2diffusion_b200 import _lib, ops
    with pytest.raises(_lib.C2DError):
        _lib.ensure_init(0)
    with pytest.raises(_lib.C2DError):
        ops.linear(torch.zeros(4, 8), torch.zeros(8, 8))
    from clap2diffusion_b200.models.audio_adapter_v4 import AudioAdapter
    with pytest.raises(_lib.C2DError):
        AudioAdapter()(torch.zeros(1, 512))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "clap2diffusion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"


def test_only_tests_smoke_and_bench_touch_the_oracle():
    """oracle/ is test infrastructure: nothing under the package, scripts/, app/ or tools/ imports it."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    offenders = []
    for sub in ("clap2diffusion_b200", "scripts", "app", "tools"):
        for dirpath, _, files in os.walk(os.path.join(root, sub)):
            for f in files:
                if f.endswith(".py"):
                    src = open(os.path.join(dirpath, f)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", src, re.M):
                        offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders
