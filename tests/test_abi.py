"""CPU-only: the C-ABI library builds, loads and exports exactly what include/tmx.h declares,
and the product path fails loudly (no fallback) when used without a GPU."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from tweediemix_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared():
    src = open(os.path.join(ROOT, "include", "tmx.h")).read()
    return sorted(set(re.findall(r"TMX_API[^;(]*?\b(tmx_\w+)\s*\(", src)))


def test_header_symbols_all_exported(lib):
    from tweediemix_b200 import _lib
    names = _declared()
    assert len(names) >= 11
    assert sorted(_lib.SIGNATURES) == names            # the ctypes table mirrors the header
    for n in names:
        assert hasattr(lib, n), f"{n} not exported by libtmx.so"


def test_version_and_error_string(lib):
    src = open(os.path.join(ROOT, "include", "tmx.h")).read()
    assert lib.tmx_version() == int(re.search(r"#define TMX_VERSION (\d+)", src).group(1))
    assert isinstance(lib.tmx_last_error(), bytes)


def test_argument_validation_without_gpu(lib):
    # argument checks run before any CUDA call, so they are testable on the CPU box
    rc = lib.tmx_tweedie_blend_ddim_fwd(None, None, None, None, None, None, 1, 3, 4, 64, 0.5, 0.6, 0.8, 0, 1, 0, None)
    assert rc == -1 and b"null pointer" in lib.tmx_last_error()
    rc = lib.tmx_resadd_fwd(16, 16, 16, 7, 1.0, 2, None)
    assert rc == -2 and b"multiple of 8" in lib.tmx_last_error()
    rc = lib.tmx_groupnorm_fwd(16, 16, 16, None, 16, 16, 1, 30, 64, 32, 1e-5, 1, 1, 2, None)
    assert rc == -2
    assert lib.tmx_groupnorm_workspace_bytes(4, 320, 16384, 32, 1) > 0


def test_no_cpu_fallback():
    from tweediemix_b200 import ops
    x = torch.zeros(1, 4, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.tweedie_blend_ddim(x, torch.zeros(4, 4, 8, 8), torch.ones(3, 1, 8, 8), 0.5, 0.6, 0.8)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.group_norm(torch.zeros(1, 32, 4, 4), torch.ones(32), torch.zeros(32), 32, 1e-5)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure; nothing under tweediemix_b200/ may import it."""
    pkg = os.path.join(ROOT, "tweediemix_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
