"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/k2b200.h declares.
No compute is attempted here (no GPU in this container)."""
import ctypes
import subprocess

import pytest

from k2transducerasr_b200 import _native


def test_library_builds_and_exports_header_symbols(built_lib):
    assert built_lib.exists()
    lib = ctypes.CDLL(str(built_lib))
    declared = _native.header_symbols()
    assert len(declared) >= 26
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"declared in k2b200.h but not exported: {missing}"
    assert set(declared) == set(_native._SIGS), "ctypes binding and header disagree"
    assert lib.k2b_abi_version() == 2


def test_only_the_abi_is_exported(built_lib):
    out = subprocess.run(["nm", "-D", "--defined-only", str(built_lib)], capture_output=True, text=True).stdout
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert syms and all(s.startswith("k2b_") for s in syms), syms


def test_sass_is_sm100a(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", str(built_lib)], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_config_struct_layout():
    assert ctypes.sizeof(_native.K2bConfig) == 16 * 4


def test_create_fails_loudly_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_native.K2bError) as e:
        _native.Handle(vocab_size=500)
    assert "no CPU fallback" in str(e.value)


def test_create_rejects_bad_config(built_lib):
    lib = _native.lib()
    cfg = _native.K2bConfig(7, 0, 500, 512, 512, 0, 2, 0, 1, 2, 0, 0, 4, 0, 0, 0)   # wrong struct_size
    h = ctypes.c_void_p()
    assert lib.k2b_create(ctypes.byref(cfg), ctypes.byref(h)) == _native.K2B_ERR_INVALID
    assert b"struct_size" in lib.k2b_last_error(None)
    cfg = _native.K2bConfig(64, 0, 500, 512, 512, 0, 3, 0, 1, 2, 0, 0, 4, 0, 0, 0)  # context_size 3 (Q9)
    assert lib.k2b_create(ctypes.byref(cfg), ctypes.byref(h)) == _native.K2B_ERR_UNSUPPORTED
    assert lib.k2b_destroy(None) == 0
