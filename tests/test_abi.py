"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/cnn_b200.h declares, and refuses to run without a device (no CPU fallback)."""
import ctypes as C
import os
import subprocess

import pytest

from cnn_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


def test_header_symbols_exported(lib):
    names = _lib.header_symbols()
    assert len(names) >= 50
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cnn_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_header_is_plain_c():
    """The boundary must be consumable from C (cgo / JNI / ctypes style bindings)."""
    src = '#include "cnn_b200.h"\nint main(void){ return cnn_version() != 0 ? 0 : 1; }\n'
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I",
                        os.path.join(_lib.ROOT, "include"), "-x", "c", "-"], input=src.encode(),
                       capture_output=True)
    assert r.returncode == 0, r.stderr.decode()


def test_no_silent_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.cnn_ctx_create(0, None, C.byref(h))
    assert rc == -2 and b"no CPU fallback" in lib.cnn_last_error()
    from cnn_b200.api import Context
    with pytest.raises(_lib.CnnError):
        Context()


def test_sass_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = {ln.split(".")[-2] for ln in out.splitlines() if ".cubin" in ln}
    assert archs == {"sm_100a"}, archs
