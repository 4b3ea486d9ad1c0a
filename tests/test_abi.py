"""CPU-only: the C-ABI library loads, exports every symbol include/vimz_gpu.h declares, and refuses to
compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import vimz_b200
from vimz_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "vimz_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(vimz_[a-z_A-Z0-9]+)\s*\(", hdr)))


def test_header_symbols_exported():
    names = declared_symbols()
    assert len(names) >= 30
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/vimz_gpu.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == names, "ctypes binding table and header disagree"


def test_version_and_error_string():
    assert _lib.lib.vimz_version() >= 100
    h = ctypes.c_void_p()
    rc = _lib.lib.vimz_ctx_create(99, 0, ctypes.byref(h))
    assert rc == _lib.VIMZ_ERR_ARG and b"curve" in _lib.lib.vimz_last_error()


def test_no_cpu_fallback_without_device():
    if _lib.lib.vimz_device_count() > 0:
        pytest.skip("a GPU is visible; the refusal path is exercised on the CPU box")
    with pytest.raises(vimz_b200.VimzError) as ei:
        vimz_b200.Engine("pallas", 0)
    assert ei.value.code == _lib.VIMZ_ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    """Only tests/, smoke() and bench.py's baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "vimz_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_accumulator_input_fast_paths():
    """Host side of the step loop (no GPU needed): small inputs may be raw little-endian bytes or (n, 4) uint64 arrays;
    anything else goes through the checked conversion, and wrong lengths are refused before the library is called."""
    import numpy as np
    from vimz_b200.nova import FoldAccumulator
    a = np.arange(8, dtype=np.uint64).reshape(2, 4)
    arr, ptr = FoldAccumulator._fr_ptr(a, 2)
    assert arr is a and ptr.value == a.ctypes.data                      # already in layout: no copy
    raw = a.tobytes()
    b, pb = FoldAccumulator._fr_ptr(raw, 2)
    assert b is raw and pb is raw                                       # bytes are handed to ctypes as they are
    with pytest.raises(ValueError):
        FoldAccumulator._fr_ptr(raw, 3)
    flat, pf = FoldAccumulator._fr_ptr(a.reshape(-1), 2)               # 1-D input is reshaped by the checked path
    assert flat.shape == (2, 4) and np.array_equal(flat, a)
    with pytest.raises(ValueError):
        FoldAccumulator._fr_ptr(np.zeros((3, 4), np.uint64), 2)


def test_dev_words_exposes_cuda_array_interface():
    from vimz_b200.sharding import _DevWords
    w = _DevWords(0x7F0000001000, 24)
    cai = w.__cuda_array_interface__
    assert cai["shape"] == (24,) and cai["typestr"] == "<i8" and cai["data"] == (0x7F0000001000, False)


def test_rust_sys_crate_binds_only_exported_symbols():
    """integration/vimz-gpu-sys (uncompiled source: no Rust toolchain here) must not declare anything the library lacks, and its
    argument COUNTS must match the header (a drifted FFI block would only show up at link / run time on the reference side)."""
    src = open(os.path.join(ROOT, "integration", "vimz-gpu-sys", "src", "lib.rs")).read()
    block = src[src.index('extern "C" {'):]
    block = block[:block.index("\n}\n")]
    decls = re.findall(r"fn (vimz_[a-zA-Z0-9_]+)\s*\((.*?)\)\s*(?:->|;)", block, flags=re.S)
    assert len(decls) >= 35
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "vimz_gpu.h")).read(), flags=re.S)
    raw = ctypes.CDLL(_lib.LIB_PATH)

    def nargs(arglist):
        arglist = re.sub(r"/\*.*?\*/", "", arglist, flags=re.S).strip()
        return 0 if arglist in ("", "void") else arglist.count(",") + 1

    for name, args in decls:
        assert hasattr(raw, name), f"{name} bound by the -sys crate but not exported"
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", hdr, flags=re.S)
        assert m, f"{name} not declared in include/vimz_gpu.h"
        assert nargs(m.group(1)) == nargs(args), f"{name}: {nargs(args)} arguments in lib.rs, {nargs(m.group(1))} in the header"
