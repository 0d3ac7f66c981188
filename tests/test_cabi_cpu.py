"""CPU tests of the boundary: the C-ABI library loads without a GPU and exports every symbol that
include/fuzzy_match_b200.h declares; host-side argument checks fail loudly (no compute calls)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import fuzzy_match_b200 as fmb
from fuzzy_match_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    capi.build_library()
    return capi.load_library()


def test_every_declared_symbol_is_exported(lib):
    header = open(os.path.join(ROOT, "include", "fuzzy_match_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(fm_[a-z_0-9]+)\s*\(", header)))
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(capi.EXPORTS) == declared


def test_struct_layouts_match_header():
    assert C.sizeof(capi.Params) == 12 * 4
    assert capi.MATCH_DTYPE.itemsize == 24
    assert capi.WIRE_DTYPE.itemsize == 16


def test_version_and_error_strings(lib):
    assert b"sm_100a" in lib.fm_version()
    assert isinstance(lib.fm_last_error(), bytes)


def test_invalid_arguments_fail_before_touching_cuda(lib):
    tok = np.array([2, 3], dtype=np.int32)
    off = np.array([0, 2], dtype=np.int64)
    h = C.c_void_p()
    rc = lib.fm_index_create(C.c_void_p(tok.ctypes.data), C.c_void_p(off.ctypes.data), 1, 10, 5000, None, 0, 0, 0, C.byref(h))
    assert rc == 1 and b"max_tokens_in_pattern" in lib.fm_last_error()
    rc = lib.fm_index_create(C.c_void_p(tok.ctypes.data), C.c_void_p(off.ctypes.data), 1, 3, 300, None, 0, 0, 0, C.byref(h))
    assert rc == 1 and b"token id" in lib.fm_last_error()


def test_no_cpu_fallback_when_library_missing(monkeypatch):
    monkeypatch.setattr(capi, "_LIB", None)
    monkeypatch.setattr(capi, "library_path", lambda: "/nonexistent/libfm_b200.so")
    with pytest.raises(fmb.FuzzyMatchError):
        capi.load_library()


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "fuzzy_match_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".hh", ".cc")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU fallback", ""), os.path.join(dirpath, f)


def test_index_load_rejects_bad_files_without_a_gpu(lib, tmp_path):
    """fm_index_load validates magic, version, every header count and the block sizes before it allocates or touches
    CUDA: foreign, outdated, corrupt or truncated files give FM_ERR_INVALID and a message, never a crash."""
    import struct
    magic = bytes([ord(c) for c in "FMB200I"] + [1])

    def load(data):
        path = tmp_path / "bad.fmb"
        path.write_bytes(data)
        h = C.c_void_p()
        rc = lib.fm_index_load(str(path).encode(), 0, C.byref(h))
        assert h.value is None
        return rc, lib.fm_last_error()

    rc, msg = load(b"not an index at all" * 10)
    assert rc == 1 and b"not a fuzzy_match_b200 index" in msg
    rc, msg = load(magic + struct.pack("<16q", 3, 100, 300, 10, 50, 64, 1023, 1023, 0, 10, 0, 13, 0, 0, 0, 0))
    assert rc == 1 and b"version" in msg  # a file of an older layout
    version = int(re.search(rb"version (\d+)", msg).group(1))
    n_blk = 14
    # negative / inconsistent counts, absurd block sizes, a header with no payload behind it
    for hdr in ((version, 100, 300, -5, 50, 64, 1023, 1023, 0, 10, 0, n_blk, 1023, 0, 0, 0),
                (version, 100, 300, 10, 50, 64, 1000, 1023, 0, 10, 0, n_blk, 1023, 0, 0, 0),
                (version, 1 << 40, 300, 10, 50, 64, 1023, 1023, 0, 10, 0, n_blk, 1023, 0, 0, 0),
                (version, 100, 300, 10, 50, 64, 1023, 1023, 0, 10, 0, n_blk, 1023, 0, 0, 0)):
        rc, msg = load(magic + struct.pack("<16q", *hdr))
        assert rc == 1 and b"corrupt or truncated" in msg, (hdr, msg)
    rc, msg = load(magic + struct.pack("<16q", version, 100, 300, 10, 50, 64, 1023, 1023, 0, 10, 0, n_blk, 1023, 0, 0, 0)
                   + struct.pack("<q", 1 << 45) + b"\0" * 64)
    assert rc == 1 and b"corrupt or truncated" in msg


def test_library_is_sm_100a_only_and_uses_256_bit_memory_ops(lib):
    """The cubins in the library are built for sm_100a alone, and the Blackwell 256-bit global loads / stores the
    walk, verify, search and prepare kernels rely on made it into the SASS (LDG/STG.E.ENL2.256)."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    path = capi.library_path()
    elfs = subprocess.run(["cuobjdump", "--list-elf", path], capture_output=True, text=True).stdout.split("\n")
    archs = {m.group(1) for line in elfs for m in [re.search(r"\.(sm_\w+)\.cubin", line)] if m}
    assert archs == {"sm_100a"}, archs
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    assert sass.count("LDG.E.ENL2.256") >= 3 and sass.count("STG.E.ENL2.256") >= 1
