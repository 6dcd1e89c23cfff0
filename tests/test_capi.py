"""The C-ABI library loads and exports every symbol include/fugue_b200.h declares
(no compute calls: there is no GPU in the build container)."""
import os
import re

import numpy as np
import pytest

from fugue_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "fugue_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fugue_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in fugue_b200/_lib.py"
    assert lib.fb_abi_version() == 1


def test_fastmod_matches_modulo():
    lib = _lib.load()
    rng = np.random.default_rng(1)
    hs = [0, 1, 2**32 - 1, 2**32, 2**63, 2**64 - 1] + [int(x) for x in rng.integers(0, 2**63, 200)] + \
         [int(x) + 2**63 for x in rng.integers(0, 2**63, 200)]
    nums = [1, 2, 3, 5, 7, 64, 255, 256, 257, 1000, 1024, 65535, 65536, 2**31 - 1, 2**31, 2**32 - 1]
    for d in nums:
        for h in hs:
            assert lib.fb_debug_fastmod_host(h, d) == h % d, (h, d)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libfugue_b200.so")
    with pytest.raises(_lib.FugueB200LibraryError):
        _lib.load()
