"""CPU: the C-ABI library builds, loads and exports every symbol include/evavos.h declares."""
import ctypes
import os
import re

from tests.helpers import ROOT


def test_library_builds_and_exports_declared_symbols():
    from evavos_b200 import _lib, build
    build.build()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "evavos.h")).read()
    declared = set(re.findall(r"\b(evavos_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.evavos_abi_version() == _lib.ABI_VERSION
    assert lib.evavos_sizeof_bank_shadow() == ctypes.sizeof(_lib.BankShadow)
    assert lib.evavos_sizeof_memread_args() == ctypes.sizeof(_lib.MemReadArgs)
    assert lib.evavos_key_tiles_bytes(1) == _lib.TILE_BYTES
    assert lib.evavos_key_tiles_bytes(129) == 2 * _lib.TILE_BYTES


def test_argument_validation_without_gpu():
    """Pure host-side checks of the entry points (no kernels launched)."""
    from evavos_b200 import _lib
    lib = _lib.load()
    a = _lib.MemReadArgs()
    a.bank.CK, a.bank.K, a.bank.CV, a.bank.capacity_pos = 64, 1, 512, 48
    a.bank.key_pm = 0x1000
    a.query = 0x1000
    a.n_pos, a.n_query, a.top_k = 48, 48, 50
    assert lib.evavos_memread(ctypes.byref(a), None) == _lib.ERR_TOPK_RANGE
    assert b"out of range" in lib.evavos_last_error()      # prop_net.py:53 behaviour
    a.bank.CK = 60
    assert lib.evavos_memread(ctypes.byref(a), None) == -2
    assert lib.evavos_aggregate_wbg(None, None, 1, 10, 0, 0, None) == -1


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under evavos_b200/ may import or execute it."""
    pkg = os.path.join(ROOT, "evavos_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "/root/reference" not in src, f
