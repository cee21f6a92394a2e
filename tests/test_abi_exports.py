"""CPU: the C-ABI library is built, loads without a GPU, and exports every symbol include/oea.h declares."""
import ctypes
import os

from openea_b200 import lib as L


def test_header_symbols_all_bound_and_exported():
    declared = L.declared_symbols()
    assert declared, "no entry points parsed from include/oea.h"
    assert sorted(L.SIGNATURES) == declared, "openea_b200.lib.SIGNATURES must mirror include/oea.h exactly"
    assert os.path.exists(L.LIB_PATH), "liboea.so not built: run python -m openea_b200.build"
    raw = ctypes.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), "liboea.so does not export " + name


def test_load_and_error_strings_without_gpu():
    lib = L.load()
    assert lib.oea_abi_version() == 1
    assert lib.oea_error_string(0) == b"ok"
    assert b"NULL" in lib.oea_error_string(1)
    # argument validation happens before any CUDA call: NULL tables are rejected with OEA_ERR_NULL
    rc = lib.oea_rowopt_apply(None, None, None)
    assert rc == 1


def test_sm100a_only():
    """The shared object carries sm_100a SASS (no multi-arch fatbin)."""
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", L.LIB_PATH], capture_output=True, text=True).stdout
    archs = {line.split(".")[-2] for line in out.splitlines() if ".cubin" in line}
    assert archs == {"sm_100a"}, archs
