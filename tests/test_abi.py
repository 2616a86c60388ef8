"""CPU: libspk.so builds for sm_100a, loads, and exports every symbol include/spk.h declares (no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "spk.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(spk_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(spk_built):
    lib = ctypes.CDLL(spk_built)
    syms = declared_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), "libspk.so does not export " + s


def test_binding_covers_header(spk_built):
    from subphaser_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.spk_version() >= 100
    assert lib.spk_packed_words(4096) * 16 >= 4096 + 4096
    assert lib.spk_count_layout(10**9, 15) == 0 and lib.spk_count_layout(10**9, 21) == 1
    assert lib.spk_count_layout(2**30 - 1, 17) == 0 and lib.spk_count_layout(2**30, 17) == 1
    assert lib.spk_map_num_lines(25_000_000, 15, 10000, 10_000_000) == 2502


def test_host_side_planners(spk_built):
    """Pure host functions of the ABI (no device work): partition bits of the counter, layout of the K9 table."""
    from subphaser_b200 import _lib
    lib = _lib.load()
    # counter: <= 4096 k-mers per partition on average, never more partition bits than 2k or 22
    assert lib.spk_pcount_pbits(300_000_000, 17) == 17
    assert lib.spk_pcount_pbits(865_000_000, 17) == 18
    assert lib.spk_pcount_pbits(10_000, 17) == 2
    assert lib.spk_pcount_pbits(10**9, 5) == 10                       # tiny k: at most 4^k words
    assert lib.spk_pcount_pbits(2**32, 17) == -1                      # >= 2^32 bases: global-table path
    for nb in (1, 4096, 10**6, 10**9):
        p = lib.spk_pcount_pbits(nb, 21)
        assert 2 <= p <= 22 and (nb >> p) <= 4096
        assert lib.spk_pcount_workspace_bytes_ex(nb, 21, p) == lib.spk_pcount_workspace_bytes(nb, 21) > 0
        assert lib.spk_pcount_workspace_bytes_ex(nb, 21, 1) == 0      # fewer bits than the automatic choice: refused
    # K9 bucketed quotient table: 16-bit slots whenever remainder + subgenome id fit, mean load <= 1.75
    sb, bb = ctypes.c_int(), ctypes.c_int()
    assert lib.spk_qtable_plan(3_600_000, 17, 3, ctypes.byref(sb), ctypes.byref(bb)) == 0
    assert (sb.value, bb.value) == (16, 22) or (sb.value, bb.value) == (16, 21)
    assert 2 * 17 - bb.value + 2 <= 16 and 3_600_000 / (1 << bb.value) <= 1.75
    assert lib.spk_qtable_plan(1000, 15, 2, ctypes.byref(sb), ctypes.byref(bb)) == 0 and sb.value == 16
    assert 2 * 15 - bb.value + 2 <= 16
    assert lib.spk_qtable_plan(3_600_000, 21, 3, ctypes.byref(sb), ctypes.byref(bb)) == 0 and sb.value == 32
    assert lib.spk_qtable_plan(1000, 32, 3, ctypes.byref(sb), ctypes.byref(bb)) != 0   # 64-bit keys: open-addressed table
    assert lib.spk_map_num_lines(0, 17, 10000, 0) == 0


def test_sass_has_bulk_copy(spk_built):
    """The sequence tiles are staged with the TMA engine: cp.async.bulk shows up as UBLKCP in SASS."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", spk_built], capture_output=True, text=True).stdout
    assert "sm_100a" in sass or "SM100a" in sass.upper() or "sm_100" in sass
    assert "UBLKCP" in sass


def test_product_path_refuses_to_run_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from subphaser_b200 import _lib, engine
    with pytest.raises(_lib.SpkError):
        engine.require_cuda()
    from subphaser_b200 import Stats
    with pytest.raises(_lib.SpkError):
        Stats.fisher_test([1, 2], [10, 20])


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "subphaser_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "/root/reference" not in txt, f
