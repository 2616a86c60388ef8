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
