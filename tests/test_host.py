"""CPU tests of the host side: the C-ABI library loads, exports every symbol
include/sb200_structured.h declares, parses the reference's HSS dump format,
reproduces the reference's flop counters, and refuses to compute without a
GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from conftest import CASES, GOLDEN, ROOT
from oracle import hss_file


def test_library_exports_every_declared_symbol(built):
    sb = built
    hdr = open(os.path.join(ROOT, "include", "sb200_structured.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b((?:SP_d|SP_s|SB200)_[A-Za-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    L = sb.lib()
    for name in declared:
        assert hasattr(L, name), name
    # the Python mirror binds exactly the declared set
    assert declared == set(sb.SYMBOLS)
    assert b"sm_100a" in L.SB200_version()


def test_default_options_match_reference(built):
    o = built.default_options(type=built.SP_TYPE_BLR)
    # reference StructuredOptions.hpp:106-162
    assert (o.rel_tol, o.abs_tol, o.leaf_size, o.max_rank) == (1e-4, 1e-10, 128, 5000)


@pytest.mark.parametrize("case", CASES)
def test_hss_file_parse_and_flop_model(built, case):
    sb = built
    path = os.path.join(GOLDEN, case + ".hss")
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    inf = sb.hss_file_info(path)
    nodes, _ = hss_file.read_hss(path)
    assert inf["rows"] == g["info"][0] and inf["cols"] == g["info"][1]
    assert inf["rank"] == g["info"][2] and inf["levels"] == g["info"][3]
    assert inf["nodes"] == len(nodes)
    # flop model == the reference's own counters (params::ULV_factor_flops,
    # params::hss_solve_flops for 3 right-hand sides), exactly
    assert inf["factor_flops"] == g["flops"][0]
    assert 3 * inf["solve_flops"] == g["flops"][1]
    assert 0 < inf["factor_flops_exec"] < inf["factor_flops"]


def test_hss_file_roundtrip(built, tmp_path):
    sb = built
    src = os.path.join(GOLDEN, CASES[1] + ".hss")
    dst = tmp_path / "copy.hss"
    sb.hss_file_copy(src, dst)
    a, _ = hss_file.read_hss(src)
    b, _ = hss_file.read_hss(dst)
    assert len(a) == len(b)
    for x, y in zip(a, b):
        for k in ("D", "Eu", "Ev", "B01", "B10"):
            assert np.array_equal(getattr(x, k), getattr(y, k)), k
        # permutations are equivalent as gathers (ipiv is not unique)
        assert np.array_equal(hss_file.ipiv_to_gather(x.Pu), hss_file.ipiv_to_gather(y.Pu))
        assert np.array_equal(hss_file.ipiv_to_gather(x.Pv), hss_file.ipiv_to_gather(y.Pv))


def test_bad_file_is_an_error_not_a_crash(built, tmp_path, capfd):
    p = tmp_path / "junk.hss"
    p.write_bytes(b"\x00" * 10)
    with pytest.raises(RuntimeError):
        built.hss_file_info(p)
    assert "Operation failed" in capfd.readouterr().err


def test_no_cpu_fallback(built, capfd):
    """Without a GPU every compute entry point must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        built.HSSMatrix.read(os.path.join(GOLDEN, CASES[0] + ".hss"))
    assert "no CUDA device" in capfd.readouterr().err


def test_truncated_and_corrupt_hss_files_are_rejected(built, tmp_path, capfd):
    """HSSHost::read_file checks every size field against what is left of the
    file before allocating (round-1 advisor finding): a truncated dump, a dump
    with a huge Psize and one with an out-of-range pivot fail with an error
    instead of a giant allocation / garbage generators."""
    import struct
    sb = built
    src = os.path.join(GOLDEN, "toeplitz_512_leaf64.hss")
    raw = open(src, "rb").read()
    assert sb.hss_file_info(src)["rows"] == 512
    for cut in (7, 40, len(raw) // 3, len(raw) - 5):
        p = tmp_path / f"cut{cut}.hss"
        p.write_bytes(raw[:cut])
        with pytest.raises(RuntimeError):
            sb.hss_file_info(str(p))
    # root record: 12 (version) + 16 (rows, cols) + 2 + 4 + 1 + 16 (ranks) = 51, then the
    # Asub DenseMatrix record (12 + 40 bytes, empty), then U's Psize (u64)
    off_psize = 51 + 52
    assert struct.unpack_from("<Q", raw, off_psize)[0] == 0        # root has no basis
    bad = bytearray(raw)
    struct.pack_into("<Q", bad, off_psize, 1 << 40)                # absurd permutation length
    p = tmp_path / "psize.hss"
    p.write_bytes(bytes(bad))
    with pytest.raises(RuntimeError):
        sb.hss_file_info(str(p))
    bad = bytearray(raw)
    struct.pack_into("<Q", bad, 12, 1 << 40)                       # absurd row count
    p = tmp_path / "rows.hss"
    p.write_bytes(bytes(bad))
    with pytest.raises(RuntimeError):
        sb.hss_file_info(str(p))
    assert "Operation failed" in capfd.readouterr().err
