"""The C++ host-side mirror (include/strumpack_b200/StructuredMatrix.hpp) of the
reference interface compiles against the C ABI; on a GPU the C++ program that
restates test/test_HSS_seq.cpp's Toeplitz ULV check passes."""
import os
import subprocess

import pytest

from conftest import ROOT


def _build(tmp_path, built):
    exe = str(tmp_path / "test_structured")
    so_dir = os.path.join(ROOT, "strumpack_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_structured.cpp"), "-o", exe,
                           "-L" + so_dir, "-lstrumpack_b200", "-Wl,-rpath," + so_dir])
    return exe


def test_cpp_mirror_compiles_and_fails_loudly_without_gpu(built, tmp_path):
    import torch
    exe = _build(tmp_path, built)
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([exe, "64"], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_cpp_mirror_toeplitz_ulv(built, tmp_path):
    exe = _build(tmp_path, built)
    r = subprocess.run([exe, "1000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "exiting" in r.stdout
