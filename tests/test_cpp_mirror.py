"""The C++ host-side mirror (include/strumpack_b200/StructuredMatrix.hpp) of the
reference interface compiles against the C ABI; on a GPU the C++ programs that
restate test/test_HSS_seq.cpp's Toeplitz ULV check and walk the HSSMatrix /
BLRMatrix / options class surface (including the front operations FrontHSS and
FrontBLR use) pass."""
import os
import subprocess

import pytest

from conftest import ROOT

PROGRAMS = ["test_structured", "test_classes", "test_front_sequence"]


def _build(tmp_path, name):
    exe = str(tmp_path / name)
    so_dir = os.path.join(ROOT, "strumpack_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-o", exe,
                           "-L" + so_dir, "-lstrumpack_b200", "-Wl,-rpath," + so_dir])
    return exe


@pytest.mark.parametrize("name", PROGRAMS)
def test_cpp_mirror_compiles_and_fails_loudly_without_gpu(built, tmp_path, name):
    import torch
    exe = _build(tmp_path, name)
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([exe, "64"], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CUDA device" in r.stderr


def test_cpp_mirror_options_and_dense_helpers(built, tmp_path):
    """CPU-only: option defaults / command-line parsing / DenseMatrix helpers of the mirror."""
    exe = _build(tmp_path, "test_options")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "options ok" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_toeplitz_ulv(built, tmp_path):
    exe = _build(tmp_path, "test_structured")
    r = subprocess.run([exe, "1000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "exiting" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_class_surface(built, tmp_path):
    exe = _build(tmp_path, "test_classes")
    r = subprocess.run([exe, "600"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "exiting" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("sizes", [("384", "256"), ("300", "217")])
def test_cpp_mirror_front_hss_call_sequence(built, tmp_path, sizes):
    """The statements of FrontHSS::multifrontal_factorization / fwd_solve_node / bwd_solve_node
    (reference src/sparse/fronts/FrontHSS.cpp:371-412, 445-496) compile against the mirror and
    solve the front's system."""
    exe = _build(tmp_path, "test_front_sequence")
    r = subprocess.run([exe, *sizes], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "exiting" in r.stdout
