"""The C++ host mirror (include/sclgpu_scl.hpp) exercised from a C++ program that uses
SCL's own types and checks every result against SCL's own CPU functions in-process
(tests/cpp/test_shim.cc; built by __graft_entry__.build() where the reference tree
exists, the binary travels to the GPU box)."""
import os
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(REPO, "tests", "cpp", "_build", "test_shim")


def _need_binary():
    if not os.path.exists(BIN):
        pytest.skip("tests/cpp/_build/test_shim not built (needs the reference tree at build time)")


def test_shim_header_cites_every_batched_function():
    src = open(os.path.join(REPO, "include", "sclgpu_scl.hpp")).read()
    for name in ("shamirSecretShare", "shamirRecoverP", "shamirRecoverD", "randomVector", "multiplyEntryWise",
                 "scalarMultiply", "beaverCombine", "multiply("):
        assert name in src
    for cite in ("shamir.h:52-68", "shamir.h:82-104", "shamir.h:117-155", "vector.h:508-519", "matrix.h:498-513"):
        assert cite in src


def test_shim_fails_loudly_without_gpu():
    """No CPU fallback: without a device the C++ Context constructor throws."""
    _need_binary()
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0
    assert "no usable sm_100 device" in r.stderr


@pytest.mark.gpu
def test_cpp_shim_against_scl_in_process():
    assert os.path.exists(BIN), "tests/cpp/_build/test_shim must travel to the GPU box"
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "SHIM_OK" in r.stdout


C_BIN = os.path.join(REPO, "tests", "cpp", "_build", "c_example")


@pytest.mark.gpu
def test_plain_c_consumer_of_the_abi():
    """examples/share_reconstruct.c: the C ABI from plain C (gcc, no CUDA / C++ headers)."""
    assert os.path.exists(C_BIN), "tests/cpp/_build/c_example must travel to the GPU box"
    r = subprocess.run([C_BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "C_EXAMPLE_OK" in r.stdout, r.stdout + r.stderr
    assert "error detected during recovery" in r.stdout


def test_copy_team_of_the_pageable_stager():
    """csrc/host_stage.h: the copy-thread team (non-temporal stores, 0..7 helpers) on the CPU -- thousands of
    back-to-back copies of random sizes and alignments, every byte and both neighbours checked."""
    exe = os.path.join(REPO, "tests", "cpp", "_build", "test_copy_team")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/_build/test_copy_team not built (run __graft_entry__.build())")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "COPY_TEAM_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
