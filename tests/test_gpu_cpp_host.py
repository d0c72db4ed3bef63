"""GPU: the C++ host layer (include/rwgpu_executor.hpp -- the C++ mirror of Execute / Message /
MockSource / HashAggExecutor / HashJoinExecutor above the C ABI) replays reference golden tests."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_host_layer_kats():
    exe = os.path.join(ROOT, "build", "test_executor_kats")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all C++ host-layer KATs passed" in r.stdout


def test_cpp_host_layer_compiles():
    """CPU: the header-only host layer compiles against include/rwgpu.h (syntax / API drift check)."""
    src = os.path.join(ROOT, "tests", "cpp", "test_executor_kats.cc")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"), src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
