"""Runs the C++ tests of the MRPT-free host mirror of the reference's Matcher/Solver plugin
interface (tests/cpp/*.cpp read like the reference's tests/test-mp2p_*.cpp). They call the C ABI,
hence need the GPU."""
import gzip
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "cpp", "_build")


def _build():
    subprocess.run(["make", "-C", os.path.join(HERE, "cpp")], check=True, capture_output=True)


def test_cpp_tests_compile():
    """CPU-side: the host mirror header and the C-ABI header are consistent (compile + link)."""
    _build()
    for t in ("test_matcher_pt2pt", "test_matcher_pt2pl", "test_optimize_and_align", "test_host_logic"):
        assert os.path.exists(os.path.join(BUILD, t))


def test_cpp_host_logic_without_gpu():
    """The host logic above the C ABI — matcher / solver gating, run_matchers, Pairings, the ICP::align loop
    with every termination reason, quality evaluation and checkpoints, parameter errors — with mock
    plugin classes: no device, no library call (tests/cpp/test_host_logic.cpp)."""
    _build()
    r = subprocess.run([os.path.join(BUILD, "test_host_logic")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "test_host_logic OK" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["test_matcher_pt2pt", "test_matcher_pt2pl"])
def test_cpp_matchers(name):
    _build()
    r = subprocess.run([os.path.join(BUILD, name)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stdout


@pytest.mark.gpu
def test_cpp_optimize_and_align(tmp_path):
    _build()
    xyz = tmp_path / "bunny.xyz"
    with gzip.open(os.path.join(HERE, "golden", "bunny_decim.xyz.gz"), "rt") as f:
        xyz.write_text(f.read())
    r = subprocess.run([os.path.join(BUILD, "test_optimize_and_align"), str(xyz)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "test_optimize_and_align OK" in r.stdout
