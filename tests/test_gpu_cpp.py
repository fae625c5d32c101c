"""Runs the C++ host-mirror parity test (tests/cpp/host_mirror_test.cpp) on the GPU."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
def test_cpp_host_mirror_parity():
    d = os.path.join(HERE, "cpp")
    subprocess.run(["make", "-s", "-C", d], check=True)
    out = subprocess.run([os.path.join(d, "host_mirror_test")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ALL OK" in out.stdout
