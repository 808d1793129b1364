"""The C++ host mirror (include/h2b200.hpp) compiles against the C ABI (CPU) and passes its checks on the GPU."""
import os
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
EXE = os.path.join(ROOT, "build", "host_mirror_test")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.join(ROOT, "halo2-lib_b200")
    subprocess.check_call([CXX, "-std=c++17", "-O1", "-Wall", SRC, "-o", EXE, f"-L{libdir}", "-lh2b200", f"-Wl,-rpath,{libdir}"])


def test_cpp_mirror_compiles_and_links():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_mirror_runs():
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all checks passed" in out.stdout
