"""C++ host mirror (include/vmb200.hpp): compiles and links against libvmb200.so on CPU; runs its reference-KAT program on
the GPU box (-m gpu)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror_test")


def _build():
    libdir = os.path.join(ROOT, "victoriametrics_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"), "-L" + libdir, "-l:libvmb200.so",
           "-Wl,-rpath," + libdir, "-o", EXE]
    subprocess.check_call(cmd)


def test_cpp_host_mirror_compiles_and_links():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_host_mirror_reference_kats_on_gpu():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host_mirror_test: OK" in r.stdout
