"""examples/c_abi_divergence.c: a plain C program against include/dzb.h + libdzb.so.  It must compile and link everywhere;
on a GPU box it must run and find div (x, 2y, 3z) = 6.  (Named to run last: it shells out to a compiler.)"""
import os
import shutil
import subprocess

import pytest

from discretize_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build(tmp_path):
    if shutil.which("gcc") is None or not os.path.exists(os.path.join(CUDA, "include", "cuda_runtime_api.h")):
        pytest.skip("needs gcc and the CUDA runtime headers")
    exe = str(tmp_path / "c_abi_divergence")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
                    os.path.join(ROOT, "examples", "c_abi_divergence.c"), "-L", libdir, "-ldzb",
                    "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-lm", "-Wl,-rpath," + libdir,
                    "-Wl,-rpath," + os.path.join(CUDA, "lib64"), "-o", exe], check=True)
    return exe


def test_c_example_compiles_and_links(tmp_path):
    assert os.path.exists(build(tmp_path))


@pytest.mark.gpu
def test_c_example_runs_on_the_gpu(tmp_path):
    out = subprocess.run([build(tmp_path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "max |div u - 6|" in out.stdout
