"""The C++ host mirror (deeplearning_b200/host/dlb200.hpp) compiled with g++ against the C ABI and run on the GPU:
reference-style client code (append_layer<CLayer>(...), learn(...)) checked against the CPU oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(real: str) -> str:
    out_dir = os.path.join(ROOT, "tests", "cpp", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, f"host_mirror_test_{real}")
    src = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    cmd = ["g++", "-std=c++20", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
           "-L", os.path.join(ROOT, "deeplearning_b200"), "-ldlb200", "-L", os.path.join(ROOT, "oracle", "_build"), f"-loracle_{real}",
           f"-Wl,-rpath,{os.path.join(ROOT, 'deeplearning_b200')}", f"-Wl,-rpath,{os.path.join(ROOT, 'oracle', '_build')}"]
    if real == "f32":
        cmd.insert(1, "-DORACLE_SINGLE_PRECISION")
    subprocess.check_call(cmd)
    return exe


@pytest.mark.parametrize("real", ["f32", "f64"])
def test_cpp_mirror_compiles_without_cuda_headers(real):
    """CPU part: the mirror only needs the C ABI header; linking against libdlb200.so resolves every symbol."""
    assert os.path.exists(build(real))


@pytest.mark.gpu
@pytest.mark.parametrize("real", ["f32", "f64"])
def test_cpp_mirror_learn_matches_oracle(real):
    exe = build(real)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
