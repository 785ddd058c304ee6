"""CPU check of the tensor-core operand producers (deeplearning_b200/csrc/tc_tiles.cuh): the index arithmetic that places
dense rows / columns, receptive-field patches (both orientations, padding, strides, mirrored window, sample gather) and
[b][f][pixel] gradients into the K-major SWIZZLE_128B tiles tcgen05.mma reads is compiled with g++ and every producer
thread is run on the host (tests/cpp/tc_tiles_test.cpp).  The GPU tests then only have to prove the MMA / epilogue side."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_operand_tiles_match_their_definition():
    out_dir = os.path.join(ROOT, "tests", "cpp", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "tc_tiles_test")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-Wno-unknown-pragmas", os.path.join(ROOT, "tests", "cpp", "tc_tiles_test.cpp"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
