"""The C++ host mirror (include/rodio_b200.hpp): compiles on CPU, runs the reference's own mixer /
conversion unit tests on the GPU through the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_reference_api.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_reference_api.bin")


SRC_SESSION = os.path.join(ROOT, "tests", "cpp", "test_session_api.cpp")
EXE_SESSION = os.path.join(ROOT, "tests", "cpp", "test_session_api.bin")


def _build(src=SRC, exe=EXE):
    lib_dir = os.path.join(ROOT, "rodio_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
           "-L", lib_dir, "-l:librodio_b200.so", f"-Wl,-rpath,{lib_dir}"]
    subprocess.run(cmd, check=True, capture_output=True)


def test_cpp_mirror_compiles_and_links(built):
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_mirror_runs_reference_tests(built):
    if not os.path.exists(EXE):
        _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all reference API tests passed" in r.stdout


def test_cpp_session_mirror_compiles_and_links(built):
    _build(SRC_SESSION, EXE_SESSION)
    assert os.path.exists(EXE_SESSION)


@pytest.mark.gpu
def test_cpp_session_mirror_streams_like_the_whole_render(built):
    if not os.path.exists(EXE_SESSION):
        _build(SRC_SESSION, EXE_SESSION)
    r = subprocess.run([EXE_SESSION], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all session API tests passed" in r.stdout


@pytest.mark.parametrize("name", ["stream_mixer", "live_player"])
def test_plain_c_example_compiles_against_the_header(built, name):
    """examples/*.c: the boundary is usable from C (plain pointers and sizes, no C++ in the header)."""
    src = os.path.join(ROOT, "examples", name + ".c")
    exe = os.path.join(ROOT, "tests", "cpp", name + ".bin")
    lib_dir = os.path.join(ROOT, "rodio_b200")
    subprocess.run(["gcc", "-std=c11", "-D_GNU_SOURCE", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                    "-L", lib_dir, "-l:librodio_b200.so", f"-Wl,-rpath,{lib_dir}", "-lm"], check=True, capture_output=True)
    assert os.path.exists(exe)


SRC_COMM = os.path.join(ROOT, "tests", "cpp", "test_comm_two_gpus.cpp")
EXE_COMM = os.path.join(ROOT, "tests", "cpp", "test_comm_two_gpus.bin")


def test_comm_test_compiles_and_links(built):
    _build(SRC_COMM, EXE_COMM)
    assert os.path.exists(EXE_COMM)


@pytest.mark.gpu
def test_comm_allreduce_through_the_c_abi(built):
    """rb_comm_init_all + rb_batch_render_mix_allreduce: two GPUs driven by one process when the box has them (gpurun --gpus 2),
    a one-rank communicator otherwise (the same entry points, NCCL loaded, no exchange)."""
    import torch
    if not os.path.exists(EXE_COMM):
        _build(SRC_COMM, EXE_COMM)
    n = 2 if torch.cuda.device_count() >= 2 else 1
    r = subprocess.run([EXE_COMM, str(n)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all communicator tests passed" in r.stdout
