"""The library's REAL session host code on the CPU: rodio_b200/csrc/rb_api.cu is compiled as plain C++ against a mock of the
CUDA runtime (tests/emu/mockcuda/cuda_runtime.h: "device" memory is host memory), the launchers of the lane kernel are the
SIMT emulator (tests/emu/hostemu.cpp), and the Python mirror (rodio_b200.Session) drives the resulting
librodio_b200_hostemu.so through the unchanged C ABI -- class order, single and packed pushes, FIFO compaction, state blobs,
held / queued sources, speed, argument errors -- bit for bit against the oracle (tests/emu/session_scenarios.py, run in a
subprocess because the product loader caches its library).  CPU only; test infrastructure, never shipped."""
import os
import subprocess
import sys
import time

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EMU = os.path.join(HERE, "emu")
LIB = os.path.join(EMU, "librodio_b200_hostemu.so")
CSRC = os.path.join(ROOT, "rodio_b200", "csrc")
DEPS = [os.path.join(EMU, "hostemu.cpp"), os.path.join(EMU, "mockcuda", "cuda_runtime.h"), os.path.join(CSRC, "rb_api.cu")] + \
       [os.path.join(CSRC, f) for f in ("rb_lanes_core.h", "rb_lanes_plan.h", "rb_session_plan.h", "rb_simt.h", "rb_lanes.h", "rb_fused.h",
                                         "rb_internal.h")] + [os.path.join(ROOT, "include", "rodio_b200.h")]


@pytest.fixture(scope="module")
def hostemu(built):
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS):
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                               "-DRB_SIMT_EMULATE=1", "-I", os.path.join(EMU, "mockcuda"), "-x", "c++", os.path.join(CSRC, "rb_api.cu"),
                               os.path.join(EMU, "hostemu.cpp"), "-o", LIB])
    return LIB


@pytest.mark.parametrize("scenario", ["errors", "mono_random_split", "mixed_with_state_blob", "held_queue_gain_speed"])
def test_session_host_code_on_the_emulator(hostemu, scenario):
    t0 = time.time()
    r = subprocess.run([sys.executable, os.path.join(EMU, "session_scenarios.py"), scenario], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and f"ok {scenario}" in r.stdout, (r.stdout + r.stderr)[-3000:]
    print(f"{scenario}: {time.time() - t0:.1f} s")
