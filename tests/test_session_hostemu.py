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


@pytest.fixture(scope="module")
def hostemu(built):
    sys.path[:0] = [EMU]
    import build_emu
    return build_emu.host_lib()


@pytest.mark.parametrize("scenario", ["errors", "mono_random_split", "mixed_with_state_blob", "held_queue_gain_speed", "follow_after_played_out", "held_across_state_blob", "skip_one", "duo_batches", "time_parallel_plan", "gain_changes",
                                      "filtered_and_plain", "batch_with_identity_conversions",
                                      "batch_unsorted_starts", "gain_in_front", "filter_in_front", "player_volume", "random:5"])
def test_session_host_code_on_the_emulator(hostemu, scenario):
    t0 = time.time()
    r = subprocess.run([sys.executable, os.path.join(EMU, "session_scenarios.py"), scenario], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and f"ok {scenario}" in r.stdout, (r.stdout + r.stderr)[-3000:]
    print(f"{scenario}: {time.time() - t0:.1f} s")


def test_lanes_batch_plan_on_the_emulator(hostemu):
    """rb_lanes_batch.cu (classes by rate pair and source channels, class-ordered rows, partial-row offsets, classification)
    compiled against the mock runtime: mono and stereo sources at four rates in a stereo mixer, one stream outside the input
    class -- bit for bit against the oracle streams summed class by class."""
    import ctypes as C

    import numpy as np

    sys.path[:0] = [HERE]
    import test_lanes_emulator as T
    from helpers import assert_bit_exact, noise

    ch_in = [1, 2, 1, 1, 2, 1, 2, 1] * 5
    rates = [44100, 44100, 48000, 22050, 48000, 44100, 32000, 48000] * 5
    pcms = [noise(ci * (250 + 5 * i), 4100 + i) for i, ci in enumerate(ch_in)]
    pcms[5][40:50] = np.float32(1e-41)
    starts = [(11 * i) % 70 for i in range(len(ch_in))]
    pres = [0.5 + 0.02 * i for i in range(len(ch_in))]
    pres[7] = 1000.0                                      # outside the fast tiles' gain range: slow tiles for its group
    c = T.make_case(pcms, rates, 48000, starts, lp=700, gain=0.9, channels=2, ch_in=ch_in, pre=pres)
    lib = C.CDLL(hostemu)
    n = len(pcms)
    arrs = [np.ascontiguousarray(p, dtype=np.float32) for p in pcms]
    ptrs = (C.POINTER(C.c_float) * n)(*[a.ctypes.data_as(C.POINTER(C.c_float)) for a in arrs])
    u64 = lambda v: (C.c_uint64 * n)(*[int(x) for x in v])
    u32 = lambda v: (C.c_uint32 * n)(*[int(x) for x in v])
    co = np.ascontiguousarray(c["coefs"], np.float32).reshape(-1)
    po = np.ascontiguousarray(c["posts"], np.float32)
    pr = np.ascontiguousarray(pres, np.float32)
    out = np.full(c["mix_len"] * 2, np.nan, np.float32)
    launches = C.c_uint32(0)
    rc = lib.hostemu_lanes_batch(ptrs, u64([a.size // ci for a, ci in zip(arrs, ch_in)]), u64(c["outs_len"]), u64(starts),
                                 co.ctypes.data_as(C.POINTER(C.c_float)), po.ctypes.data_as(C.POINTER(C.c_float)), pr.ctypes.data_as(C.POINTER(C.c_float)),
                                 C.c_uint32(n), C.c_uint32(2), u32(ch_in), u32(c["from_"]), u32(c["to"]), C.c_uint64(c["mix_len"]), 1, 1, 1,
                                 out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(launches))
    assert rc == 0
    assert launches.value == 6 + 1            # six classes (rate pair x source channels) and the sum
    want = T.expected_mix_classes(c["per_stream"], [s * 2 for s in starts], c["mix_len"] * 2, c["from_"], list(zip(c["to"], ch_in)))
    assert_bit_exact(out, want, "batch plan of the lane kernel")


def test_plain_c_example_runs_on_the_emulator(hostemu):
    """examples/stream_mixer.c linked against the host-emulated library: 60 ms of a filtered stereo 44.1 kHz source and a
    plain mono 48 kHz one through a stereo session, from C."""
    exe = os.path.join(HERE, "cpp", "stream_mixer_emu.bin")
    subprocess.run(["gcc", "-std=c11", "-D_GNU_SOURCE", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "stream_mixer.c"),
                    "-o", exe, "-L", EMU, "-l:" + os.path.basename(hostemu), f"-Wl,-rpath,{EMU}", "-lm"], check=True, capture_output=True)
    r = subprocess.run([exe, "6"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    frames, peak = int(r.stdout.split()[0]), float(r.stdout.split()[-1])
    assert frames == 2880 and 0.3 < peak < 0.9, r.stdout      # 6 x 10 ms at 48 kHz; 0.8 * 0.5 low-passed music + 0.25 voice


def test_live_player_example_runs_on_the_emulator(hostemu):
    """examples/live_player.c: a queue of two sounds (the first low-passed at its own rate), a pause, a volume change -- the
    Player controls on a session, from C, against the host-emulated library."""
    exe = os.path.join(HERE, "cpp", "live_player_emu.bin")
    subprocess.run(["gcc", "-std=c11", "-D_GNU_SOURCE", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "live_player.c"),
                    "-o", exe, "-L", EMU, "-l:" + os.path.basename(hostemu), f"-Wl,-rpath,{EMU}", "-lm"], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    w = r.stdout.replace(",", " ").split()
    frames, silent, loud, quiet = int(w[0]), int(w[2]), float(w[5]), float(w[7])
    # 4400 frames of 22.05 kHz (9578 at 48 kHz), then (30 + 5) x 441 frames of 44.1 kHz (16800); 5 x 10 ms of pause; 0.6 -> 0.6 * 0.25
    assert frames == 9578 + 16800 and 2400 <= silent <= 2420 and abs(loud - 0.6) < 0.01 and abs(quiet - 0.15) < 0.005, r.stdout


def test_gpu_lane_and_session_tests_hold_on_the_emulator(hostemu):
    """Every GPU test of the lane kernel and of the sessions that needs nothing but the C ABI and the oracle, run unchanged
    against the host-emulated library -- their expectations are known to be right before they meet a device.  (Left out: the
    two tests that compare with the default kernels, which exist on the device only.)"""
    code = ("import os, sys; sys.path[:0] = [%r, %r]; import rodio_b200._capi as c; c.LIB_PATH = %r; import pytest; "
            "sys.exit(pytest.main([%r, '-q', '-m', 'gpu', '-x', '-p', 'no:cacheprovider', '-k', "
            "'(test_lanes or test_session) and not fallback and not full_size']))"
            % (ROOT, HERE, hostemu, os.path.join(HERE, "test_parity_gpu.py")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and " passed" in r.stdout and "failed" not in r.stdout, (r.stdout + r.stderr)[-3000:]
    assert int(r.stdout.rsplit(" passed", 1)[0].split()[-1]) >= 31, r.stdout[-500:]


def test_cpp_live_mixer_on_the_emulator(hostemu):
    """tests/cpp/test_session_api.cpp (rodio::mixer::LiveMixer of include/rodio_b200.hpp: uneven pushes, sample-by-sample pull,
    mono and stereo, against the whole-stream lane render) linked against the host-emulated library."""
    exe = os.path.join(HERE, "cpp", "test_session_api_emu.bin")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(HERE, "cpp", "test_session_api.cpp"), "-o", exe,
                    "-L", EMU, "-l:" + os.path.basename(hostemu), f"-Wl,-rpath,{EMU}"], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "all session API tests passed" in r.stdout, r.stdout + r.stderr
