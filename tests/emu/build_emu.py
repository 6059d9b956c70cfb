"""Builds the two CPU harnesses of the lane kernel (test infrastructure): liblanes_emu.so (the warp program alone) and
librodio_b200_hostemu.so (the library's host code over the mock CUDA runtime).  The instantiations of the warp program are
compiled as thirteen objects in parallel (warp_variants.cpp) and shared by both; everything is rebuilt only when a source
it is made from is newer.    python tests/emu/build_emu.py   builds both."""
import os
import subprocess
import sys

EMU = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(EMU))
CSRC = os.path.join(ROOT, "rodio_b200", "csrc")
OBJ = os.path.join(EMU, "obj")
CXX = ["g++", "-std=c++20", "-O1", "-pthread", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-DRB_SIMT_EMULATE=1"]
# Geometry variants of the kernel (the A/B knobs of rb_lanes_core.h) get libraries of their own, so the whole emulator suite can
# be run against them:   RB_EMU_VARIANT=slots3 python -m pytest tests/test_lanes_emulator.py tests/test_session_hostemu.py
VARIANTS = {"": [], "slots3": ["-DRB_LANES_UP_SLOTS=3"], "stereo32": ["-DRB_LANES_STEREO_CHW=32"],
            "slots3_stereo32": ["-DRB_LANES_UP_SLOTS=3", "-DRB_LANES_STEREO_CHW=32"]}
VARIANT = os.environ.get("RB_EMU_VARIANT", "")
CXX += VARIANTS[VARIANT]
_SFX = ("_" + VARIANT) if VARIANT else ""
KERNEL_DEPS = [os.path.join(CSRC, f) for f in ("rb_lanes_core.h", "rb_duo_core.h", "rb_lanes_plan.h", "rb_simt.h")]
VARIANT_DEPS = KERNEL_DEPS + [os.path.join(EMU, "warp_variants.cpp"), os.path.join(EMU, "warp_variants.h")]
LANES_LIB = os.path.join(EMU, f"liblanes_emu{_SFX}.so")
HOST_LIB = os.path.join(EMU, f"librodio_b200_hostemu{_SFX}.so")
HOST_DEPS = [os.path.join(EMU, "hostemu.cpp"), os.path.join(EMU, "mockcuda", "cuda_runtime.h"), os.path.join(EMU, "warp_variants.h"),
             os.path.join(CSRC, "rb_api.cu"), os.path.join(CSRC, "rb_lanes_batch.cu"), os.path.join(ROOT, "include", "rodio_b200.h")] + KERNEL_DEPS + \
            [os.path.join(CSRC, f) for f in ("rb_session_plan.h", "rb_lanes.h", "rb_fused.h", "rb_fused_rows.h", "rb_internal.h")]


def _stale(target, deps):
    return not os.path.exists(target) or any(os.path.getmtime(d) > os.path.getmtime(target) for d in deps)


def _run_all(cmds):
    procs = [(c, subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)) for c in cmds]
    for c, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("failed: " + " ".join(c))


def variant_objects():
    os.makedirs(OBJ, exist_ok=True)
    objs = [os.path.join(OBJ, f"warp_variants{_SFX}_{k}.o") for k in range(13)]
    _run_all([CXX + ["-c", f"-DRB_EMU_PART={k}", os.path.join(EMU, "warp_variants.cpp"), "-o", o]
              for k, o in enumerate(objs) if _stale(o, VARIANT_DEPS)])
    return objs


def lanes_lib():
    objs = variant_objects()
    src = os.path.join(EMU, "lanes_emu.cpp")
    if _stale(LANES_LIB, objs + [src, os.path.join(CSRC, "rb_session_plan.h"), os.path.join(EMU, "warp_variants.h")] + KERNEL_DEPS):
        subprocess.check_call(CXX + ["-shared", src] + objs + ["-o", LANES_LIB])
    return LANES_LIB


def host_lib():
    objs = variant_objects()
    if _stale(HOST_LIB, objs + HOST_DEPS):
        units = [os.path.join(CSRC, "rb_api.cu"), os.path.join(CSRC, "rb_lanes_batch.cu"), os.path.join(EMU, "hostemu.cpp")]
        uobjs = [os.path.join(OBJ, "host" + _SFX + "_" + os.path.basename(u).split(".")[0] + ".o") for u in units]
        _run_all([CXX + ["-I", os.path.join(EMU, "mockcuda"), "-x", "c++", "-c", u, "-o", o] for u, o in zip(units, uobjs)])
        subprocess.check_call(CXX + ["-shared"] + uobjs + objs + ["-o", HOST_LIB])
    return HOST_LIB


if __name__ == "__main__":
    print(lanes_lib())
    print(host_lib())
