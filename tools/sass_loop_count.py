#!/usr/bin/env python
"""Static look at the hot loop of k_fused_lanes (no GPU needed): compile rb_lanes.cu for sm_100a, disassemble, find the
steady-state loop of the <biquad, FF2, one gain> instantiation (the backward branch whose body holds the butterfly
shuffles of the mixer sum) and print its instruction mix per tile of 8 samples x 32 lanes.  The refill block (LDGSTS)
sits inside the loop but is branched over on most iterations; it is listed separately.
    python tools/sass_loop_count.py [mangled-name-fragment]
    default ILi1ELi1ELb1ELb1ELi1ELb0E = <CI 1, CO 1, biquad, FF2, one gain, interpolating> (the first match is the variant
    without PRE / FRONT / DOWN: ...Lb0ELb0ELb0E); stereo: ILi2ELi2ELb1ELb1ELi1ELb0E; gain in front: ...Lb0ELb1ELb0ELb0E; filter in
    front: ILi1ELi1ELb1ELb0ELi1ELb0ELb0ELb1ELb0E; sources above the mixer's rate: ILi1ELi1ELb1ELb1ELi1ELb0ELb0ELb0ELb1E"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "rodio_b200", "csrc", "rb_lanes.cu")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--fmad=false", "-Xcompiler",
         "-fPIC,-ffp-contract=off,-fno-fast-math"]


def main():
    frag = sys.argv[1] if len(sys.argv) > 1 else "ILi1ELi1ELb1ELb1ELi1ELb0E"
    kernel = sys.argv[2] if len(sys.argv) > 2 else "k_fused_lanes"   # k_fused_duo ILb1ELb1ELi1E: the lane-pair kernel
    per = 16 if kernel == "k_fused_duo" else 8                       # samples per lane and tile
    with tempfile.TemporaryDirectory() as td:
        obj = os.path.join(td, "rb_lanes.o")
        r = subprocess.run(["nvcc"] + FLAGS + ["-Xptxas", "-v", "-c", SRC, "-o", obj], capture_output=True, text=True, check=True)
        regs = [l for l in (r.stdout + r.stderr).splitlines() if "registers" in l or "Compiling entry" in l]
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)
    body = next(f for f in funcs if kernel in f.split("\n", 1)[0] and frag in f.split("\n", 1)[0])
    ins = []   # (address, predicate, opcode, text)
    for line in body.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4})\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_.]+)(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), (m.group(2) or "").strip(), m.group(3), m.group(4)))
    addr = {a: i for i, (a, _, _, _) in enumerate(ins)}
    best = None
    for i, (a, _, op, rest) in enumerate(ins):
        if op.startswith("BRA"):
            t = re.search(r"0x([0-9a-f]+)", rest)
            if t and int(t.group(1), 16) < a and int(t.group(1), 16) in addr:
                lo = addr[int(t.group(1), 16)]
                seg = ins[lo:i + 1]
                if any(o.startswith("SHFL.BFLY") for _, _, o, _ in seg) and any(o == "LDS" and p for _, p, o, _ in seg):
                    if best is None or len(seg) < len(best):
                        best = seg
    assert best, "hot loop not found"
    # the refill block: from the DEPBAR (cp.async.wait_group) to the LDGDEPBAR (commit) that follows the LDGSTS
    idx_dep = next(i for i, x in enumerate(best) if x[2].startswith("DEPBAR"))
    idx_end = max(i for i, x in enumerate(best) if x[2].startswith("LDGDEPBAR"))
    refill = best[idx_dep:idx_end + 1]
    steady = best[:idx_dep] + best[idx_end + 1:]
    hist = collections.Counter(o.split(".")[0] for _, _, o, _ in steady)
    n = len(steady)
    for l in regs:
        if frag in l or "registers" in l:
            pass
    print(f"{kernel}<{frag}>: loop of {len(best)} instructions, {len(refill)} of them in the ring-refill block")
    print(f"steady state: {n} warp instructions per tile of 8 steps = {n / per:.1f} issue slots per sample and lane")
    for op, c in hist.most_common():
        print(f"  {op:10s} {c:4d}  {c / per:5.2f} / sample")
    print("(static count; the measured issue rate decides what fraction of 4 x 1 warp-instruction/clk/SM it reaches)")


if __name__ == "__main__":
    main()
