#!/usr/bin/env python
"""Back-of-the-envelope timing of ONE warp running the steady-state loop of k_fused_lanes, from the SASS alone (no GPU):
in-order issue (one instruction per cycle at best), every instruction waits for its source registers / predicates, results
become available a fixed latency after issue.  It answers "how long is the dependent chain through one tile, and through
which instructions" -- the quantity that bounds the kernel when few warps share a sub-partition (a lone
warp was measured at about 800 cycles per tile).  Latencies are round numbers (dependent FP32 / integer issue 4 cycles,
shared-memory load 30, shuffle 26, constant load 8): a model to compare schedules with, not a prediction.
    python tools/sass_inorder_sim.py [mangled-name-fragment]"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
LAT = {"LDS": 30, "SHFL": 26, "LDC": 8, "S2R": 20, "LDG": 400, "STG": 1, "LDGSTS": 1, "BRA": 1, "DEPBAR": 1, "LDGDEPBAR": 1,
       "WARPSYNC": 1, "NOP": 1}
DEFAULT_LAT = 4


def loop_instructions(frag, csrc=None):
    src = os.path.join(csrc or os.path.join(ROOT, "rodio_b200", "csrc"), "rb_lanes.cu")
    flags = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "--fmad=false"]
    with tempfile.TemporaryDirectory() as td:
        obj = os.path.join(td, "o.o")
        subprocess.run(["nvcc"] + flags + ["-c", src, "-o", obj], check=True, capture_output=True)
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    body = next(f for f in re.split(r"\n\s*Function : ", sass) if "k_fused_lanes" in f.split("\n", 1)[0] and frag in f.split("\n", 1)[0])
    ins = []
    for line in body.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4})\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_.]+)\s*(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), (m.group(2) or "").strip(), m.group(3), m.group(4)))
    addr = {a: i for i, (a, _, _, _) in enumerate(ins)}
    best = None
    for i, (a, _, op, rest) in enumerate(ins):
        if op.startswith("BRA"):
            t = re.search(r"0x([0-9a-f]+)", rest)
            if t and int(t.group(1), 16) < a and int(t.group(1), 16) in addr:
                seg = ins[addr[int(t.group(1), 16)]:i + 1]
                if any(o.startswith("SHFL.BFLY") for _, _, o, _ in seg) and any(o == "LDS" and p for _, p, o, _ in seg):
                    if best is None or len(seg) < len(best):
                        best = seg
    # drop the refill block (branched over on most iterations)
    i0 = next(i for i, x in enumerate(best) if x[2].startswith("DEPBAR"))
    i1 = max(i for i, x in enumerate(best) if x[2].startswith("LDGDEPBAR"))
    return best[:i0] + best[i1 + 1:]


REG = re.compile(r"\b(U?R\d+|U?P\d+)\b")


def operands(pred, op, rest):
    """(dest registers, source registers) of one SASS instruction -- first operand(s) written, the rest read."""
    parts = [p.strip() for p in rest.split(",")]
    regs = [REG.findall(p.replace(".reuse", "")) for p in parts]
    opc = op.split(".")[0]
    n_dst = 1
    if opc in ("STG", "STS", "BRA", "DEPBAR", "LDGDEPBAR", "NOP", "WARPSYNC", "LDGSTS", "EXIT"):
        n_dst = 0
    elif opc in ("FSETP", "ISETP", "UISETP", "PLOP3", "SHFL", "LEA", "IADD3", "UIADD3", "VIADD", "IMAD") and len(regs) > 1 and regs[1] and \
            regs[1][0].lstrip("U").startswith("P") and regs[0] and regs[0][0].lstrip("U").startswith(("P", "R")):
        n_dst = 2 if opc in ("FSETP", "ISETP", "UISETP", "PLOP3", "SHFL") else 1
    dst = [r for g in regs[:n_dst] for r in g]
    srcs = [r for g in regs[n_dst:] for r in g]
    if op.endswith(".64") or ".WIDE" in op:
        dst += [re.sub(r"\d+$", lambda m: str(int(m.group()) + 1), r) for r in dst if r.startswith("R")]
    if pred:
        p = pred.lstrip("@!")
        srcs.append(p)
        srcs += dst                 # a predicated write keeps the old value when the predicate is off
    return [r for r in dst if r not in ("RZ", "PT", "URZ", "UPT")], [r for r in srcs if r not in ("RZ", "PT", "URZ", "UPT")]


def simulate(loop, iterations=6):
    ready = {}
    t = 0
    starts = []
    last_writer = {}
    crit = None
    for it in range(iterations):
        starts.append(t)
        for (addr, pred, op, rest) in loop:
            dst, srcs = operands(pred, op, rest)
            need = max([ready.get(r, 0) for r in srcs] + [t])
            blocker = max(srcs, key=lambda r: ready.get(r, 0)) if srcs else None
            t_issue = need
            lat = LAT.get(op.split(".")[0], DEFAULT_LAT)
            for r in dst:
                ready[r] = t_issue + lat
                last_writer[r] = (addr, op, t_issue, blocker)
            t = t_issue + 1
        crit = dict(last_writer)
    per_iter = [b - a for a, b in zip(starts, starts[1:])]
    return per_iter, crit


def main():
    frag = sys.argv[1] if len(sys.argv) > 1 else "ILi1ELi1ELb1ELb1ELi1ELb0E"
    loop = loop_instructions(frag, sys.argv[2] if len(sys.argv) > 2 else None)   # argv[2]: another csrc directory to compare
    per_iter, _ = simulate(loop)
    n = len(loop)
    print(f"k_fused_lanes<{frag}>: {n} instructions per tile; in-order single-warp model: {per_iter[-1]} cycles per tile "
          f"({per_iter[-1] / n:.2f} cycles per instruction, {n / per_iter[-1]:.2f} IPC); iterations: {per_iter}")
    for name, over in (("shared-memory load 60 instead of 30", {"LDS": 60}), ("shuffle 50 instead of 26", {"SHFL": 50}),
                       ("dependent FP32/int issue 6 instead of 4", None)):
        global DEFAULT_LAT
        old, old_d = dict(LAT), DEFAULT_LAT
        if over:
            LAT.update(over)
        else:
            DEFAULT_LAT = 6
        p2, _ = simulate(loop)
        print(f"  sensitivity, {name}: {p2[-1]} cycles per tile")
        LAT.clear(), LAT.update(old)
        DEFAULT_LAT = old_d


if __name__ == "__main__":
    main()
