#!/usr/bin/env python
"""Turn one `ncu --set full` capture (.ncu-rep, kept in gpurun_out/) into a tracked summary under profiles/:
    python tools/ncu_summary.py gpurun_out/r2_pass1/lanes_full.ncu-rep profiles/r2_ncu_k_fused_lanes_65536.json "command line of the capture"
Selected raw metrics of the first kernel in the report, plus derived figures (warp instructions per output sample when
`--samples N` is given, DRAM traffic).  Runs here (no GPU needed)."""
import csv
import io
import json
import os
import subprocess
import sys

KEEP = ("dram__bytes", "dram__throughput", "gpu__time_duration", "launch__registers", "launch__occupancy_limit", "launch__shared_mem_per_block",
        "launch__waves", "launch__grid_size", "launch__block_size", "smsp__issue_active", "smsp__inst_executed.sum", "smsp__inst_issued.sum",
        "sm__inst_executed_pipe_alu.avg", "sm__inst_executed_pipe_fma.avg", "sm__inst_executed_pipe_lsu.avg", "sm__inst_executed_pipe_uniform.avg",
        "sm__pipe_fma_cycles_active.avg", "sm__pipe_alu_cycles_active.avg", "sm__pipe_fmaheavy_cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__throughput.avg",
        "lts__throughput.avg", "smsp__average_warps_issue_stalled", "smsp__average_warp_latency", "sm__warps_active.avg", "smsp__warps_eligible.avg",
        "smsp__warps_active.avg", "sm__cycles_elapsed.max", "sm__throughput.avg", "lts__t_bytes.sum", "lts__t_sector_hit_rate")
UNITS = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main():
    rep, out_path = sys.argv[1], sys.argv[2]
    command = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else ""
    samples = float(sys.argv[sys.argv.index("--samples") + 1]) if "--samples" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {}
    for h, u, v in zip(hdr, units, vals):
        if h.startswith(KEEP) or h in ("Kernel Name", "Grid Size", "Block Size"):
            out[h] = {"unit": u, "value": v}

    def num(key):
        m = out[key]
        return float(m["value"].replace(",", "")) * UNITS.get(m["unit"], 1.0)
    derived = {"dram_bytes_read": int(round(num("dram__bytes_read.sum"))), "dram_bytes_write": int(round(num("dram__bytes_write.sum")))}
    derived["traffic"] = derived["dram_bytes_read"] + derived["dram_bytes_write"]
    if samples:
        derived["output_samples"] = samples
        derived["warp_instructions_per_sample_x32"] = num("smsp__inst_executed.sum") * 32.0 / samples
    out["_derived"] = derived
    out["_capture"] = {"report": os.path.basename(rep), "command": command}
    with open(out_path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps({"kernel": out.get("Kernel Name", {}).get("value"), "duration": out.get("gpu__time_duration.sum"), **derived}))


if __name__ == "__main__":
    main()
