#!/usr/bin/env python
"""Turn one `ncu --set full` capture of k_fused_hot into the tracked summaries under profiles/:
    python tools/ncu_summary.py gpurun_out/prof_fused_r1_final.ncu-rep
writes profiles/r1_ncu_k_fused_hot_full.json (selected raw metrics) and profiles/r1_traffic.json (the
dram__bytes_read/write figures bench.py reports as roofline.traffic).  Runs here (no GPU needed)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ("dram__", "gpu__time_duration", "gpu__dram_throughput", "launch__", "smsp__issue", "smsp__inst_executed.sum",
        "smsp__inst_issued.sum", "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma.", "sm__inst_executed_pipe_lsu",
        "sm__inst_executed_pipe_tma", "l1tex__data_pipe_lsu_wavefronts_mem_shared", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared",
        "smsp__average_warps_issue_stalled", "sm__warps_active", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate")
UNITS = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {}
    for h, u, v in zip(hdr, units, vals):
        if h.startswith(KEEP) or h in ("Kernel Name", "Grid Size", "Block Size"):
            out[h] = {"unit": u, "value": v}
    out["_capture"] = {"report": os.path.basename(rep), "command": "ncu --set full --clock-control none --import-source on "
                       "-k regex:k_fused_hot -s 3 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"}
    with open(os.path.join(ROOT, "profiles", "r1_ncu_k_fused_hot_full.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)

    def nbytes(key):
        m = out[key]
        return float(m["value"].replace(",", "")) * UNITS[m["unit"]]
    rd, wr = int(round(nbytes("dram__bytes_read.sum"))), int(round(nbytes("dram__bytes_write.sum")))
    traffic = {"cfg3_pipeline": {"streams_per_gpu": 4096, "seconds": 2.0, "kernel": "k_fused_hot", "dram_bytes_read": rd,
                                 "dram_bytes_write": wr, "traffic": rd + wr,
                                 "source": "profiles/r1_ncu_k_fused_hot_full.json (ncu --set full --clock-control none, one launch)"}}
    with open(os.path.join(ROOT, "profiles", "r1_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    print(json.dumps({"duration": out.get("gpu__time_duration.sum"), "dram_read": rd, "dram_write": wr, "metrics": len(out)}))


if __name__ == "__main__":
    main()
