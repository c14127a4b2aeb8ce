#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the few numbers DESIGN.md / bench.py quote.
   python tools/ncu_summary.py gpurun_out/x.ncu-rep [more.ncu-rep ...] > profiles/x_summary.json
One entry per distinct (kernel, grid) -- the LAST captured instance."""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.sum", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__cycles_active.avg", "sm__cycles_elapsed.max",
]


def summarize(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    H, U = rows[0], rows[1]
    ki = H.index("Kernel Name")
    out = {}
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("dgm::", "").replace("void ", "")
        ent = {}
        for i, h in enumerate(H):
            if h in KEEP and i < len(r):
                ent[h] = f"{r[i]} {U[i]}".strip()
        stalls = []
        for i, h in enumerate(H):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")],
                                   float(r[i].replace(",", ""))))
                except ValueError:
                    pass
        ent["top_stalls"] = sorted(stalls, key=lambda x: -x[1])[:4]
        out[f"{name} grid={ent.get('launch__grid_size', '?').split()[0]}"] = ent
    return out


if __name__ == "__main__":
    res = {}
    for p in sys.argv[1:]:
        res.update(summarize(p))
    print(json.dumps(res, indent=1))
