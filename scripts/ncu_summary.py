"""Turns an ncu report (--set full) into the JSON summary committed under profiles/: one entry per
profiled launch with the metrics the rooflines are read from.
    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rXX_ncu_full_summary.json"""
import csv, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
        "smsp__cycles_active.avg", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
res = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")]}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            d[k] = f"{r[i]} {units[i]}".strip()
    res.append(d)
json.dump(res, sys.stdout, indent=1)
print()
