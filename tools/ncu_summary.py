"""Print a markdown table of the key `ncu --set full` metrics of the first kernel in each .ncu-rep given.
usage: python tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units, r = rows[0], rows[1], rows[2]
    name = r[h.index("Kernel Name")].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    print(f"### `{name}`  ({rep.split('/')[-1]})\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in h:
            i = h.index(k)
            print(f"| `{k}` | {r[i]} | {units[i]} |")
    print()
