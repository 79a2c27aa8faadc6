"""tools/ncu_summary.py REPORT.ncu-rep OUT.csv -- selected metrics of an `ncu --set full` report, one column per kernel."""
import csv, re, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
keep = re.compile(r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|dram__throughput.avg.pct|lts__throughput.avg.pct|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|"
                  r"l1tex__throughput.avg.pct|launch__registers|launch__occupancy_limit|launch__grid_size|launch__block_size|sm__warps_active.avg.pct|"
                  r"sm__(inst_executed_pipe_fp64|pipe_fp64_cycles_active)\.avg\.pct|sm__issue_active.avg.pct|smsp__inst_executed.sum$|sm__throughput.avg.pct|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active|smsp__sass_thread_inst_executed_op_dfma_pred_on.sum$|smsp__sass_inst_executed_op_shared|"
                  r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|sm__cycles_active.avg$")
seen = {}
for r in data:
    name = r[ki].split("(")[0].strip()
    seen.setdefault(name, r)          # first launch of each kernel
names = list(seen)
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + names)
    for i, h in enumerate(hdr):
        if keep.search(h):
            w.writerow([h, units[i]] + [seen[n][i] for n in names])
print("kernels:", names)
