"""Text summary of an ncu report for profiles/: per kernel the headline metrics, stall reasons and the opcode mix of the
hot loop.    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x_summary.txt"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sass__inst_executed_local_loads", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
print(f"# ncu --set full summary of {rep}")
for k, r in enumerate(rows[2:]):
    print(f"\n== launch {k}")
    for key in KEYS:
        if key in hdr:
            i = hdr.index(key)
            print(f"  {key:70s} {r[i]:>18s} {units[i]}")
    stalls = [(hdr[i].replace("smsp__pcsamp_warps_issue_stalled_", ""), int(r[i])) for i in range(len(hdr))
              if hdr[i].startswith("smsp__pcsamp_warps_issue_stalled_") and not hdr[i].endswith("_not_issued") and r[i].isdigit()]
    tot = sum(v for _, v in stalls) or 1
    print("  warp-state samples: " + ", ".join(f"{n} {100.0 * v / tot:.0f}%" for n, v in sorted(stalls, key=lambda x: -x[1])[:7]))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{k + 1}"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) > 3 and "Source" in srows[1]:
        h = srows[1]
        ia, ie = h.index("Source"), h.index("Instructions Executed")
        data = [x for x in srows[2:] if len(x) > ie and x[ie].isdigit()]
        mx = max(int(x[ie]) for x in data)
        hot = [x for x in data if int(x[ie]) > 0.8 * mx]
        c = Counter()
        for x in hot:
            parts = x[ia].split()
            op = parts[1] if parts[0].startswith("@") else parts[0]
            c[op.split(".")[0]] += 1
        total = sum(int(x[ie]) for x in data)
        print(f"  hot loop: {len(hot)} SASS instructions executed {mx} times per SM-warp set ({100.0 * sum(int(x[ie]) for x in hot) / total:.0f}% of all "
              f"executed instructions); opcode mix: " + ", ".join(f"{o} {n}" for o, n in c.most_common(14)))
