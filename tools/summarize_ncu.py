"""Summarise an ncu report (raw + source pages) into text for profiles/.

    python tools/summarize_ncu.py gpurun_out/knn_c5_r01.ncu-rep > profiles/r01_knn_c5_full.txt
"""
import csv
import subprocess
import sys
from collections import Counter

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fp32_pred_on.sum", "smsp__sass_thread_inst_executed_op_fp64_pred_on.sum",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for launch in rows[2:]:
        name = launch[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"== kernel: {name}")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"{m:75s} {launch[i]:>18s} {units[i]}")
        st = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(launch[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        print("warp stalls per issued instruction: " + ", ".join(f"{h}={v:.2f}" for v, h in sorted(st, reverse=True)[:9]))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    heads = [i for i, r in enumerate(rows) if "Address" in r and "Source" in r]
    if not heads:
        return
    hi = heads[0]
    hdr, data = rows[hi], rows[hi + 1:]
    isrc, iss, ie = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    tot = sum(int(r[iss] or 0) for r in data if len(r) > iss)
    tote = sum(int(r[ie] or 0) for r in data if len(r) > ie)
    c, ce = Counter(), Counter()
    for r in data:
        if len(r) <= max(iss, ie):
            continue
        t = r[isrc].split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
        c[op] += int(r[iss] or 0)
        ce[op] += int(r[ie] or 0)
    print(f"\nSASS opcode mix (first kernel): warp instructions executed {tote:.4e}, stall samples {tot}")
    for op, v in ce.most_common(18):
        print(f"  {op:10s} inst {100 * v / max(tote, 1):5.1f}% ({v:.3e})   stall samples {100 * c[op] / max(tot, 1):5.1f}%")


if __name__ == "__main__":
    main()
