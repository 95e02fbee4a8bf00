#!/usr/bin/env python
"""`ncu -i X.ncu-rep --page raw --csv` -> the handful of per-launch numbers DESIGN.md / profiles/ quote.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv | python tools/ncu_summary.py > profiles/prof.txt
"""
import csv
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block", "launch__cluster", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max", "sm__cycles_active.avg", "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform",
    "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled",
    "l1tex__t_bytes.sum", "sm__sass_inst_executed_op_shared", "smsp__pcsamp_warps_issue_stalled",
]


def main():
    rows = list(csv.reader(l for l in sys.stdin if not l.startswith("==")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    print("# launches captured: %d  (%s)" % (len(data), ", ".join(sorted({r[name_i].split("(")[0] for r in data}))))
    for i, h in enumerate(hdr):
        if any(h.endswith(w) or w in h for w in WANT):
            vals = [r[i] for r in data]
            if all(v in ("0", "", "n/a") for v in vals):
                continue
            print("%-100s %-14s %s" % (h, units[i], "  ".join(vals)))


if __name__ == "__main__":
    main()
