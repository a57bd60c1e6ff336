#!/usr/bin/env python
"""Key counters of an `ncu --set full` capture, read with no GPU: python tools/ncu_summary.py file.ncu-rep [...]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "smsp__cycles_active.avg",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(path, "no data")
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            print("==", path, "|", d.get("Kernel Name"), "| grid", d.get("launch__grid_size"), "block", d.get("launch__block_size"))
            for k in KEYS:
                if k in d:
                    print("  %-68s %s %s" % (k, d[k], units[hdr.index(k)]))
            st = sorted(((float(v.replace(",", "")), k) for k, v in d.items() if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")), reverse=True)
            tot = sum(v for v, _ in st) or 1.0
            print("  warp cycles per issued instruction %s; stall shares:" % d.get("smsp__average_warp_latency_per_inst_issued.ratio"),
                  ", ".join("%s %.0f%%" % (k[len(STALL):-len("_per_issue_active.ratio")], 100 * v / tot) for v, k in st[:6]))


if __name__ == "__main__":
    main()
