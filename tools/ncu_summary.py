"""Key metrics per kernel from an .ncu-rep (raw page). Usage: python tools/ncu_summary.py REP"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H, U = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
        "smsp__inst_executed_op_shared_st.sum", "smsp__inst_executed_op_global_ld.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    print("==", r[H.index("Kernel Name")][:90])
    for w in want:
        if w in H:
            print("   %-70s %s %s" % (w, r[H.index(w)], U[H.index(w)]))
    st = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): int(float(r[i])) for i, h in enumerate(H)
          if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and r[i] and float(r[i]) > 0}
    tot = sum(st.values()) or 1
    print("   stalls:", {k: "%.0f%%" % (100.0 * v / tot) for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]})
