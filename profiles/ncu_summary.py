"""Summarise an .ncu-rep (raw page + source page) into the handful of numbers DESIGN.md / bench.py quote."""
import collections, csv, io, re, subprocess, sys

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]

def main(rep):
    hdr, units, rows = raw(rep)
    for row in rows:
        d = dict(zip(hdr, row))
        print("kernel:", d.get("Kernel Name", "?")[:60], "grid", d.get("launch__grid_size"))
        for k in KEYS:
            if k in d:
                print(f"  {k:70s} {d[k]:>16s} {units[hdr.index(k)]}")
        for k in hdr:
            if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
                try:
                    v = float(d[k])
                except ValueError:
                    continue
                if v >= 0.15:
                    print(f"  stall {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:40s} {v:8.3f} warps/issue")

if __name__ == "__main__":
    main(sys.argv[1])
