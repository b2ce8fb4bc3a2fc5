"""Summarise ncu reports (read here on the CPU box) into profiles/*.md / *.csv.
usage: python scripts/summarize_ncu.py <report.ncu-rep> <out.md> [title]"""
import csv, io, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]

def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    units = rows[1]
    lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none), read with `ncu -i --page raw --csv`", ""]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines.append(f"## {d.get('Kernel Name', '?')}  (id {d.get('ID')}, grid {d.get('Grid Size')}, block {d.get('Block Size')})")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for k in KEYS:
            if k in d:
                lines.append(f"| {k} | {d[k]} | {u.get(k, '')} |")
        try:
            t = float(d["gpu__time_duration.sum"].replace(",", ""))
            by = float(d["dram__bytes_read.sum"].replace(",", "")) + float(d["dram__bytes_write.sum"].replace(",", ""))
            lines.append(f"| dram traffic (read+write) | {by:.6g} | {u.get('dram__bytes_read.sum','')} |")
        except Exception:
            pass
        lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out, len(rows) - 2, "kernels")

if __name__ == "__main__":
    main()
