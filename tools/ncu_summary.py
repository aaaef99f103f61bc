#!/usr/bin/env python
"""Condense an `ncu --set full` report into the per-kernel numbers DESIGN.md / bench.py quote.

usage: tools/ncu_summary.py <report.ncu-rep> [out.md]
       tools/ncu_summary.py --traffic <workload> <report.ncu-rep> <summary.md name> <kernel regex>   (updates profiles/ncu_traffic.json)
One block per profiled launch: duration, issue/IPC, thread efficiency, occupancy, L1/L2 hit rates, DRAM bytes
(= bench.py's roofline.traffic), top stall reasons.  Reads the report with `ncu -i ... --page raw --csv`."""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), CTAs/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), CTAs/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC per SM (max 4)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per warp instruction (max 32)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sectors_op_atom.sum", "L2 atomic sectors"),
    ("lts__t_sectors_op_red.sum", "L2 reduction sectors"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
]


def traffic(workload, rep, summary, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of the first launch matching `kernel`, per launch, into
    profiles/ncu_traffic.json (bench.py's roofline.traffic reads it)."""
    import json, os, re
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    iN, iR, iW = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = [(float(r[iR].replace(",", "")) * scale[units[iR]] + float(r[iW].replace(",", "")) * scale[units[iW]], r[iN])
            for r in rows[2:] if re.search(kernel, r[iN])]
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    data = json.load(open(path)) if os.path.isfile(path) else {}
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    data[workload] = {"dram_bytes_per_launch": sum(v for v, _ in vals) / len(vals), "launches_averaged": len(vals),
                      "kernel": vals[0][1].replace("<unnamed>::", ""), "summary": summary, "captured_at_commit": commit}
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    print(path, data[workload])


def main():
    if sys.argv[1] == "--traffic":
        return traffic(*sys.argv[2:6])
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    lines = ["# ncu summary of `%s`" % rep.split("/")[-1], "",
             "Captured with `ncu --set full --clock-control none --import-source on` on a B200 (sm_100a); per-launch "
             "values are cold-cache and serialised (see B200_PROFILING.md), so shares, not absolutes, compare with "
             "bench.py.", ""]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines.append("## %s" % name.replace("<unnamed>::", ""))
        lines.append("")
        lines.append("| metric | value |")
        lines.append("|---|---|")
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                v = r[i]
                try:
                    v = "%.4g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                lines.append("| %s | %s %s |" % (label, v, units[i]))
        lines.append("")
    text = "\n".join(lines)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")
    else:
        print(text)


if __name__ == "__main__":
    main()
