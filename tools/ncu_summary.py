#!/usr/bin/env python
"""Markdown summary of `ncu --page raw --csv` exports (one row per captured kernel).

    python tools/ncu_summary.py gpurun_out/r02f_grid_raw.csv [more.csv ...] > profiles/r02f_ncu_tables.md

The counters are the ones the roofline argument rests on (FP64 pipe, issue slots, occupancy, DRAM bytes,
the top stall reasons); profiles/*.md quote these tables."""
import csv
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "FP64 pipe busy (elapsed)"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 instr issue vs peak (active)"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy"),
    ("sm__warps_active.avg.per_cycle_active", "warps active per SM cycle"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic shared memory / block"),
    ("launch__occupancy_limit_registers", "occupancy limit: registers (blocks)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit: shared memory (blocks)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier / issue"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving / issue"),
]


def main():
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        names = [r[hdr.index("Kernel Name")] for r in data]
        print("### %s\n" % path.split("/")[-1])
        print("| metric | " + " | ".join(n.replace("|", "/")[:48] for n in names) + " | unit |")
        print("|---|" + "---|" * (len(names) + 1))
        for key, label in METRICS:
            if key not in hdr:
                continue
            j = hdr.index(key)
            vals = []
            for r in data:
                try:
                    v = float(r[j].replace(",", ""))
                    vals.append("%.4g" % v)
                except ValueError:
                    vals.append(r[j])
            print("| %s (`%s`) | %s | %s |" % (label, key, " | ".join(vals), units[j]))
        print()


if __name__ == "__main__":
    main()
