"""One row per profiled launch from an `ncu --set full` report exported with
    ncu -i <report>.ncu-rep --page raw --csv > raw.csv
    python profiles/summarize_ncu_full.py raw.csv <capture label>  > profiles/<name>.csv
Columns: duration, DRAM bytes, DRAM throughput, L2 hit rate, achieved warps, registers, occupancy limits, issue
utilisation, and the five largest warp-stall reasons (stalled warps per issue-active cycle)."""
import csv
import re
import sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "dram__sectors_read.sum"]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio")


def main():
    path, label = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "full"
    rows = list(csv.reader(open(path)))
    head, units, body = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(head)}
    out = csv.writer(sys.stdout)
    out.writerow(["capture", "Kernel Name", "Grid Size", "Block Size"] +
                 [f"{c} [{units[ix[c]]}]" if c in ix and units[ix[c]] else c for c in COLS] + ["top stalls (warps per issue-active cycle)"])
    for r in body:
        stalls = []
        for h, i in ix.items():
            m = STALL.fullmatch(h)
            if m and r[i] not in ("", "n/a"):
                try:
                    stalls.append((float(r[i].replace(",", "")), m.group(1)))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        top = "; ".join(f"{n} {v:.1f}" for v, n in stalls[:5])
        out.writerow([label, r[ix["Kernel Name"]], r[ix["Grid Size"]], r[ix["Block Size"]]] +
                     [r[ix[c]] if c in ix else "" for c in COLS] + [top])


if __name__ == "__main__":
    main()
