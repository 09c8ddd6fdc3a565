"""Condense an ncu report (`ncu -i X.ncu-rep --page raw --csv`) into the handful of lines DESIGN.md quotes.

usage: ncu -i gpurun_out/X.ncu-rep --page raw --csv | python tools/ncu_summary.py > profiles/X_summary.txt
"""
import csv
import sys

WANT = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu.sum",
]


def main():
    rows = list(csv.reader(sys.stdin))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        for w in WANT:
            if w in d and d[w] != "":
                print(f"{w:88s} {d[w]:>18s} {units[hdr.index(w)]}")
        stalls = []
        for h in hdr:
            if "warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(d[h].replace(",", "")), h))
                except ValueError:
                    pass
        for v, h in sorted(stalls, reverse=True)[:6]:
            print(f"stall {h:82s} {v:18.2f} warps per issue")
        print()


if __name__ == "__main__":
    main()
