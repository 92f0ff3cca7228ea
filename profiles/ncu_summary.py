"""Turn an `ncu --set full` report into the per-kernel text summary kept under profiles/.

usage: python profiles/ncu_summary.py REPORT.ncu-rep "header line" > profiles/rNN_ncu_*.txt
"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
]


def main():
    rep, header = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    names, units = rows[0], rows[1]
    ix = {n: i for i, n in enumerate(names)}
    print("# " + header + "\n")
    for r in rows[2:]:
        print(f"{'Kernel Name':<76}{r[ix['Kernel Name']][:90]} ")
        print(f"{'Grid Size':<76}{r[ix['Grid Size']]:>30} ")
        print(f"{'Block Size':<76}{r[ix['Block Size']]:>30} ")
        for m in METRICS:
            if m in ix:
                print(f"{m:<76}{r[ix[m]]:>30} {units[ix[m]]}")
        print()


if __name__ == "__main__":
    main()
