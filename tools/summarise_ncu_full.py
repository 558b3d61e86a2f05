"""Selected metrics per kernel launch from an `ncu --set full` report.

    ncu -i X.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/summarise_ncu_full.py /tmp/raw.csv "header line" > profiles/rNN_ncu_full_top_kernels.txt
"""
import csv
import sys

METRICS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for line in sys.argv[2:]:
        print("# " + line)
    for r in rows[2:]:
        print("---")
        for m in METRICS:
            if m in idx:
                u = units[idx[m]]
                print(f"{m} = {r[idx[m]][:130]} {u}".rstrip())


if __name__ == "__main__":
    main()
