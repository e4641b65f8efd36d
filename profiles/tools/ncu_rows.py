"""Selected columns of an `ncu --page raw --csv` export, one row per captured launch (markdown).

    ncu -i x.ncu-rep --page raw --csv > x.csv;  python profiles/tools/ncu_rows.py x.csv
"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "time us", 1.0), ("dram__bytes_read.sum", "DRAM rd MB", None), ("dram__bytes_write.sum", "DRAM wr MB", None),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1.0),
        ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue %", 1.0),
        ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor %", 1.0),
        ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor %", 1.0),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %", 1.0),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem LSU %", 1.0),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem TC %", 1.0),
        ("launch__registers_per_thread", "regs", 1.0), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %", 1.0)]


def to_mb(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    h, units = rows[0], rows[1]
    ki, gi = h.index("Kernel Name"), h.index("Grid Size")
    seen, cols = set(), []
    for name, label, _ in COLS:
        if name in h and label not in seen:
            seen.add(label)
            cols.append((h.index(name), label))
    print("| kernel | grid | " + " | ".join(l for _, l in cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for r in rows[2:]:
        out = []
        for i, label in cols:
            v = r[i]
            if "MB" in label:
                out.append(f"{to_mb(v, units[i]):.1f}")
            else:
                try:
                    out.append(f"{float(v.replace(',', '')):.1f}")
                except ValueError:
                    out.append(v or "-")
        print(f"| `{r[ki][:58]}` | {r[gi]} | " + " | ".join(out) + " |")


if __name__ == "__main__":
    main()
