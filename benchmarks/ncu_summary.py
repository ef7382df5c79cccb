"""Markdown table of the judged numbers out of `ncu --page raw --csv` exports (benchmarks/ncu_export.sh):
per launch - duration, SM clock, tensor-pipe activity (and which sub-pipe), L2 / DRAM throughput, DRAM bytes,
registers, occupancy.   python benchmarks/ncu_summary.py <raw.csv> [<raw.csv> ...]"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__cycles_elapsed.avg.per_second", "SM clk"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "of which DMMA %"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "HMMA %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
]


def main():
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        name_i = hdr.index("Kernel Name")
        grid_i = hdr.index("Grid Size")
        idx = [(hdr.index(k), lab) for k, lab in COLS if k in hdr]
        print(f"\n### {path}\n")
        print("| kernel | grid | " + " | ".join(lab for _, lab in idx) + " |")
        print("|---|---|" + "---|" * len(idx))
        for r in rows[2:]:
            cells = []
            for i, _ in idx:
                v = r[i]
                try:
                    f = float(v.replace(",", ""))
                    v = f"{f:.3f}".rstrip("0").rstrip(".") if abs(f) < 1000 else f"{f:.0f}"
                except ValueError:
                    pass
                cells.append(f"{v} {units[i]}".strip())
            kn = r[name_i].split("(")[0].replace("void ", "")
            print(f"| `{kn}` | {r[grid_i]} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
