"""Summarise an `ncu --set full` report (exported with --page raw --csv) into a markdown table of roofline-relevant metrics."""
import csv
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(n for _, n in WANT) + " |")
    print("|---|" + "---:|" * len(WANT))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")[:48]
        cells = []
        for key, _ in WANT:
            if key in idx:
                v, u = r[idx[key]], units[idx[key]]
                try:
                    v = f"{float(v):.3g}"
                except ValueError:
                    pass
                cells.append(f"{v} {u}".strip() if u not in ("%", "") else v)
            else:
                cells.append("-")
        print(f"| `{name}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
