"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares."""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("void ", "").replace("wiski::", "")
    return name[:90]


def main(path, skip=0, count=None):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for row in csv.DictReader(lines):
        if row.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((short(row["Kernel Name"]), float(row["Metric Value"]), row["Grid Size"], row["Block Size"]))
    rows = rows[skip:] if count is None else rows[skip:skip + count]
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for n, t, _, _ in rows:
        tot[n] += t
        cnt[n] += 1
    total = sum(tot.values())
    print(f"launches: {len(rows)}  total device time: {total / 1e6:.3f} ms  (cold-cache, serialised: compare shares)")
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for n in sorted(tot, key=lambda k: -tot[k])[:25]:
        print(f"| `{n}` | {cnt[n]} | {tot[n] / 1e6:.3f} | {100 * tot[n] / total:.1f}% | {tot[n] / cnt[n] / 1e3:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else None)
