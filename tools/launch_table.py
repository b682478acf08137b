"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (launches, total us, share)."""
import collections
import csv
import io
import re
import sys


def main(path, top=40):
    lines = [l for l in open(path) if l.startswith('"')]
    per = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
        per[name][0] += 1
        per[name][1] += v
    tot = sum(v[1] for v in per.values())
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"| `{k[:88]}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")
    print(f"| **total** | {sum(v[0] for v in per.values())} | {tot:.1f} | 100% |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
