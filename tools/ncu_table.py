"""Print one row per profiled kernel from an .ncu-rep (raw page): duration, DRAM bytes, DRAM %, L2 %, tensor pipe %, XU %, warps, regs."""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB_r"), ("dram__bytes_write.sum", "MB_w"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "xu%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("launch__registers_per_thread", "regs")]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    print("| kernel | " + " | ".join(k[1] for k in KEYS) + " |")
    print("|---|" + "---:|" * len(KEYS))
    for r in rows[2:]:
        vals = []
        for key, _ in KEYS:
            if key in hdr:
                i = hdr.index(key)
                v = r[i].replace(",", "")
                try:
                    f = float(v)
                    u = units[i]
                    if u == "ns": f /= 1e3
                    if u == "byte": f /= 1e6
                    if u == "Kbyte": f /= 1e3
                    if u == "Gbyte": f *= 1e3
                    vals.append(f"{f:.1f}")
                except ValueError:
                    vals.append(v)
            else:
                vals.append("-")
        name = r[kn].replace("void ", "").replace("<unnamed>::", "")
        name = name[:name.index("(")] if "(" in name else name
        print(f"| `{name[:70]}` | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
