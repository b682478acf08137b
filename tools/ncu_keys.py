"""Print selected raw metrics (pipes, LSU wavefronts, stall reasons) of every kernel in an .ncu-rep, one column per report."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg", "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[2:]


def main(paths):
    cols = []
    for p in paths:
        hdr, rows = load(p)
        for r in rows:
            cols.append((p.split("/")[-1][:14] + ":" + r[hdr.index("Kernel Name")][16:40], hdr, r))
    keys = list(WANT)
    hdr0 = cols[0][1]
    keys += sorted(h for h in hdr0 if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"))
    print(" " * 70, *[f"{c[0][:26]:>27s}" for c in cols])
    for k in keys:
        vals = []
        for _, hdr, r in cols:
            i = next((j for j, h in enumerate(hdr) if h == k or h.endswith("." + k)), None)
            vals.append(r[i] if i is not None else "-")
        print(f"{k[-70:]:70s}", *[f"{v:>27s}" for v in vals])


if __name__ == "__main__":
    main(sys.argv[1:])
