"""Summaries of ncu output for profiles/.

    python tools/summarize_ncu.py launches <launches.csv>          per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list
    python tools/summarize_ncu.py full <report.ncu-rep> [regex]    key metrics per launch of an `ncu --set full` report (needs ncu on PATH)
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
    tot = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"^void ", "", r[4]).replace("<unnamed>::", "").split("(")[0][:60]
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += float(r[14]) / 1000.0
    total = sum(v[1] for v in tot.values())
    print(f"# per-kernel totals of {path} (ncu --metrics gpu__time_duration.sum, cold-cache serialised launches; compare SHARES)")
    for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:60s} n={n:3d} total_us={us:9.1f} share={100 * us / total:5.1f}%")


def full(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    for r in rows[2:]:
        if pattern and not re.search(pattern, r[ik]):
            continue
        print(r[ik].replace("<unnamed>::", "")[:110])
        for k in KEYS:
            if k in hdr:
                print(f"    {k} {r[hdr.index(k)]} {units[hdr.index(k)]}")


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif len(sys.argv) >= 3 and sys.argv[1] == "full":
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        print(__doc__)
