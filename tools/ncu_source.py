"""Per-source-line hot spots of one kernel of an ncu report captured with --import-source on (kernels built with -lineinfo).

    python tools/ncu_source.py <report.ncu-rep> <kernel regex> [top N] [launch index among the matching launches]

Prints executed warp instructions and stall samples per CUDA source line, largest sample count first.
"""
import csv
import io
import os
import re
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    launches, cur, fpath, hdr = [], None, "", None       # a launch = consecutive file blocks of one function
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == "File Path":
            fpath = os.path.basename(row[1])
        elif row[0] == "Function Name":
            if cur is None or cur["name"] != row[1] or fpath in cur["files"]:
                cur = dict(name=row[1], rows=[], files=set())
                launches.append(cur)
            cur["files"].add(fpath)
        elif row[0] == "Line No":
            hdr = row
        elif cur is not None and hdr is not None and len(row) == len(hdr) and row[2] == "-":
            cur["rows"].append((fpath, row))
    sel = [b for b in launches if re.search(pat, b["name"])]
    b = sel[which]
    ci = {}
    for i, h in enumerate(hdr):
        ci.setdefault(h, i)
    ie, isamp = ci["Instructions Executed"], ci["# Samples"]
    rows = b["rows"]
    tot_i = sum(int(r[ie] or 0) for _, r in rows)
    tot_s = sum(int(r[isamp] or 0) for _, r in rows)
    print(f"# {b['name']}: {tot_i} warp instructions, {tot_s} stall samples; {len(sel)} matching launch(es)")
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    rows.sort(key=lambda fr: -int(fr[1][isamp] or 0))
    for f, r in rows[:top]:
        n, s = int(r[ie] or 0), int(r[isamp] or 0)
        st = sorted(((int(r[ci[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
        st = " ".join(f"{h}:{v}" for v, h in st if v)
        print(f"{100 * n / max(tot_i, 1):5.1f}% inst {100 * s / max(tot_s, 1):5.1f}% samp {f}:{r[0]:>4s} {r[1].strip()[:100]:100s} | {st}")


if __name__ == "__main__":
    main()
