#!/usr/bin/env python3
"""Key metrics of an `ncu --set full` report as 'metric,unit,value' lines, one file per kernel
(the format of profiles/*_ncu_full_*_summary.csv).

  python tools/ncu_summary.py gpurun_out/r01f_full.ncu-rep profiles/r01f_ncu_full

The metric list is the one of profiles/r01d_ncu_full_k_var_base_summary.csv.  Development aid: runs
`ncu -i <rep> --page raw --csv` here (no GPU needed to read a report)."""
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [l.split(",")[0] for l in open(os.path.join(ROOT, "profiles", "r01d_ncu_full_k_var_base_summary.csv"))]


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    names, units = rows[0], rows[1]
    for r in rows[2:]:
        rec = dict(zip(names, r))
        kname = re.match(r"\w+", rec["Kernel Name"]).group(0)
        out = f"{prefix}_{kname}_summary.csv"
        with open(out, "w") as f:
            for m in WANT:
                if m in rec:
                    f.write(f"{m},{units[names.index(m)]},{rec[m]}\n")
        print(out, rec.get("gpu__time_duration.sum"), rec.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"))


if __name__ == "__main__":
    main()
