#!/usr/bin/env python3
"""Executed warp instructions of one kernel by opcode, from the SASS page of an `ncu --set full
--import-source on` report (the format of profiles/*_opcode_mix.txt).

  python tools/ncu_opcode_mix.py gpurun_out/r01f_full.ncu-rep k_var_base 1048576

Development aid: runs `ncu -i <rep> --page source --csv -k regex:<kernel>` here (no GPU needed)."""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, kernel, items = sys.argv[1], sys.argv[2], int(sys.argv[3])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kernel}"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    cols = rows[hdr]
    ia, isrc, iex = cols.index("Address"), cols.index("Source"), cols.index("Instructions Executed")
    by_op, total, seen = collections.Counter(), 0, set()
    for r in rows[hdr + 1:]:
        if len(r) <= iex or not r[ia] or r[0] == "Address":
            continue
        if r[ia] in seen:  # the first launch of the kernel only
            break
        seen.add(r[ia])
        src = re.sub(r"^@!?U?P\w+\s+", "", r[isrc].strip())
        op = src.split()[0] if src else "?"
        fam = ".".join(op.split(".")[:2]) if op.startswith("IMAD") else op.split(".")[0]
        n = int(float(r[iex] or 0))
        by_op[fam] += n
        total += n
    print(f"{kernel}: total warp instructions {total}")
    for op, n in by_op.most_common(18):
        print(f"  {op:14s} {n:12d}  {100.0 * n / total:5.1f} %")
    wide = sum(n for op, n in by_op.items() if op.startswith("IMAD.WIDE"))
    mov = sum(n for op, n in by_op.items() if op.startswith("IMAD.MOV"))
    oth = sum(n for op, n in by_op.items() if op.startswith("IMAD")) - wide - mov
    print(f"IMAD.WIDE per item (thread level): {32.0 * wide / items:.0f};  instructions per item: {32.0 * total / items:.0f}")
    cyc = 4 * wide + 2 * (mov + oth)
    print(f"multiplier-pipe cycles at 4 per IMAD.WIDE, 2 per other IMAD form: WIDE {400.0 * wide / cyc:.1f} %, "
          f"IMAD.MOV {200.0 * mov / cyc:.1f} %, other {200.0 * oth / cyc:.1f} %")


if __name__ == "__main__":
    main()
