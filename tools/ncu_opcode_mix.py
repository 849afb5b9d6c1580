#!/usr/bin/env python3
"""Executed warp instructions of one kernel by opcode, from the SASS page of an `ncu --set full
--import-source on` report (the format of profiles/*_opcode_mix.txt).

  python tools/ncu_opcode_mix.py gpurun_out/r01f_full.ncu-rep k_var_base 1048576 [--json profiles/kernel_work.json]

With --json the per-item IMAD.WIDE and instruction counts, the DRAM bytes of the launch (dram__bytes_read.sum +
dram__bytes_write.sum of the same report) and the multiplier-pipe figures are merged into that file under the kernel's
name: bench.py reads its roofline numerators and `traffic` from there instead of from constants in its source.

Development aid: runs `ncu -i <rep> --page source --csv -k regex:<kernel>` here (no GPU needed)."""
import collections
import csv
import io
import re
import subprocess
import sys


def raw_metrics(rep, kernel):
    """metric name -> value of the first launch of `kernel` in the report's raw page"""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "-k", f"regex:{kernel}"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
             "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}
    out = {}
    for n, u, v in zip(names, units, vals):
        try:
            out[n] = float(v.replace(",", "")) * scale.get(u, 1.0)   # bytes and seconds in base units, everything else as printed
        except ValueError:
            out[n] = v
    return out


def main():
    rep, kernel, items = sys.argv[1], sys.argv[2], int(sys.argv[3])
    json_out = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
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
    if json_out:
        import json
        import os
        m = raw_metrics(rep, kernel)
        num = lambda k: m.get(k, 0.0) if isinstance(m.get(k, 0.0), float) else 0.0
        d = json.load(open(json_out)) if os.path.exists(json_out) else {}
        d[kernel] = {"imad_wide_per_item": round(32.0 * wide / items), "instr_per_item": round(32.0 * total / items), "items": items,
                     "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
                     "fmaheavy_pct": num("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed") or None,
                     "duration_ms_under_ncu": num("gpu__time_duration.sum") * 1e3,
                     "pipe_cycles_share": {"imad_wide": 4.0 * wide / cyc, "imad_mov": 2.0 * mov / cyc, "imad_other": 2.0 * oth / cyc},
                     "source": os.path.basename(rep) + " (ncu --set full, first launch), tools/ncu_opcode_mix.py"}
        d[kernel]["ncu_note"] = (f"fmaheavy pipe {d[kernel]['fmaheavy_pct']} % active; {d[kernel]['instr_per_item']} instructions per item, "
                                 f"{d[kernel]['imad_wide_per_item']} of them IMAD.WIDE ({os.path.basename(rep)})")
        json.dump(d, open(json_out, "w"), indent=1)
        print("wrote", json_out)


if __name__ == "__main__":
    main()
