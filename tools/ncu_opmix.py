#!/usr/bin/env python
"""tools/ncu_opmix.py REPORT.ncu-rep [BYTES] -- executed warp instructions per opcode (from the source page of a capture made with
--import-source on), optionally per output byte: which instructions a kernel actually spends its issue slots on."""
import csv
import io
import re
import subprocess
import sys
from collections import Counter


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if "Instructions Executed" in r)
    i_src, i_ex, i_st = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    ops, stalls = Counter(), Counter()
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) <= i_ex:
            continue
        t = re.sub(r"^\s*@!?U?P\d+\s+", "", r[i_src].strip())
        op = t.split()[0].split(".")[0] if t else "?"
        try:
            ops[op] += int(r[i_ex])
            stalls[op] += int(r[i_st])
        except ValueError:
            pass
    total = sum(ops.values())
    nbytes = float(sys.argv[2]) if len(sys.argv) > 2 else 0
    print("warp instructions executed: %d%s" % (total, "  = %.2f thread instructions per byte" % (32 * total / nbytes) if nbytes else ""))
    for op, n in ops.most_common(24):
        print("  %-10s %10d  %5.1f %%%s   stall samples %d" % (op, n, 100.0 * n / total, "  %.2f/B" % (32 * n / nbytes) if nbytes else "", stalls[op]))


if __name__ == "__main__":
    main()
