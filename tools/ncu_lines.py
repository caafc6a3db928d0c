#!/usr/bin/env python
"""tools/ncu_lines.py REPORT.ncu-rep BYTES [N] -- executed thread instructions per byte by CUDA source line (capture made with
--import-source on, library built with -lineinfo): where a kernel's issue slots go."""
import csv
import io
import subprocess
import sys


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    nbytes = float(sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    cur, hdr, agg = "?", None, {}
    for r in csv.reader(io.StringIO(out)):
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if "Instructions Executed" in r:
            hdr = r
            i_ex = r.index("Instructions Executed")
            continue
        if hdr is None or len(r) <= i_ex or not r[0]:  # (rows without a line number are the SASS lines under a source line)
            continue
        try:
            agg[(cur, int(r[0]), r[1].strip()[:100])] = agg.get((cur, int(r[0]), r[1].strip()[:100]), 0) + int(r[i_ex])
        except ValueError:
            pass
    total = sum(agg.values())
    print("%.2f thread instructions per byte in all" % (32 * total / nbytes))
    for k, v in sorted(agg.items(), key=lambda x: -x[1])[:top]:
        print("%6.2f/B  %s:%d  %s" % (32 * v / nbytes, k[0], k[1], k[2]))


if __name__ == "__main__":
    main()
