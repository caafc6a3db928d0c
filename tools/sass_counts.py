#!/usr/bin/env python
"""tools/sass_counts.py [library.so] -- per kernel of the built library: how many SASS instructions of the kinds that
show what the kernel is made of (bulk copies UBLKCP, mbarrier SYNCS, cp.async LDGSTS, dp4a IDP, byte permutes PRMT,
FP64 DMUL/DADD/DFMA, conversions I2F/F2I, tensor-core UTC*MMA / HMMA, 128-bit global accesses).  No GPU needed."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "imageprocessingtools_b200", "libppmx_gpu.so")
KINDS = [("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("ACQBULK", r"\bACQBULK"), ("UTMALDG/STG", r"\bUTMA(LDG|STG)"),
         ("LDGSTS", r"\bLDGSTS"), ("LDG.128", r"\bLDG\.E(\.\w+)*\.128"), ("STG.128", r"\bSTG\.E(\.\w+)*\.128"),
         ("IDP.4A", r"\bIDP\.4A"), ("IDP.2A", r"\bIDP\.2A"), ("PRMT", r"\bPRMT"), ("SHFL", r"\bSHFL"), ("ATOMS/RED", r"\b(ATOMS|RED|ATOMG)"),
         ("DMUL", r"\bDMUL"), ("DADD", r"\bDADD"), ("DFMA", r"\bDFMA"), ("I2F", r"\bI2F"), ("MMA", r"\b(UTC\w*MMA|HMMA|IMMA)")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, name, order = {}, None, []
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ppmx::", "").replace("ppmx::", "")
            counts[name] = collections.Counter()
            order.append(name)
            continue
        if name and re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            counts[name]["total"] += 1
            for k, pat in KINDS:
                if re.search(pat, ln):
                    counts[name][k] += 1
    cols = ["total"] + [k for k, _ in KINDS]
    print("# %s: SASS instruction counts per kernel (cuobjdump -sass)" % os.path.basename(LIB))
    print("%-58s" % "kernel" + "".join("%12s" % c for c in cols))
    tot = collections.Counter()
    for n in sorted(order):
        print("%-58s" % n[:57] + "".join("%12d" % counts[n][c] for c in cols))
        tot.update(counts[n])
    print("%-58s" % ("ALL (%d kernels)" % len(order)) + "".join("%12d" % tot[c] for c in cols))


if __name__ == "__main__":
    main()
