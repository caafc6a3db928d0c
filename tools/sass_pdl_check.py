#!/usr/bin/env python
"""tools/sass_pdl_check.py [OBJECT.o | LIB.so ...] -- every kernel of the library starts with griddepcontrol.launch_dependents /
griddepcontrol.wait (programmatic dependent launch): nothing may touch global memory before the wait (SASS: ACQBULK), or a kernel
reads its predecessor's output while that is still being written.  The compiler is free to hoist ld.global.nc (__ldg, const
__restrict__) above the wait; this scan of the SASS is the guard.  Prints the offenders, exit code 1 if there are any."""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GLOBAL_ACCESS = re.compile(r"\b(LDG|LDGSTS|STG|ATOMG|ATOM|RED|UBLKCP|UTMALDG|UTMASTG)\b")


def offenders(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    fn, waited, kernels, bad = None, False, 0, {}
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn, waited = m.group(1), False
            kernels += 1
            continue
        if "ACQBULK" in line:
            waited = True
        elif fn and not waited and GLOBAL_ACCESS.search(line):
            bad.setdefault(fn, line.strip()[:100])
    return kernels, bad


def main():
    paths = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "imageprocessingtools_b200", "build", "*.cu.o")))
    total, rc = 0, 0
    for p in paths:
        kernels, bad = offenders(p)
        total += kernels
        for fn, line in bad.items():
            rc = 1
            print("%s: %s\n    %s" % (os.path.basename(p), fn, line))
    print("%d kernels scanned in %d files, %s" % (total, len(paths), "global access before griddepcontrol.wait FOUND" if rc else
                                                   "none touches global memory before griddepcontrol.wait"))
    return rc


if __name__ == "__main__":
    sys.exit(main())
