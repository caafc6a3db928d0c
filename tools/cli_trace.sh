#!/bin/bash
# tools/cli_trace.sh -- where the wall time of one ppmx-b200 run goes (PPMX_TRACE=1), beside the reference binary
set -e
cd "$(dirname "$0")/.."
T=$(mktemp -d)
python - "$T" <<'PY'
import sys
sys.path.insert(0, ".")
import oracle
from imageprocessingtools_b200 import ppmx as pp
oracle.write_p6(sys.argv[1] + "/a.ppm", pp.synth_lcg(4096, 4096, 1))
for i in range(8):
    oracle.write_p6(sys.argv[1] + "/b%d.ppm" % i, pp.synth_lcg(4096, 4096, 2 + i))
PY
wall() { local t0=$(date +%s%N); "${@:2}"; local t1=$(date +%s%N); echo "$1: wall $(( (t1 - t0) / 1000000 )) ms"; }
export PPMX_TRACE=1
for i in 1 2; do
  wall "ppmx-b200 -gray (4096x4096, 50 MB)" imageprocessingtools_b200/ppmx-b200 -gray $T/a.ppm
done
wall "ppmx-b200 -batch -gray, 8 files" imageprocessingtools_b200/ppmx-b200 -batch -gray $T/b?.ppm
wall "ppmx-b200 -gray, 8 runs of one file" bash -c "for f in $T/b?.ppm; do imageprocessingtools_b200/ppmx-b200 -gray \$f; done"
[ -x oracle/_ref/ppmx-edward ] && wall "reference ppmx-edward -gray" oracle/_ref/ppmx-edward -gray $T/a.ppm
[ -x oracle/_ref/ppmx-edward ] && wall "reference, 8 runs of one file" bash -c "for f in $T/b?.ppm; do oracle/_ref/ppmx-edward -gray \$f; done"
rm -rf "$T"
