"""Static checks on the built SASS (CPU only: cuobjdump of the objects build() left under imageprocessingtools_b200/build)."""
import glob
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_no_global_access_before_the_pdl_wait():
    """Every kernel is launched with programmatic stream serialization: it may start while its predecessor still runs, and only
    griddepcontrol.wait orders it behind the predecessor's writes.  The compiler may hoist ld.global.nc above the wait (it did,
    once: ppmx_common.cuh ld_global_u32), so the SASS of every kernel is scanned for global accesses before ACQBULK."""
    import sass_pdl_check
    from imageprocessingtools_b200 import build
    build.build_all()
    objs = sorted(glob.glob(os.path.join(ROOT, "imageprocessingtools_b200", "build", "*.cu.o")))
    assert len(objs) >= 16, objs  # release + tuning flavour of every translation unit
    kernels = 0
    for o in objs:
        n, bad = sass_pdl_check.offenders(o)
        kernels += n
        assert not bad, (os.path.basename(o), bad)
    assert kernels > 200


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_instruction_set_claims():
    """What DESIGN.md says about the instruction mix, checked on the release objects: no tensor-core instruction of any kind
    (nothing on this path is a dense contraction), integer dot products (IDP.4A / IDP.2A) for the convolutions, the bulk-copy engine
    (UBLKCP) with mbarriers (SYNCS) in the transposers, cp.async (LDGSTS) staging, and every kernel built for sm_100a."""
    import re
    import subprocess
    objs = [o for o in sorted(glob.glob(os.path.join(ROOT, "imageprocessingtools_b200", "build", "*.cu.o")))
            if not os.path.basename(o).startswith("tuning.")]
    assert len(objs) >= 8, objs
    text = ""
    for o in objs:
        out = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
        assert "arch = sm_100a" in out or "sm_100a" in out, o
        text += out
    assert not re.search(r"\b(HMMA|IMMA|DMMA|QMMA|OMMA|UTCHMMA|UTCIMMA|UTCQMMA|UTCOMMA|HGMMA|IGMMA)\b", text)
    for needle in ("IDP.4A", "IDP.2A", "UBLKCP", "SYNCS", "LDGSTS", "ACQBULK"):
        assert needle in text, needle
