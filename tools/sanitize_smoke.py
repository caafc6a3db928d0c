#!/usr/bin/env python
"""tools/sanitize_smoke.py -- a small pass over every kernel (aligned fast paths and generic paths),
meant to be run under ONE compute-sanitizer tool:

    compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_smoke.py

Results are also compared with the oracle, so a clean exit means: no memory error AND bit-exact."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
import patterns as P  # noqa: E402
import imageprocessingtools_b200 as ip  # noqa: E402


def main():
    orc = oracle.orc()
    g = ip.Ppmx(0)
    n = 0
    for (w, h) in [(1, 1), (5, 3), (17, 9), (48, 16), (64, 64), (100, 37), (128, 80)]:
        for img in (P.lcg(w, h, 5), P.mixed(w, h)):
            assert np.array_equal(g.gray(img), orc.gray(img))
            assert np.array_equal(g.mono(img), orc.mono(img))
            assert np.array_equal(g.mono_bits(img), orc.pack_pbm(orc.mono(img)))
            for d in (0, 1):
                assert np.array_equal(g.flip(img, d), orc.flip(img, d))
            for a in (90, 180, 270, 33):
                assert np.array_equal(g.rotate(img, a), orc.rotate(img, a))
            for new in (max(1, w // 2), w + 3, 2 * w):
                for dim, n_in in ((1, w), (0, h)):
                    wt, ix = g.calc_contributions(n_in, new, float(new) / n_in)
                    assert np.array_equal(g.imresize(img, new, dim, wt, ix), orc.imresize(img, new, dim, wt, ix))
            for k in (3, 5, 7, 9):
                coef = np.arange(k * k).reshape(k, k) % 5 - 2
                assert np.array_equal(g.conv(img, coef, 7, 3), orc.conv(img, coef, 7, 3))
            gr, bins = g.gray_hist(img)
            assert np.array_equal(bins, orc.hist_gray(img)) and np.array_equal(gr, orc.gray(img))
            for kw in (dict(gray=True, fliph=True), dict(resize_w=w + 5, angle=90, mono=True, flipv=True),
                       dict(angle=45, gray=True)):
                exp = orc.process(img, **kw)
                got = g.process(img, **kw)
                assert got[1:] == exp[1:] and np.array_equal(got[0], exp[0])
            n += 1
    g.close()
    print("sanitize_smoke ok:", n, "rasters,", "all operators bit-exact")


if __name__ == "__main__":
    main()
