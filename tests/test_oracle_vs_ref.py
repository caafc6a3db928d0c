"""Pins the oracle (oracle/ppmx_oracle.c) to the UNMODIFIED reference compiled from its
own source (oracle/_ref, built by oracle/Makefile).  CPU only.  Skipped where the compiled
reference is absent; tests/test_golden.py then still pins the oracle to recorded digests.
"""
import os
import subprocess

import numpy as np
import pytest

import oracle
import patterns as P

ANGLES = [1, 7, 30, 45, 77, 89, 91, 135, 179, 181, 200, 269, 271, 300, 359]


def _cases(sizes):
    for (w, h) in sizes:
        for name, img in P.all_patterns(w, h).items():
            yield w, h, name, img


def test_struct_layout(ref):
    assert ref.lib.ref_sizeof_pixel() == 3  # a reference row is byte-identical to a P6 row


def test_gray_mono_flip(orc, ref):
    for w, h, name, img in _cases(P.SMALL_SIZES[:-1] + P.ODD_WIDTHS[::3]):
        g, ft = ref.gray(img)
        assert ft == oracle.FT_PGM
        assert not g[..., 1:].any(), "reference leaves .g/.b zero"
        assert np.array_equal(orc.gray(img), g[..., 0]), (w, h, name)
        m, ft = ref.mono(img)
        assert ft == oracle.FT_PBM
        assert np.array_equal(orc.mono(img), m[..., 0]), (w, h, name)
        for d in (0, 1):
            assert np.array_equal(orc.flip(img, d), ref.flip(img, d)), (w, h, name, d)


def test_mono_integer_thresholds(orc):
    """The double compare of ref:967 equals g < thr with the integer table used on the GPU."""
    for gval in range(256):
        img = P.const(4, 4, gval)
        m = orc.mono(img)
        for y in range(4):
            for x in range(4):
                assert m[y, x] == (1 if gval < P.BAYER_THR[(x % 4) * 4 + (y % 4)] else 0)


def test_cubic_and_mod(orc, ref):
    xs = np.concatenate([np.linspace(-3, 3, 1201), np.random.default_rng(0).uniform(-2.5, 2.5, 2000)])
    for x in xs:
        assert orc.cubic(x) == ref.cubic(x)
    for a in range(-40, 40):
        for b in (0, 1, 2, 7, 16):
            assert orc.lib.orc_mod(a, b) == ref.lib.ref_mod(a, b)


def test_rotate_sizes(orc, ref):
    for (w, h) in [(1, 1), (2, 3), (37, 23), (512, 512), (1920, 1080), (4096, 4096), (333, 4001)]:
        for a in range(0, 360):
            assert orc.rotate_size(a, w, h) == ref.rotate_size(a, w, h), (w, h, a)


def test_rotate_orthogonal(orc, ref):
    for w, h, name, img in _cases([(1, 1), (2, 3), (13, 7), (16, 4), (37, 23), (64, 48)]):
        for a in (0, 90, 180, 270):
            assert np.array_equal(orc.rotate(img, a), ref.rotate(img, a)), (w, h, name, a)


def test_rotate_bicubic(orc, ref):
    for (w, h) in [(1, 1), (2, 2), (3, 5), (5, 4), (6, 6), (37, 23), (64, 48)]:
        pats = P.all_patterns(w, h)
        for name in ("lcg", "c200", "c255", "checker", "xramp", "mixed"):
            for a in ANGLES:
                assert np.array_equal(orc.rotate(pats[name], a), ref.rotate(pats[name], a)), (w, h, name, a)


def test_rotate_bicubic_larger(orc, ref):
    img = P.lcg(301, 211, 99)
    flat = P.const(301, 211, 200)
    for a in (30, 123, 331):
        assert np.array_equal(orc.rotate(img, a), ref.rotate(img, a))
        assert np.array_equal(orc.rotate(flat, a), ref.rotate(flat, a))


RESIZE_IO = [(37, 74), (37, 55), (37, 37), (37, 18), (37, 7), (37, 3), (64, 96), (64, 32), (64, 13), (5, 64),
             (211, 300), (211, 100), (3, 2), (2, 1), (1, 3), (4096, 6144), (4096, 2048), (1000, 999)]


def test_contributions(orc, ref):
    for n_in, n_out in RESIZE_IO:
        scale = float(n_out) / n_in
        w0, i0 = orc.calc_contributions(n_in, n_out, scale)
        w1, i1 = ref.calc_contributions(n_in, n_out, scale)
        assert w0.shape == w1.shape, (n_in, n_out)
        assert np.array_equal(i0, i1), (n_in, n_out)
        assert np.array_equal(w0.view(np.uint64), w1.view(np.uint64)), (n_in, n_out)  # bit-identical doubles


def test_imresize_passes(orc, ref):
    for (w, h) in [(37, 23), (64, 48), (5, 3)]:
        pats = P.all_patterns(w, h)
        for name in ("lcg", "c200", "c255", "checker", "mixed"):
            img = pats[name]
            for new in (1, 2, 3, 7, w // 2, w - 1, w, w + 1, w * 3 // 2, 2 * w, 5 * w):
                if new < 1:
                    continue
                for dim, n_in in ((1, w), (0, h)):
                    scale = float(new) / n_in
                    wt, ix = ref.calc_contributions(n_in, new, scale)
                    a = orc.imresize(img, new, dim, wt, ix)
                    b = ref.imresize(img, new, dim, wt, ix)
                    assert np.array_equal(a, b), (w, h, name, new, dim)


def _cli(tmp_path, img, args):
    p = str(tmp_path / "in.ppm")
    oracle.write_p6(p, img)
    if os.path.exists(p + ".out"):
        os.remove(p + ".out")
    rc, out = oracle.ref_cli(args, p)
    data = open(p + ".out", "rb").read() if os.path.exists(p + ".out") else None
    return rc, out, data


def _flags(args):
    kw = {}
    for a in args:
        if a == "-gray": kw["gray"] = True
        elif a == "-mono": kw["mono"] = True
        elif a == "-fv": kw["flipv"] = True
        elif a == "-fh": kw["fliph"] = True
        elif a.startswith("-w"): kw["resize_w"] = int(a[2:])
        elif a.startswith("-r"): kw["angle"] = int(a[2:])
    return kw


CHAINS = [["-gray"], ["-mono"], ["-fv"], ["-fh"], ["-r90"], ["-r180"], ["-r270"], ["-r30"], ["-r0"],
          ["-w74"], ["-w20"], ["-w37"], ["-w55", "-r90"], ["-w20", "-r45", "-gray"], ["-w50", "-mono"],
          ["-r90", "-mono", "-fh"], ["-w18", "-r90", "-gray", "-fv"], ["-r270", "-gray", "-fh"],
          ["-gray", "-fh"], ["-gray", "-fv"], ["-mono", "-fh"], ["-mono", "-fv"],  # the leaked-result quirks
          ["-w40", "-fv"], ["-r200", "-fh"], ["-w100", "-r359", "-mono", "-fv"]]


def test_full_chain_against_reference_cli(orc, tmp_path):
    """Whole output FILE (header + raster) of the reference binary == oracle header + orc_process."""
    if not os.path.exists(oracle.REF_CLI):
        pytest.skip("reference CLI not built")
    for (w, h) in [(37, 23), (16, 8), (5, 7)]:
        for name in ("lcg", "mixed", "c200"):
            img = P.all_patterns(w, h)[name]
            for args in CHAINS:
                rc, out, data = _cli(tmp_path, img, args)
                assert rc == 0 and data is not None, (args, out)
                raster, ow, oh, ft = orc.process(img, **_flags(args))
                assert data == orc.header(ft, ow, oh, 255) + raster.tobytes(), (w, h, name, args)


def test_no_op_is_an_error(tmp_path):
    if not os.path.exists(oracle.REF_CLI):
        pytest.skip("reference CLI not built")
    rc, out, data = _cli(tmp_path, P.lcg(8, 8, 1), [])
    assert rc == 255 and "no data to write" in out


def test_resize_driver(orc, tmp_path):
    if not os.path.exists(oracle.REF_CLI):
        pytest.skip("reference CLI not built")
    for (w, h) in [(37, 23), (64, 48), (23, 37)]:
        img = P.lcg(w, h, 7)
        for new in (1 if h >= w else 2, 9, w // 2, w - 1, w, w + 1, w * 3 // 2, 2 * w, 3 * w + 1):
            try:
                exp = orc.resize(img, new)
            except ValueError:
                continue  # height truncates to 0: undefined in the reference
            rc, out, data = _cli(tmp_path, img, ["-w%d" % new])
            assert rc == 0, out
            assert data == orc.header(0, exp.shape[1], exp.shape[0], 255) + exp.tobytes(), (w, h, new)
