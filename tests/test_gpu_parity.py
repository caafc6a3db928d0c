"""GPU parity: the CUDA path, called through the C ABI (libppmx_gpu.so / libppmx_host.so),
against the oracle on the same seeded inputs, against the golden vectors recorded from the
compiled reference, and -- at BASELINE.json's full sizes -- through size-independent properties.
Integer/byte work and the FP64 bicubic operators alike must be BIT-EXACT (tolerance 0).
Run on a B200 with:  python -m pytest tests -m gpu
"""
import os
import subprocess

import numpy as np
import pytest

import golden_replay
import oracle
import patterns as P

pytestmark = pytest.mark.gpu

ANGLES = [1, 7, 30, 45, 77, 89, 91, 135, 179, 181, 200, 269, 271, 300, 359]


@pytest.fixture(scope="module")
def gpu():
    import imageprocessingtools_b200 as ip
    g = ip.Ppmx(0)
    yield g
    g.close()


@pytest.fixture(scope="module")
def gpu_tuning():
    """The tuning build of the same sources (-DPPMX_TUNING): the kernel variants the release library leaves out."""
    import imageprocessingtools_b200 as ip
    g = ip.Ppmx(0, tuning=True)
    yield g
    g.set_tuning("variant", 0)
    g.set_tuning("pdl", 1)
    g.close()


def test_native_library_is_the_one_running(gpu):
    import imageprocessingtools_b200.ppmx as pp
    assert os.path.exists(pp.GPU_SO)
    assert b"sm_100a" in gpu.L.ppmx_gpu_version()
    n0 = gpu.launch_count()
    gpu.gray(P.lcg(64, 64, 1))
    assert gpu.launch_count() > n0, "no kernel was launched"


def test_golden_vectors(gpu):
    """Digests recorded from the compiled reference (tests/golden/make_golden.py)."""
    n = golden_replay.replay(gpu, gpu.header)
    assert n > 1000


def test_integer_ops_sweep(gpu, orc):
    for (w, h) in P.SMALL_SIZES + P.ODD_WIDTHS:
        for name, img in P.all_patterns(w, h).items():
            tag = (w, h, name)
            assert np.array_equal(gpu.gray(img), orc.gray(img)), tag
            m = orc.mono(img)
            assert np.array_equal(gpu.mono(img), m), tag
            assert np.array_equal(gpu.mono_bits(img), orc.pack_pbm(m)), tag
            for d in (0, 1):
                assert np.array_equal(gpu.flip(img, d), orc.flip(img, d)), tag + (d,)
            for a in (0, 90, 180, 270):
                assert np.array_equal(gpu.rotate(img, a), orc.rotate(img, a)), tag + (a,)


def test_pack_pbm_raw_bytes(gpu, orc):
    rng = np.random.default_rng(5)
    for (w, h) in [(1, 1), (7, 3), (8, 2), (9, 2), (37, 5), (64, 16), (100, 7)]:
        raw = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert np.array_equal(gpu.pack_pbm(raw), orc.pack_pbm(raw)), (w, h)


def test_rotate_orth_bulk_copy_tiles(gpu_tuning, orc):
    gpu = gpu_tuning
    """90 / 270 degrees on rasters whose sides are multiples of 16 take the bulk-copy (cp.async.bulk) transposer
    (64 x 64 tiles, narrower / shorter at the right and bottom edges);
    the register-path kernel (variant 6) and the generic one (variant 1) must give the same bytes."""
    try:
        for (w, h) in [(64, 64), (128, 64), (64, 192), (320, 256), (1024, 576), (16, 16), (80, 48), (144, 208), (48, 400)]:
            for name in ("lcg", "xramp", "yramp"):
                img = P.all_patterns(w, h).get(name, P.lcg(w, h, 3))
                for a in (90, 270):
                    exp = orc.rotate(img, a)
                    for v in (0, 6, 1):
                        gpu.set_tuning("variant", v)
                        assert np.array_equal(gpu.rotate(img, a), exp), (w, h, name, a, v)
    finally:
        gpu.set_tuning("variant", 0)


def test_rotate_bicubic_sweep(gpu, orc):
    for (w, h) in [(1, 1), (2, 2), (3, 5), (5, 4), (6, 6), (37, 23), (64, 48), (100, 37)]:
        pats = P.all_patterns(w, h)
        for name in ("lcg", "c200", "c255", "checker", "xramp", "mixed"):
            for a in ANGLES:
                assert np.array_equal(gpu.rotate(pats[name], a), orc.rotate(pats[name], a)), (w, h, name, a)


def test_rotate_bicubic_flat_image_truncation(gpu, orc):
    """A constant-200 image rotated 30 degrees gives a 199/200 speckle in the reference
    ((int) of 199.99999999999997, ref:779); an FMA-contracted build differs in ~4% of bytes."""
    for (w, h, a) in [(301, 211, 30), (301, 211, 77), (512, 512, 123)]:
        img = P.const(w, h, 200)
        exp = orc.rotate(img, a)
        assert 199 in exp and 200 in exp
        assert np.array_equal(gpu.rotate(img, a), exp), (w, h, a)


def test_imresize_sweep(gpu, orc):
    for (w, h) in [(37, 23), (64, 48), (5, 3), (100, 64)]:
        pats = P.all_patterns(w, h)
        for name in ("lcg", "c200", "c255", "checker", "mixed"):
            img = pats[name]
            for new in (1, 2, 3, 7, w // 2, w - 1, w, w + 1, w * 3 // 2, 2 * w, 5 * w):
                if new < 1:
                    continue
                for dim, n_in in ((1, w), (0, h)):
                    wt, ix = gpu.calc_contributions(n_in, new, float(new) / n_in)
                    a = gpu.imresize(img, new, dim, wt, ix)
                    b = orc.imresize(img, new, dim, wt, ix)
                    assert np.array_equal(a, b), (w, h, name, new, dim)


def test_imresize_tap_count_specialisations(gpu, orc):
    """Height pass with 4..8 taps as compile-time constants, 9+ through the loop; width pass templates K = 4..8."""
    img = P.lcg(64, 120, 21)
    seen = set()
    for new in (180, 121, 100, 90, 84, 80, 72, 70, 66, 60, 50, 40):
        wt, ix = gpu.calc_contributions(120, new, new / 120.0)
        seen.add(wt.shape[1])
        assert np.array_equal(gpu.imresize(img, new, 0, wt, ix), orc.imresize(img, new, 0, wt, ix)), new
        t = np.ascontiguousarray(img.transpose(1, 0, 2))  # 120 wide: the same tables drive the width pass
        assert np.array_equal(gpu.imresize(t, new, 1, wt, ix), orc.imresize(t, new, 1, wt, ix)), new
    assert {4, 5, 6, 7, 8} <= seen and max(seen) > 8


CHAINS = [dict(gray=True), dict(mono=True), dict(flipv=True), dict(fliph=True), dict(angle=90), dict(angle=180),
          dict(angle=270), dict(angle=30), dict(angle=0), dict(resize_w=74), dict(resize_w=20), dict(resize_w=37),
          dict(resize_w=55, angle=90), dict(resize_w=20, angle=45, gray=True), dict(resize_w=50, mono=True),
          dict(angle=90, mono=True, fliph=True), dict(resize_w=18, angle=90, gray=True, flipv=True),
          dict(angle=270, gray=True, fliph=True),
          dict(gray=True, fliph=True), dict(gray=True, flipv=True), dict(mono=True, fliph=True),
          dict(mono=True, flipv=True),  # the reference's leaked-result quirks
          dict(resize_w=40, flipv=True), dict(angle=200, fliph=True),
          dict(resize_w=100, angle=359, mono=True, flipv=True)]


def test_op_chains(gpu, orc):
    for (w, h) in [(37, 23), (16, 8), (5, 7), (64, 64)]:
        for name in ("lcg", "mixed", "c200"):
            img = P.all_patterns(w, h)[name]
            for kw in CHAINS:
                exp, ew, eh, eft = orc.process(img, **kw)
                got, gw, gh, gft = gpu.process(img, **kw)
                assert (gw, gh, gft) == (ew, eh, eft), (w, h, name, kw)
                assert np.array_equal(got, exp), (w, h, name, kw)


def test_no_op_chain_is_an_error(gpu):
    import imageprocessingtools_b200 as ip
    with pytest.raises(ip.PpmxError):
        gpu.process(P.lcg(8, 8, 1))


def test_config1_512(gpu, orc):
    """BASELINE config 1: one synthetic 512x512 P6 through -gray and every other operator."""
    img = P.lcg(512, 512, 0xC0FFEE ^ 1)
    assert np.array_equal(gpu.gray(img), orc.gray(img))
    assert np.array_equal(gpu.mono_bits(img), orc.pack_pbm(orc.mono(img)))
    for d in (0, 1):
        assert np.array_equal(gpu.flip(img, d), orc.flip(img, d))
    for a in (90, 180, 270, 30):
        assert np.array_equal(gpu.rotate(img, a), orc.rotate(img, a))
    for kw in (dict(resize_w=768), dict(resize_w=256), dict(resize_w=300, angle=90, gray=True, flipv=True)):
        exp = orc.process(img, **kw)
        got = gpu.process(img, **kw)
        assert got[1:] == exp[1:] and np.array_equal(got[0], exp[0]), kw


def test_full_size_gray_mono_4096(gpu, orc):
    """BASELINE config 2 size: direct comparison (the oracle does 16.8 Mpix in well under a second)."""
    img = P.lcg(4096, 4096, 0xC0FFEE ^ 2)
    assert np.array_equal(gpu.gray(img), orc.gray(img))
    assert np.array_equal(gpu.mono_bits(img), orc.pack_pbm(orc.mono(img)))
    flat = P.const(4096, 4096, 77)
    assert np.array_equal(gpu.gray(flat), orc.gray(flat))


def test_full_size_properties_8192(gpu, orc):
    """BASELINE config 3 size through size-independent properties."""
    img = P.lcg(8192, 8192, 0xC0FFEE ^ 3)
    fh = gpu.flip(img, 0)
    assert np.array_equal(fh, img[:, ::-1])
    assert np.array_equal(gpu.flip(fh, 0), img)                      # involution
    fv = gpu.flip(img, 1)
    assert np.array_equal(fv, img[::-1])
    r90 = gpu.rotate(img, 90)
    assert np.array_equal(r90, np.rot90(img, -1))                     # clockwise
    assert np.array_equal(gpu.rotate(r90, 270), img)                  # 90 then 270 = identity
    assert np.array_equal(gpu.rotate(img, 180), img[::-1, ::-1])
    g = gpu.gray(img)
    assert np.array_equal(gpu.gray(fh), g[:, ::-1])                   # gray commutes with flips
    assert int(g.astype(np.uint64).sum()) == int((img.astype(np.uint32).sum(axis=2) // 3).sum())


def test_resize_and_rotate_mid_size(gpu, orc):
    """1920x1080 (config 5 frame): full resize and rotate against the oracle."""
    img = P.lcg(1920, 1080, 0xC0FFEE ^ 5)
    for kw in (dict(resize_w=960), dict(resize_w=2880)):
        exp = orc.process(img, **kw)
        got = gpu.process(img, **kw)
        assert got[1:] == exp[1:] and np.array_equal(got[0], exp[0]), kw
    small = img[:540, :960].copy()
    exp = orc.rotate(small, 30)
    assert np.array_equal(gpu.rotate(small, 30), exp)


def test_cli_against_reference_binary(gpu, tmp_path):
    """ppmx-b200 and the compiled reference CLI on the same file: identical .out files."""
    import imageprocessingtools_b200.ppmx as pp
    if not os.path.exists(oracle.REF_CLI):
        pytest.skip("compiled reference CLI (oracle/_ref) did not travel")
    img = P.lcg(64, 40, 11)
    for args in (["-gray"], ["-mono"], ["-fv"], ["-fh"], ["-r90"], ["-r30"], ["-w100"], ["-w30", "-r90", "-gray", "-fv"],
                 ["-gray", "-fh"], ["-r270", "-mono", "-fh"]):
        a, b = str(tmp_path / "a.ppm"), str(tmp_path / "b.ppm")
        oracle.write_p6(a, img, comment="made by test")
        oracle.write_p6(b, img, comment="made by test")
        rc_ref, _ = oracle.ref_cli(args, a)
        p = subprocess.run([pp.CLI] + args + [b], capture_output=True, text=True)
        assert rc_ref == 0 and p.returncode == 0, (args, p.stdout)
        assert open(a + ".out", "rb").read() == open(b + ".out", "rb").read(), args
    # error behaviour: no operator -> "no data to write", exit 255
    p = subprocess.run([pp.CLI, str(tmp_path / "b.ppm")], capture_output=True, text=True)
    assert p.returncode == 255 and "no data to write" in p.stdout


# ---- extensions: self-oracle only, parity UNPINNED (no reference counterpart) ----------------

KERNELS = {
    "blur3": (np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]]), 16, 0),
    "sharpen3": (np.array([[0, -1, 0], [-1, 5, -1], [0, -1, 0]]), 1, 0),
    "edge3": (np.array([[-1, -1, -1], [-1, 8, -1], [-1, -1, -1]]), 1, 0),
    "box7": (np.ones((7, 7), np.int64), 49, 0),
    "emboss5": (np.array([[-2, -1, 0, 0, 0], [-1, -1, 0, 0, 0], [0, 0, 1, 0, 0], [0, 0, 0, 1, 1], [0, 0, 0, 1, 2]]), 1, 128),
}


KERNELS["mix3_div8_biasneg"] = (np.array([[-1, 2, -1], [2, 4, 2], [-1, 2, -1]]), 8, -5)   # power-of-two divisor + bias
KERNELS["neg7_div64_bias"] = (np.tile(np.array([[3, -2, 1, 4, 1, -2, 3]]), (7, 1)), 64, 17)
KERNELS["gauss7"] = (np.outer([1, 6, 15, 20, 15, 6, 1], [1, 6, 15, 20, 15, 6, 1]), 4096, 0)       # rank 1, v fits int8
KERNELS["sep5_signed"] = (np.outer([2, -4, 6, -4, 2], [-1, 3, 5, 3, -1]), 32, 9)                    # rank 1 with signs + gcd
KERNELS["sobel3"] = (np.outer([1, 2, 1], [-1, 0, 1]), 1, 128)                                        # rank 1, zero column
KERNELS["box5"] = (np.ones((5, 5), np.int64), 25, 0)                                                  # running-sum box kernel
KERNELS["sep5_asym"] = (np.outer([1, 2, 3, 2, 1], [-1, -2, 0, 2, 1]), 8, 128)                      # rank 1, v not symmetric (Sobel-like)
KERNELS["sep7_div3"] = (np.outer([1, 1, 2, 3, 2, 1, 1], [1, 2, 3, 4, 3, 2, 1]), 3, 0)               # rank 1, general divisor, saturates
KERNELS["sep7_div1_neg"] = (np.outer([0, -1, 2, -3, 2, -1, 0], [1, 0, -2, 3, -2, 0, 1]), 1, 100)    # rank 1, zeros and signs, div 1
KERNELS["gauss5"] = (np.outer([1, 4, 6, 4, 1], [1, 4, 6, 4, 1]), 256, 0)
KERNELS["wide3"] = (np.array([[-300, 200, 129], [-129, 5000, -128], [127, 128, -16320]]), 997, 40)   # 3x3 beyond int8: split dp4a chains
KERNELS["wide3_limits"] = (np.array([[16319, -16320, 0], [0, 16320, 1], [-64, 63, 128]]), 30011, 0)  # the split's limits; 16320 is one beyond (generic kernel)
KERNELS["wide3_max"] = (np.array([[16319, -16320, 16319], [-16320, 16319, -16320], [64, -65, 0]]), 8192, 1)
KERNELS["wide3_pow2"] = (np.array([[256, 512, 256], [512, 1024, 512], [256, 512, 256]]), 4096, 0)
KERNELS["box9"] = (np.ones((9, 9), np.int64), 81, 0)
KERNELS["box11"] = (np.ones((11, 11), np.int64), 121, 0)
KERNELS["box11_bias"] = (np.ones((11, 11), np.int64), 121, 3)                                          # constants overflow 32 bits: generic kernel
KERNELS["box7_twos_bias"] = (2 * np.ones((7, 7), np.int64), 196, 64)                                  # box with a factor and a bias
KERNELS["box7_sat"] = (np.ones((7, 7), np.int64), 40, 0)                                              # would pass 255: not the box kernel
KERNELS["blur3_div256"] = (np.array([[16, 32, 16], [32, 64, 32], [16, 32, 16]]), 256, 0)   # normalised, power-of-two divisor:
KERNELS["cross3_div4"] = (np.array([[0, 1, 0], [1, 0, 1], [0, 1, 0]]), 4, 0)                  #   the quotient is byte 1 of the scaled sum
KERNELS["pair3_div2"] = (np.array([[0, 0, 0], [0, 1, 1], [0, 0, 0]]), 2, 0)                   # scaled coefficient 128 > 127: shift form
KERNELS["big5"] = (np.array([[300, -200, 0, 5, 1]] * 5), 7, -3)  # coefficients beyond int8: generic kernel
_rng7 = np.random.RandomState(7)
KERNELS["dense7_mix"] = (_rng7.randint(-128, 128, (7, 7)), 173, 5)          # dense 7x7, full signed-byte range: vertical-word strip kernel
KERNELS["dense7_pos_pow2"] = (_rng7.randint(0, 6, (7, 7)), 128, 0)           # dense 7x7, power-of-two divisor (shift form)
KERNELS["dense5_div1"] = (_rng7.randint(-3, 4, (5, 5)), 1, 64)               # dense 5x5, div 1
KERNELS["sep7_u16_edge"] = (np.outer([36, 36, 37, 37, 37, 37, 37], [1, 2, 3, 4, 3, 2, 1]), 4112, 0)  # rank 1, column sums up to 65535: the 16-bit limit
KERNELS["sep7_u16_over"] = (np.outer([36, 37, 37, 37, 37, 37, 37], [1, 2, 3, 4, 3, 2, 1]), 4128, 0)  # one beyond it: the 32-bit rank-1 kernel
_g7, _g5 = np.outer([1, 6, 15, 20, 15, 6, 1], [1, 6, 15, 20, 15, 6, 1]), np.outer([1, 4, 6, 4, 1], [1, 4, 6, 4, 1])
KERNELS["unsharp7"] = (-_g7 + 2 * 4096 * (np.arange(49).reshape(7, 7) == 24), 4096, 0)      # rank 1 + centre: 2 I - G (7x7 sharpen)
KERNELS["unsharp5_bias"] = (-3 * _g5 + 4 * 256 * (np.arange(25).reshape(5, 5) == 12), 256, -7)  # 4 I - 3 G, bias
KERNELS["blurid5_div3"] = (_g5 + 77 * (np.arange(25).reshape(5, 5) == 12), 333, 0)            # G + 77 I, general divisor
_d7w = _rng7.randint(-300, 301, (7, 7)) * (np.arange(49).reshape(7, 7) % 5 != 0)
_d7w[3, 3], _d7w[0, 6] = 16319, -16320                                    # the limits of the split (16320 would need hi = 128)
KERNELS["dense7_wide"] = (_d7w, 9973, 3)                                   # dense 7x7 beyond int8: two chains of dot products
KERNELS["dense5_wide"] = (_rng7.randint(-2000, 2001, (5, 5)), 4096, 0)   # dense 5x5, every coefficient split
_rng9 = np.random.RandomState(9)
KERNELS["gauss9"] = (np.outer([1, 8, 28, 56, 70, 56, 28, 8, 1], [1, 8, 28, 56, 70, 56, 28, 8, 1]) // 64, 1044, 0)  # 9x9 blur (dense after the floor)
KERNELS["dense11_mix"] = (_rng9.randint(-128, 128, (11, 11)), 977, -2)
KERNELS["dense13_pow2"] = (_rng9.randint(-3, 9, (13, 13)), 512, 0)
KERNELS["dense15_div1"] = (_rng9.randint(-2, 3, (15, 15)) * (_rng9.randint(0, 4, (15, 15)) == 0), 1, 90)
_b9 = [1, 8, 28, 56, 70, 56, 28, 8, 1]
KERNELS["binom9"] = (np.outer(_b9, _b9), 65536, 0)                                  # rank-1 9x9: 16-bit column sums, byte-2 quotient
KERNELS["gauss11_div"] = (np.outer([1, 2, 4, 7, 10, 11, 10, 7, 4, 2, 1], [1, 3, 6, 10, 14, 16, 14, 10, 6, 3, 1]), 4957, 1)  # rank-1 11x11, general divisor
KERNELS["gauss13_asym"] = (np.outer([1, 1, 2, 3, 5, 8, 13, 8, 5, 3, 2, 1, 1], [0, 1, 2, 4, 8, 16, 32, 16, 8, 4, 2, 1, 0]), 4096, 0)
KERNELS["sep15_sat"] = (np.outer([1] * 15, [17] * 15), 2000, 0)                     # box would apply only up to 11; rank-1, saturates
KERNELS["binom11"] = (np.outer([1, 10, 45, 120, 210, 252, 210, 120, 45, 10, 1], [1, 10, 45, 120, 210, 252, 210, 120, 45, 10, 1]), 1 << 20, 0)  # column sums beyond 16 bits: not that kernel
KERNELS["sep5_s16_edge"] = (np.outer([-25, 26, -26, 26, -25], [1, -2, 3, -2, 1]), 9, 128)             # signed rank 1, |column sums| up to 32640


def test_extension_conv(gpu, orc):
    for (w, h) in [(1, 1), (2, 3), (5, 4), (37, 23), (64, 48), (130, 70), (301, 211), (16, 1), (16, 40), (128, 32),
                   (144, 37), (256, 64), (400, 33), (64, 200), (48, 131), (2064, 67)]:
        for pname in ("lcg", "checker", "mixed"):
            img = P.all_patterns(w, h)[pname]
            for kname, (coef, div, bias) in KERNELS.items():
                assert np.array_equal(gpu.conv(img, coef, div, bias), orc.conv(img, coef, div, bias)), (w, h, pname, kname)


def test_extension_conv3_any_width(gpu_tuning, orc):
    """3x3 at widths that are no multiple of 16 (ppmx_conv_ua.cu): the row's end at every position of a 16-byte chunk,
    rows of one warp, several warps and several CTAs, strips cut short by the raster's end; the first any-alignment kernel
    (tuning variant 20) gives the same bytes."""
    gpu = gpu_tuning
    widths = list(range(16, 33)) + [47, 63, 65, 170, 171, 172, 173, 341, 683, 684, 685, 1367, 2731]
    try:
        for v in (0, 20):
            gpu.set_tuning("variant", v)
            for w in widths:
                h = max(9, -(-4096 // w) + (w % 4))
                img = P.lcg(w, h, 1000 + w)
                for kname in ("blur3", "edge3", "wide3", "mix3_div8_biasneg"):
                    coef, div, bias = KERNELS[kname]
                    assert np.array_equal(gpu.conv(img, coef, div, bias), orc.conv(img, coef, div, bias)), (v, w, h, kname)
    finally:
        gpu.set_tuning("variant", 0)


def test_extension_conv_large_k_odd_widths(gpu, orc):
    """9x9 .. 15x15 at widths that are no multiple of 16: the padded copy (up to 22 mirrored pixels behind every row) in front of
    the shared-memory vertical-word kernels."""
    for (w, h) in [(130, 70), (301, 41), (1000, 33), (77, 60)]:
        img = P.lcg(w, h, 4000 + w)
        for kname in ("binom9", "gauss11_div", "gauss13_asym", "dense15_div1", "dense11_mix", "sep15_sat", "box11"):
            coef, div, bias = KERNELS[kname]
            assert np.array_equal(gpu.conv(img, coef, div, bias), orc.conv(img, coef, div, bias)), (w, h, kname)


def test_extension_conv_fuzz(gpu, orc):
    """Seeded random sweep over the convolution dispatch: every kernel size 3..15, widths with and without the 16-byte layout,
    coefficient classes that select each vector kernel (box, rank-1 unsigned / signed / with a centre term, dense within and beyond a
    signed byte, sparse, oversized: the scalar kernel), divisors 1 / power of two / normalising / arbitrary, biases; the whole
    raster against the self-oracle, and for a third of the cases cut into row bands with halo rows."""
    import torch
    import imageprocessingtools_b200.ppmx as pp
    rng = np.random.RandomState(20261018)
    dev = torch.device("cuda", 0)

    def coefs(k, cls):
        if cls == "box":
            return np.full((k, k), rng.randint(1, 4), np.int64)
        if cls in ("r1u", "r1s", "r1c"):
            lo = 0 if cls == "r1u" else -6
            u, v = rng.randint(lo, 12, k), rng.randint(lo, 12, k)
            u[k // 2] += 1
            v[k // 2] += 1
            c = np.outer(u, v).astype(np.int64)
            if cls == "r1c":
                c[k // 2, k // 2] += rng.randint(-3000, 9000)
            return c
        if cls == "dense":
            return rng.randint(-128, 128, (k, k)).astype(np.int64)
        if cls == "wide":
            return rng.randint(-16320, 16320, (k, k)).astype(np.int64)
        if cls == "sparse":
            return (rng.randint(-9, 10, (k, k)) * (rng.randint(0, 5, (k, k)) == 0)).astype(np.int64)
        return rng.randint(-40000, 40000, (k, k)).astype(np.int64)  # "huge"

    n = 0
    for it in range(260):
        k = int(rng.choice([3, 3, 5, 5, 7, 7, 9, 11, 13, 15]))
        cls = str(rng.choice(["box", "r1u", "r1u", "r1s", "r1c", "dense", "dense", "wide", "sparse", "huge"]))
        w = int(rng.randint(1, 45)) * 16 if rng.randint(0, 3) else int(rng.randint(16, 700))
        h = int(rng.randint(1, 90))
        coef = coefs(k, cls)
        tot = int(coef.sum())
        div = int(rng.choice([1, 2 ** int(rng.randint(1, 13)), max(1, abs(tot)), int(rng.randint(1, 5000))]))
        bias = int(rng.choice([0, 0, rng.randint(-40, 200)]))
        if 255 * int(np.abs(coef).sum()) >= 2 ** 30:
            continue
        img = P.lcg(w, h, 7000 + it)
        exp = orc.conv(img, coef, div, bias)
        assert np.array_equal(gpu.conv(img, coef, div, bias), exp), (it, k, cls, w, h, div, bias)
        n += 1
        r = k // 2
        if it % 3 == 0 and h >= 2 * r + 2:  # two or three row bands with halo rows through the band pointers
            cuts = sorted(set([0, h] + [int(c) for c in rng.randint(r, h - r + 1, 2)]))
            if any(b - a < r for a, b in zip(cuts[:-1], cuts[1:])):
                continue
            bands = [torch.from_numpy(img[a:b].copy()).to(dev) for a, b in zip(cuts[:-1], cuts[1:])]
            outs = [torch.zeros_like(b) for b in bands]
            op = gpu.conv_op(coef, div, bias)
            for i, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
                band = pp.PpmxBand(full_h=h, y0=a, halo=r)
                if i > 0:
                    band.d_top = bands[i - 1].data_ptr() + (bands[i - 1].shape[0] - r) * w * 3
                if i < len(bands) - 1:
                    band.d_bottom = bands[i + 1].data_ptr()
                gpu.launch(op, bands[i].data_ptr(), w, b - a, pp.LAYOUT_RGB8, outs[i].data_ptr(), band)
            torch.cuda.synchronize()
            got = np.concatenate([o.cpu().numpy() for o in outs], axis=0)
            assert np.array_equal(got, exp), ("bands", it, k, cls, w, h, cuts)
    assert n > 200


def test_extension_conv_row_bands(gpu, orc):
    """A raster cut into row bands, each convolved separately with halo rows read through the band
    pointers (here: the neighbour band in the same HBM), equals the whole-raster result."""
    import torch
    import imageprocessingtools_b200.ppmx as pp
    for (w, h, k, cuts, kname) in [(128, 96, 3, [0, 32, 64, 96], None), (128, 96, 7, [0, 24, 48, 72, 96], None),
                                   (64, 50, 5, [0, 7, 13, 50], "emboss5"), (37, 23, 3, [0, 10, 23], None),
                                   (301, 60, 3, [0, 13, 30, 31, 60], "edge3"), (1000, 21, 3, [0, 2, 9, 21], "blur3"),   # any-width strip kernel
                                   (2048, 70, 7, [0, 17, 40, 70], "dense7_mix"), (2048, 70, 5, [0, 35, 70], "sep5_s16_edge"),
                                   (1024, 45, 7, [0, 16, 33, 45], "unsharp7"), (512, 90, 9, [0, 30, 61, 90], "gauss9"),
                                   (256, 64, 15, [0, 7, 40, 64], "dense15_div1"), (512, 90, 9, [0, 30, 61, 90], "binom9"),
                                   (256, 64, 13, [0, 6, 40, 64], "gauss13_asym"),
                                   (256, 40, 7, [0, 3, 6, 40], None),
                                   (128, 96, 5, [0, 32, 64, 96], "gauss5"), (256, 40, 7, [0, 3, 6, 40], "gauss7"),   # rank-1 kernel
                                   (64, 50, 5, [0, 7, 13, 50], "sep5_asym"), (128, 200, 3, [0, 67, 134, 200], "edge3"),
                                   (128, 96, 5, [0, 24, 48, 72, 96], "box5"), (64, 64, 11, [0, 16, 37, 64], "box11")]:
        coef, div, bias = (np.ones((k, k), np.int64), k * k, 0) if kname is None else KERNELS[kname]
        img = P.lcg(w, h, 77)
        exp = orc.conv(img, coef, div, bias)
        r = k // 2
        dev = torch.device("cuda", 0)
        bands = [torch.from_numpy(img[a:b].copy()).to(dev) for a, b in zip(cuts[:-1], cuts[1:])]
        outs = [torch.zeros_like(b) for b in bands]
        op = gpu.conv_op(coef, div, bias)
        for i, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
            band = pp.PpmxBand(full_h=h, y0=a, halo=r)
            # halo rows: the last r rows of the band above, the first r rows of the band below
            if i > 0:
                prev = bands[i - 1]
                assert prev.shape[0] >= r
                band.d_top = prev.data_ptr() + (prev.shape[0] - r) * w * 3
            if i < len(bands) - 1:
                assert bands[i + 1].shape[0] >= r
                band.d_bottom = bands[i + 1].data_ptr()
            gpu.launch(op, bands[i].data_ptr(), w, b - a, pp.LAYOUT_RGB8, outs[i].data_ptr(), band)
        torch.cuda.synchronize()
        got = np.concatenate([o.cpu().numpy() for o in outs], axis=0)
        assert np.array_equal(got, exp), (w, h, k)
    # a band whose halo is thinner than k/2, or missing, is refused instead of read out of bounds
    mid = torch.zeros((32, 128, 3), dtype=torch.uint8, device=dev)
    op7 = gpu.conv_op(np.ones((7, 7), np.int64), 49, 0)
    for bad in (pp.PpmxBand(full_h=96, y0=32, halo=1, d_top=mid.data_ptr(), d_bottom=mid.data_ptr()),
                pp.PpmxBand(full_h=96, y0=32, halo=3, d_top=mid.data_ptr()),
                pp.PpmxBand(full_h=40, y0=32, halo=3, d_top=mid.data_ptr())):
        with pytest.raises(pp.PpmxError):
            gpu.launch(op7, mid.data_ptr(), 128, 32, pp.LAYOUT_RGB8, mid.data_ptr(), bad)


def test_imresize_height_pass_row_bands(gpu, orc):
    """ref:820-838 cut into row bands: every band owns a slice of the SOURCE rows and produces a slice of
    the OUTPUT rows; taps that fall into a neighbour's rows are read through the halo pointers."""
    import torch
    import imageprocessingtools_b200.ppmx as pp
    dev = torch.device("cuda", 0)
    for (w, h, new_h, n) in [(64, 96, 144, 3), (64, 96, 40, 2), (48, 50, 75, 4), (37, 60, 90, 2), (128, 64, 64, 2)]:
        img = P.lcg(w, h, 123)
        wt, ix = gpu.calc_contributions(h, new_h, float(new_h) / h)
        exp = orc.imresize(img, new_h, 0, wt, ix)
        op = gpu.imresize_op(new_h, 0, wt, ix)
        tables = gpu.tables_upload(op)
        in_bands = [pp.band_plan(h, n, r, 1) for r in range(n)]
        out_bands = [pp.band_plan(new_h, n, r, 1) for r in range(n)]
        srcs = [torch.from_numpy(img[a:a + c].copy()).to(dev) for a, c in in_bands]
        got = []
        for r in range(n):
            (y0, rows), (oy0, orows) = in_bands[r], out_bands[r]
            need = ix[oy0:oy0 + orows]
            halo = int(max(0, y0 - need.min(), need.max() - (y0 + rows - 1))) if orows else 0
            band = pp.PpmxBand(full_h=h, y0=y0, halo=halo, out_y0=oy0, out_rows=orows)
            if r > 0 and halo:
                assert in_bands[r - 1][1] >= halo
                band.d_top = srcs[r - 1].data_ptr() + (in_bands[r - 1][1] - halo) * w * 3
            if r < n - 1 and halo:
                assert in_bands[r + 1][1] >= halo
                band.d_bottom = srcs[r + 1].data_ptr()
            dst = torch.zeros((max(orows, 1), w, 3), dtype=torch.uint8, device=dev)
            gpu.launch(op, srcs[r].data_ptr(), w, rows, pp.LAYOUT_RGB8, dst.data_ptr(), band, 0, tables)
            torch.cuda.synchronize()
            got.append(dst.cpu().numpy()[:orows])
        gpu.tables_free(tables)
        assert np.array_equal(np.concatenate(got, axis=0), exp), (w, h, new_h, n)


def test_mono_row_bands_keep_bayer_phase(gpu, orc):
    import torch
    import imageprocessingtools_b200.ppmx as pp
    w, h = 64, 23
    img = P.bayer_edges(w, h)
    exp = orc.pack_pbm(orc.mono(img))
    dev = torch.device("cuda", 0)
    got = []
    for a, b in [(0, 5), (5, 6), (6, 23)]:  # cuts that are not multiples of 4
        src = torch.from_numpy(img[a:b].copy()).to(dev)
        dst = torch.zeros(((b - a) * (w // 8),), dtype=torch.uint8, device=dev)
        gpu.launch(pp.PpmxOp(kind=pp.OP_MONO_BITS), src.data_ptr(), w, b - a, pp.LAYOUT_RGB8, dst.data_ptr(),
                   pp.PpmxBand(full_h=h, y0=a))
        torch.cuda.synchronize()
        got.append(dst.cpu().numpy())
    assert np.array_equal(np.concatenate(got), exp)


def test_batch_and_gray_hist_chain(gpu, orc):
    """ppmx_gpu_apply_batch: rasters go round-robin over the context's streams; results stay per raster."""
    import ctypes as C
    import imageprocessingtools_b200.ppmx as pp
    imgs = np.stack([P.lcg(96, 40, 100 + i) for i in range(7)])
    out, w, h, ft = gpu.apply_ops(imgs, [pp.PpmxOp(kind=pp.OP_FLIP, flip_direction=0, renew_before=0)])
    assert (w, h, ft) == (96, 40, 0)
    for i in range(7):
        assert np.array_equal(out[i].reshape(40, 96, 3), orc.flip(imgs[i], 0))
    ph = pp._PlanHolder(resize_w=60, angle=90, mono=True, w=96, h=40)
    ops = [ph.plan.ops[i] for i in range(ph.plan.nops)]
    out, w, h, ft = gpu.apply_ops(imgs, ops)
    for i in range(7):
        exp, ew, eh, eft = orc.process(imgs[i], resize_w=60, angle=90, mono=True)
        assert (w, h, ft) == (ew, eh, eft) and np.array_equal(out[i], exp), i
    ph.close()
    bins = np.zeros((7, 256), np.uint64)
    op = pp.PpmxOp(kind=pp.OP_GRAY_HIST, hist_out=bins.ctypes.data_as(C.POINTER(C.c_uint64)))
    out, w, h, ft = gpu.apply_ops(imgs, [op])
    assert ft == 1
    for i in range(7):
        assert np.array_equal(out[i].reshape(40, 96), orc.gray(imgs[i]))
        assert np.array_equal(bins[i], orc.hist_gray(imgs[i]))


def test_extension_histogram(gpu, orc):
    for (w, h) in [(1, 1), (13, 7), (64, 64), (301, 211), (1024, 1024)]:
        for pname in ("lcg", "c200", "bayer"):
            img = P.all_patterns(w, h)[pname]
            exp = orc.hist_gray(img)
            assert np.array_equal(gpu.hist_gray(img), exp), (w, h, pname)
            g, bins = gpu.gray_hist(img)
            assert np.array_equal(bins, exp) and np.array_equal(g, orc.gray(img)), (w, h, pname)
            assert int(bins.sum()) == w * h


def test_no_out_of_bounds_writes(gpu, orc):
    """compute-sanitizer is closed on this GPU pool, so out-of-bounds WRITES are checked by hand: every
    operator writes into the middle of a canary-filled device buffer (raw launch on caller-owned memory)
    and the canaries on both sides must survive.  Outputs are compared with the oracle as well."""
    import torch
    import imageprocessingtools_b200.ppmx as pp
    dev = torch.device("cuda", 0)
    PAD = 4096

    def run(op, img, out_bytes, layout=pp.LAYOUT_RGB8, tables=0, src_off=0, dst_off=0):
        h, w = img.shape[0], img.shape[1]
        flat = np.ascontiguousarray(img).reshape(-1)
        srcbuf = torch.zeros((flat.size + 64,), dtype=torch.uint8, device=dev)
        srcbuf[src_off:src_off + flat.size] = torch.from_numpy(flat).to(dev)
        buf = torch.full((PAD + out_bytes + PAD,), 0xA5, dtype=torch.uint8, device=dev)
        gpu.launch(op, srcbuf.data_ptr() + src_off, w, h, layout, buf.data_ptr() + PAD + dst_off, None, 0, tables)
        torch.cuda.synchronize()
        host = buf.cpu().numpy()
        assert (host[:PAD + dst_off] == 0xA5).all() and (host[PAD + dst_off + out_bytes:] == 0xA5).all(), "canary overwritten"
        return host[PAD + dst_off:PAD + dst_off + out_bytes]

    # rasters large enough for the tile / row / unaligned-strip kernels, at odd sizes AND odd pointers on both sides
    for (w, h) in [(130, 70), (257, 33), (1000, 9), (67, 130), (128, 80), (640, 24)]:
        img = P.lcg(w, h, 19)
        for (so, do) in [(0, 0), (5, 3), (1, 15), (16, 8)]:
            kw = dict(src_off=so, dst_off=do)
            assert np.array_equal(run(pp.PpmxOp(kind=pp.OP_GRAY), img, w * h, **kw), orc.gray(img).reshape(-1)), (w, h, so, do)
            assert np.array_equal(run(pp.PpmxOp(kind=pp.OP_MONO_BITS), img, ((w + 7) // 8) * h, **kw), orc.pack_pbm(orc.mono(img)))
            assert np.array_equal(run(pp.PpmxOp(kind=pp.OP_EXTRACT_R), img, w * h, **kw), img[:, :, 0].reshape(-1))
            for d in (0, 1):
                assert np.array_equal(run(pp.PpmxOp(kind=pp.OP_FLIP, flip_direction=d), img, w * h * 3, **kw),
                                      orc.flip(img, d).reshape(-1)), (w, h, so, do, d)
            for a in (90, 180, 270):
                exp = orc.rotate(img, a)
                assert np.array_equal(run(gpu.rotate_op(a, w, h), img, exp.size, **kw), exp.reshape(-1)), (w, h, so, do, a)
            for kname in ("blur3", "edge3", "wide3", "box7"):
                coef, div, bias = KERNELS[kname]
                assert np.array_equal(run(gpu.conv_op(coef, div, bias), img, w * h * 3, **kw),
                                      orc.conv(img, coef, div, bias).reshape(-1)), (w, h, so, do, kname)

    for (w, h) in [(1, 1), (5, 3), (17, 9), (33, 7), (48, 16), (64, 64), (100, 37), (128, 80), (256, 16)]:
        img = P.lcg(w, h, 9)
        assert np.array_equal(run(pp.PpmxOp(kind=pp.OP_GRAY), img, w * h), orc.gray(img).reshape(-1))
        assert np.array_equal(run(pp.PpmxOp(kind=pp.OP_MONO_BITS), img, ((w + 7) // 8) * h), orc.pack_pbm(orc.mono(img)))
        assert np.array_equal(run(pp.PpmxOp(kind=pp.OP_MONO), img, w * h), orc.mono(img).reshape(-1))
        for d in (0, 1):
            assert np.array_equal(run(pp.PpmxOp(kind=pp.OP_FLIP, flip_direction=d), img, w * h * 3),
                                  orc.flip(img, d).reshape(-1))
        for a in (90, 180, 270, 31):
            op = gpu.rotate_op(a, w, h)
            exp = orc.rotate(img, a)
            assert np.array_equal(run(op, img, exp.size), exp.reshape(-1)), (w, h, a)
        for new in (max(1, w // 2), w + 3):
            for dim, n_in in ((1, w), (0, h)):
                wt, ix = gpu.calc_contributions(n_in, new, float(new) / n_in)
                op = gpu.imresize_op(new, dim, wt, ix)
                t = gpu.tables_upload(op)
                exp = orc.imresize(img, new, dim, wt, ix)
                assert np.array_equal(run(op, img, exp.size, tables=t), exp.reshape(-1)), (w, h, new, dim)
                gpu.tables_free(t)
        for k in (3, 7):
            coef = np.ones((k, k), np.int64)
            assert np.array_equal(run(gpu.conv_op(coef, k * k, 0), img, w * h * 3), orc.conv(img, coef, k * k, 0).reshape(-1))


def test_full_size_bicubic_4096(gpu, orc):
    """The reference's real stencil work at BASELINE size: -w6144 / -w2048 and -r30 on a 4096x4096 raster,
    compared directly (the oracle needs a few seconds each), plus a flat image for the truncation speckle."""
    img = P.lcg(4096, 4096, 0xC0FFEE ^ 3)
    for new_w in (6144, 2048):
        exp = orc.process(img, resize_w=new_w)
        got = gpu.process(img, resize_w=new_w)
        assert got[1:] == exp[1:] and np.array_equal(got[0], exp[0]), new_w
    exp = orc.rotate(img, 30)
    assert np.array_equal(gpu.rotate(img, 30), exp)
    flat = P.const(2048, 2048, 200)
    assert np.array_equal(gpu.rotate(flat, 77), orc.rotate(flat, 77))


def test_tuning_variants_are_bit_identical(gpu_tuning, orc):
    """Every alternative kernel kept for benchmarking must give the default's bytes."""
    gpu = gpu_tuning
    img = P.lcg(256, 192, 31)
    wt, ix = gpu.calc_contributions(192, 288, 1.5)
    wt2, ix2 = gpu.calc_contributions(256, 128, 0.5)
    coef = np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]])
    try:
        for v in (0, 1, 2, 3, 4, 5, 6, 7, 8):
            for pdl in (1, 0):
                gpu.set_tuning("variant", v)
                gpu.set_tuning("pdl", pdl)
                assert np.array_equal(gpu.gray(img), orc.gray(img)), (v, pdl)
                g, bins = gpu.gray_hist(img)
                assert np.array_equal(g, orc.gray(img)) and np.array_equal(bins, orc.hist_gray(img)), (v, pdl)
                assert np.array_equal(gpu.rotate(img, 90), orc.rotate(img, 90)), (v, pdl)
                assert np.array_equal(gpu.rotate(img, 33), orc.rotate(img, 33)), (v, pdl)
                assert np.array_equal(gpu.imresize(img, 288, 0, wt, ix), orc.imresize(img, 288, 0, wt, ix)), (v, pdl)
                assert np.array_equal(gpu.imresize(img, 128, 1, wt2, ix2), orc.imresize(img, 128, 1, wt2, ix2)), (v, pdl)
                assert np.array_equal(gpu.conv(img, coef, 16, 0), orc.conv(img, coef, 16, 0)), (v, pdl)
    finally:
        gpu.set_tuning("variant", 0)
        gpu.set_tuning("pdl", 1)


def test_extension_conv3_strip_variants(gpu_tuning, orc):
    """3x3: the strip kernel (default; other strip heights / CTA sizes = variants 9-13; variant 8 = without the
    scaled-coefficient byte extraction) and the row-wise planar kernel (variant 7) give the self-oracle's bytes,
    on flat extremes too (saturation at both ends)."""
    gpu = gpu_tuning
    imgs = [P.lcg(64, 200, 5), P.const(64, 131, 255), P.const(32, 70, 0), P.all_patterns(48, 133)["checker"]]
    names = ("blur3", "sharpen3", "edge3", "mix3_div8_biasneg", "sobel3", "wide3", "wide3_pow2")
    exp = {(i, k): orc.conv(img, *KERNELS[k]) for i, img in enumerate(imgs) for k in names}
    exp_big = orc.conv(imgs[0], np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]]) * 4, 64, 0)  # scaled coefficients would pass 127
    try:
        for v in (0, 7, 8, 9, 10, 11, 12, 13):
            gpu.set_tuning("variant", v)
            for i, img in enumerate(imgs):
                for k in names:
                    assert np.array_equal(gpu.conv(img, *KERNELS[k]), exp[(i, k)]), (v, i, k)
            assert np.array_equal(gpu.conv(imgs[0], np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]]) * 4, 64, 0), exp_big), v
    finally:
        gpu.set_tuning("variant", 0)


def test_extension_conv_vertical_word_kernels(gpu_tuning, orc):
    """5x5 / 7x7 on the vertical-word strip kernels (ppmx_conv_sep.cu): rank-1 with 16-bit column sums (unsigned, scaled
    "byte 2" quotient, signed) and dense, at sizes with many strips, inner and edge warps and an odd last row; the strip
    geometries of the tuning build (variants 15-18) and the older kernels (variant 14) give the same bytes."""
    gpu = gpu_tuning
    names = ("gauss7", "gauss5", "sep5_signed", "sep7_div3", "sep7_div1_neg", "sep7_u16_edge", "sep7_u16_over", "sep5_s16_edge",
             "dense7_mix", "dense7_pos_pow2", "dense5_div1", "emboss5", "neg7_div64_bias", "box7_sat", "unsharp7", "unsharp5_bias",
             "blurid5_div3", "dense7_wide", "dense5_wide", "gauss9", "dense11_mix", "dense13_pow2", "dense15_div1", "binom9", "gauss11_div", "gauss13_asym",
             "sep15_sat", "binom11")
    imgs = [P.lcg(2048, 301, 21), P.const(496, 70, 255), P.all_patterns(1008, 37)["mixed"]]
    exp = {(i, k): orc.conv(img, *KERNELS[k]) for i, img in enumerate(imgs) for k in names}
    try:
        for v in (0, 14, 15, 16, 17, 18):
            gpu.set_tuning("variant", v)
            for i, img in enumerate(imgs):
                for k in names:
                    assert np.array_equal(gpu.conv(img, *KERNELS[k]), exp[(i, k)]), (v, i, k)
    finally:
        gpu.set_tuning("variant", 0)


def test_extension_full_size_conv_levels_8192(gpu, orc):
    """BASELINE config 3 size (8192x8192): the strip and box kernels against the self-oracle directly (the C loops
    need a few seconds per filter), the levels table against numpy indexing, and the identity filter."""
    img = P.lcg(8192, 8192, 0xC0FFEE ^ 3)
    for kname in ("blur3", "edge3", "box7"):
        coef, div, bias = KERNELS[kname]
        assert np.array_equal(gpu.conv(img, coef, div, bias), orc.conv(img, coef, div, bias)), kname
    ident = np.zeros((3, 3), np.int64)
    ident[1, 1] = 1
    assert np.array_equal(gpu.conv(img, ident, 1, 0), img)
    lut = gpu.levels_lut_linear(16, 235)
    assert np.array_equal(gpu.levels(img, lut), lut[img])


def test_extension_cli_conv_presets(gpu, orc, tmp_path):
    """EXTENSION flags of ppmx-b200 (-blur, -blur7, -sharpen, -edge; the reference rejects them): the stage
    sits after resize/rotate and before gray/mono/flip.  Self-oracle only (parity unpinned)."""
    import imageprocessingtools_b200.ppmx as pp
    img = P.lcg(64, 40, 5)
    presets = {"-blur": KERNELS["blur3"], "-blur7": (np.ones((7, 7), np.int64), 49, 0), "-sharpen": KERNELS["sharpen3"],
               "-edge": KERNELS["edge3"], "-gauss5": KERNELS["gauss5"], "-gauss7": KERNELS["gauss7"], "-sharpen7": KERNELS["unsharp7"],
               "-gauss9": KERNELS["binom9"]}
    path = str(tmp_path / "x.ppm")
    for flag, (coef, div, bias) in presets.items():
        oracle.write_p6(path, img)
        p = subprocess.run([pp.CLI, flag, "-gray", path], capture_output=True, text=True)
        assert p.returncode == 0, p.stdout
        exp = orc.header(oracle.FT_PGM, 64, 40) + orc.gray(orc.conv(img, coef, div, bias)).tobytes()
        assert open(path + ".out", "rb").read() == exp, flag
        p = subprocess.run([pp.CLI, "-r90", flag, "-fh", path], capture_output=True, text=True)
        assert p.returncode == 0, p.stdout
        exp = orc.flip(orc.conv(orc.rotate(img, 90), coef, div, bias), 0)
        assert open(path + ".out", "rb").read() == orc.header(oracle.FT_PPM, 40, 64) + exp.tobytes(), flag
    p = subprocess.run([pp.CLI, "-blur", "-edge", path], capture_output=True, text=True)
    assert p.returncode == 255 and "Duplicate" in p.stdout


def test_extension_levels(gpu, orc, tmp_path):
    """EXTENSION levels (no reference counterpart, self-oracle): every byte through a 256-entry table, on aligned,
    ragged and tiny rasters; tables from the host helper; auto points from the device histogram; the CLI flag."""
    import imageprocessingtools_b200.ppmx as pp
    rng = np.random.default_rng(3)
    luts = [np.arange(256, dtype=np.uint8)[::-1].copy(), rng.integers(0, 256, 256).astype(np.uint8),
            gpu.levels_lut_linear(16, 235), gpu.levels_lut_linear(0, 255), gpu.levels_lut_linear(100, 101)]
    for (w, h) in [(1, 1), (5, 3), (16, 16), (37, 23), (256, 64), (301, 211), (1024, 33), (1024, 700), (1001, 707)]:  # last two: > 1 MB, the lane-column kernel
        img = P.lcg(w, h, 9)
        for lut in luts:
            assert np.array_equal(gpu.levels(img, lut), orc.levels(img, lut)), (w, h)
    # an R8 plane goes through the same kernel
    plane = orc.gray(P.lcg(48, 31, 2))
    d = gpu.upload(plane, pp.LAYOUT_R8)
    out = gpu.op(gpu.levels_op(luts[1]), d, False)
    assert np.array_equal(gpu.download(out, oracle.FT_PGM).reshape(31, 48), orc.levels(plane, luts[1]))
    gpu.release(out)
    gpu.release(d)
    # auto levels: histogram on the device -> points on the host -> table -> levels
    img = (P.lcg(128, 96, 4) // 2 + 40).astype(np.uint8)
    bins = gpu.hist_gray(img)
    lo, hi = gpu.levels_points_from_hist(bins, 5)
    cum = np.cumsum(orc.hist_gray(img))
    assert 40 <= lo < hi <= 167 and cum[lo - 1] * 1000 <= 5 * cum[-1] < cum[lo] * 1000
    assert np.array_equal(gpu.levels(img, gpu.levels_lut_linear(lo, hi)), orc.levels(img, orc.levels_lut_linear(lo, hi)))
    # chain placement: resize/rotate -> conv -> levels -> gray/mono/flip
    got = gpu.process(img, angle=90, conv_preset=1, levels=(50, 150), gray=True)
    blur = orc.conv(orc.rotate(img, 90), *KERNELS["blur3"])
    assert np.array_equal(got[0].reshape(128, 96), orc.gray(orc.levels(blur, orc.levels_lut_linear(50, 150))))
    path = str(tmp_path / "l.ppm")
    oracle.write_p6(path, img)
    p = subprocess.run([pp.CLI, "-levels50-150", "-fv", path], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout
    exp = orc.flip(orc.levels(img, orc.levels_lut_linear(50, 150)), 1)
    assert open(path + ".out", "rb").read() == orc.header(oracle.FT_PPM, 128, 96) + exp.tobytes()
    for bad in ("-levels", "-levels10", "-levels200-100", "-levels0-256", "-levels1-2x"):
        p = subprocess.run([pp.CLI, bad, path], capture_output=True, text=True)
        assert p.returncode == 255 and "levels" in p.stdout, bad


def test_degenerate_rasters_cli(gpu, tmp_path):
    """Zero-height / zero-width rasters and 1-pixel rasters go through the same code in both programs."""
    import imageprocessingtools_b200.ppmx as pp
    if not os.path.exists(oracle.REF_CLI):
        pytest.skip("compiled reference CLI (oracle/_ref) did not travel")
    cases = [(b"P6\n5 0\n255\n", ["-gray"]), (b"P6\n0 4\n255\n", ["-fh"]), (b"P6\n1 1\n255\n\x10\x80\xf0", ["-mono"]),
             (b"P6\n1 1\n255\n\x10\x80\xf0", ["-r90"]), (b"P6\n2 1\n255\n\x01\x02\x03\x04\x05\x06", ["-fv"]),
             (b"P6\n3 3\n255\n" + bytes(range(27)), ["-w7"]), (b"P6\n3 3\n255\n" + bytes(range(27)), ["-r45", "-gray"])]
    for data, args in cases:
        a, b = str(tmp_path / "a.ppm"), str(tmp_path / "b.ppm")
        for p in (a, b):
            open(p, "wb").write(data)
            if os.path.exists(p + ".out"):
                os.remove(p + ".out")
        r = subprocess.run([oracle.REF_CLI] + args + [a], capture_output=True, text=True)
        o = subprocess.run([pp.CLI] + args + [b], capture_output=True, text=True)
        assert r.returncode == o.returncode, (data[:12], args, r.stdout, o.stdout)
        if r.returncode == 0:
            assert open(a + ".out", "rb").read() == open(b + ".out", "rb").read(), (data[:12], args)


def test_operator_level_host_api_stepwise(gpu, tmp_path):
    """the reference's doProcessPPM control flow over the operator-level C
    functions (ppmx_getImageInfo, ppmx_imresize, ppmx_renewBuffer, ppmx_rotate, ppmx_gray, ppmx_mono,
    ppmx_flip, ppmx_putImageToFile), compiled here from tests/c/stepwise.c, must write the same files as
    the one-shot CLI and as the compiled reference."""
    import imageprocessingtools_b200.ppmx as pp
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "stepwise")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-std=gnu99", "-I" + os.path.join(root, "include"), "-o", exe,
                    os.path.join(root, "tests", "c", "stepwise.c"), "-L" + pp.PKG, "-lppmx_host", "-lppmx_gpu", "-lm",
                    "-Wl,-rpath," + pp.PKG], check=True)
    img = P.lcg(64, 40, 21)
    chains = [["-gray"], ["-mono"], ["-fv"], ["-fh"], ["-r90"], ["-r30"], ["-r0"], ["-w100"], ["-w30"], ["-w64"],
              ["-w30", "-r90", "-gray", "-fv"], ["-w100", "-r45", "-mono", "-fh"], ["-gray", "-fh"], ["-mono", "-fv"],
              ["-r270", "-mono", "-fh"], ["-w20", "-fv"]]
    for args in chains:
        a, b, c = (str(tmp_path / n) for n in ("a.ppm", "b.ppm", "c.ppm"))
        for p in (a, b, c):
            oracle.write_p6(p, img)
        r1 = subprocess.run([exe] + args + [a], capture_output=True, text=True)
        r2 = subprocess.run([pp.CLI] + args + [b], capture_output=True, text=True)
        assert r1.returncode == 0 and r2.returncode == 0, (args, r1.stdout, r2.stdout)
        assert open(a + ".out", "rb").read() == open(b + ".out", "rb").read(), args
        if os.path.exists(oracle.REF_CLI):
            rc, _ = oracle.ref_cli(args, c)
            assert rc == 0 and open(c + ".out", "rb").read() == open(a + ".out", "rb").read(), args


# ---- round 2: row parts, row bands, several devices, CUDA graphs ------------------------------------------

PART_CHAINS = [dict(conv_preset=1), dict(conv_preset=2, mono=True, flipv=True), dict(resize_w=300, gray=True),
               dict(conv_preset=6), dict(conv_preset=7, mono=True), dict(conv_preset=8, flipv=True), dict(resize_w=208, conv_preset=5, gray=True),
               dict(resize_w=97, conv_preset=4, fliph=True), dict(flipv=True), dict(angle=180, levels=(16, 235)),
               dict(gray=True, flipv=True), dict(mono=True), dict(resize_w=201, angle=180, conv_preset=3, mono=True, flipv=True),
               dict(angle=90, gray=True), dict(angle=33)]


def _orc_chain(orc, img, resize_w=None, angle=None, gray=False, mono=False, flipv=False, fliph=False, conv_preset=0,
               levels=None):
    """The oracle's chain with the extension stages where ppmx_plan_chain_ext2 puts them (after resize / rotate,
    before gray / mono / flip); chains without -w/-r and with an extension stage behave as if renewBuffer ran."""
    cur = img
    if resize_w is not None:
        cur = orc.resize(cur, resize_w)
    if angle is not None:
        cur = orc.rotate(cur, angle)
    if conv_preset:
        coef, div = {1: (KERNELS["blur3"][0], 16), 2: (np.ones((7, 7), np.int64), 49), 3: (KERNELS["sharpen3"][0], 1),
                     4: (KERNELS["edge3"][0], 1), 5: KERNELS["gauss5"][:2], 6: KERNELS["gauss7"][:2], 7: KERNELS["unsharp7"][:2],
                     8: KERNELS["binom9"][:2]}[conv_preset]
        cur = orc.conv(cur, coef, div, 0)
    if levels is not None:
        cur = orc.levels(cur, orc.levels_lut_linear(*levels))
    renewed = resize_w is not None or angle is not None or conv_preset or levels is not None
    if renewed:
        return orc.process(cur, angle=0, gray=gray, mono=mono, flipv=flipv, fliph=fliph) if (gray or mono or flipv or fliph) \
            else (cur.reshape(-1), cur.shape[1], cur.shape[0], 0)
    return orc.process(cur, gray=gray, mono=mono, flipv=flipv, fliph=fliph)


def test_row_parts_equal_whole_raster(gpu, orc, monkeypatch):
    """ppmx_gpu_apply cuts a raster into row parts that overlap upload / kernels / download; with PPMX_PART_BYTES
    small rasters are cut into many uneven parts: same bytes as the oracle's whole-raster chain."""
    for part_bytes in ("3000", "20000"):
        monkeypatch.setenv("PPMX_PART_BYTES", part_bytes)
        for (w, h) in [(64, 200), (48, 131), (37, 90)]:
            img = P.lcg(w, h, 900 + w)
            for kw in PART_CHAINS:
                if kw.get("resize_w") and kw["resize_w"] > 4 * w:
                    continue
                exp, ew, eh, eft = _orc_chain(orc, img, **kw)
                got, gw, gh, gft = gpu.process(img, **kw)
                assert (gw, gh, gft) == (ew, eh, eft), (part_bytes, w, h, kw)
                assert np.array_equal(got, exp), (part_bytes, w, h, kw)
    monkeypatch.delenv("PPMX_PART_BYTES")
    img = P.lcg(2048, 1500, 77)  # 9.2 MB: cut by the default part size
    got, gw, gh, gft = gpu.process(img, conv_preset=1, gray=True)
    exp = orc.process(orc.conv(img, KERNELS["blur3"][0], 16, 0), angle=0, gray=True)
    assert (gw, gh, gft) == exp[1:] and np.array_equal(got, exp[0])


def test_apply_band_stitches_to_whole(gpu, orc):
    """ppmx_gpu_apply_band: every band writes its rows of ONE shared output; all bands together = the whole result.
    Each call gets a virtual whole-raster pointer behind which only the rows ppmx_gpu_band_rows names exist."""
    import ctypes as C
    import imageprocessingtools_b200.ppmx as pp
    w, h = 64, 150
    img = P.lcg(w, h, 4711)
    for kw in [dict(conv_preset=1), dict(conv_preset=2, mono=True, flipv=True), dict(resize_w=96, gray=True), dict(flipv=True),
               dict(angle=180), dict(gray=True), dict(resize_w=40, conv_preset=4)]:
        ph = pp._PlanHolder(w=w, h=h, **kw)
        ops = [ph.plan.ops[i] for i in range(ph.plan.nops)]
        exp, ew, eh, eft = _orc_chain(orc, img, **kw)
        for nb in (1, 2, 3, 5, 8):
            ow, oh, ft, nbytes, split, _ = pp.chain_info(ops, w, h)
            assert split and (ow, oh, ft) == (ew, eh, eft)
            out = np.zeros(nbytes, np.uint8)
            row_bytes = nbytes // oh
            arr = (pp.PpmxOp * len(ops))(*ops)
            for b in range(nb):
                oy0, orows, sy0, srows = pp.band_rows(ops, w, h, b, nb)
                src = img[sy0:sy0 + srows].copy()  # nothing else of the raster exists for this call
                n, rw, rh, rft, y0, rows = C.c_size_t(), C.c_uint32(), C.c_uint32(), C.c_int(), C.c_uint32(), C.c_uint32()
                rc = gpu.L.ppmx_gpu_apply_band(gpu.ctx, arr, len(ops), C.c_void_p(src.ctypes.data - sy0 * w * 3), w, h, b, nb,
                                               C.c_void_p(out.ctypes.data), out.size, C.byref(n), C.byref(rw), C.byref(rh),
                                               C.byref(rft), C.byref(y0), C.byref(rows))
                assert rc == 0 and (y0.value, rows.value) == (oy0, orows)
            assert np.array_equal(out, exp), (kw, nb)
        ph.close()
    # a chain with a 90 degree rotation can not be cut into bands: refused, not computed wrongly
    ph = pp._PlanHolder(w=w, h=h, angle=90)
    ops = [ph.plan.ops[i] for i in range(ph.plan.nops)]
    assert pp.chain_info(ops, w, h)[4] is False
    with pytest.raises(pp.PpmxError):
        gpu.apply_band(img, ops, 0, 2)
    ph.close()


def test_multi_device_context_row_bands(orc):
    """ppmx_gpu_init_multi: ONE raster over all visible GPUs as row bands (config 4), a batch image-parallel."""
    import torch
    import imageprocessingtools_b200 as ip
    import imageprocessingtools_b200.ppmx as pp
    n = torch.cuda.device_count()
    devs = list(range(n)) if n > 1 else [0, 0]  # one GPU: two contexts on it still exercise the split
    g = ip.Ppmx(devs)
    try:
        assert g.device_count() == len(devs)
        img = P.lcg(256, 301, 8)
        for kw in [dict(conv_preset=2), dict(conv_preset=1, mono=True, flipv=True), dict(resize_w=384, gray=True), dict(angle=90)]:
            exp, ew, eh, eft = _orc_chain(orc, img, **kw)
            got, gw, gh, gft = g.process(img, **kw)
            assert (gw, gh, gft) == (ew, eh, eft) and np.array_equal(got, exp), kw
        imgs = np.stack([P.lcg(96, 40, 300 + i) for i in range(9)])
        bins = np.zeros((9, 256), np.uint64)
        import ctypes as C
        op = pp.PpmxOp(kind=pp.OP_GRAY_HIST, hist_out=bins.ctypes.data_as(C.POINTER(C.c_uint64)))
        out, w, h, ft = g.apply_ops(imgs, [op])
        for i in range(9):
            assert np.array_equal(out[i].reshape(40, 96), orc.gray(imgs[i])) and np.array_equal(bins[i], orc.hist_gray(imgs[i]))
        big = P.lcg(1024, 515, 9)
        bins1 = np.zeros(256, np.uint64)
        op = pp.PpmxOp(kind=pp.OP_GRAY_HIST, hist_out=bins1.ctypes.data_as(C.POINTER(C.c_uint64)))
        out, w, h, ft = g.apply_ops(big[None], [op])  # one raster: bands over the devices, bins summed on the host
        assert np.array_equal(out[0].reshape(515, 1024), orc.gray(big)) and np.array_equal(bins1, orc.hist_gray(big))
    finally:
        g.close()


def test_full_size_16384_square(gpu, orc):
    """BASELINE config 4 at its full size, one GPU: 16384 x 16384 (805 MB) gray against the oracle, mono / flips
    through size-independent properties, and the 3x3 convolution as "8 bands == whole raster" plus oracle slices at
    the raster's top, middle and bottom."""
    import imageprocessingtools_b200.ppmx as pp
    w = h = 16384
    img = pp.synth_lcg(w, h, 0xC0FFEE ^ 4)
    g = gpu.gray(img)
    assert np.array_equal(g, orc.gray(img))
    bits = gpu.mono_bits(img).reshape(h, w // 8)
    assert np.array_equal(bits[4096:4100], orc.pack_pbm(orc.mono(img[4096:4100])).reshape(4, w // 8))   # row phase 4096 % 4 == 0
    assert np.array_equal(bits[-4:], orc.pack_pbm(orc.mono(img[-4:])).reshape(4, w // 8))
    fv = gpu.flip(img, 1)
    assert np.array_equal(fv, img[::-1])
    del fv
    r180 = gpu.rotate(img, 180)
    assert np.array_equal(r180[:64], img[::-1, ::-1][:64]) and np.array_equal(r180[-64:], img[::-1, ::-1][-64:])
    assert int(r180.astype(np.uint8).sum(dtype=np.uint64)) == int(img.sum(dtype=np.uint64))
    del r180, g, bits
    coef, div, bias = KERNELS["blur3"]
    op = gpu.conv_op(coef, div, bias)
    whole = gpu.conv(img, coef, div, bias)
    out = np.zeros(w * h * 3, np.uint8)
    for b in range(8):
        gpu.apply_band(img, [op], b, 8, out)
    assert np.array_equal(out.reshape(h, w, 3), whole)
    del out
    assert np.array_equal(whole[:40], orc.conv(img[:41], coef, div, bias)[:40])                   # mirror border at the top
    assert np.array_equal(whole[8000:8040], orc.conv(img[7999:8041], coef, div, bias)[1:-1])      # across a band cut (8192 is one)
    assert np.array_equal(whole[8180:8200], orc.conv(img[8179:8201], coef, div, bias)[1:-1])
    assert np.array_equal(whole[-40:], orc.conv(img[-41:], coef, div, bias)[1:])                  # mirror border at the bottom


def test_full_size_bands_16384(gpu, orc):
    """BASELINE config 4 size: a 16384-wide raster; N bands == whole raster for the convolutions (property), gray /
    mono / flips of a 16384 x 2048 slice against the oracle directly."""
    import imageprocessingtools_b200.ppmx as pp
    w, h = 16384, 2048
    img = pp.synth_lcg(w, h, 0xC0FFEE ^ 4)
    assert np.array_equal(gpu.gray(img), orc.gray(img))
    assert np.array_equal(gpu.mono_bits(img), orc.pack_pbm(orc.mono(img)))
    assert np.array_equal(gpu.flip(img, 1), img[::-1])
    assert np.array_equal(gpu.rotate(img, 180), img[::-1, ::-1])
    for kname in ("blur3", "box7"):
        coef, div, bias = KERNELS[kname]
        op = gpu.conv_op(coef, div, bias)
        whole = gpu.conv(img, coef, div, bias)
        out = np.zeros(w * h * 3, np.uint8)
        for b in range(8):
            gpu.apply_band(img, [op], b, 8, out)
        assert np.array_equal(out.reshape(h, w, 3), whole), kname
        sl = slice(1000, 1040)  # a slice of it against the oracle (rows far from the raster's edges)
        r = coef.shape[0] // 2
        exp = orc.conv(img[sl.start - r:sl.stop + r], coef, div, bias)[r:-r]
        assert np.array_equal(whole[sl], exp), kname


def test_cuda_graph_replays_launches(gpu, orc):
    """ppmx_gpu_graph_*: launches recorded once and replayed give the bytes of direct launches."""
    import torch
    import imageprocessingtools_b200.ppmx as pp
    dev = torch.device("cuda", 0)
    img = P.lcg(256, 128, 5)
    src = torch.from_numpy(img).to(dev)
    mid = torch.zeros_like(src)
    out = torch.zeros((128, 256), dtype=torch.uint8, device=dev)
    coef, div, bias = KERNELS["blur3"]
    op1, op2 = gpu.conv_op(coef, div, bias), pp.PpmxOp(kind=pp.OP_GRAY)
    s = torch.cuda.Stream()
    n0 = gpu.launch_count()
    with torch.cuda.stream(s):
        gpu.graph_begin(s.cuda_stream)
        gpu.launch(op1, src.data_ptr(), 256, 128, pp.LAYOUT_RGB8, mid.data_ptr(), None, 0, 0, s.cuda_stream)
        gpu.launch(op2, mid.data_ptr(), 256, 128, pp.LAYOUT_RGB8, out.data_ptr(), None, 0, 0, s.cuda_stream)
        graph, nodes = gpu.graph_end(s.cuda_stream)
        assert nodes == 2
        assert int(out.sum().item()) == 0  # recording runs nothing
        gpu.graph_launch(graph, s.cuda_stream)
        gpu.graph_launch(graph, s.cuda_stream)
        s.synchronize()
    assert gpu.launch_count() - n0 == 2 + 4
    gpu.graph_free(graph)
    assert np.array_equal(out.cpu().numpy(), orc.gray(orc.conv(img, coef, div, bias)))


def test_cli_batch_p3_and_multi_device(gpu, orc, tmp_path):
    """EXTENSION flags of the command line: -batch (several files, one device context, file I/O overlapped with the device
    work) gives the same files as one run per file; a P3 / 16-bit input gives the P6 result; PPMX_DEVICE=all spreads one
    raster over every visible GPU as row bands."""
    import imageprocessingtools_b200.ppmx as pp
    imgs = [P.lcg(96 + 16 * i, 40 + 3 * i, 700 + i) for i in range(5)]
    paths = []
    for i, im in enumerate(imgs):
        p = str(tmp_path / ("b%d.ppm" % i))
        oracle.write_p6(p, im)
        paths.append(p)
    r = subprocess.run([pp.CLI, "-batch", "-r90", "-gray", "-fv"] + paths, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    for p, im in zip(paths, imgs):
        exp, ew, eh, eft = orc.process(im, angle=90, gray=True, flipv=True)
        assert open(p + ".out", "rb").read() == orc.header(eft, ew, eh, 255) + exp.tobytes(), p
    # without -batch a second file name is refused like the reference does (ref:180)
    r = subprocess.run([pp.CLI, "-gray", paths[0], paths[1]], capture_output=True, text=True)
    assert r.returncode != 0 and "Error: invalid options" in r.stdout
    # one bad file in a batch fails the run but the others are still written
    bad = str(tmp_path / "bad.ppm")
    open(bad, "wb").write(b"P6\n4 4\n255\nxx")
    for p in paths:
        os.remove(p + ".out")
    r = subprocess.run([pp.CLI, "-batch", "-mono", paths[0], bad, paths[1]], capture_output=True, text=True)
    assert r.returncode != 0 and os.path.exists(paths[0] + ".out") and os.path.exists(paths[1] + ".out")
    # P3 and 16-bit P6 of the same picture
    im = imgs[0]
    h, w, _ = im.shape
    p3 = str(tmp_path / "t.p3.ppm")
    open(p3, "wb").write(b"P3\n%d %d\n255\n" % (w, h) + b" ".join(b"%d" % v for v in im.reshape(-1)) + b"\n")
    p16 = str(tmp_path / "t.p16.ppm")
    open(p16, "wb").write(b"P6\n%d %d\n65535\n" % (w, h) + (im.astype(np.uint16) * 257).astype(">u2").tobytes())
    exp, ew, eh, eft = orc.process(im, gray=True)
    for p in (p3, p16):
        r = subprocess.run([pp.CLI, "-gray", p], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout
        assert open(p + ".out", "rb").read() == orc.header(eft, ew, eh, 255) + exp.tobytes(), p
    # every visible GPU on one raster
    big = str(tmp_path / "big.ppm")
    bimg = P.lcg(512, 700, 99)
    oracle.write_p6(big, bimg)
    env = dict(os.environ, PPMX_DEVICE="all", PPMX_PART_BYTES="100000")
    r = subprocess.run([pp.CLI, "-blur", "-mono", "-fv", big], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout
    exp, ew, eh, eft = _orc_chain(orc, bimg, conv_preset=1, mono=True, flipv=True)
    assert open(big + ".out", "rb").read() == orc.header(eft, ew, eh, 255) + exp.tobytes()
