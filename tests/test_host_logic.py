"""CPU-only checks of the host layer (C, libppmx_host.so) and of the ABI surface: no compute calls,
no GPU.  Contribution tables and rotation sizes must be bit-identical to the oracle's."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import patterns as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pp():
    from imageprocessingtools_b200 import build, ppmx
    build.build_all()
    return ppmx


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ppmx_[a-zA-Z0-9_]+)\s*\(", src)))


def test_abi_exports_every_declared_symbol(pp):
    """include/*.h is the contract: each declared function must be exported by the built libraries."""
    for header, lib in (("ppmx_gpu.h", pp.GPU_SO), ("ppmx_host.h", pp.HOST_SO)):
        names = _declared(header)
        assert len(names) >= 15
        out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
        exported = set(re.findall(r" T (\w+)", out))
        missing = [n for n in names if n not in exported]
        assert not missing, (header, missing)
    L = pp.gpu_lib()
    pp.host_lib()
    assert b"sm_100a" in L.ppmx_gpu_version()


def test_op_struct_layout(pp):
    assert C.sizeof(pp.PpmxOp) == 112 and C.sizeof(pp.PpmxBand) == 40


def test_library_holds_sm100a_code_only(pp):
    out = subprocess.run(["cuobjdump", "-lelf", pp.GPU_SO], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_gpu_means_loud_failure(pp):
    """The product has no CPU fallback: without a device ppmx_gpu_init must fail, not emulate."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import imageprocessingtools_b200 as ip
    with pytest.raises(ip.PpmxError):
        ip.Ppmx(0)


def test_div3_multiplier_exact():
    s = np.arange(0, 766, dtype=np.uint64)
    assert np.array_equal((s * 43691) >> 17, s // 3)          # used by the kernels
    assert not np.array_equal((s * 171) >> 9, s // 3)          # the shorter constant is NOT exact


def test_mono_nibble_multiplier():
    for m in range(16):
        b = [(m >> (3 - i)) & 1 for i in range(4)]
        word = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24)
        assert ((word * 0x08040201) & 0xFFFFFFFF) >> 24 & 0xF == m


def test_contributions_match_oracle(pp, orc):
    for n_in, n_out in [(37, 74), (37, 55), (37, 37), (37, 18), (37, 7), (37, 3), (64, 96), (64, 13), (5, 64), (3, 2),
                        (2, 1), (1, 3), (4096, 6144), (4096, 2048), (1920, 960), (1080, 540), (1000, 999)]:
        s = float(n_out) / n_in
        w0, i0 = orc.calc_contributions(n_in, n_out, s)
        w1, i1 = pp.calc_contributions(n_in, n_out, s)
        assert w0.shape == w1.shape and np.array_equal(i0, i1), (n_in, n_out)
        assert np.array_equal(w0.view(np.uint64), w1.view(np.uint64)), (n_in, n_out)


def test_rotation_sizes_and_cubic(pp, orc):
    H = pp.host_lib()
    for (w, h) in [(1, 1), (37, 23), (512, 512), (1920, 1080), (4096, 4096)]:
        for a in range(360):
            assert pp.rotate_size(a, w, h) == orc.rotate_size(a, w, h), (w, h, a)
    for x in np.linspace(-3, 3, 601):
        assert H.ppmx_cubic(float(x)) == orc.cubic(float(x))
    for a in range(-20, 20):
        for b in (0, 1, 2, 7):
            assert H.ppmx_mod(a, b) == orc.lib.orc_mod(a, b)


def test_header_parse_and_format(pp, orc):
    raster = bytes(range(24))
    assert pp.parse_header(b"P6\n4 2\n255\n" + raster) == (4, 2, 255, 11)
    assert pp.parse_header(b"P6\n# a comment\n4 2\n# another\n255\n" + raster)[:3] == (4, 2, 255)
    assert pp.parse_header(b"P6 4 2 255 " + raster)[3] == 11
    for bad in (b"P3\n4 2\n255\n" + raster, b"P6\n4 2\n255\n" + raster + b"x", b"P6\n4 2\n255\n" + raster[:-4],
                b"P6\n4 x\n255\n" + raster, b"\n", b"P6\n4 2\n65535\n" + raster * 2):
        with pytest.raises(pp.PpmxError):
            pp.parse_header(bad)
    for ft in (0, 1, 2):
        assert pp.format_header(ft, 37, 23, 255) == orc.header(ft, 37, 23, 255)


def test_plan_chain_follows_reference_order(pp):
    """ref:1084-1155: resize(2 passes) -> rotate -> gray|mono -> flipv|fliph; renew iff -w or -r."""
    ph = pp._PlanHolder(resize_w=55, angle=90, gray=True, flipv=True, w=37, h=23)
    ops = [ph.plan.ops[i] for i in range(ph.plan.nops)]
    assert [o.kind for o in ops] == [pp.OP_IMRESIZE, pp.OP_IMRESIZE, pp.OP_ROTATE, pp.OP_GRAY, pp.OP_FLIP]
    assert [o.renew_before for o in ops] == [0, 1, 1, 1, 1]
    # scale[0] < scale[1] here, so the height pass runs first (ref:1102)
    assert ops[0].dim == 0 and ops[1].dim == 1 and ops[1].out_size == 55 and ops[0].out_size == int(23 * (55 / 37))
    ph.close()
    ph = pp._PlanHolder(gray=True, fliph=True, w=8, h=8)  # the leaked-gray quirk: no renew without -w/-r
    assert [ph.plan.ops[i].renew_before for i in range(ph.plan.nops)] == [0, 0]
    ph.close()
    ph = pp._PlanHolder(resize_w=74, w=37, h=23)  # exact scale: width pass first
    assert ph.plan.ops[0].dim == 1 and ph.plan.ops[1].dim == 0
    ph.close()
    ph = pp._PlanHolder(w=8, h=8)
    assert ph.plan.nops == 0
    ph.close()


def test_cli_flag_errors_match_reference(pp, tmp_path):
    """Flag parsing happens before any device work, so these run without a GPU (ref:125-187)."""
    import oracle
    img = str(tmp_path / "x.ppm")
    oracle.write_p6(img, P.lcg(4, 4, 1))
    cases = [["-gray", "-mono", img], ["-fv", "-fh", img], ["-gray", "-gray", img], ["-r360", img], ["-r", img],
             ["-wabc", img], ["-q", img], ["-fx", img], [img, img], []]
    for args in cases:
        ours = subprocess.run([pp.CLI] + args, capture_output=True, text=True)
        assert ours.returncode == 255, args
        if os.path.exists(oracle.REF_CLI):
            ref = subprocess.run([oracle.REF_CLI] + args, capture_output=True, text=True)
            assert ref.returncode == 255 and ref.stdout == ours.stdout, (args, ref.stdout, ours.stdout)


def test_levels_table_matches_oracle(pp, orc):
    """The integer table of ppmx_levels_lut_linear equals the oracle's floor(x + 0.5) in doubles for every (lo, hi),
    and the histogram clip picks the documented points."""
    for lo in range(0, 255, 7):
        for hi in list(range(lo + 1, 256, 5)) + [255]:
            assert np.array_equal(pp.Ppmx.levels_lut_linear(lo, hi), orc.levels_lut_linear(lo, hi)), (lo, hi)
    with pytest.raises(pp.PpmxError):
        pp.Ppmx.levels_lut_linear(10, 10)
    bins = np.zeros(256, np.uint64)
    bins[30], bins[31], bins[100], bins[200], bins[201] = 4, 6, 1980, 7, 3          # 2000 pixels, 5 permille = 10
    with pytest.raises(pp.PpmxError):                                                # both points land on grey 100: flat
        pp.Ppmx.levels_points_from_hist(bins, 5)
    assert pp.Ppmx.levels_points_from_hist(bins, 4) == (31, 200)                     # 8 pixels: 4 below, 3 above fit
    bins2 = np.zeros(256, np.uint64)
    bins2[10], bins2[20], bins2[240], bins2[250] = 5, 995, 990, 10                   # 2000 pixels
    assert pp.Ppmx.levels_points_from_hist(bins2, 5) == (20, 240)                    # 5 and 10 pixels may be clipped
    assert pp.Ppmx.levels_points_from_hist(bins2, 0) == (10, 250)
    with pytest.raises(pp.PpmxError):
        pp.Ppmx.levels_points_from_hist(np.zeros(256, np.uint64), 5)


def test_conv_rounding_magic():
    """The convolution's exact floor((2*acc+div)/(2*div)) via one multiply-high (ppmx_conv.cu
    make_conv_round / ConvRound): n/d == (n*M) >> (31+l) for all 0 <= n < 2^31."""
    rng = np.random.default_rng(0)
    for div in [1, 2, 3, 7, 9, 16, 25, 49, 81, 100, 255, 256, 1000, 4096, 65535, 12345678]:
        d = 2 * div
        l = 0
        while (1 << l) < d:
            l += 1
        M = ((1 << (31 + l)) + d - 1) // d
        assert M < (1 << 32)
        ns = np.concatenate([np.arange(0, 70000, dtype=np.uint64), rng.integers(0, 1 << 31, 200000).astype(np.uint64),
                             np.array([(1 << 31) - 1 - i for i in range(1000)], np.uint64),
                             (np.arange(1, 5000, dtype=np.uint64) * np.uint64(d)) - np.uint64(1),
                             np.arange(1, 5000, dtype=np.uint64) * np.uint64(d)])
        ns = ns[ns < (1 << 31)]
        hi = (ns.astype(object) * M) >> 32
        q = np.array([int(v) >> (l - 1) for v in hi], dtype=object)
        assert all(int(a) == int(b) // d for a, b in zip(q[::37], ns[::37]))
        exp = (ns // np.uint64(d)).astype(object)
        assert (q == exp).all(), div


# ---- round 2 ---------------------------------------------------------------------------------------------

def test_bayer_thresholds_are_ceil_of_matrix_times_255():
    """c_bayer (ppmx_common.cuh) holds ceil(matrix * 255); for an integer grey, g < c_bayer <=> !(g >= matrix * 255)
    (ref:954, 967) -- all 256 x 16 cases."""
    src = open(os.path.join(ROOT, "imageprocessingtools_b200", "csrc", "ppmx_common.cuh")).read()
    thr = [int(x) for x in re.search(r"c_bayer\[16\] = \{([^}]*)\}", src).group(1).split(",")]
    matrix = [.125, 1, .1875, .8125, .625, .375, .6875, .4375, .25, .875, .0625, .9375, .75, .5, .5625, .3125]  # ref:954
    assert thr == [int(np.ceil(m * 255)) for m in matrix]
    for i in range(16):
        for gv in range(256):
            assert (gv < thr[i]) == (not (gv >= matrix[i] * 255.0))


def test_planner_refuses_conflicting_flags_and_stays_in_bounds(pp):
    """ref:130,135,166,171: -gray x -mono and -fv x -fh are refused; every accepted flag set fits PPMX_PLAN_MAX_OPS."""
    for kw in (dict(gray=True, mono=True), dict(flipv=True, fliph=True),
               dict(resize_w=10, angle=30, gray=True, mono=True, flipv=True, fliph=True, conv_preset=1, levels=(1, 200))):
        with pytest.raises(pp.PpmxError):
            pp._PlanHolder(w=16, h=16, **kw)
    ph = pp._PlanHolder(w=16, h=16, resize_w=10, angle=30, mono=True, flipv=True, conv_preset=1, levels=(1, 200))
    assert ph.plan.nops == 7 <= pp.PLAN_MAX_OPS
    ph.close()


def test_chain_linearisation_and_band_rows(pp):
    """ppmx_gpu_chain_info / ppmx_gpu_band_rows need no device: the chain is reduced to the operators that feed the
    writer (ref:1084-1155 hand-over rules) and every band knows which source rows it reads."""
    def ops_of(**kw):
        ph = pp._PlanHolder(w=64, h=40, **kw)
        return ph, [ph.plan.ops[i] for i in range(ph.plan.nops)]
    ph, ops = ops_of(gray=True, fliph=True)         # the leaked grey raster: flip + .r extraction only,
    assert pp.chain_info(ops, 64, 40) == (64, 40, pp.FT_PGM, 64 * 40, True, 1)   # ... fused into one kernel
    os.environ["PPMX_NO_FUSE"] = "1"
    try:
        assert pp.chain_info(ops, 64, 40) == (64, 40, pp.FT_PGM, 64 * 40, True, 2)
    finally:
        del os.environ["PPMX_NO_FUSE"]
    ph.close()
    ph, ops = ops_of(angle=90, mono=True, fliph=True)   # config 5: one kernel
    assert pp.chain_info(ops, 1920, 1080)[5] == 1
    ph.close()
    ph, ops = ops_of(resize_w=960, angle=90, gray=True, flipv=True)
    assert pp.chain_info(ops, 64, 40)[5] == 3           # two resize passes + one fused rotate/grey/flip
    ph.close()
    ph, ops = ops_of(mono=True)                      # mono + P4 packer in one kernel
    assert pp.chain_info(ops, 64, 40) == (64, 40, pp.FT_PBM, 8 * 40, True, 1)
    ph.close()
    ph, ops = ops_of(angle=90, mono=True, fliph=True)
    assert pp.chain_info(ops, 64, 40)[:5] == (40, 64, pp.FT_PBM, 5 * 64, False)
    with pytest.raises(pp.PpmxError):
        pp.band_rows(ops, 64, 40, 0, 2)
    ph.close()
    ph, ops = ops_of(conv_preset=2)                  # 7x7: three halo rows on either side, none beyond the raster
    cover = []
    for b in range(3):
        oy0, orows, sy0, srows = pp.band_rows(ops, 64, 40, b, 3)
        assert sy0 == max(0, oy0 - 3) and sy0 + srows == min(40, oy0 + orows + 3)
        cover += list(range(oy0, oy0 + orows))
    assert cover == list(range(40))
    ph.close()
    ph, ops = ops_of(flipv=True)                     # the mirrored band, no halo
    for b in range(4):
        oy0, orows, sy0, srows = pp.band_rows(ops, 64, 40, b, 4)
        assert (sy0, srows) == (40 - oy0 - orows, orows)
    ph.close()
    ph, ops = ops_of(resize_w=96)                    # x1.5: the height pass reads the rows its table names
    wt, ix = pp.calc_contributions(40, 60, 1.5)
    for b in range(3):
        oy0, orows, sy0, srows = pp.band_rows(ops, 64, 40, b, 3)
        need = ix[oy0:oy0 + orows]
        assert (sy0, sy0 + srows) == (int(need.min()), int(need.max()) + 1)
    ph.close()


def test_synthetic_lcg_rows(pp):
    """ppmx_synth_lcg: any row range of the LCG raster equals the same rows of the whole raster (jump-ahead)."""
    whole = pp.synth_lcg(37, 23, 0xC0FFEE)
    assert np.array_equal(whole, P.lcg(37, 23, 0xC0FFEE))
    for y0, rows in [(0, 1), (5, 7), (22, 1), (11, 12)]:
        assert np.array_equal(pp.synth_lcg(37, rows, 0xC0FFEE, y0=y0), whole[y0:y0 + rows])


def test_release_library_carries_no_tuning_variants(pp):
    """The alternative kernel variants live in libppmx_gpu_tuning.so only."""
    L = pp.gpu_lib()
    assert L.ppmx_gpu_set_tuning(b"variant", 0) == 0 and L.ppmx_gpu_set_tuning(b"variant", 3) != 0
    assert L.ppmx_gpu_set_tuning(b"pdl", 1) == 0
    T = pp.gpu_lib(tuning=True)
    assert T.ppmx_gpu_set_tuning(b"variant", 3) == 0 and T.ppmx_gpu_set_tuning(b"variant", 0) == 0


def test_extension_p3_and_16bit_decoding(pp):
    """EXTENSION (the reference rejects P3, ref:386, and 16-bit samples, ref:453): both decode to the packed 8-bit
    raster; maxval <= 255 keeps bytes and maxval, larger maxvals scale with round(v * 255 / maxval)."""
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (5, 7, 3), dtype=np.uint8)
    txt = b"P3\n# made by a test\n7 5\n255\n" + b"\n".join(b" ".join(b"%d" % v for v in row.reshape(-1)) for row in img) + b"\n"
    got, mx = pp.decode_pnm(txt)
    assert mx == 255 and np.array_equal(got, img)
    low = (img // 16).astype(np.uint8)   # maxval 15: samples stay as they are, the header keeps 15
    txt = b"P3 7 5 15 " + b" ".join(b"%d" % v for v in low.reshape(-1)) + b" # trailing comment\n"
    got, mx = pp.decode_pnm(txt)
    assert mx == 15 and np.array_equal(got, low)
    wide = rng.integers(0, 65536, (5, 7, 3), dtype=np.uint32)
    exp = ((2 * wide.astype(np.uint64) * 255 + 65535) // (2 * 65535)).astype(np.uint8)
    got, mx = pp.decode_pnm(b"P6\n7 5\n65535\n" + wide.astype(">u2").tobytes())
    assert mx == 255 and np.array_equal(got, exp)
    got, mx = pp.decode_pnm(b"P3\n7 5\n1000\n" + b" ".join(b"%d" % min(v, 1000) for v in wide.reshape(-1)))
    w1000 = np.minimum(wide, 1000).astype(np.uint64)
    assert mx == 255 and np.array_equal(got, ((2 * w1000 * 255 + 1000) // 2000).astype(np.uint8))
    for bad in (b"P3\n7 5\n255\n1 2 3", b"P3\n7 5\n255\n" + b"1 " * 105 + b"x", b"P6\n7 5\n65535\n" + b"\0" * 209,
                b"P5\n7 5\n255\n" + b"\0" * 35, b"P3\n7 5\n70000\n" + b"1 " * 105):
        with pytest.raises(pp.PpmxError):
            pp.decode_pnm(bad)
