"""GPU parity of the fused "geometry + pointwise tail" kernel (csrc/ppmx_fused.cu) and of the chain fusion pass:
every orientation x every tail, at aligned and ragged sizes, against the oracle's stage-by-stage result; the fused
chain must equal the unfused one byte for byte."""
import os

import numpy as np
import pytest

import patterns as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import imageprocessingtools_b200 as ip
    g = ip.Ppmx(0)
    yield g
    g.close()


SIZES = [(64, 64), (128, 64), (64, 192), (16, 16), (80, 48), (1, 1), (3, 5), (37, 23), (130, 70), (200, 56), (96, 1080 // 8),
         (257, 129), (72, 40), (8, 8), (24, 136)]
FUSED_CHAINS = [dict(angle=90, mono=True, fliph=True), dict(angle=90, gray=True, flipv=True), dict(angle=270, gray=True, fliph=True),
                dict(angle=180, mono=True, flipv=True), dict(angle=90, flipv=True), dict(angle=270, fliph=True),
                dict(angle=180, gray=True), dict(angle=90, gray=True), dict(angle=270, mono=True), dict(angle=90, mono=True),
                dict(angle=90, mono=True, flipv=True), dict(angle=270, mono=True, fliph=True), dict(angle=180, mono=True, fliph=True),
                dict(angle=90), dict(angle=270), dict(gray=True, fliph=True), dict(gray=True, flipv=True),
                dict(angle=180, fliph=True), dict(angle=180, flipv=True)]


def test_fused_chains_match_oracle(gpu, orc):
    import imageprocessingtools_b200.ppmx as pp
    fused_seen = 0
    for (w, h) in SIZES:
        for name in ("lcg", "bayer"):
            img = P.all_patterns(w, h)[name]
            for kw in FUSED_CHAINS:
                exp, ew, eh, eft = orc.process(img, **kw)
                got, gw, gh, gft = gpu.process(img, **kw)
                assert (gw, gh, gft) == (ew, eh, eft), (w, h, name, kw)
                assert np.array_equal(got, exp), (w, h, name, kw)
                ph = pp._PlanHolder(w=w, h=h, **kw)
                ops = [ph.plan.ops[i] for i in range(ph.plan.nops)]
                fused_seen += pp.chain_info(ops, w, h)[5] == 1
                ph.close()
    assert fused_seen > 100  # most of these chains are ONE kernel


def test_fused_chain_fuzz(gpu, orc):
    """Seeded random sweep: sizes from 1 x 1 to a few hundred pixels a side (rows at every alignment, ragged tiles on both axes),
    every combination of right-angle rotation, grey / mono and flips the command line accepts, against the oracle's chain."""
    rng = np.random.RandomState(19)
    for it in range(220):
        w, h = int(rng.randint(1, 330)), int(rng.randint(1, 260))
        if it % 7 == 0:
            w = 16 * int(rng.randint(1, 20))
        if it % 11 == 0:
            h = 8 * int(rng.randint(1, 30))
        kw = {}
        a = int(rng.choice([0, 90, 180, 270, 90, 270]))
        if a:
            kw["angle"] = a
        t = int(rng.randint(0, 3))
        if t == 1:
            kw["gray"] = True
        elif t == 2:
            kw["mono"] = True
        f = int(rng.randint(0, 3))
        if f == 1:
            kw["flipv"] = True
        elif f == 2:
            kw["fliph"] = True
        if not kw:
            kw["flipv"] = True
        img = P.lcg(w, h, 500 + it)
        exp, ew, eh, eft = orc.process(img, **kw)
        got, gw, gh, gft = gpu.process(img, **kw)
        assert (gw, gh, gft) == (ew, eh, eft), (it, w, h, kw)
        assert np.array_equal(got, exp), (it, w, h, kw)


def test_config5_chains_kernel_counts(gpu, orc):
    """BASELINE config 5 chains on a 1920x1080 frame: "-r90 -mono -fh" is one kernel, "-w960 -r90 -gray -fv" three."""
    import imageprocessingtools_b200.ppmx as pp
    img = P.lcg(1920, 1080, 0xC0FFEE ^ 5)
    for kw, kernels in ((dict(angle=90, mono=True, fliph=True), 1), (dict(resize_w=960, angle=90, gray=True, flipv=True), 3)):
        ph = pp._PlanHolder(w=1920, h=1080, **kw)
        ops = [ph.plan.ops[i] for i in range(ph.plan.nops)]
        assert pp.chain_info(ops, 1920, 1080)[5] == kernels, kw
        ph.close()
        n0 = gpu.launch_count()
        got = gpu.process(img, **kw)
        assert gpu.launch_count() - n0 == kernels, kw
        exp = orc.process(img, **kw)
        assert got[1:] == exp[1:] and np.array_equal(got[0], exp[0]), kw


def test_fused_equals_unfused(gpu, orc, monkeypatch):
    for (w, h) in [(200, 56), (37, 23), (64, 64)]:
        img = P.lcg(w, h, 5 + w)
        for kw in FUSED_CHAINS:
            a = gpu.process(img, **kw)
            monkeypatch.setenv("PPMX_NO_FUSE", "1")
            b = gpu.process(img, **kw)
            monkeypatch.delenv("PPMX_NO_FUSE")
            assert a[1:] == b[1:] and np.array_equal(a[0], b[0]), (w, h, kw)


def test_odd_size_rotations_take_the_tile_kernel(gpu, orc):
    """90 / 270 degrees at sizes the bulk-copy transposer can not take (ragged tiles, unaligned rows)."""
    for (w, h) in [(1920, 1080), (1080, 1920), (4090 // 8, 4090 // 8 + 3), (1000, 37), (33, 1000)]:
        img = P.lcg(w, h, 77)
        for a in (90, 270):
            assert np.array_equal(gpu.rotate(img, a), orc.rotate(img, a)), (w, h, a)


def test_prepared_chain_on_device_rasters(gpu, orc):
    """ppmx_gpu_chain_prepare / _run: the chain on a raster already in HBM, recorded into a CUDA graph and replayed."""
    import torch
    import imageprocessingtools_b200.ppmx as pp
    dev = torch.device("cuda", 0)
    for (w, h, kw) in [(1920, 1080, dict(angle=90, mono=True, fliph=True)), (200, 56, dict(resize_w=100, angle=90, gray=True, flipv=True)),
                       (130, 70, dict(conv_preset=1, mono=True, flipv=True)), (64, 64, dict(angle=0))]:
        img = P.lcg(w, h, 31 + w)
        ph = pp._PlanHolder(w=w, h=h, **kw)
        ops = [ph.plan.ops[i] for i in range(ph.plan.nops)]
        ch, meta = gpu.chain_prepare(ops, w, h)
        src = torch.from_numpy(img).to(dev)
        dst = torch.zeros(meta["out_bytes"] + 16, dtype=torch.uint8, device=dev)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            gpu.graph_begin(s.cuda_stream)
            gpu.chain_run(ch, src.data_ptr(), dst.data_ptr(), s.cuda_stream)
            graph, nodes = gpu.graph_end(s.cuda_stream)
            # one kernel per stage (an odd-width convolution adds its padding copies; "-r0" alone is a plain copy)
            assert nodes >= meta["kernels"] or kw == dict(angle=0)
            if "conv_preset" not in kw and kw != dict(angle=0):
                assert nodes == meta["kernels"]
            gpu.graph_launch(graph, s.cuda_stream)
            s.synchronize()
        gpu.graph_free(graph)
        if "conv_preset" in kw:
            exp = orc.process(orc.conv(img, np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]]), 16, 0), angle=0, mono=True, flipv=True)
        else:
            exp = orc.process(img, **kw)
        got = dst.cpu().numpy()[:meta["out_bytes"]]
        assert (meta["out_w"], meta["out_h"], meta["file_type"]) == exp[1:] and np.array_equal(got, exp[0]), kw
        gpu.chain_free(ch)
        ph.close()
