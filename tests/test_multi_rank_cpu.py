"""The N>1 host logic on CPU: two gloo ranks each own a row band (ppmx_band_plan), exchange halo
rows with their neighbour, run the per-band operator (here the ORACLE stands in for the device,
this test is about partitioning and halo bookkeeping) and rank 0 stitches the result, which must
equal the whole-raster answer.  Also: the histogram reduce (sum of per-band bins)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def test_band_plan_covers_every_row_once():
    from imageprocessingtools_b200 import ppmx as pp
    for full_h in [0, 1, 3, 4, 5, 23, 64, 1080, 4096, 16384, 16385]:
        for n in [1, 2, 3, 4, 8]:
            for align in [1, 4, 16]:
                nxt = 0
                sizes = []
                for r in range(n):
                    y0, rows = pp.band_plan(full_h, n, r, align)
                    assert y0 == nxt and (y0 % align == 0 or rows == 0)
                    nxt = y0 + rows
                    sizes.append(rows)
                assert nxt == full_h
                units = [-(-s // align) for s in sizes]
                assert max(units) - min(units) <= 1


def _worker(rank, world, port, w, h, k, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    import patterns as P
    from imageprocessingtools_b200 import ppmx as pp
    orc = oracle.orc()
    img = P.lcg(w, h, 4242)          # every rank can regenerate the raster; it only USES its band (+ halos)
    r = k // 2
    y0, rows = pp.band_plan(h, world, rank, 4)
    band = torch.from_numpy(img[y0:y0 + rows].copy())
    # halo exchange with the neighbours (device path: peer pointers; here: gloo send/recv)
    top = torch.zeros((r, w, 3), dtype=torch.uint8)
    bot = torch.zeros((r, w, 3), dtype=torch.uint8)
    reqs = []
    if rank > 0:
        reqs += [dist.isend(band[:r].contiguous(), rank - 1), dist.irecv(top, rank - 1)]
    if rank < world - 1:
        reqs += [dist.isend(band[rows - r:].contiguous(), rank + 1), dist.irecv(bot, rank + 1)]
    for x in reqs:
        x.wait()
    # per-band convolution = convolution of [halo; band; halo] with the halo rows cropped away; at the
    # raster's top/bottom the mirror border applies instead of a halo
    coef = np.ones((k, k), np.int64)
    parts = ([top.numpy()] if rank > 0 else []) + [band.numpy()] + ([bot.numpy()] if rank < world - 1 else [])
    ext = np.concatenate(parts, axis=0)
    res = orc.conv(ext, coef, k * k, 0)
    a = r if rank > 0 else 0
    out = torch.from_numpy(res[a:a + rows].copy())
    # mono with the global Bayer phase: band starts are multiples of 4, so the band alone is enough
    mono = torch.from_numpy(orc.mono(band.numpy()))
    hist = torch.from_numpy(orc.hist_gray(band.numpy()).astype(np.int64))
    dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    gathered = [None] * world
    dist.gather_object((y0, rows, out.numpy(), mono.numpy()), gathered if rank == 0 else None, dst=0)
    if rank == 0:
        full = np.concatenate([g[2] for g in sorted(gathered, key=lambda t: t[0])], axis=0)
        fullm = np.concatenate([g[3] for g in sorted(gathered, key=lambda t: t[0])], axis=0)
        ok = (np.array_equal(full, orc.conv(img, coef, k * k, 0)) and np.array_equal(fullm, orc.mono(img)) and
              np.array_equal(hist.numpy().astype(np.uint64), orc.hist_gray(img)))
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("w,h,k", [(48, 37, 3), (64, 50, 7)])
def test_two_rank_bands_with_halo_exchange(w, h, k):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() * 7 + k) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, w, h, k, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs)
    assert q.get(timeout=5) is True


def _band_worker(rank, world, port, q):
    """Round 2: the partition of ppmx_gpu_apply_band (ppmx_gpu_band_rows, no device work) with two real ranks: each rank
    takes ONLY the source rows the call would read, runs the chain on them with the oracle standing in for the device,
    and rank 0 stitches; the result must equal the whole-raster chain."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    import patterns as P
    from imageprocessingtools_b200 import ppmx as pp
    orc = oracle.orc()
    w, h = 48, 61
    img = P.lcg(w, h, 777)
    ok = True
    # chain: 7x7 box convolution, then vertical flip (the mirrored band + 3 halo rows either side)
    ph = pp._PlanHolder(w=w, h=h, conv_preset=2, flipv=True)
    ops = [ph.plan.ops[i] for i in range(ph.plan.nops)]
    oy0, orows, sy0, srows = pp.band_rows(ops, w, h, rank, world)
    slab = img[sy0:sy0 + srows]                      # nothing else of the raster is touched
    conv = orc.conv(slab, np.ones((7, 7), np.int64), 49, 0)
    # output rows [oy0, oy0 + orows) of the flipped raster = convolved rows h-oy0-orows .. h-oy0-1, upside down
    a = (h - oy0 - orows) - sy0
    band = conv[a:a + orows][::-1]
    gathered = [None] * world
    dist.gather_object((oy0, np.ascontiguousarray(band)), gathered if rank == 0 else None, dst=0)
    if rank == 0:
        full = np.concatenate([g[1] for g in sorted(gathered, key=lambda t: t[0])], axis=0)
        exp = orc.flip(orc.conv(img, np.ones((7, 7), np.int64), 49, 0), 1)
        ok = ok and np.array_equal(full, exp)
    ph.close()
    # chain: resize to 1.5x (height pass reads the rows its table names), grey
    ph = pp._PlanHolder(w=w, h=h, resize_w=72, gray=True)
    ops = [ph.plan.ops[i] for i in range(ph.plan.nops)]
    oy0, orows, sy0, srows = pp.band_rows(ops, w, h, rank, world)
    whole = orc.process(img, resize_w=72, gray=True)
    new_h = whole[2]
    wt, ix = pp.calc_contributions(h, new_h, float(new_h) / h)
    need = ix[oy0:oy0 + orows]
    ok = ok and (sy0, sy0 + srows) == (int(need.min()), int(need.max()) + 1)
    covered = [None] * world
    dist.all_gather_object(covered, (oy0, orows))
    ok = ok and sorted(covered)[0][0] == 0 and sum(c[1] for c in covered) == new_h
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok))
    if rank == 0:
        q.put(all(flags))
    ph.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_apply_band_partition():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + (os.getpid() * 11) % 2000
    procs = [ctx.Process(target=_band_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs)
    assert q.get(timeout=5) is True
