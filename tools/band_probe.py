#!/usr/bin/env python
"""tools/band_probe.py -- how close to the HBM roofline a SEQUENCE of band-sized 3x3 launches runs on one B200,
by band height (16384 rows = the whole 16384^2 raster = 1 GPU ... 2048 rows = one of 8 bands), launch mode
(direct launches / CUDA graph) and number of streams the launches alternate over.  Development aid."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import imageprocessingtools_b200 as ip
    torch.cuda.set_device(0)
    torch.cuda.set_stream(torch.cuda.Stream())
    g = ip.Ppmx(0)
    peak, _ = bench.peaks()
    ops = {"blur3": g.conv_op(*bench.conv_spec("blur3"), 0), "box7": g.conv_op(*bench.conv_spec("box7"), 0)}
    for rows in (16384, 4096, 2048):
        B = bench.Bands(torch, g, None, 0, 1, bench.FULL, rows, 4, 0xC0FFEE ^ 4)
        for name, op in ops.items():
            for graph, lanes in ((False, 1), (True, 1), (True, 2), (True, 3), (False, 2)):
                n = 64
                run = bench.StepRunner(torch, g, B.calls(op, 0, n), 1, graph, lanes)
                t, _, _ = bench.time_runner(torch, run, 3, 2)
                nst = max(3, int(150.0 / max(t / 3, 1e-3)))
                t, _, _ = bench.time_runner(torch, run, nst, 1)
                us = t * 1e3 / (nst * n)
                gbs = 6.0 * rows * bench.FULL / (us * 1e-6) / 1e9
                print(json.dumps({"rows": rows, "op": name, "graph": graph, "lanes": lanes, "us_per_launch": round(us, 2),
                                  "gbs": round(gbs, 1), "frac": round(gbs / peak, 4), "mode": run.mode}), flush=True)
                run.close()
        B.close()
    g.close()


if __name__ == "__main__":
    main()
