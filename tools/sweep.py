#!/usr/bin/env python
"""tools/sweep.py -- measure operator variants on one B200 (development aid, not the bench).

    python tools/sweep.py --ops gray,mono --variants 0,3,4 --pdl 0,1 [--release] [--lanes 3] [--direct]

For each (op, variant, pdl) prints Mpix/s, GB/s, fraction of the measured HBM peak, and checks that the variant's
output equals variant 0's (bit for bit, on the device).  Loads libppmx_gpu_tuning.so (the release library carries
the default kernels only); --release loads the release library instead (variant 0 only; what ncu captures use).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="gray")
    ap.add_argument("--variants", default="0")
    ap.add_argument("--pdl", default="1")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--lanes", type=int, default=1)
    ap.add_argument("--release", action="store_true")
    ap.add_argument("--direct", action="store_true", help="direct launches instead of a CUDA graph")
    args = ap.parse_args()
    import torch
    import imageprocessingtools_b200 as ip
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    torch.cuda.set_stream(torch.cuda.Stream())
    g = ip.Ppmx(0, tuning=not args.release)
    peak, _ = bench.peaks()
    for op in args.ops.split(","):
        w, h, batch, bpp, _ = bench.WORKLOADS[op]
        r = bench.OpRunner(torch, g, op, dev, seed=11)
        g.set_tuning("variant", 0)
        g.set_tuning("pdl", 1)
        base = bench.StepRunner(torch, g, r.calls, 1, False, 1)
        base.step()
        torch.cuda.synchronize()
        ref_out = [d.clone() for d in r.dst]
        ref_hist = r.hist.clone() if r.hist is not None else None
        for v in [int(x) for x in args.variants.split(",")]:
            for pdl in [int(x) for x in args.pdl.split(",")]:
                g.set_tuning("variant", v)
                g.set_tuning("pdl", pdl)
                for d in r.dst:
                    d.zero_()
                if r.hist is not None:
                    r.hist.zero_()
                run = bench.StepRunner(torch, g, r.calls, 1, not args.direct, args.lanes if len(r.ops) == 1 else 1)
                run.step()
                torch.cuda.synchronize()
                same = all(torch.equal(a, b) for a, b in zip(ref_out, r.dst))
                if ref_hist is not None:
                    same = same and torch.equal(ref_hist, r.hist)
                ms, _, _ = bench.time_runner(torch, run, args.steps, 3)
                mp = args.steps * r.pixels_per_step / (ms / 1e3) / 1e6
                row = {"op": op, "variant": v, "pdl": pdl, "mpix_s": round(mp, 1), "same_as_v0": bool(same),
                       "us_per_launch": round(ms * 1e3 / (args.steps * run.launches_per_step), 2), "mode": run.mode[:28]}
                if bpp:
                    row["gbs"] = round(bpp * mp / 1e3, 1)
                    row["frac"] = round(bpp * mp / 1e3 / peak, 4)
                print(json.dumps(row), flush=True)
                run.close()
        r.close()
        del r, ref_out
        torch.cuda.empty_cache()
    g.set_tuning("variant", 0)
    g.close()


if __name__ == "__main__":
    main()
