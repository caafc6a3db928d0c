"""Replays tests/golden/golden.json against any implementation exposing the reference's
operator names (the oracle on CPU, the CUDA binding on the GPU box)."""
import hashlib
import json
import os

import numpy as np

import patterns as P

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load():
    return json.load(open(os.path.join(HERE, "golden", "golden.json")))


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


def replay(impl, header_fn, max_pixels=None, tables_from=None):
    """impl: object with gray/mono/flip/rotate/calc_contributions/imresize/process.
    header_fn(ft, w, h, maxval) -> bytes.  tables_from: where resize tables come from
    (defaults to impl).  Returns the number of vectors checked."""
    G = load()
    tables = tables_from or impl
    n = 0
    for rec in G["cases"]:
        w, h = rec["w"], rec["h"]
        if max_pixels and w * h > max_pixels:
            continue
        img = P.all_patterns(w, h)[rec["pattern"]]
        assert sha(img) == rec["input"], "pattern generator drifted"
        for op, want in rec["ops"].items():
            tag = (w, h, rec["pattern"], op)
            if op == "gray":
                assert sha(impl.gray(img)) == want, tag
            elif op == "mono_plane":
                assert sha(impl.mono(img)) == want, tag
            elif op == "flipv":
                assert sha(impl.flip(img, 1)) == want, tag
            elif op == "fliph":
                assert sha(impl.flip(img, 0)) == want, tag
            elif op.startswith("rotate"):
                r = impl.rotate(img, int(op[6:]))
                assert [r.shape[1], r.shape[0], sha(r)] == want, tag
            elif op.startswith("imresize"):
                _, d, nw = op.split("_")
                dim, nw = int(d[1:]), int(nw)
                n_in = w if dim == 1 else h
                wt, ix = tables.calc_contributions(n_in, nw, float(nw) / n_in)
                assert [wt.shape[1], sha(wt), sha(ix)] == want[:3], tag
                assert sha(impl.imresize(img, nw, dim, wt, ix)) == want[3], tag
            elif op.startswith("cli "):
                raster, ow, oh, ft = impl.process(img, **_flags(op[4:].split()))
                data = header_fn(ft, ow, oh, 255) + np.ascontiguousarray(raster).tobytes()
                assert hashlib.sha256(data).hexdigest() == want, tag
            else:
                raise AssertionError("unknown golden op " + op)
            n += 1
    return n
