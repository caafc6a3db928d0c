"""Records golden vectors from the UNMODIFIED reference (oracle/_ref, compiled from
/root/reference/ppmx-edward.c by oracle/Makefile).  Run in the build container:

    python tests/golden/make_golden.py

Writes tests/golden/golden.json (sha256 digests of reference outputs for seeded inputs that
tests/patterns.py regenerates) and tests/golden/small_vectors.npz (a few complete
input/output pairs small enough to read by eye).  tests/test_golden.py replays them against
the oracle on CPU and against the CUDA path on the GPU box, where /root/reference does not exist.
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle  # noqa: E402
import patterns as P  # noqa: E402

SIZES = [(1, 1), (5, 3), (13, 7), (16, 4), (37, 23), (64, 48), (100, 37), (301, 211), (512, 512)]
PATS = ["lcg", "c200", "checker", "bayer", "mixed"]
ANGLES = [0, 90, 180, 270, 30, 45, 77, 135, 200, 331]
WIDTHS = lambda w: sorted({max(1, w // 3), max(1, w // 2), w, w + 1, w * 3 // 2, 2 * w})
CHAINS = [["-w%d", "-r90", "-gray", "-fv"], ["-r90", "-mono", "-fh"], ["-gray", "-fh"], ["-mono", "-fv"],
          ["-w%d", "-r30", "-mono"], ["-r270", "-fh"]]


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = oracle.ref()
    assert ref is not None, "needs the compiled reference"
    G = {"source": "oracle/_ref built from /root/reference/ppmx-edward.c (gcc -O2 -ffp-contract=off)", "cases": []}
    small = {}
    tmp = tempfile.mkdtemp()
    for (w, h) in SIZES:
        for name in PATS:
            img = P.all_patterns(w, h)[name]
            rec = {"w": w, "h": h, "pattern": name, "input": sha(img), "ops": {}}
            g, _ = ref.gray(img)
            rec["ops"]["gray"] = sha(g[..., 0])
            m, _ = ref.mono(img)
            rec["ops"]["mono_plane"] = sha(m[..., 0])
            rec["ops"]["flipv"] = sha(ref.flip(img, 1))
            rec["ops"]["fliph"] = sha(ref.flip(img, 0))
            big = w * h > 100000
            for a in (ANGLES[:6] if big else ANGLES):
                r = ref.rotate(img, a)
                rec["ops"]["rotate%d" % a] = [r.shape[1], r.shape[0], sha(r)]
            for nw in ([w // 2, w * 3 // 2] if big else WIDTHS(w)):
                for dim, n_in in ((1, w), (0, h)):
                    wt, ix = ref.calc_contributions(n_in, nw, float(nw) / n_in)
                    r = ref.imresize(img, nw, dim, wt, ix)
                    rec["ops"]["imresize_d%d_%d" % (dim, nw)] = [wt.shape[1], sha(wt), sha(ix), sha(r)]
            # whole-file outputs of the reference CLI (header + raster), incl. chaining quirks
            p = os.path.join(tmp, "in.ppm")
            oracle.write_p6(p, img)
            if w >= 4 and h >= 4 and not big:
                for ch in CHAINS + [["-w%d"]]:
                    for nw in (max(2, w // 2), w * 3 // 2):
                        args = [a % nw if "%d" in a else a for a in ch]
                        if os.path.exists(p + ".out"):
                            os.remove(p + ".out")
                        rc, _ = oracle.ref_cli(args, p)
                        if rc != 0 or not os.path.exists(p + ".out"):
                            continue
                        rec["ops"]["cli " + " ".join(args)] = sha(np.frombuffer(open(p + ".out", "rb").read(), np.uint8))
                        if "%d" not in "".join(ch):
                            break
            G["cases"].append(rec)
            if (w, h) in ((5, 3), (13, 7)) and name in ("lcg", "mixed"):
                k = "%dx%d_%s" % (w, h, name)
                small[k + "_in"] = img
                small[k + "_gray"] = g[..., 0]
                small[k + "_mono"] = m[..., 0]
                small[k + "_rot30"] = ref.rotate(img, 30)
                small[k + "_rot90"] = ref.rotate(img, 90)
                wt, ix = ref.calc_contributions(w, 2 * w, 2.0)
                small[k + "_w2x_weights"] = wt
                small[k + "_w2x_indices"] = ix
                small[k + "_w2x"] = ref.imresize(img, 2 * w, 1, wt, ix)
    json.dump(G, open(os.path.join(HERE, "golden.json"), "w"), indent=0, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "small_vectors.npz"), **small)
    print("cases:", len(G["cases"]), "ops:", sum(len(c["ops"]) for c in G["cases"]))


if __name__ == "__main__":
    main()
