"""The oracle against the golden vectors recorded from the compiled reference
(tests/golden/make_golden.py).  CPU only; needs neither /root/reference nor oracle/_ref."""
import os

import numpy as np

import golden_replay
import patterns as P


def test_oracle_matches_golden_digests(orc):
    n = golden_replay.replay(orc, orc.header)
    assert n > 1000


def test_oracle_matches_small_vectors(orc):
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "small_vectors.npz"))
    keys = sorted({k.rsplit("_in", 1)[0] for k in z.files if k.endswith("_in")})
    assert keys
    for k in keys:
        img = z[k + "_in"]
        assert np.array_equal(orc.gray(img), z[k + "_gray"])
        assert np.array_equal(orc.mono(img), z[k + "_mono"])
        assert np.array_equal(orc.rotate(img, 30), z[k + "_rot30"])
        assert np.array_equal(orc.rotate(img, 90), z[k + "_rot90"])
        w = img.shape[1]
        wt, ix = orc.calc_contributions(w, 2 * w, 2.0)
        assert np.array_equal(wt.view(np.uint64), z[k + "_w2x_weights"].view(np.uint64))
        assert np.array_equal(ix, z[k + "_w2x_indices"])
        assert np.array_equal(orc.imresize(img, 2 * w, 1, wt, ix), z[k + "_w2x"])


def test_pack_pbm_closed_form(orc):
    """MSB-first, rows padded to a byte; arbitrary .r bytes spill like the reference's int shift."""
    rng = np.random.default_rng(3)
    for (w, h) in [(1, 1), (7, 3), (8, 2), (9, 2), (16, 1), (37, 5)]:
        bits = rng.integers(0, 2, (h, w), dtype=np.uint8)
        got = orc.pack_pbm(bits)
        exp = np.packbits(bits, axis=1).reshape(-1)
        assert np.array_equal(got, exp), (w, h)
        raw = rng.integers(0, 256, (h, w), dtype=np.uint8)  # the "-mono -fh" quirk packs raw red bytes
        got = orc.pack_pbm(raw)
        exp = np.zeros((h, (w + 7) // 8), np.uint8)
        for y in range(h):
            for x in range(w):
                exp[y, x // 8] |= (int(raw[y, x]) << (7 - x % 8)) & 0xFF
        assert np.array_equal(got, exp.reshape(-1)), (w, h)
