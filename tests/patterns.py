"""Synthetic rasters shared by the CPU and GPU parity tests."""
import numpy as np

# 4x4 Bayer thresholds in the reference's x-major index order
BAYER_THR = [32, 255, 48, 208, 160, 96, 176, 112, 64, 224, 16, 240, 192, 128, 144, 80]

SMALL_SIZES = [(1, 1), (2, 2), (3, 5), (5, 3), (13, 7), (16, 4), (17, 9), (31, 33), (48, 16), (64, 64),
               (100, 37), (301, 211), (512, 512)]
ODD_WIDTHS = [(w, 6) for w in list(range(1, 34)) + [47, 63, 65, 127, 129, 255, 257]]


def lcg(w, h, seed):
    """s = s*1664525 + 1013904223 mod 2^32 per pixel; r = s>>24, g = s>>16, b = s>>8."""
    n = w * h
    out = np.empty((n, 3), np.uint8)
    # vectorised by block doubling: the map for a jump of k steps is affine (A, Cc) mod 2^32
    a = np.uint32(1664525)
    c = np.uint32(1013904223)
    st = np.empty(n, np.uint32)
    cur = np.uint32(seed & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        if n:
            st[0] = cur * a + c
        filled = 1
        A, Cc = a, c  # affine map for a jump of `filled` steps
        while filled < n:
            m = min(filled, n - filled)
            st[filled:filled + m] = st[:m] * A + Cc
            # compose jump by `filled` with itself
            Cc = np.uint32(A * Cc + Cc)
            A = np.uint32(A * A)
            filled += m
    out[:, 0] = (st >> 24).astype(np.uint8)
    out[:, 1] = (st >> 16).astype(np.uint8)
    out[:, 2] = (st >> 8).astype(np.uint8)
    return out.reshape(h, w, 3)


def const(w, h, v):
    return np.full((h, w, 3), v, np.uint8)


def xramp(w, h):
    x = (np.arange(w) * 255 // max(w - 1, 1)).astype(np.uint8)
    return np.broadcast_to(x[None, :, None], (h, w, 3)).copy()


def yramp(w, h):
    y = (np.arange(h) * 255 // max(h - 1, 1)).astype(np.uint8)
    return np.broadcast_to(y[:, None, None], (h, w, 3)).copy()


def checker(w, h, cell=4):
    yy, xx = np.mgrid[0:h, 0:w]
    v = (((xx // cell) + (yy // cell)) % 2 * 255).astype(np.uint8)
    return np.repeat(v[:, :, None], 3, axis=2)


def bayer_edges(w, h):
    """Greys sitting at each Bayer threshold -1/0/+1, cycling over the image."""
    vals = []
    for t in BAYER_THR:
        vals += [max(t - 1, 0), t, min(t + 1, 255)]
    vals = np.array(vals, np.uint8)
    idx = (np.arange(w * h) // 3) % len(vals)
    g = vals[idx].reshape(h, w)
    return np.repeat(g[:, :, None], 3, axis=2)


def mixed(w, h, seed=1):
    """Random image whose channels differ strongly (catches channel-order mistakes)."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    a[..., 1] //= 2
    a[..., 2] = 255 - a[..., 2] // 3
    return a


def all_patterns(w, h, seed=0xC0FFEE):
    return {
        "lcg": lcg(w, h, seed ^ (w * 7919 + h)),
        "zero": const(w, h, 0),
        "c200": const(w, h, 200),
        "c255": const(w, h, 255),
        "xramp": xramp(w, h),
        "yramp": yramp(w, h),
        "checker": checker(w, h),
        "bayer": bayer_edges(w, h),
        "mixed": mixed(w, h),
    }
