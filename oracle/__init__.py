"""oracle -- TEST INFRASTRUCTURE ONLY.

ctypes front ends for the two checkers:

* ``orc``  -- oracle/libppmx_oracle.so, the plain-C restatement (ppmx_oracle.c) of the
  reference operators (/root/reference/ppmx-edward.c).
* ``ref``  -- oracle/_ref/libppmx_ref.so, the UNMODIFIED reference compiled from its own
  source by oracle/Makefile (present in the build container; travels to the GPU box as a
  built file).  ``ref()`` returns None when it is absent.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  The product (imageprocessingtools_b200) never does.

All images are numpy uint8 arrays of shape (h, w, 3) unless stated otherwise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libppmx_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libppmx_ref.so")
REF_CLI = os.path.join(_HERE, "_ref", "ppmx-edward")
REF_SOURCE = "/root/reference/ppmx-edward.c"

FT_PPM, FT_PGM, FT_PBM = 0, 1, 2

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_dblp = C.POINTER(C.c_double)
_intp = C.POINTER(C.c_int)


def build(ref: bool = True) -> None:
    """Compile the checkers (oracle always; oracle/_ref only where the reference source exists)."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", _HERE] + targets, check=True)


def _ptr(a: np.ndarray, typ=_u8p):
    return a.ctypes.data_as(typ)


def _img(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 3 and a.shape[2] == 3, a.shape
    return a


class OrcFlags(C.Structure):
    _fields_ = [("resize_enable", C.c_int), ("rotate_enable", C.c_int), ("flipv_enable", C.c_int),
                ("fliph_enable", C.c_int), ("gray_enable", C.c_int), ("mono_enable", C.c_int),
                ("resize_w", C.c_uint32), ("angle", C.c_int)]


class Oracle:
    """The C restatement.  Function names follow the reference's."""

    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        self.lib = L = C.CDLL(path)
        L.orc_gray.argtypes = [_u8p, C.c_uint32, C.c_uint32, _u8p]
        L.orc_mono.argtypes = [_u8p, C.c_uint32, C.c_uint32, _u8p]
        L.orc_pack_pbm.argtypes = [_u8p, C.c_uint32, C.c_uint32, _u8p]
        L.orc_pack_pbm.restype = C.c_size_t
        L.orc_flip.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.c_int]
        L.orc_cubic.argtypes = [C.c_double]
        L.orc_cubic.restype = C.c_double
        L.orc_mod.argtypes = [C.c_int, C.c_int]
        L.orc_rotate_size.argtypes = [C.c_double, C.c_uint32, C.c_uint32, _u32p, _u32p]
        L.orc_rotate_size.restype = None
        L.orc_rotate.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.c_double, _u8p]
        L.orc_calc_contributions.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, _intp,
                                             C.POINTER(_dblp), C.POINTER(_intp)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_free.restype = None
        L.orc_imresize.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, _dblp, _intp, C.c_int, _u8p]
        L.orc_resize.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(_u8p), _u32p, _u32p]
        L.orc_process.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.POINTER(OrcFlags), C.POINTER(_u8p),
                                  C.POINTER(C.c_size_t), _u32p, _u32p, _intp]
        L.orc_header.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orx_conv.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_int32), C.c_int32,
                               C.c_int32, _u8p]
        L.orx_hist_gray.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]
        L.orx_levels.argtypes = [_u8p, C.c_size_t, _u8p, _u8p]
        L.orx_levels_lut_linear.argtypes = [C.c_int, C.c_int, _u8p]
        L.orc_lcg_fill.argtypes = [_u8p, C.c_size_t, C.c_uint32]
        L.orc_lcg_fill.restype = None

    # -- reference operators ------------------------------------------------
    def gray(self, img) -> np.ndarray:
        img = _img(img)
        h, w, _ = img.shape
        out = np.empty((h, w), np.uint8)
        assert self.lib.orc_gray(_ptr(img), w, h, _ptr(out)) == 0
        return out

    def mono(self, img) -> np.ndarray:
        """(h, w) plane of 0/1 -- the .r member the reference stores."""
        img = _img(img)
        h, w, _ = img.shape
        out = np.empty((h, w), np.uint8)
        assert self.lib.orc_mono(_ptr(img), w, h, _ptr(out)) == 0
        return out

    def pack_pbm(self, plane) -> np.ndarray:
        plane = np.ascontiguousarray(plane, np.uint8)
        h, w = plane.shape
        out = np.zeros(h * ((w + 7) // 8) + 8, np.uint8)
        n = self.lib.orc_pack_pbm(_ptr(plane), w, h, _ptr(out))
        return out[:n].copy()

    def flip(self, img, direction: int) -> np.ndarray:
        out = _img(img).copy()
        h, w, _ = out.shape
        assert self.lib.orc_flip(_ptr(out), w, h, int(direction)) == 0
        return out

    def cubic(self, x: float) -> float:
        return self.lib.orc_cubic(float(x))

    def rotate_size(self, angle: float, w: int, h: int) -> Tuple[int, int]:
        nw, nh = C.c_uint32(), C.c_uint32()
        self.lib.orc_rotate_size(float(angle), w, h, C.byref(nw), C.byref(nh))
        return nw.value, nh.value

    def rotate(self, img, angle: float) -> np.ndarray:
        img = _img(img)
        h, w, _ = img.shape
        nw, nh = self.rotate_size(angle, w, h)
        if angle == 0:
            nw, nh = w, h
        out = np.empty((nh, nw, 3), np.uint8)
        assert self.lib.orc_rotate(_ptr(img), w, h, float(angle), _ptr(out)) == 0
        return out

    def calc_contributions(self, in_size: int, out_size: int, scale: float, k_width: float = 4.0):
        taps = C.c_int()
        wp, ip = _dblp(), _intp()
        rc = self.lib.orc_calc_contributions(in_size, out_size, scale, k_width, C.byref(taps), C.byref(wp), C.byref(ip))
        if rc != 0:
            raise ValueError("calc_contributions failed")
        k = taps.value
        w = np.ctypeslib.as_array(wp, shape=(out_size, max(k, 1))).copy()[:, :k]
        i = np.ctypeslib.as_array(ip, shape=(out_size, max(k, 1))).copy()[:, :k]
        self.lib.orc_free(wp)
        self.lib.orc_free(ip)
        return np.ascontiguousarray(w), np.ascontiguousarray(i.astype(np.int32))

    def imresize(self, img, out_size: int, dim: int, weights, indices) -> np.ndarray:
        img = _img(img)
        h, w, _ = img.shape
        weights = np.ascontiguousarray(weights, np.float64)
        indices = np.ascontiguousarray(indices, np.int32)
        taps = weights.shape[1]
        out = np.empty((out_size, w, 3) if dim == 0 else (h, out_size, 3), np.uint8)
        assert self.lib.orc_imresize(_ptr(img), w, h, out_size, dim, _ptr(weights, _dblp), _ptr(indices, _intp),
                                     taps, _ptr(out)) == 0
        return out

    def resize(self, img, new_w: int) -> np.ndarray:
        img = _img(img)
        h, w, _ = img.shape
        op = _u8p()
        ow, oh = C.c_uint32(), C.c_uint32()
        rc = self.lib.orc_resize(_ptr(img), w, h, new_w, C.byref(op), C.byref(ow), C.byref(oh))
        if rc != 0:
            raise ValueError("resize failed")
        out = np.ctypeslib.as_array(op, shape=(oh.value, ow.value, 3)).copy()
        self.lib.orc_free(op)
        return out

    def process(self, img, resize_w: Optional[int] = None, angle: Optional[int] = None, gray=False, mono=False,
                flipv=False, fliph=False):
        """Whole chain; returns (raster bytes after the header, out_w, out_h, file_type)."""
        img = _img(img)
        h, w, _ = img.shape
        f = OrcFlags(int(resize_w is not None), int(angle is not None), int(flipv), int(fliph), int(gray), int(mono),
                     int(resize_w or 0), int(angle or 0))
        op = _u8p()
        nb = C.c_size_t()
        ow, oh = C.c_uint32(), C.c_uint32()
        ft = C.c_int()
        rc = self.lib.orc_process(_ptr(img), w, h, C.byref(f), C.byref(op), C.byref(nb), C.byref(ow), C.byref(oh),
                                  C.byref(ft))
        if rc != 0:
            raise ValueError("process failed")
        out = np.ctypeslib.as_array(op, shape=(max(nb.value, 1),)).copy()[:nb.value]
        self.lib.orc_free(op)
        return out, ow.value, oh.value, ft.value

    def header(self, file_type: int, w: int, h: int, maxval: int = 255) -> bytes:
        buf = C.create_string_buffer(128)
        n = self.lib.orc_header(buf, 128, file_type, w, h, maxval)
        return buf.raw[:n]

    # -- extensions (no reference counterpart; parity unpinned) --------------
    def conv(self, img, coef, div: int = 1, bias: int = 0) -> np.ndarray:
        img = _img(img)
        h, w, _ = img.shape
        coef = np.ascontiguousarray(coef, np.int32)
        k = coef.shape[0]
        out = np.empty_like(img)
        assert self.lib.orx_conv(_ptr(img), w, h, k, _ptr(coef, C.POINTER(C.c_int32)), div, bias, _ptr(out)) == 0
        return out

    def hist_gray(self, img) -> np.ndarray:
        img = _img(img)
        h, w, _ = img.shape
        bins = np.zeros(256, np.uint64)
        assert self.lib.orx_hist_gray(_ptr(img), w, h, _ptr(bins, C.POINTER(C.c_uint64))) == 0
        return bins

    def levels(self, img, lut) -> np.ndarray:
        img = np.ascontiguousarray(img, np.uint8)
        lut = np.ascontiguousarray(lut, np.uint8)
        assert lut.size == 256
        out = np.empty_like(img)
        assert self.lib.orx_levels(_ptr(img), img.size, _ptr(lut), _ptr(out)) == 0
        return out

    def levels_lut_linear(self, lo: int, hi: int) -> np.ndarray:
        lut = np.zeros(256, np.uint8)
        assert self.lib.orx_levels_lut_linear(lo, hi, _ptr(lut)) == 0
        return lut

    def lcg_image(self, w: int, h: int, seed: int) -> np.ndarray:
        out = np.empty((h, w, 3), np.uint8)
        self.lib.orc_lcg_fill(_ptr(out), w * h, seed & 0xFFFFFFFF)
        return out


class Ref:
    """The compiled, unmodified reference (oracle/_ref/libppmx_ref.so)."""

    def __init__(self, path: str = REF_SO):
        self.lib = L = C.CDLL(path)
        L.ref_gray.argtypes = [_u8p, C.c_uint32, C.c_uint32, _u8p, _intp, _dblp]
        L.ref_mono.argtypes = [_u8p, C.c_uint32, C.c_uint32, _u8p, _intp, _dblp]
        L.ref_flip.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.c_int, _u8p, _dblp]
        L.ref_rotate_size.argtypes = [C.c_double, C.c_uint32, C.c_uint32, _u32p, _u32p]
        L.ref_rotate_size.restype = None
        L.ref_rotate.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.c_double, _u8p, _u32p, _u32p, _dblp]
        L.ref_cubic.argtypes = [C.c_double]
        L.ref_cubic.restype = C.c_double
        L.ref_mod.argtypes = [C.c_int, C.c_int]
        L.ref_calc_contributions.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, _intp, C.POINTER(_dblp),
                                             C.POINTER(_intp)]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_free.restype = None
        L.ref_imresize.argtypes = [_u8p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, _dblp, _intp, C.c_int, _u8p, _dblp]
        self.last_seconds = 0.0

    def _t(self):
        return C.c_double(0.0)

    def gray(self, img):
        """Full pixel buffer (h, w, 3): .r = grey, .g = .b = 0; and the file type set."""
        img = _img(img)
        h, w, _ = img.shape
        out = np.empty_like(img)
        ft, t = C.c_int(), self._t()
        assert self.lib.ref_gray(_ptr(img), w, h, _ptr(out), C.byref(ft), C.byref(t)) == 0
        self.last_seconds = t.value
        return out, ft.value

    def mono(self, img):
        img = _img(img)
        h, w, _ = img.shape
        out = np.empty_like(img)
        ft, t = C.c_int(), self._t()
        assert self.lib.ref_mono(_ptr(img), w, h, _ptr(out), C.byref(ft), C.byref(t)) == 0
        self.last_seconds = t.value
        return out, ft.value

    def flip(self, img, direction: int):
        img = _img(img)
        h, w, _ = img.shape
        out = np.empty_like(img)
        t = self._t()
        assert self.lib.ref_flip(_ptr(img), w, h, int(direction), _ptr(out), C.byref(t)) == 0
        self.last_seconds = t.value
        return out

    def rotate_size(self, angle: float, w: int, h: int):
        nw, nh = C.c_uint32(), C.c_uint32()
        self.lib.ref_rotate_size(float(angle), w, h, C.byref(nw), C.byref(nh))
        return nw.value, nh.value

    def rotate(self, img, angle: float):
        img = _img(img)
        h, w, _ = img.shape
        nw, nh = self.rotate_size(angle, w, h)
        if angle == 0:
            nw, nh = w, h
        out = np.empty((nh, nw, 3), np.uint8)
        rw, rh, t = C.c_uint32(), C.c_uint32(), self._t()
        assert self.lib.ref_rotate(_ptr(img), w, h, float(angle), _ptr(out), C.byref(rw), C.byref(rh), C.byref(t)) == 0
        assert (rw.value, rh.value) == (nw, nh)
        self.last_seconds = t.value
        return out

    def cubic(self, x: float) -> float:
        return self.lib.ref_cubic(float(x))

    def calc_contributions(self, in_size: int, out_size: int, scale: float, k_width: float = 4.0):
        taps = C.c_int()
        wp, ip = _dblp(), _intp()
        rc = self.lib.ref_calc_contributions(in_size, out_size, scale, k_width, C.byref(taps), C.byref(wp), C.byref(ip))
        if rc != 0:
            raise ValueError("calc_contributions failed")
        k = taps.value
        w = np.ctypeslib.as_array(wp, shape=(out_size, max(k, 1))).copy()[:, :k]
        i = np.ctypeslib.as_array(ip, shape=(out_size, max(k, 1))).copy()[:, :k]
        self.lib.ref_free(wp)
        self.lib.ref_free(ip)
        return np.ascontiguousarray(w), np.ascontiguousarray(i.astype(np.int32))

    def imresize(self, img, out_size: int, dim: int, weights, indices):
        img = _img(img)
        h, w, _ = img.shape
        weights = np.ascontiguousarray(weights, np.float64)
        indices = np.ascontiguousarray(indices, np.int32)
        taps = weights.shape[1]
        out = np.empty((out_size, w, 3) if dim == 0 else (h, out_size, 3), np.uint8)
        t = self._t()
        assert self.lib.ref_imresize(_ptr(img), w, h, out_size, dim, _ptr(weights, _dblp), _ptr(indices, _intp), taps,
                                     _ptr(out), C.byref(t)) == 0
        self.last_seconds = t.value
        return out


_orc: Optional[Oracle] = None
_ref: Optional[Ref] = None


def orc() -> Oracle:
    global _orc
    if _orc is None:
        _orc = Oracle()
    return _orc


def ref() -> Optional[Ref]:
    """The compiled reference, built on demand where its source exists, else None."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO) and os.path.exists(REF_SOURCE):
            build(ref=True)
        if os.path.exists(REF_SO):
            _ref = Ref()
    return _ref


def ref_cli(args, ppm_path: str) -> Tuple[int, str]:
    """Run the compiled reference CLI (writes <ppm_path>.out); returns (exit code, stdout)."""
    p = subprocess.run([REF_CLI] + list(args) + [ppm_path], capture_output=True, text=True)
    return p.returncode, p.stdout


def write_p6(path: str, img, maxval: int = 255, comment: Optional[str] = None) -> None:
    img = _img(img)
    h, w, _ = img.shape
    with open(path, "wb") as f:
        f.write(b"P6\n")
        if comment is not None:
            f.write(b"# " + comment.encode() + b"\n")
        f.write(b"%d %d\n%d\n" % (w, h, maxval))
        f.write(img.tobytes())
