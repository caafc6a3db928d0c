"""ctypes binding of the C ABI (include/ppmx_gpu.h) and the C host layer (include/ppmx_host.h).

Method names follow the reference's functions (/root/reference/ppmx-edward.c: gray ref:986,
mono ref:949, flip ref:888, rotate ref:673, imresize ref:808, calc_contributions ref:516) so the
parity tests read like calls into the reference.  Nothing here computes pixels.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
GPU_SO = os.path.join(PKG, "libppmx_gpu.so")
TUNING_SO = os.path.join(PKG, "libppmx_gpu_tuning.so")  # same sources, -DPPMX_TUNING: alternative kernel variants
HOST_SO = os.path.join(PKG, "libppmx_host.so")
CLI = os.path.join(PKG, "ppmx-b200")

FT_PPM, FT_PGM, FT_PBM = 0, 1, 2
LAYOUT_RGB8, LAYOUT_R8, LAYOUT_BITS = 0, 1, 2
OP_GRAY, OP_MONO, OP_FLIP, OP_ROTATE, OP_IMRESIZE, OP_MONO_BITS, OP_PACK_PBM, OP_EXTRACT_R = range(8)
OP_CONV, OP_HIST_GRAY, OP_GRAY_HIST, OP_LEVELS = 16, 17, 18, 19

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_dblp = C.POINTER(C.c_double)


class PpmxError(RuntimeError):
    pass


class PpmxOp(C.Structure):
    """struct ppmx_op of include/ppmx_gpu.h"""
    _fields_ = [("kind", C.c_int32), ("renew_before", C.c_int32), ("flip_direction", C.c_int32),
                ("angle_deg", C.c_int32), ("cos_t", C.c_double), ("sin_t", C.c_double),
                ("new_width", C.c_uint32), ("new_height", C.c_uint32),
                ("dim", C.c_int32), ("out_size", C.c_int32), ("weights_sz", C.c_int32),
                ("weights", _dblp), ("indices", _i32p),
                ("conv_k", C.c_int32), ("conv_div", C.c_int32), ("conv_bias", C.c_int32), ("conv_coef", _i32p),
                ("hist_out", C.POINTER(C.c_uint64)), ("levels_lut", C.POINTER(C.c_uint8))]


class PpmxBand(C.Structure):
    """struct ppmx_band of include/ppmx_gpu.h"""
    _fields_ = [("full_h", C.c_uint32), ("y0", C.c_uint32), ("d_top", C.c_void_p), ("d_bottom", C.c_void_p),
                ("halo", C.c_uint32), ("out_y0", C.c_uint32), ("out_rows", C.c_uint32)]


class _ArgsFlag(C.Structure):
    _fields_ = [(n, C.c_char) for n in ("resize_enable", "rotate_enable", "flipv_enable", "fliph_enable",
                                        "gray_enable", "mono_enable")]


class _Contrib(C.Structure):
    _fields_ = [("weights", _dblp), ("indices", _i32p), ("weights_sz", C.c_int), ("out_size", C.c_int)]


PLAN_MAX_OPS = 12  # PPMX_PLAN_MAX_OPS of include/ppmx_host.h


class _Plan(C.Structure):
    _fields_ = [("ops", PpmxOp * PLAN_MAX_OPS), ("nops", C.c_int), ("contrib", _Contrib * 2), ("levels_lut", C.c_uint8 * 256)]


_gpu = None
_tuning = None
_host = None


def gpu_lib(tuning: bool = False) -> C.CDLL:
    """libppmx_gpu.so (or the tuning build); raises if it has not been built (no fallback)."""
    global _gpu, _tuning
    if tuning:
        if _tuning is None:
            if not os.path.exists(TUNING_SO):
                raise PpmxError("libppmx_gpu_tuning.so is not built")
            _tuning = _declare(C.CDLL(TUNING_SO, mode=C.RTLD_LOCAL))
        return _tuning
    if _gpu is None:
        if not os.path.exists(GPU_SO):
            raise PpmxError("libppmx_gpu.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _gpu = _declare(C.CDLL(GPU_SO, mode=C.RTLD_GLOBAL))
    return _gpu


def _declare(L: C.CDLL) -> C.CDLL:
    if True:
        vp = C.c_void_p
        L.ppmx_gpu_init.argtypes = [C.POINTER(vp), C.c_int]
        L.ppmx_gpu_init_multi.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int]
        L.ppmx_gpu_device_count.argtypes = [vp]
        L.ppmx_gpu_host_register.argtypes = [vp, vp, C.c_size_t]
        L.ppmx_gpu_host_unregister.argtypes = [vp, vp]
        L.ppmx_gpu_apply_band.argtypes = [vp, C.POINTER(PpmxOp), C.c_int, vp, C.c_uint32, C.c_uint32, C.c_int, C.c_int, vp,
                                          C.c_size_t, C.POINTER(C.c_size_t), _u32p, _u32p, C.POINTER(C.c_int), _u32p, _u32p]
        L.ppmx_gpu_chain_info.argtypes = [C.POINTER(PpmxOp), C.c_int, C.c_uint32, C.c_uint32, _u32p, _u32p,
                                          C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ppmx_gpu_band_rows.argtypes = [C.POINTER(PpmxOp), C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int, _u32p, _u32p,
                                         _u32p, _u32p]
        L.ppmx_gpu_chain_prepare.argtypes = [vp, C.POINTER(PpmxOp), C.c_int, C.c_uint32, C.c_uint32, C.POINTER(vp)]
        L.ppmx_gpu_chain_run.argtypes = [vp, vp, vp, vp]
        L.ppmx_gpu_chain_info2.argtypes = [vp, _u32p, _u32p, C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_int),
                                           C.POINTER(C.c_size_t)]
        L.ppmx_gpu_chain_free.argtypes = [vp]
        L.ppmx_gpu_chain_free.restype = None
        L.ppmx_gpu_graph_begin.argtypes = [vp]
        L.ppmx_gpu_graph_end.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
        L.ppmx_gpu_graph_launch.argtypes = [vp, vp]
        L.ppmx_gpu_graph_free.argtypes = [vp]
        L.ppmx_gpu_graph_free.restype = None
        L.ppmx_gpu_free.argtypes = [vp]
        L.ppmx_gpu_free.restype = None
        L.ppmx_gpu_host_alloc.argtypes = [vp, C.c_size_t]
        L.ppmx_gpu_host_alloc.restype = vp
        L.ppmx_gpu_host_free.argtypes = [vp, vp]
        L.ppmx_gpu_host_free.restype = None
        L.ppmx_gpu_apply.argtypes = [vp, C.POINTER(PpmxOp), C.c_int, vp, C.c_uint32, C.c_uint32, vp, C.c_size_t,
                                     C.POINTER(C.c_size_t), _u32p, _u32p, C.POINTER(C.c_int)]
        L.ppmx_gpu_apply_batch.argtypes = [vp, C.POINTER(PpmxOp), C.c_int, vp, C.c_uint32, C.c_uint32, C.c_int, vp,
                                           C.c_size_t, C.POINTER(C.c_size_t), _u32p, _u32p, C.POINTER(C.c_int)]
        L.ppmx_gpu_upload.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(vp)]
        L.ppmx_gpu_image_alloc.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(vp)]
        L.ppmx_gpu_image_free.argtypes = [vp, vp]
        L.ppmx_gpu_image_free.restype = None
        L.ppmx_gpu_image_info.argtypes = [vp, _u32p, _u32p, C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(vp)]
        L.ppmx_gpu_op.argtypes = [vp, C.POINTER(PpmxOp), vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
        L.ppmx_gpu_download.argtypes = [vp, vp, C.c_int, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.ppmx_gpu_sync.argtypes = [vp]
        L.ppmx_gpu_launch.argtypes = [C.POINTER(PpmxOp), vp, C.c_uint32, C.c_uint32, C.c_int, vp, C.POINTER(PpmxBand),
                                      vp, vp, vp]
        L.ppmx_gpu_tables_upload.argtypes = [C.POINTER(PpmxOp), C.POINTER(vp)]
        L.ppmx_gpu_tables_free.argtypes = [vp]
        L.ppmx_gpu_tables_free.restype = None
        L.ppmx_gpu_layout_bytes.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
        L.ppmx_gpu_layout_bytes.restype = C.c_size_t
        L.ppmx_gpu_op_output.argtypes = [C.POINTER(PpmxOp), C.c_uint32, C.c_uint32, C.c_int, _u32p, _u32p,
                                         C.POINTER(C.c_int)]
        L.ppmx_gpu_device_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
        L.ppmx_gpu_device_free.argtypes = [vp, vp]
        L.ppmx_gpu_device_free.restype = None
        L.ppmx_gpu_copy.argtypes = [vp, vp, vp, C.c_size_t, C.c_int]
        L.ppmx_gpu_ipc_export.argtypes = [vp, vp, _u8p]
        L.ppmx_gpu_ipc_open.argtypes = [vp, _u8p, C.POINTER(vp)]
        L.ppmx_gpu_ipc_close.argtypes = [vp, vp]
        L.ppmx_gpu_set_tuning.argtypes = [C.c_char_p, C.c_int]
        L.ppmx_gpu_launch_count.restype = C.c_uint64
        L.ppmx_gpu_version.restype = C.c_char_p
    return L


def host_lib() -> C.CDLL:
    """libppmx_host.so (C host layer); raises if it has not been built."""
    global _host
    if _host is None:
        gpu_lib()
        if not os.path.exists(HOST_SO):
            raise PpmxError("libppmx_host.so is not built")
        H = C.CDLL(HOST_SO)
        H.ppmx_cubic.argtypes = [C.c_double]
        H.ppmx_cubic.restype = C.c_double
        H.ppmx_mod.argtypes = [C.c_int, C.c_int]
        H.ppmx_calc_contributions.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(_Contrib)]
        H.ppmx_contributions_free.argtypes = [C.POINTER(_Contrib)]
        H.ppmx_contributions_free.restype = None
        H.ppmx_calc_rot_size.argtypes = [C.c_double, C.c_uint, C.c_uint, _u32p, _u32p]
        H.ppmx_calc_rot_size.restype = None
        H.ppmx_plan_chain.argtypes = [C.POINTER(_ArgsFlag), C.c_uint, C.c_double, C.c_uint, C.c_uint, C.POINTER(_Plan)]
        H.ppmx_plan_chain_ext.argtypes = [C.POINTER(_ArgsFlag), C.c_uint, C.c_double, C.c_uint, C.c_uint, C.c_int,
                                          C.POINTER(_Plan)]
        H.ppmx_plan_chain_ext2.argtypes = [C.POINTER(_ArgsFlag), C.c_uint, C.c_double, C.c_uint, C.c_uint, C.c_int, C.c_int,
                                           C.c_int, C.POINTER(_Plan)]
        H.ppmx_levels_lut_linear.argtypes = [C.c_int, C.c_int, C.c_void_p]
        H.ppmx_levels_points_from_hist.argtypes = [C.c_void_p, C.c_uint, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        H.ppmx_plan_free.argtypes = [C.POINTER(_Plan)]
        H.ppmx_plan_free.restype = None
        H.ppmx_band_plan.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_uint, _u32p, _u32p]
        H.ppmx_parse_header.argtypes = [C.c_char_p, C.c_size_t, _u32p, _u32p, _u32p, C.POINTER(C.c_size_t)]
        H.ppmx_format_header.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_uint, C.c_uint, C.c_uint]
        H.ppmx_probe_pnm.argtypes = [C.c_char_p, C.c_size_t, _u32p, _u32p, _u32p, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
        H.ppmx_decode_pnm.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_void_p,
                                      _u32p]
        H.ppmx_synth_lcg.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint32]
        H.ppmx_synth_lcg.restype = None
        _host = H
    return _host


def _img(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if a.ndim != 3 or a.shape[2] != 3:
        raise ValueError("expected an (h, w, 3) uint8 raster, got %r" % (a.shape,))
    return a


def _vp(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


# ---- host-only helpers (no GPU needed) ----------------------------------------------------

def calc_contributions(in_size: int, out_size: int, scale: float, k_width: float = 4.0):
    """ppmx_calc_contributions (host C, ref:516-641) -> (weights[out][K] f64, indices[out][K] i32)."""
    H = host_lib()
    c = _Contrib()
    if H.ppmx_calc_contributions(in_size, out_size, float(scale), float(k_width), C.byref(c)) != 0:
        raise PpmxError("ppmx_calc_contributions failed")
    k = c.weights_sz
    w = np.ctypeslib.as_array(c.weights, shape=(out_size, max(k, 1))).copy()[:, :k]
    i = np.ctypeslib.as_array(c.indices, shape=(out_size, max(k, 1))).copy()[:, :k]
    H.ppmx_contributions_free(C.byref(c))
    return np.ascontiguousarray(w), np.ascontiguousarray(i)


def rotate_size(angle: float, w: int, h: int) -> Tuple[int, int]:
    """Output size rotate() uses: angle folded into [0,90] (ref:687-691) then ppmx_calc_rot_size."""
    a = float(angle)
    if a >= 270:
        a = 360 - a
    elif a > 180:
        a = a - 180
    elif a > 90:
        a = 180 - a
    nw, nh = C.c_uint32(), C.c_uint32()
    host_lib().ppmx_calc_rot_size(a, w, h, C.byref(nw), C.byref(nh))
    return nw.value, nh.value


def band_plan(full_h: int, nranks: int, rank: int, align: int = 1) -> Tuple[int, int]:
    """ppmx_band_plan (host C): rows [y0, y0+rows) of a full_h-row raster owned by `rank`."""
    y0, rows = C.c_uint32(), C.c_uint32()
    if host_lib().ppmx_band_plan(full_h, nranks, rank, align, C.byref(y0), C.byref(rows)) != 0:
        raise PpmxError("ppmx_band_plan failed")
    return y0.value, rows.value


def synth_lcg(w: int, rows: int, seed: int, y0: int = 0, out: Optional[np.ndarray] = None) -> np.ndarray:
    """Rows [y0, y0 + rows) of the w-wide LCG raster of SURVEY.md 8d (ppmx_synth_lcg, host C)."""
    if out is None:
        out = np.empty((rows, w, 3), np.uint8)
    assert out.size == rows * w * 3 and out.flags["C_CONTIGUOUS"]
    host_lib().ppmx_synth_lcg(C.c_void_p(out.ctypes.data), y0 * w, rows * w, seed & 0xFFFFFFFF)
    return out


def chain_info(ops, w: int, h: int):
    """ppmx_gpu_chain_info (no device work): (out_w, out_h, file_type, out_bytes, splittable, kernels_per_part)."""
    arr = (PpmxOp * len(ops))(*ops)
    ow, oh, ft, nb, sp, k = C.c_uint32(), C.c_uint32(), C.c_int(), C.c_size_t(), C.c_int(), C.c_int()
    if gpu_lib().ppmx_gpu_chain_info(arr, len(ops), w, h, C.byref(ow), C.byref(oh), C.byref(ft), C.byref(nb), C.byref(sp),
                                     C.byref(k)) != 0:
        raise PpmxError("ppmx_gpu_chain_info failed")
    return ow.value, oh.value, ft.value, nb.value, bool(sp.value), k.value


def band_rows(ops, w: int, h: int, band: int, nbands: int):
    """ppmx_gpu_band_rows (no device work): (out_y0, out_rows, src_y0, src_rows) of one band of ppmx_gpu_apply_band."""
    arr = (PpmxOp * len(ops))(*ops)
    v = [C.c_uint32() for _ in range(4)]
    if gpu_lib().ppmx_gpu_band_rows(arr, len(ops), w, h, band, nbands, *[C.byref(x) for x in v]) != 0:
        raise PpmxError("ppmx_gpu_band_rows failed")
    return tuple(x.value for x in v)


def parse_header(data: bytes):
    w, h, m = C.c_uint32(), C.c_uint32(), C.c_uint32()
    off = C.c_size_t()
    rc = host_lib().ppmx_parse_header(data, len(data), C.byref(w), C.byref(h), C.byref(m), C.byref(off))
    if rc != 0:
        raise PpmxError("ppmx_parse_header failed")
    return w.value, h.value, m.value, off.value


def decode_pnm(data: bytes):
    """EXTENSION (ppmx_probe_pnm + ppmx_decode_pnm): a P3 or P6 file of any maxval -> ((h, w, 3) uint8, out maxval)."""
    H = host_lib()
    w, h, m, fmt = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_int()
    off = C.c_size_t()
    if H.ppmx_probe_pnm(data, len(data), C.byref(w), C.byref(h), C.byref(m), C.byref(off), C.byref(fmt)) != 0:
        raise PpmxError("ppmx_probe_pnm failed")
    out = np.empty((h.value, w.value, 3), np.uint8)
    om = C.c_uint32()
    if H.ppmx_decode_pnm(data, len(data), off.value, fmt.value, w.value, h.value, m.value,
                         C.c_void_p(out.ctypes.data), C.byref(om)) != 0:
        raise PpmxError("ppmx_decode_pnm failed")
    return out, om.value


def format_header(file_type: int, w: int, h: int, maxval: int = 255) -> bytes:
    buf = C.create_string_buffer(128)
    n = host_lib().ppmx_format_header(buf, 128, file_type, w, h, maxval)
    return buf.raw[:n]


class _PlanHolder:
    def __init__(self, resize_w=None, angle=None, gray=False, mono=False, flipv=False, fliph=False, w=0, h=0,
                 conv_preset=0, levels=None):
        self.plan = _Plan()
        f = _ArgsFlag(bytes([int(resize_w is not None)]), bytes([int(angle is not None)]), bytes([int(flipv)]),
                      bytes([int(fliph)]), bytes([int(gray)]), bytes([int(mono)]))
        lo, hi = levels if levels is not None else (-1, -1)
        rc = host_lib().ppmx_plan_chain_ext2(C.byref(f), int(resize_w or 0), float(angle or 0), w, h, int(conv_preset),
                                             int(lo), int(hi), C.byref(self.plan))
        if rc != 0:
            raise PpmxError("ppmx_plan_chain failed")

    def close(self):
        host_lib().ppmx_plan_free(C.byref(self.plan))


# ---- the device context --------------------------------------------------------------------

class Ppmx:
    """One ppmx_gpu_ctx.  Every method goes through the C ABI; rasters are numpy (h, w, 3) uint8."""

    def __init__(self, device=0, tuning: bool = False):
        """device: an int (one GPU), or a list of device numbers / "all" for a multi-device context
        (ppmx_gpu_init_multi).  tuning=True loads libppmx_gpu_tuning.so (kernel variants for sweeps)."""
        self.L = gpu_lib(tuning)
        self.ctx = C.c_void_p()
        if isinstance(device, int):
            rc = self.L.ppmx_gpu_init(C.byref(self.ctx), device)
        elif device == "all":
            rc = self.L.ppmx_gpu_init_multi(C.byref(self.ctx), None, 0)
        else:
            devs = (C.c_int * len(device))(*device)
            rc = self.L.ppmx_gpu_init_multi(C.byref(self.ctx), devs, len(device))
        if rc != 0:
            raise PpmxError("ppmx_gpu_init failed (no B200 visible?) -- there is no CPU fallback")
        self.device = device

    def device_count(self) -> int:
        return int(self.L.ppmx_gpu_device_count(self.ctx))

    def close(self):
        if self.ctx:
            self.L.ppmx_gpu_free(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- device rasters ---------------------------------------------------------------------
    def upload(self, arr: np.ndarray, layout: int = LAYOUT_RGB8) -> C.c_void_p:
        arr = np.ascontiguousarray(arr, np.uint8)
        h, w = arr.shape[0], arr.shape[1]
        img = C.c_void_p()
        if self.L.ppmx_gpu_upload(self.ctx, _vp(arr), w, h, layout, C.byref(img)) != 0:
            raise PpmxError("ppmx_gpu_upload failed")
        self.L.ppmx_gpu_sync(self.ctx)  # arr may be a temporary
        return img

    def info(self, img):
        w, h, lay = C.c_uint32(), C.c_uint32(), C.c_int()
        nb = C.c_size_t()
        ptr = C.c_void_p()
        self.L.ppmx_gpu_image_info(img, C.byref(w), C.byref(h), C.byref(lay), C.byref(nb), C.byref(ptr))
        return w.value, h.value, lay.value, nb.value, ptr.value

    def op(self, op: PpmxOp, img, hist: bool = False):
        out = C.c_void_p()
        bins = (C.c_uint64 * 256)()
        if self.L.ppmx_gpu_op(self.ctx, C.byref(op), img, C.byref(out), bins if hist else None) != 0:
            raise PpmxError("ppmx_gpu_op failed (kind %d)" % op.kind)
        return (out, np.array(bins[:], np.uint64)) if hist else out

    def download(self, img, file_type: int) -> np.ndarray:
        w, h, lay, nb, _ = self.info(img)
        out = np.empty(max(w * h * 3, 1) + 16, np.uint8)
        n = C.c_size_t()
        if self.L.ppmx_gpu_download(self.ctx, img, file_type, _vp(out), out.size, C.byref(n)) != 0:
            raise PpmxError("ppmx_gpu_download failed")
        return out[:n.value].copy()

    def release(self, img):
        self.L.ppmx_gpu_image_free(self.ctx, img)

    def _one(self, img: np.ndarray, op: PpmxOp, file_type: int, hist: bool = False):
        img = _img(img)
        d = self.upload(img)
        try:
            res = self.op(op, d, hist)
            out, bins = (res if hist else (res, None))
            try:
                if out:
                    w, h, lay, nb, _ = self.info(out)
                    data = self.download(out, file_type)
                else:
                    w = h = 0
                    lay, data = -1, None
            finally:
                if out:
                    self.release(out)
        finally:
            self.release(d)
        return data, w, h, lay, bins

    # -- reference operators (same names as the oracle) --------------------------------------
    def gray(self, img) -> np.ndarray:
        data, w, h, _, _ = self._one(img, PpmxOp(kind=OP_GRAY), FT_PGM)
        return data.reshape(h, w)

    def mono(self, img) -> np.ndarray:
        """(h, w) plane of 0/1 (the .r members)."""
        data, w, h, _, _ = self._one(img, PpmxOp(kind=OP_MONO), FT_PGM)
        return data.reshape(h, w)

    def mono_bits(self, img) -> np.ndarray:
        """mono fused with the P4 packer: the bytes of a P4 raster."""
        data, _, _, _, _ = self._one(img, PpmxOp(kind=OP_MONO_BITS), FT_PBM)
        return data

    def pack_pbm(self, plane) -> np.ndarray:
        plane = np.ascontiguousarray(plane, np.uint8)
        d = self.upload(plane, LAYOUT_R8)
        try:
            return self.download(d, FT_PBM)
        finally:
            self.release(d)

    def flip(self, img, direction: int) -> np.ndarray:
        data, w, h, _, _ = self._one(img, PpmxOp(kind=OP_FLIP, flip_direction=int(direction)), FT_PPM)
        return data.reshape(h, w, 3)

    def rotate_op(self, angle: int, w: int, h: int) -> PpmxOp:
        import math
        nw, nh = rotate_size(angle, w, h)
        th = (float(angle) * 3.14159265358979323846) / 180.0  # ref:692, host libm for cos/sin
        return PpmxOp(kind=OP_ROTATE, angle_deg=int(angle), cos_t=math.cos(th), sin_t=math.sin(th), new_width=nw,
                      new_height=nh)

    def rotate(self, img, angle: int) -> np.ndarray:
        img = _img(img)
        op = self.rotate_op(angle, img.shape[1], img.shape[0])
        data, w, h, _, _ = self._one(img, op, FT_PPM)
        return data.reshape(h, w, 3)

    @staticmethod
    def calc_contributions(in_size, out_size, scale, k_width=4.0):
        return calc_contributions(in_size, out_size, scale, k_width)

    @staticmethod
    def imresize_op(out_size: int, dim: int, weights: np.ndarray, indices: np.ndarray) -> PpmxOp:
        return PpmxOp(kind=OP_IMRESIZE, dim=dim, out_size=out_size, weights_sz=weights.shape[1],
                      weights=weights.ctypes.data_as(_dblp), indices=indices.ctypes.data_as(_i32p))

    def imresize(self, img, out_size: int, dim: int, weights, indices) -> np.ndarray:
        weights = np.ascontiguousarray(weights, np.float64)
        indices = np.ascontiguousarray(indices, np.int32)
        data, w, h, _, _ = self._one(img, self.imresize_op(out_size, dim, weights, indices), FT_PPM)
        return data.reshape(h, w, 3)

    def process(self, img, resize_w: Optional[int] = None, angle: Optional[int] = None, gray=False, mono=False,
                flipv=False, fliph=False, conv_preset: int = 0, levels=None):
        """The whole chain through ppmx_plan_chain + ppmx_gpu_apply (host raster in, writer bytes out)."""
        img = _img(img)
        h, w, _ = img.shape
        ph = _PlanHolder(resize_w, angle, gray, mono, flipv, fliph, w, h, conv_preset, levels)
        try:
            if ph.plan.nops == 0:
                raise PpmxError("Error: no data to write")
            ow, oh = w, h
            for i in range(ph.plan.nops):
                o = ph.plan.ops[i]
                if o.kind in (OP_IMRESIZE, OP_ROTATE):
                    a, b, c = C.c_uint32(), C.c_uint32(), C.c_int()
                    self.L.ppmx_gpu_op_output(C.byref(o), ow, oh, LAYOUT_RGB8, C.byref(a), C.byref(b), C.byref(c))
                    ow, oh = a.value, b.value
            out = np.empty(ow * oh * 3 + 16, np.uint8)
            n = C.c_size_t()
            rw, rh, ft = C.c_uint32(), C.c_uint32(), C.c_int()
            rc = self.L.ppmx_gpu_apply(self.ctx, ph.plan.ops, ph.plan.nops, _vp(img), w, h, _vp(out), out.size,
                                       C.byref(n), C.byref(rw), C.byref(rh), C.byref(ft))
            if rc != 0:
                raise PpmxError("ppmx_gpu_apply failed")
            return out[:n.value].copy(), rw.value, rh.value, ft.value
        finally:
            ph.close()

    def apply_ops(self, imgs, ops, out_bytes_each: Optional[int] = None):
        """ppmx_gpu_apply_batch on a stack of equally sized rasters (n, h, w, 3) with an explicit op list.
        Returns (outputs (n, bytes_each) uint8, out_w, out_h, file_type)."""
        imgs = np.ascontiguousarray(imgs, np.uint8)
        if imgs.ndim == 3:
            imgs = imgs[None]
        n, h, w, _ = imgs.shape
        arr = (PpmxOp * len(ops))(*ops)
        cap = out_bytes_each or (max(w, h) ** 2 * 6 + 64)
        out = np.empty((n, cap), np.uint8)
        each, rw, rh, ft = C.c_size_t(), C.c_uint32(), C.c_uint32(), C.c_int()
        rc = self.L.ppmx_gpu_apply_batch(self.ctx, arr, len(ops), _vp(imgs), w, h, n, _vp(out), cap, C.byref(each),
                                         C.byref(rw), C.byref(rh), C.byref(ft))
        if rc != 0:
            raise PpmxError("ppmx_gpu_apply_batch failed")
        return out[:, :each.value].copy(), rw.value, rh.value, ft.value

    def apply_band(self, img, ops, band: int, nbands: int, out: Optional[np.ndarray] = None):
        """ppmx_gpu_apply_band: the rows of row band `band` of `nbands` of the chain's output, written into `out`
        (the WHOLE output, shared by all bands).  Returns (out, out_w, out_h, file_type, band_y0, band_rows)."""
        img = _img(img)
        h, w, _ = img.shape
        arr = (PpmxOp * len(ops))(*ops)
        ow, oh, ft, nb = self.chain_info(ops, w, h)[:4]
        if out is None:
            out = np.zeros(nb, np.uint8)
        n = C.c_size_t()
        rw, rh, rft, y0, rows = C.c_uint32(), C.c_uint32(), C.c_int(), C.c_uint32(), C.c_uint32()
        rc = self.L.ppmx_gpu_apply_band(self.ctx, arr, len(ops), _vp(img), w, h, band, nbands, _vp(out), out.size,
                                        C.byref(n), C.byref(rw), C.byref(rh), C.byref(rft), C.byref(y0), C.byref(rows))
        if rc != 0:
            raise PpmxError("ppmx_gpu_apply_band failed")
        return out, rw.value, rh.value, rft.value, y0.value, rows.value

    @staticmethod
    def chain_info(ops, w: int, h: int):
        return chain_info(ops, w, h)

    # -- a chain prepared for device-resident rasters ----------------------------------------------
    def chain_prepare(self, ops, w: int, h: int):
        arr = (PpmxOp * len(ops))(*ops)
        ch = C.c_void_p()
        if self.L.ppmx_gpu_chain_prepare(self.ctx, arr, len(ops), w, h, C.byref(ch)) != 0:
            raise PpmxError("ppmx_gpu_chain_prepare failed")
        ow, oh, ft, nb, k, moved = C.c_uint32(), C.c_uint32(), C.c_int(), C.c_size_t(), C.c_int(), C.c_size_t()
        self.L.ppmx_gpu_chain_info2(ch, C.byref(ow), C.byref(oh), C.byref(ft), C.byref(nb), C.byref(k), C.byref(moved))
        return ch, dict(out_w=ow.value, out_h=oh.value, file_type=ft.value, out_bytes=nb.value, kernels=k.value,
                        bytes_moved=moved.value)

    def chain_run(self, chain, d_src: int, d_dst: int, stream: int = 0) -> None:
        if self.L.ppmx_gpu_chain_run(chain, C.c_void_p(d_src), C.c_void_p(d_dst), C.c_void_p(stream)) != 0:
            raise PpmxError("ppmx_gpu_chain_run failed")

    def chain_free(self, chain) -> None:
        self.L.ppmx_gpu_chain_free(chain)

    # -- CUDA graphs over raw launches ---------------------------------------------------------
    def graph_begin(self, stream: int) -> None:
        if self.L.ppmx_gpu_graph_begin(C.c_void_p(stream)) != 0:
            raise PpmxError("ppmx_gpu_graph_begin failed")

    def graph_end(self, stream: int):
        g, n = C.c_void_p(), C.c_uint64()
        if self.L.ppmx_gpu_graph_end(C.c_void_p(stream), C.byref(g), C.byref(n)) != 0:
            raise PpmxError("ppmx_gpu_graph_end failed")
        return g, int(n.value)

    def graph_launch(self, graph, stream: int) -> None:
        if self.L.ppmx_gpu_graph_launch(graph, C.c_void_p(stream)) != 0:
            raise PpmxError("ppmx_gpu_graph_launch failed")

    def graph_free(self, graph) -> None:
        self.L.ppmx_gpu_graph_free(graph)

    @staticmethod
    def header(file_type: int, w: int, h: int, maxval: int = 255) -> bytes:
        return format_header(file_type, w, h, maxval)

    # -- extensions (no reference counterpart; parity unpinned) -------------------------------
    @staticmethod
    def conv_op(coef, div: int = 1, bias: int = 0) -> PpmxOp:
        coef = np.ascontiguousarray(coef, np.int32)
        op = PpmxOp(kind=OP_CONV, conv_k=coef.shape[0], conv_div=div, conv_bias=bias,
                    conv_coef=coef.ctypes.data_as(_i32p))
        op._keep = coef
        return op

    def conv(self, img, coef, div: int = 1, bias: int = 0) -> np.ndarray:
        data, w, h, _, _ = self._one(img, self.conv_op(coef, div, bias), FT_PPM)
        return data.reshape(h, w, 3)

    @staticmethod
    def levels_op(lut) -> PpmxOp:
        lut = np.ascontiguousarray(lut, np.uint8)
        assert lut.size == 256
        op = PpmxOp(kind=OP_LEVELS, levels_lut=lut.ctypes.data_as(C.POINTER(C.c_uint8)))
        op._keep = lut
        return op

    def levels(self, img, lut) -> np.ndarray:
        """Every byte through a 256-entry table (extension); RGB (h, w, 3) in, same shape out."""
        data, w, h, _, _ = self._one(img, self.levels_op(lut), FT_PPM)
        return data.reshape(h, w, 3)

    @staticmethod
    def levels_lut_linear(lo: int, hi: int) -> np.ndarray:
        lut = np.zeros(256, np.uint8)
        if host_lib().ppmx_levels_lut_linear(int(lo), int(hi), lut.ctypes.data_as(C.c_void_p)) != 0:
            raise PpmxError("ppmx_levels_lut_linear: need 0 <= lo < hi <= 255")
        return lut

    @staticmethod
    def levels_points_from_hist(bins, clip_permille: int = 5):
        bins = np.ascontiguousarray(bins, np.uint64)
        lo, hi = C.c_int(), C.c_int()
        if host_lib().ppmx_levels_points_from_hist(bins.ctypes.data_as(C.c_void_p), int(clip_permille), C.byref(lo),
                                                   C.byref(hi)) != 0:
            raise PpmxError("ppmx_levels_points_from_hist: empty or flat histogram")
        return lo.value, hi.value

    def hist_gray(self, img) -> np.ndarray:
        _, _, _, _, bins = self._one(img, PpmxOp(kind=OP_HIST_GRAY), FT_PGM, hist=True)
        return bins

    def gray_hist(self, img):
        data, w, h, _, bins = self._one(img, PpmxOp(kind=OP_GRAY_HIST), FT_PGM, hist=True)
        return data.reshape(h, w), bins

    # -- raw launch on caller-owned device memory (bench.py, multi-GPU bands) -----------------
    def launch(self, op: PpmxOp, d_src: int, w: int, h: int, layout: int, d_dst: int, band: Optional[PpmxBand] = None,
               d_hist: int = 0, d_tables: int = 0, stream: int = 0) -> None:
        rc = self.L.ppmx_gpu_launch(C.byref(op), C.c_void_p(d_src), w, h, layout, C.c_void_p(d_dst),
                                    C.byref(band) if band is not None else None, C.c_void_p(d_hist),
                                    C.c_void_p(d_tables), C.c_void_p(stream))
        if rc != 0:
            raise PpmxError("ppmx_gpu_launch failed (kind %d)" % op.kind)

    # -- exportable HBM + CUDA IPC (one process per GPU reads its neighbours' bands over NVLink) --
    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        if self.L.ppmx_gpu_device_alloc(self.ctx, nbytes, C.byref(p)) != 0:
            raise PpmxError("ppmx_gpu_device_alloc failed")
        return p.value

    def device_free(self, p: int) -> None:
        self.L.ppmx_gpu_device_free(self.ctx, C.c_void_p(p))

    def copy(self, dst: int, src: int, nbytes: int, kind: int) -> None:
        if self.L.ppmx_gpu_copy(self.ctx, C.c_void_p(dst), C.c_void_p(src), nbytes, kind) != 0:
            raise PpmxError("ppmx_gpu_copy failed")

    def ipc_export(self, p: int) -> bytes:
        h = (C.c_uint8 * 64)()
        if self.L.ppmx_gpu_ipc_export(self.ctx, C.c_void_p(p), h) != 0:
            raise PpmxError("ppmx_gpu_ipc_export failed")
        return bytes(h)

    def ipc_open(self, handle: bytes) -> int:
        h = (C.c_uint8 * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        if self.L.ppmx_gpu_ipc_open(self.ctx, h, C.byref(p)) != 0:
            raise PpmxError("ppmx_gpu_ipc_open failed")
        return p.value

    def ipc_close(self, p: int) -> None:
        self.L.ppmx_gpu_ipc_close(self.ctx, C.c_void_p(p))

    def tables_upload(self, op: PpmxOp) -> int:
        p = C.c_void_p()
        if self.L.ppmx_gpu_tables_upload(C.byref(op), C.byref(p)) != 0:
            raise PpmxError("ppmx_gpu_tables_upload failed")
        return p.value

    def tables_free(self, p: int) -> None:
        self.L.ppmx_gpu_tables_free(C.c_void_p(p))

    def set_tuning(self, key: str, value: int) -> None:
        if self.L.ppmx_gpu_set_tuning(key.encode(), int(value)) != 0:
            raise PpmxError("unknown tuning key " + key)

    def launch_count(self) -> int:
        return int(self.L.ppmx_gpu_launch_count())
