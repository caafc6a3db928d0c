#!/usr/bin/env python
"""bench.py -- megapixels/s of the per-pixel hot path of ppmx-edward.c on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One JSON line on stdout (rank 0).  A "step" is one pass of the operator over a batch of
BATCH distinct synthetic rasters (the batch is larger than the 126 MB L2, so every step reads
its input from HBM).  `value` is device-resident throughput (inputs already in HBM, CUDA events
on the launching stream, max over ranks); `e2e` is the same operator through the C ABI call
ppmx_gpu_apply_batch with pinned HOST buffers, H2D and D2H copies inside the timed region.
`roofline` is for the operator's kernel: algorithmic bytes / measured duration
against MEASURED_PEAKS.json.  `cpu_baseline` is the compiled reference (oracle/_ref) timed on
this box's host cores on a bounded sample.  Multi-GPU: one process per GPU under torchrun,
rasters sharded across ranks with no data-path collective ("weak"); the fused gray+histogram
workload adds the one real exchange, an NCCL all-reduce of the 256 bins.

--impl reference times the reference's own CPU implementation (oracle/_ref) of the same
operator on the host cores (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "megapixels/sec per op"
UNIT = "Mpix/s"

# workload -> (w, h, batch, algorithmic bytes per input pixel, description)
WORKLOADS = {
    "gray": (4096, 4096, 32, 4.0, "4096x4096 P6 RGB->greyscale, 32 distinct rasters per step"),
    "gray_hist": (4096, 4096, 32, 4.0,
                  "4096x4096 P6 RGB->greyscale + histogram in one pass (config 2), 32 distinct rasters per step"),
    "mono": (4096, 4096, 32, 3.125, "4096x4096 P6 -> Bayer bilevel P4 bits, 32 rasters per step"),
    "fliph": (4096, 4096, 8, 6.0, "4096x4096 horizontal flip, 8 rasters per step"),
    "flipv": (4096, 4096, 8, 6.0, "4096x4096 vertical flip, 8 rasters per step"),
    "rot90": (4096, 4096, 8, 6.0, "4096x4096 rotate 90, 8 rasters per step"),
    "rot180": (4096, 4096, 8, 6.0, "4096x4096 rotate 180, 8 rasters per step"),
    "gray_16k": (16384, 16384, 2, 4.0, "16384x16384 RGB->greyscale, two rasters (805 MB each) per step"),
    "mono_16k": (16384, 16384, 2, 3.125, "16384x16384 -> Bayer bilevel P4 bits, two rasters per step"),
    "fliph_16k": (16384, 16384, 2, 6.0, "16384x16384 horizontal flip, two rasters per step"),
    "rot90_16k": (16384, 16384, 2, 6.0, "16384x16384 rotate 90, two rasters per step"),
    "mono_4090": (4090, 4090, 8, 3.125, "4090x4090 -> Bayer bilevel bits: sides no multiple of 16 (generic kernel), 8 rasters per step"),
    "fliph_4090": (4090, 4090, 8, 6.0, "4090x4090 horizontal flip: sides no multiple of 16 (generic kernel), 8 rasters per step"),
    "flipv_4090": (4090, 4090, 8, 6.0, "4090x4090 vertical flip: sides no multiple of 16 (generic kernel), 8 rasters per step"),
    "rot90_4090": (4090, 4090, 8, 6.0, "4090x4090 rotate 90: sides no multiple of 16 (generic kernel), 8 rasters per step"),
    "rot180_4090": (4090, 4090, 8, 6.0, "4090x4090 rotate 180: sides no multiple of 16 (generic kernel), 8 rasters per step"),
    "gray_4090": (4090, 4090, 8, 4.0, "4090x4090 RGB->greyscale: sides no multiple of 16, 8 rasters per step"),
    "levels": (4096, 4096, 8, 6.0, "4096x4096 levels (256-entry table on every byte, extension), 8 rasters per step"),
    "conv3": (8192, 8192, 2, 6.0, "8192x8192 3x3 blur (extension, config 3), 2 rasters per step"),
    "conv7": (8192, 8192, 2, 6.0, "8192x8192 7x7 box (extension, config 3), 2 rasters per step"),
    "conv3_odd": (4090, 4090, 4, 6.0, "4090x4090 3x3 blur (extension): a width that is no multiple of 16 goes through a padded copy, 4 rasters per step"),
    "gauss7": (8192, 8192, 2, 6.0, "8192x8192 7x7 binomial blur (1 6 15 20 15 6 1)^2 / 4096 (extension), 2 rasters per step"),
    "gauss5": (8192, 8192, 2, 6.0, "8192x8192 5x5 binomial blur (1 4 6 4 1)^2 / 256 (extension), 2 rasters per step"),
    "sharpen3": (8192, 8192, 2, 6.0, "8192x8192 3x3 sharpen (0 -1 0; -1 5 -1; 0 -1 0) (extension, config 3), 2 rasters per step"),
    "edge3": (8192, 8192, 2, 6.0, "8192x8192 3x3 edge (-1 .. 8 .. -1) (extension, config 3), 2 rasters per step"),
    "resize_up": (4096, 4096, 2, None, "4096x4096 -w6144 bicubic resize (FP64), 2 rasters per step"),
    "resize_down": (4096, 4096, 2, None, "4096x4096 -w2048 bicubic resize (FP64), 2 rasters per step"),
    "rot30": (4096, 4096, 2, None, "4096x4096 -r30 bicubic rotate (FP64), 2 rasters per step"),
}
DEFAULT_WORKLOAD = "gray_hist"
PER_OP = ["gray", "gray_hist", "mono", "fliph", "flipv", "rot90", "rot180", "gray_16k", "mono_16k", "fliph_16k",
          "rot90_16k", "levels", "conv3", "conv7", "sharpen3", "edge3", "gauss5", "gauss7", "conv3_odd", "resize_up", "resize_down", "rot30"]


def traffic_for(name):
    """DRAM bytes per launch from the committed ncu --set full capture of this kernel, or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p)).get(name)
        return int(t["dram_read"] + t["dram_write"]) if t else None
    except Exception:
        return None


def dp_peak():
    """Measured non-fused FP64 instruction rate (tools/dp_peak.cu), the roofline of the bicubic operators."""
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "dp_peak.json")))["dmul_dadd_inst_per_s"])
    except Exception:
        return 1.84e13


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock and throttle reasons while the GPU works (pynvml).  Samples taken between
    mark_timed(True) and mark_timed(False) belong to the timed region; if that region is too short to
    catch three samples the summary falls back to every sample taken under load and says so."""

    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
             0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.index, self.samples, self.max_mhz = index, [], None
        self.timed = False
        self._stop = threading.Event()
        self._t = None

    def _loop(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self._stop.is_set():
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((self.timed, mhz, r))
                time.sleep(0.002)
        except Exception:
            pass

    def mark_timed(self, on: bool):
        self.timed = on

    def __enter__(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=2)

    def summary(self):
        timed = [x for x in self.samples if x[0]]
        use, window = (timed, "timed region") if len(timed) >= 3 else (self.samples, "whole run under load")
        mhz = sorted(x[1] for x in use)
        reasons = set()
        for _, _, r in use:
            for bit, n in self.NAMES.items():
                if r & bit:
                    reasons.add(n)
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
                "samples": len(use), "window": window}


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------

class Runner:
    """Builds the op + device buffers of one workload and launches it through ppmx_gpu_launch."""

    def __init__(self, torch, g, name, device, seed):
        import numpy as np
        from imageprocessingtools_b200 import ppmx as pp
        self.torch, self.g, self.name, self.pp = torch, g, name, pp
        w, h, batch, bpp, desc = WORKLOADS[name]
        self.w, self.h, self.batch, self.bpp, self.desc = w, h, batch, bpp, desc
        gen = torch.Generator(device=device)
        gen.manual_seed(seed)
        self.src = torch.randint(0, 256, (batch, h, w, 3), dtype=torch.uint8, device=device, generator=gen)
        self.tables = 0
        self.hist = None
        self.mid = None
        self.ops = []  # (op, out_w, out_h, out_bytes_per_raster, src_layout)
        self.keep = []
        L = pp
        base = name.split("_16k")[0]
        if name.endswith("_16k"):
            name = base
        if name.endswith("_4090"):  # same operator on a raster whose sides are no multiple of 16 (generic kernels)
            name = name[:-5]
        if name in ("gray", "gray_hist"):
            kind = L.OP_GRAY if name == "gray" else L.OP_GRAY_HIST
            self.ops = [(L.PpmxOp(kind=kind), w, h, w * h)]
            if name == "gray_hist":
                self.hist = torch.zeros(256, dtype=torch.int64, device=device)
        elif name == "mono":
            self.ops = [(L.PpmxOp(kind=L.OP_MONO_BITS), w, h, ((w + 7) // 8) * h)]
        elif name in ("fliph", "flipv"):
            self.ops = [(L.PpmxOp(kind=L.OP_FLIP, flip_direction=int(name == "flipv")), w, h, w * h * 3)]
        elif name in ("rot90", "rot180", "rot30"):
            op = g.rotate_op(int(name[3:]), w, h)
            self.ops = [(op, op.new_width, op.new_height, op.new_width * op.new_height * 3)]
        elif name == "levels":
            self.ops = [(g.levels_op(g.levels_lut_linear(16, 235)), w, h, w * h * 3)]
        elif name in ("conv3", "conv7", "sharpen3", "edge3", "gauss5", "gauss7", "conv3_odd"):
            if name in ("conv3", "conv3_odd"):
                coef, div = np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], np.int32), 16
            elif name == "sharpen3":
                coef, div = np.array([[0, -1, 0], [-1, 5, -1], [0, -1, 0]], np.int32), 1
            elif name == "gauss7":
                coef, div = np.outer([1, 6, 15, 20, 15, 6, 1], [1, 6, 15, 20, 15, 6, 1]).astype(np.int32), 4096
            elif name == "gauss5":
                coef, div = np.outer([1, 4, 6, 4, 1], [1, 4, 6, 4, 1]).astype(np.int32), 256
            elif name == "edge3":
                coef, div = np.array([[-1, -1, -1], [-1, 8, -1], [-1, -1, -1]], np.int32), 1
            else:
                coef, div = np.ones((7, 7), np.int32), 49
            self.ops = [(g.conv_op(coef, div, 0), w, h, w * h * 3)]
        elif name in ("resize_up", "resize_down"):
            new_w = w * 3 // 2 if name == "resize_up" else w // 2
            ph = pp._PlanHolder(resize_w=new_w, w=w, h=h)
            self.keep.append(ph)
            cw, chh = w, h
            for i in range(ph.plan.nops):
                op = ph.plan.ops[i]
                ow, oh = (cw, op.out_size) if op.dim == 0 else (op.out_size, chh)
                self.ops.append((op, ow, oh, ow * oh * 3))
                cw, chh = ow, oh
        else:
            raise SystemExit("unknown workload " + name)
        self.out_w, self.out_h = self.ops[-1][1], self.ops[-1][2]
        self.dst = [torch.empty((batch, o[3]), dtype=torch.uint8, device=device) for o in self.ops]
        self.tabs = [g.tables_upload(o[0]) if o[0].kind == L.OP_IMRESIZE else 0 for o in self.ops]
        self.launches_per_step = batch * len(self.ops)
        self.pixels_per_step = batch * w * h  # input pixels (resize/rotate: also reported per output px)
        # The argument lists of every ppmx_gpu_launch of a step are built ONCE: a 4096^2 launch lasts ~12 us on the
        # device, and slicing tensors + boxing ctypes values per call costs about as much on a slow host core, which
        # would make the step a measurement of Python, not of the kernels.
        import ctypes as C
        hist_ptr = C.c_void_p(self.hist.data_ptr() if self.hist is not None else 0)
        self._calls = []
        for b in range(batch):
            src, cw, chh = self.src[b].data_ptr(), w, h
            for i, (op, ow, oh, nbytes) in enumerate(self.ops):
                dst = self.dst[i][b].data_ptr()
                self._calls.append((C.byref(op), C.c_void_p(src), cw, chh, L.LAYOUT_RGB8, C.c_void_p(dst), None, hist_ptr,
                                    C.c_void_p(self.tabs[i])))
                src, cw, chh = dst, ow, oh
        self._launch = g.L.ppmx_gpu_launch
        self._stream = (None, None)

    def step(self, stream):
        if self._stream[0] != stream:
            import ctypes as C
            self._stream = (stream, C.c_void_p(stream))
        st, fn = self._stream[1], self._launch
        for a in self._calls:
            if fn(*a, st) != 0:
                raise SystemExit("ppmx_gpu_launch failed")

    def close(self):
        for t in self.tabs:
            if t:
                self.g.tables_free(t)
        for k in self.keep:
            k.close()


def time_steps(torch, runner, steps, warmup, dist=None, after_step=None, clk=None, before_end=None):
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(warmup):
        runner.step(stream)
        if after_step:
            after_step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if clk is not None:
        clk.mark_timed(True)
    e0.record()
    for _ in range(steps):
        runner.step(stream)
        if after_step:
            after_step()
    if before_end:
        before_end()  # e.g. make the stream wait for collectives still in flight: they belong to the region
    e1.record()
    torch.cuda.synchronize()
    if clk is not None:
        clk.mark_timed(False)
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def e2e_run(torch, g, name, steps, warmup, device, dist=None):
    """Same operator through ppmx_gpu_apply_batch: pinned host rasters in, pinned host results out."""
    import ctypes as C
    from imageprocessingtools_b200 import ppmx as pp
    w, h, batch, _, _ = WORKLOADS[name]
    batch = min(batch, 8)
    if name == "gray":
        ops = (pp.PpmxOp * 1)(pp.PpmxOp(kind=pp.OP_GRAY))
        out_each = w * h
    elif name == "gray_hist":
        hist_host = torch.zeros((batch, 256), dtype=torch.int64).pin_memory()
        ops = (pp.PpmxOp * 1)(pp.PpmxOp(kind=pp.OP_GRAY_HIST,
                                        hist_out=C.cast(C.c_void_p(hist_host.data_ptr()), C.POINTER(C.c_uint64))))
        out_each = w * h
    elif name == "mono":
        ops = (pp.PpmxOp * 1)(pp.PpmxOp(kind=pp.OP_MONO))
        out_each = ((w + 7) // 8) * h
    elif name in ("fliph", "flipv"):
        ops = (pp.PpmxOp * 1)(pp.PpmxOp(kind=pp.OP_FLIP, flip_direction=int(name == "flipv")))
        out_each = w * h * 3
    elif name in ("rot90", "rot180", "rot30"):
        op = g.rotate_op(int(name[3:]), w, h)
        ops = (pp.PpmxOp * 1)(op)
        out_each = op.new_width * op.new_height * 3
    else:
        return None
    src = torch.randint(0, 256, (batch, h, w, 3), dtype=torch.uint8).pin_memory()
    dst = torch.empty((batch, out_each + 16), dtype=torch.uint8).pin_memory()
    each, ow, oh, ft = C.c_size_t(), C.c_uint32(), C.c_uint32(), C.c_int()

    def one():
        rc = g.L.ppmx_gpu_apply_batch(g.ctx, ops, 1, C.c_void_p(src.data_ptr()), w, h, batch, C.c_void_p(dst.data_ptr()),
                                      out_each + 16, C.byref(each), C.byref(ow), C.byref(oh), C.byref(ft))
        if rc != 0:
            raise SystemExit("ppmx_gpu_apply_batch failed")

    for _ in range(max(1, warmup)):
        one()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()  # returns after the last D2H copy has landed (the call synchronises its streams)
    dt = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    world = dist.get_world_size() if dist is not None else 1
    return {"value": round(world * steps * batch * w * h / dt / 1e6, 1), "unit": UNIT,
            "h2d_bytes_per_step": batch * w * h * 3, "d2h_bytes_per_step": batch * int(each.value),
            "rasters_per_step": batch, "api": "ppmx_gpu_apply_batch (pinned host in/out)"}


def cpu_baseline_sample(name, threads, seconds_budget=12.0):
    """The compiled reference (oracle/_ref) on this box's host cores; bounded sample."""
    import numpy as np
    import oracle
    ref = oracle.ref()
    kind = "reference"
    if ref is None:
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": "oracle/_ref missing"}
    w, h, _, _, _ = WORKLOADS[name]
    rng = np.random.default_rng(1)

    def call(r, img):
        if name in ("gray", "gray_hist"):
            r.gray(img)
        elif name == "mono":
            r.mono(img)
        elif name in ("fliph", "flipv"):
            r.flip(img, int(name == "flipv"))
        elif name in ("rot90", "rot180", "rot30"):
            r.rotate(img, int(name[3:]))
        else:
            r.gray(img)
        return r.last_seconds

    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(threads)]
    refs = [oracle.Ref() for _ in range(threads)]
    op_time = [0.0] * threads
    count = [0] * threads
    t_end = time.perf_counter() + seconds_budget

    def work(i):
        while True:
            op_time[i] += call(refs[i], imgs[i])
            count[i] += 1
            if time.perf_counter() > t_end or count[i] >= 64:
                break

    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    total_px = sum(count) * w * h
    fn = {"gray_hist": "gray", "fliph": "flip", "flipv": "flip", "rot90": "rotate", "rot180": "rotate",
          "rot30": "rotate"}.get(name, name)
    return {"value": round(total_px / max(op_time) / 1e6, 1), "unit": UNIT, "cores": threads, "kind": kind,
            "sample": "%d calls of the reference's %s() on %dx%d rasters over %d thread(s); time inside the "
                      "reference function only (its own image_buff_alloc included)" % (sum(count), fn, w, h, threads)}


def band_split_ref_ops(torch, g, dist, rank, world, device, full=16384, iters=10):
    """The same 16384 x 16384 raster cut into row bands for three REFERENCE operators (bit-exact ones):
    gray and mono need no exchange (mono takes the band's y0 for the Bayer phase); the resize height pass
    (x1.5, K = 4) reads the few source rows beyond its band from the neighbours' HBM over NVLink."""
    import numpy as np
    from imageprocessingtools_b200 import ppmx as pp
    w = full
    y0, rows = pp.band_plan(full, world, rank, 4)
    nbytes = rows * w * 3
    gen = torch.Generator(device=device)
    gen.manual_seed(0x5EED ^ rank)
    srcs, handles = [], []
    for i in range(2):
        p = g.device_alloc(nbytes)
        t = torch.randint(0, 256, (nbytes,), dtype=torch.uint8, device=device, generator=gen)
        g.copy(p, t.data_ptr(), nbytes, 2)
        del t
        srcs.append(p)
        handles.append(g.ipc_export(p))
    new_h = full * 3 // 2
    wt, ix = pp.calc_contributions(full, new_h, 1.5)
    oy0, orows = pp.band_plan(new_h, world, rank, 1)
    need = ix[oy0:oy0 + orows]
    halo = int(max(0, y0 - need.min(), need.max() - (y0 + rows - 1)))
    rop = g.imresize_op(new_h, 0, wt, ix)
    tables = g.tables_upload(rop)
    dst = g.device_alloc(max(orows * w * 3, nbytes))
    peers = {}
    bands = [pp.PpmxBand(full_h=full, y0=y0, halo=halo, out_y0=oy0, out_rows=orows) for _ in range(2)]
    if world > 1:
        info = [None] * world
        dist.all_gather_object(info, (rows, handles))
        for nb in (rank - 1, rank + 1):
            if 0 <= nb < world:
                peers[nb] = [g.ipc_open(hd) for hd in info[nb][1]]
        for i in range(2):
            if rank > 0:
                bands[i].d_top = peers[rank - 1][i] + (info[rank - 1][0] - halo) * w * 3
            if rank < world - 1:
                bands[i].d_bottom = peers[rank + 1][i]
        dist.barrier()
    stream = torch.cuda.current_stream().cuda_stream
    plain = pp.PpmxBand(full_h=full, y0=y0)
    results = []
    for label, op, bnd, tab, bpp in (("gray", pp.PpmxOp(kind=pp.OP_GRAY), [plain, plain], 0, 4.0),
                                     ("mono", pp.PpmxOp(kind=pp.OP_MONO_BITS), [plain, plain], 0, 3.125),
                                     ("imresize height pass x1.5", rop, bands, tables, None)):
        def step(i):
            g.launch(op, srcs[i & 1], w, rows, pp.LAYOUT_RGB8, dst, bnd[i & 1], 0, tab, stream)
        for i in range(4):
            step(i)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            dist.barrier()
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        mp = iters * full * full / (ms / 1e3) / 1e6
        ent = {"workload": "%dx%d raster, %s (reference operator), row bands over %d GPU(s)" % (full, full, label, world),
               "scaling": "strong", "mpix_s": round(mp, 1), "ms_per_raster": round(ms / iters, 4), "rows_per_gpu": rows}
        if bpp:
            ent["gbs_total"] = round(bpp * mp / 1e3, 1)
        else:
            ent.update({"halo_rows": halo, "out_mpix_s": round(mp * 1.5, 1)})
        results.append(ent)
    # flipv / rot180: output band r is input band N-1-r upside down -- read it straight
    # from that rank's HBM over NVLink (one launch, no copy); with one rank it is a local flip
    mirror = world - 1 - rank
    m_rows = rows
    m_ptr = None
    if world > 1:
        m_rows = info[mirror][0]
        if mirror == rank:
            m_ptr = srcs
        elif mirror in peers:
            m_ptr = peers[mirror]
        else:
            m_ptr = [g.ipc_open(hd) for hd in info[mirror][1]]
            peers[mirror] = m_ptr
    else:
        m_ptr = srcs
    for label, op in (("flipv", pp.PpmxOp(kind=pp.OP_FLIP, flip_direction=1)), ("rot180", g.rotate_op(180, w, m_rows))):
        def step(i):
            g.launch(op, m_ptr[i & 1], w, m_rows, pp.LAYOUT_RGB8, dst, None, 0, 0, stream)
        if m_rows * w * 3 > max(orows * w * 3, nbytes):
            continue
        for i in range(3):
            step(i)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            dist.barrier()
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        mp = iters * full * full / (ms / 1e3) / 1e6
        results.append({"workload": "%dx%d raster, %s (reference operator), row bands over %d GPU(s); a rank reads the "
                                    "mirrored band from its owner's HBM over NVLink" % (full, full, label, world),
                        "scaling": "strong", "mpix_s": round(mp, 1), "ms_per_raster": round(ms / iters, 4),
                        "peer_read_gbs_per_gpu": round(3.0 * mp / 1e3 / world, 1) if world > 1 else None})
    for lst in peers.values():
        for p in lst:
            g.ipc_close(p)
    if dist is not None:
        dist.barrier()
    g.tables_free(tables)
    for p in srcs + [dst]:
        g.device_free(p)
    return results


def band_split_run(torch, g, dist, rank, world, device, k=3, full=16384, iters=10):
    """BASELINE config 4: ONE full x full raster cut into row bands (ppmx_band_plan), one band per rank, k x k
    convolution (extension op) with the r = k/2 halo rows read straight from the neighbours' HBM over NVLink
    (CUDA IPC peer pointers, no copy, no collective).  Total work is fixed: strong scaling."""
    import numpy as np
    from imageprocessingtools_b200 import ppmx as pp
    w = full
    r = k // 2
    y0, rows = pp.band_plan(full, world, rank, 1)
    nbytes = rows * w * 3
    gen = torch.Generator(device=device)
    gen.manual_seed(0xBA2D ^ rank)
    srcs, handles = [], []
    for i in range(2):  # two distinct rasters alternate so that a band never stays resident in the 126 MB L2
        p = g.device_alloc(nbytes)
        t = torch.randint(0, 256, (nbytes,), dtype=torch.uint8, device=device, generator=gen)
        g.copy(p, t.data_ptr(), nbytes, 2)
        del t
        srcs.append(p)
        handles.append(g.ipc_export(p))
    dst = g.device_alloc(nbytes)
    peers = {}
    bands = []
    if world > 1:
        info = [None] * world
        dist.all_gather_object(info, (rows, handles))
        for nb in (rank - 1, rank + 1):
            if 0 <= nb < world:
                peers[nb] = [g.ipc_open(hd) for hd in info[nb][1]]
        for i in range(2):
            b = pp.PpmxBand(full_h=full, y0=y0, halo=r)
            if rank > 0:
                b.d_top = peers[rank - 1][i] + (info[rank - 1][0] - r) * w * 3
            if rank < world - 1:
                b.d_bottom = peers[rank + 1][i]
            bands.append(b)
        dist.barrier()
    else:
        bands = [pp.PpmxBand(full_h=full, y0=0, halo=r)] * 2
    coef = np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], np.int32) if k == 3 else np.ones((k, k), np.int32)
    op = g.conv_op(coef, 16 if k == 3 else k * k, 0)
    stream = torch.cuda.current_stream().cuda_stream

    def step(i):
        g.launch(op, srcs[i & 1], w, rows, pp.LAYOUT_RGB8, dst, bands[i & 1], 0, 0, stream)

    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        dist.barrier()
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    for lst in peers.values():
        for p in lst:
            g.ipc_close(p)
    if dist is not None:
        dist.barrier()  # nobody frees a band a neighbour may still have mapped
    for p in srcs + [dst]:
        g.device_free(p)
    mp = iters * full * full / (ms / 1e3) / 1e6
    return {"workload": "%dx%d raster, %dx%d convolution (extension), row bands over %d GPU(s), halo rows read from "
                        "peer HBM over NVLink" % (full, full, k, k, world),
            "scaling": "strong", "mpix_s": round(mp, 1), "ms_per_raster": round(ms / iters, 4),
            "gbs_total": round(6.0 * mp / 1e3, 1), "rows_per_gpu": rows, "halo_rows": r}


def batch_chain_run(torch, g, dist, world, images=128):
    """BASELINE config 5: a batch of 1920x1080 PPM rasters through full op chains, image-parallel (every rank
    takes `images` rasters), end to end through ppmx_gpu_apply_batch with pinned host buffers."""
    import ctypes as C
    from imageprocessingtools_b200 import ppmx as pp
    w, h = 1920, 1080
    src = torch.randint(0, 256, (images, h, w, 3), dtype=torch.uint8).pin_memory()
    out = []
    for label, kw in (("-w960 -r90 -gray -fv", dict(resize_w=960, angle=90, gray=True, flipv=True)),
                      ("-r90 -mono -fh", dict(angle=90, mono=True, fliph=True))):
        ph = pp._PlanHolder(w=w, h=h, **kw)
        cap = w * h * 3 + 64
        dst = torch.empty((images, cap), dtype=torch.uint8).pin_memory()
        each, ow, oh, ft = C.c_size_t(), C.c_uint32(), C.c_uint32(), C.c_int()

        def one():
            rc = g.L.ppmx_gpu_apply_batch(g.ctx, ph.plan.ops, ph.plan.nops, C.c_void_p(src.data_ptr()), w, h, images,
                                          C.c_void_p(dst.data_ptr()), cap, C.byref(each), C.byref(ow), C.byref(oh),
                                          C.byref(ft))
            if rc != 0:
                raise SystemExit("ppmx_gpu_apply_batch failed")

        one()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        one()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        out.append({"chain": label, "images": images * world, "frame": "%dx%d" % (w, h),
                    "out": "%dx%d type %d" % (ow.value, oh.value, ft.value),
                    "e2e_mpix_s": round(world * images * w * h / dt / 1e6, 1),
                    "images_per_s": round(world * images / dt, 1)})
        ph.close()
        del dst
    return out


def cpu_reference_per_op(names):
    """One bounded, single-threaded call (two for the cheap ones) of the compiled reference per operator,
    on a 4096x4096 raster of the same kind; time inside the reference function only."""
    import numpy as np
    import oracle
    ref = oracle.ref()
    if ref is None:
        return {}
    img = np.random.default_rng(2).integers(0, 256, (4096, 4096, 3), dtype=np.uint8)
    out = {}
    for name in names:
        base = name.split("_16k")[0]
        try:
            if base in ("gray", "gray_hist"):
                ref.gray(img); t = ref.last_seconds
            elif base == "mono":
                ref.mono(img); t = ref.last_seconds
            elif base in ("fliph", "flipv"):
                ref.flip(img, int(base == "flipv")); t = ref.last_seconds
            elif base in ("rot90", "rot180", "rot30"):
                ref.rotate(img, int(base[3:])); t = ref.last_seconds
            elif base in ("resize_up", "resize_down"):
                new_w = 6144 if base == "resize_up" else 2048
                # both passes in the reference's order (ref:1098-1120): width then height for exact scales
                wt1, ix1 = ref.calc_contributions(4096, new_w, new_w / 4096.0)
                mid = ref.imresize(img, new_w, 1, wt1, ix1); t = ref.last_seconds
                mid = ref.imresize(mid, new_w, 0, wt1, ix1); t += ref.last_seconds
            else:
                continue  # extension operators have no reference implementation
            out[name] = round(4096 * 4096 / t / 1e6, 1)
        except Exception:
            pass
    return out


def run_ours(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=device)
        dist = dist_mod
    import imageprocessingtools_b200 as ip
    g = ip.Ppmx(local)  # raises if libppmx_gpu.so is missing or no B200 is visible: no fallback

    name = args.workload
    w, h, batch, bpp, desc = WORKLOADS[name]
    peak, peak_src = peaks()
    runner = Runner(torch, g, name, device, seed=0xC0FFEE ^ (rank + 2))

    after = None
    drain = None
    if name == "gray_hist" and dist is not None:
        # The one real exchange of this path is the job's FINAL histogram reduction (north star: "NCCL only for the
        # final histogram reduction"): every rank accumulates the 256 bins of all its rasters on the device (u64) and
        # one NCCL all-reduce, issued in stream order inside the timed region, sums them over the ranks.
        # PPMX_BENCH_ALLREDUCE_EVERY_STEP=1 reduces after every step instead (2500 reductions/s: at N=8 the step
        # grows from 0.39 to 0.45 ms, since the NCCL kernel has to find an SM among 148 resident 1024-thread CTAs
        # and the ranks then wait for each other every 0.4 ms).
        def reduce_bins():
            dist.all_reduce(runner.hist, op=dist.ReduceOp.SUM)
        if os.environ.get("PPMX_BENCH_ALLREDUCE_EVERY_STEP") == "1":
            after = reduce_bins
        else:
            drain = reduce_bins

    n0 = g.launch_count()
    with ClockSampler(local) as clk:
        ms = time_steps(torch, runner, args.steps, args.warmup, dist, after, clk, drain)
    launches = (g.launch_count() - n0) - args.warmup * runner.launches_per_step
    px = world * args.steps * runner.pixels_per_step
    value = px / (ms / 1e3) / 1e6

    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": desc, "op": name, "raster": "%dx%d" % (w, h), "rasters_per_step_per_gpu": batch,
                   "l2": "inputs larger than L2 (%d MB per step per GPU)" % (batch * w * h * 3 // 1000000),
                   "sharding": "rasters sharded across ranks, no data-path collective" +
                               ("; NCCL all-reduce of the 256 histogram bins per step" if after else
                                "; one NCCL all-reduce of the 256 accumulated histogram bins at the end of the timed region"
                                if drain else "")},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
    }
    if bpp is not None:
        per_launch_ms = ms / (args.steps * runner.launches_per_step)
        achieved = bpp * w * h / (per_launch_ms / 1e3) / 1e9
        line["roofline"] = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                            "frac": round(achieved / peak, 4), "traffic": traffic_for(name), "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": int(bpp * w * h),
                            "kernel": name, "algorithmic_bytes_per_pixel": bpp,
                            "avg_launch_us": round(per_launch_ms * 1e3, 2)}
    else:
        line["roofline"] = {"bound": "fp64 issue (not HBM, not tensor)", "achieved": None, "peak": None, "unit": "GB/s",
                            "frac": None, "traffic": None, "kernel": name}

    e2e = e2e_run(torch, g, name, max(2, args.steps // 4), 1, device, dist)
    if e2e is not None:
        line["e2e"] = e2e

    if not args.no_per_op and world == 1:
        per = {}
        for op_name in PER_OP:
            if op_name == name:
                per[op_name] = {"mpix_s": round(value, 1), "gbs": line["roofline"].get("achieved"),
                                "frac": line["roofline"].get("frac")}
                continue
            ow, oh, ob, obpp, _ = WORKLOADS[op_name]
            r = Runner(torch, g, op_name, device, seed=7)
            t = time_steps(torch, r, 3, 3)
            nst = 3
            if t / 3 < 10.0:  # short steps: time at least ~30 ms of them, as tools/sweep.py does
                nst = max(3, min(100, int(30.0 / max(t / 3, 1e-3))))
                t = time_steps(torch, r, nst, 1)
            mp = nst * r.pixels_per_step / (t / 1e3) / 1e6
            ent = {"mpix_s": round(mp, 1), "raster": "%dx%d" % (ow, oh)}
            if obpp is not None:
                gbs = obpp * mp * 1e6 / 1e9
                ent.update({"gbs": round(gbs, 1), "frac": round(gbs / peak, 4)})
            else:
                ent["out_mpix_s"] = round(nst * r.batch * r.out_w * r.out_h / (t / 1e3) / 1e6, 1)
                # algorithmic FP64 instructions: imresize 6*K per output pixel per pass;
                # bicubic rotate ~190 per output pixel that maps inside the source (~ w*h of them)
                if op_name.startswith("resize"):
                    dp = sum(6.0 * o[0].weights_sz * o[1] * o[2] for o in r.ops)
                else:
                    dp = 190.0 * ow * oh
                rate = nst * r.batch * dp / (t / 1e3)
                ent.update({"bound": "fp64 issue", "dp_inst_per_s": float("%.4g" % rate),
                            "frac_of_dp_peak": round(rate / dp_peak(), 4)})
            per[op_name] = ent
            r.close()
            del r
            torch.cuda.empty_cache()
        line["per_op"] = per

    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline_sample(name, 1)
        if "per_op" in line:  # the reference's own function, 1 thread, 4096x4096, beside every operator
            for k, v in cpu_reference_per_op(list(line["per_op"])).items():
                line["per_op"][k]["cpu_reference_mpix_s_1thread"] = v
    runner.close()
    del runner
    torch.cuda.empty_cache()
    if not args.no_band:
        line["band_split"] = band_split_ref_ops(torch, g, dist, rank, world, device) + [
            band_split_run(torch, g, dist, rank, world, device, k=3),
            band_split_run(torch, g, dist, rank, world, device, k=7)]
        line["batch_chain"] = batch_chain_run(torch, g, dist, world)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)
    g.close()


# --------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU code on the host cores
# --------------------------------------------------------------------------------------------

def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    w, h, batch, _, desc = WORKLOADS[name]
    threads = min(os.cpu_count() or 1, 64)
    t0 = time.perf_counter()
    # steps*: each "step" is a bounded sample; total CPU time is capped to stay within minutes
    budget = min(60.0, max(4.0, 2.0 * (args.steps + args.warmup)))
    cb = cpu_baseline_sample(name, threads, seconds_budget=budget)
    dt = time.perf_counter() - t0
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT,
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(dt * 1e3 / max(args.steps, 1), 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": desc, "op": name, "raster": "%dx%d" % (w, h)},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-per-op", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-band", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
