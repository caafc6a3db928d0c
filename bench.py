#!/usr/bin/env python
"""bench.py -- megapixels/s of the per-pixel hot path of ppmx-edward.c on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One JSON line on stdout (rank 0).

Default workload = BASELINE config 4 at every N (strong scaling): ONE 16384 x 16384 raster cut into N row bands
(ppmx_band_plan), one band per rank, 3x3 convolution; the k/2 halo rows of a band are read from the neighbour
ranks' HBM over NVLink by the kernel itself (CUDA IPC peer pointers, no copy, no collective).  A "step" is one
pass over a batch of RASTERS_PER_STEP such rasters (drawn in turn from NSRC distinct resident rasters, each 6x
the L2).  `value` = device-resident Mpix/s over all ranks: CUDA events on the launching stream, a barrier and a
synchronize on both sides, max over ranks.  The launches of a step are recorded once into a CUDA graph
(ppmx_gpu_graph_*) and replayed, so the host is off the critical path.  `e2e` = the same operator through the
reference-facing call ppmx_gpu_apply_band with pinned HOST rasters: every rank uploads its band plus halo rows
from the host raster, convolves, downloads its rows -- H2D and D2H inside the timed region.  `band_parity`:
before anything is timed every rank compares the bytes of its band -- device-resident launch with peer halos AND
the host call -- with the oracle on a 16384-wide slice, for every banded operator; a mismatch fails the run.

`roofline` is for the convolution kernel: algorithmic bytes (6 B/px) / measured launch duration
against MEASURED_PEAKS.json.  `cpu_baseline` / --impl reference: the CPU implementation on this box's host
cores.  The 3x3 convolution is an EXTENSION operator (the reference has none), so its CPU arm is
the repo's plain-C oracle port (kind "port"); the reference's own operators are timed from the compiled reference
(oracle/_ref) in `per_op` and with --workload gray|mono|... .
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

# ---- the headline workload (config 4) ----------------------------------------------------------
FULL = 16384              # raster side
RASTERS_PER_STEP = 512    # fixed whatever N is: strong scaling
NSRC = 4                  # distinct resident rasters a step cycles through
GRAPH_LAUNCHES = 64       # kernel nodes per recorded graph
BLUR3 = ([[1, 2, 1], [2, 4, 2], [1, 2, 1]], 16)

# ---- per-operator extras: workload -> (w, h, batch, algorithmic bytes per input pixel, description) ----
WORKLOADS = {
    "conv3_bands": (FULL, FULL, RASTERS_PER_STEP, 6.0,
                    "16384x16384 raster, 3x3 convolution (extension), row bands over the ranks, halo rows read from the "
                    "neighbours' HBM over NVLink (config 4)"),
    "gray": (4096, 4096, 32, 4.0, "4096x4096 P6 RGB->greyscale, 32 distinct rasters per step"),
    "gray_hist": (4096, 4096, 32, 4.0,
                  "4096x4096 P6 RGB->greyscale + histogram in one pass (config 2), 32 distinct rasters per step"),
    "gray_hist_const": (4096, 4096, 32, 4.0,
                        "4096x4096 CONSTANT image (every pixel one bin: worst-case histogram contention), gray+hist"),
    "mono": (4096, 4096, 32, 3.125, "4096x4096 P6 -> Bayer bilevel P4 bits, 32 rasters per step"),
    "fliph": (4096, 4096, 8, 6.0, "4096x4096 horizontal flip, 8 rasters per step"),
    "flipv": (4096, 4096, 8, 6.0, "4096x4096 vertical flip, 8 rasters per step"),
    "rot90": (4096, 4096, 8, 6.0, "4096x4096 rotate 90, 8 rasters per step"),
    "rot180": (4096, 4096, 8, 6.0, "4096x4096 rotate 180, 8 rasters per step"),
    "rot90_1080p": (1920, 1080, 64, 6.0, "1920x1080 rotate 90 (config 5 frame: height no multiple of 16), 64 frames per step"),
    "gray_16k": (16384, 16384, 2, 4.0, "16384x16384 RGB->greyscale, two rasters (805 MB each) per step"),
    "mono_16k": (16384, 16384, 2, 3.125, "16384x16384 -> Bayer bilevel P4 bits, two rasters per step"),
    "fliph_16k": (16384, 16384, 2, 6.0, "16384x16384 horizontal flip, two rasters per step"),
    "rot90_16k": (16384, 16384, 2, 6.0, "16384x16384 rotate 90, two rasters per step"),
    "mono_4090": (4090, 4090, 8, 3.125, "4090x4090 -> Bayer bilevel bits: sides no multiple of 16, 8 rasters per step"),
    "fliph_4090": (4090, 4090, 8, 6.0, "4090x4090 horizontal flip: sides no multiple of 16, 8 rasters per step"),
    "flipv_4090": (4090, 4090, 8, 6.0, "4090x4090 vertical flip: sides no multiple of 16, 8 rasters per step"),
    "rot90_4090": (4090, 4090, 8, 6.0, "4090x4090 rotate 90: sides no multiple of 16, 8 rasters per step"),
    "rot180_4090": (4090, 4090, 8, 6.0, "4090x4090 rotate 180: sides no multiple of 16, 8 rasters per step"),
    "gray_4090": (4090, 4090, 8, 4.0, "4090x4090 RGB->greyscale: sides no multiple of 16, 8 rasters per step"),
    "levels": (4096, 4096, 8, 6.0, "4096x4096 levels (256-entry table on every byte, extension), 8 rasters per step"),
    "conv3": (8192, 8192, 2, 6.0, "8192x8192 3x3 blur (extension, config 3), 2 rasters per step"),
    "conv7": (8192, 8192, 2, 6.0, "8192x8192 7x7 box (extension, config 3), 2 rasters per step"),
    "conv3_odd": (4090, 4090, 4, 6.0, "4090x4090 3x3 blur (extension): a width that is no multiple of 16, 4 rasters per step"),
    "gauss7": (8192, 8192, 2, 6.0, "8192x8192 7x7 binomial blur (1 6 15 20 15 6 1)^2 / 4096 (extension), 2 rasters per step"),
    "gauss5": (8192, 8192, 2, 6.0, "8192x8192 5x5 binomial blur (1 4 6 4 1)^2 / 256 (extension), 2 rasters per step"),
    "dense5": (8192, 8192, 2, 6.0, "8192x8192 dense (rank > 1) 5x5 kernel (extension), 2 rasters per step"),
    "dense7": (8192, 8192, 2, 6.0, "8192x8192 dense (rank > 1) 7x7 kernel (extension), 2 rasters per step"),
    "gauss9": (8192, 8192, 2, 6.0, "8192x8192 9x9 binomial blur (1 8 28 56 70 56 28 8 1)^2 / 65536 (extension), 2 rasters per step"),
    "dense9": (8192, 8192, 2, 6.0, "8192x8192 dense 9x9 kernel (floored binomial blur, extension), 2 rasters per step"),
    "sharpen7": (8192, 8192, 2, 6.0, "8192x8192 7x7 sharpen = unsharp mask 2 I - binomial 7x7 (extension, config 3), 2 rasters per step"),
    "sharpen3": (8192, 8192, 2, 6.0, "8192x8192 3x3 sharpen (0 -1 0; -1 5 -1; 0 -1 0) (extension, config 3), 2 rasters per step"),
    "edge3": (8192, 8192, 2, 6.0, "8192x8192 3x3 edge (-1 .. 8 .. -1) (extension, config 3), 2 rasters per step"),
    "resize_up": (4096, 4096, 2, None, "4096x4096 -w6144 bicubic resize (FP64), 2 rasters per step"),
    "resize_down": (4096, 4096, 2, None, "4096x4096 -w2048 bicubic resize (FP64), 2 rasters per step"),
    "rot30": (4096, 4096, 2, None, "4096x4096 -r30 bicubic rotate (FP64), 2 rasters per step"),
}
DEFAULT_WORKLOAD = "conv3_bands"
PER_OP = ["gray", "gray_hist", "gray_hist_const", "mono", "fliph", "flipv", "rot90", "rot180", "rot90_1080p", "gray_16k",
          "mono_16k", "fliph_16k", "rot90_16k", "levels", "conv3", "conv7", "sharpen3", "edge3", "gauss5", "gauss7", "sharpen7", "dense5",
          "dense7", "gauss9", "dense9", "gray_4090", "mono_4090", "fliph_4090", "flipv_4090", "rot90_4090", "rot180_4090", "conv3_odd",
          "resize_up", "resize_down", "rot30"]


def static_traffic(name):
    """DRAM bytes per launch of this kernel from a committed `ncu --set full` capture (profiles/traffic.json) --
    a constant read from a file, NOT measured in this run (bench numbers are never taken under a profiler)."""
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


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


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
                time.sleep(0.005)
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
# launching: a list of prebuilt ppmx_gpu_launch argument tuples, replayed directly or as a CUDA graph
# --------------------------------------------------------------------------------------------

class StepRunner:
    """`calls` = the ppmx_gpu_launch argument tuples (everything but the stream) of ONE recording; a step replays
    the recording `replays` times.  graph=True: the recording is captured once into a CUDA graph through the C ABI
    (ppmx_gpu_graph_begin/_end) and a replay is one ppmx_gpu_graph_launch; else every replay issues the launches
    one by one (the argument lists are prebuilt, so a launch costs one ctypes call).  lanes > 1: the launches of a
    recording go round-robin over that many streams forked from / joined into the timing stream (independent
    rasters: their ramps and tails overlap)."""

    def __init__(self, torch, g, calls, replays=1, graph=True, lanes=1, fn=None, kernels_per_call=1):
        import ctypes as C
        self.torch, self.g, self.calls, self.replays, self.lanes = torch, g, calls, replays, max(1, lanes)
        self.C = C
        self.fn = fn if fn is not None else g.L.ppmx_gpu_launch  # any ABI call whose LAST argument is the stream
        self.main = torch.cuda.current_stream()
        self.side = [torch.cuda.Stream() for _ in range(self.lanes - 1)]
        self.fork = torch.cuda.Event()
        self.join = [torch.cuda.Event() for _ in self.side]
        self.graph = None
        self.mode = "direct launches (prebuilt argument lists)"
        self.launches_per_step = len(calls) * replays * kernels_per_call
        if graph:
            try:
                g.graph_begin(self.main.cuda_stream)
                try:
                    self._issue()
                finally:
                    self.graph, n = g.graph_end(self.main.cuda_stream)
                if n < len(calls) * kernels_per_call:  # (an operator may launch a remainder kernel beside its main one)
                    raise RuntimeError("graph holds %d kernel nodes, expected %d" % (n, len(calls) * kernels_per_call))
                self.launches_per_step = n * replays
                self.mode = "CUDA graph of %d kernel nodes (ppmx_gpu_graph_*), %d replay(s) per step" % (n, replays)
            except Exception as e:  # stay measurable without graphs; say so in the line
                self.graph = None
                self.mode = "direct launches (graph capture failed: %s)" % (str(e)[:80],)
        if self.lanes > 1:
            self.mode += ", %d streams" % self.lanes

    def _issue(self):
        C, fn = self.C, self.fn
        streams = [self.main] + self.side
        ptrs = [C.c_void_p(s.cuda_stream) for s in streams]
        if self.side:
            self.fork.record(self.main)
            for s in self.side:
                s.wait_event(self.fork)
        n = len(streams)
        for i, a in enumerate(self.calls):
            if fn(*a, ptrs[i % n]) != 0:
                raise RuntimeError("launch failed")
        for s, e in zip(self.side, self.join):
            e.record(s)
            self.main.wait_event(e)

    def step(self):
        if self.graph is not None:
            st = self.main.cuda_stream
            for _ in range(self.replays):
                self.g.graph_launch(self.graph, st)
        else:
            for _ in range(self.replays):
                self._issue()

    def close(self):
        if self.graph is not None:
            self.g.graph_free(self.graph)
            self.graph = None


def time_runner(torch, runner, steps, warmup, dist=None, clk=None, before_end=None):
    """W untimed steps, then K steps between two CUDA events on the launching stream, barrier + synchronize on
    both sides; returns (max, min, median) of the ranks' milliseconds."""
    for _ in range(warmup):
        runner.step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if clk is not None:
        clk.mark_timed(True)
    e0.record()
    for _ in range(steps):
        runner.step()
    if before_end:
        before_end()
    e1.record()
    torch.cuda.synchronize()
    if clk is not None:
        clk.mark_timed(False)
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is None:
        return ms, ms, ms
    all_ms = [None] * dist.get_world_size()
    dist.all_gather_object(all_ms, ms)
    s = sorted(all_ms)
    return s[-1], s[0], s[len(s) // 2]


# --------------------------------------------------------------------------------------------
# the banded 16384^2 raster (config 4), device-resident: bands in exportable HBM, halos through IPC
# --------------------------------------------------------------------------------------------

def conv_spec(name):
    import numpy as np
    if name in ("conv3", "conv3_bands", "conv3_odd", "blur3"):
        return np.array(BLUR3[0], np.int32), BLUR3[1]
    if name == "sharpen3":
        return np.array([[0, -1, 0], [-1, 5, -1], [0, -1, 0]], np.int32), 1
    if name == "edge3":
        return np.array([[-1, -1, -1], [-1, 8, -1], [-1, -1, -1]], np.int32), 1
    if name == "gauss7":
        return np.outer([1, 6, 15, 20, 15, 6, 1], [1, 6, 15, 20, 15, 6, 1]).astype(np.int32), 4096
    if name == "gauss5":
        return np.outer([1, 4, 6, 4, 1], [1, 4, 6, 4, 1]).astype(np.int32), 256
    if name == "gauss9":
        return np.outer([1, 8, 28, 56, 70, 56, 28, 8, 1], [1, 8, 28, 56, 70, 56, 28, 8, 1]).astype(np.int32), 65536
    if name == "dense9":  # a 9x9 blur whose integer coefficients are floored: not rank 1 any more
        c = (np.outer([1, 8, 28, 56, 70, 56, 28, 8, 1], [1, 8, 28, 56, 70, 56, 28, 8, 1]) // 64).astype(np.int32)
        return c, int(c.sum())
    if name == "sharpen7":  # unsharp mask: rank 1 plus a centre term
        c = -np.outer([1, 6, 15, 20, 15, 6, 1], [1, 6, 15, 20, 15, 6, 1]).astype(np.int32)
        c[3, 3] += 2 * 4096
        return c, 4096
    if name == "dense5":  # rank 2: neither a box nor an outer product
        c = np.outer([1, 4, 6, 4, 1], [1, 4, 6, 4, 1]).astype(np.int32)
        c[0, 0] += 3
        c[4, 4] += 1
        return c, int(c.sum())
    if name == "dense7":
        c = np.outer([1, 3, 5, 7, 5, 3, 1], [1, 3, 5, 7, 5, 3, 1]).astype(np.int32)
        c[0, 0] += 2
        c[6, 6] += 1
        return c, int(c.sum())
    if name in ("conv7", "box7"):
        return np.ones((7, 7), np.int32), 49
    raise KeyError(name)


class Bands:
    """One w x full_h raster (nsrc distinct copies of it) cut into row bands over the ranks.  Rank r owns rows
    [y0, y0 + rows) of each in exportable HBM (ppmx_gpu_device_alloc), filled from the LCG raster of SURVEY.md 8d;
    the neighbours' bands are mapped through CUDA IPC so that kernels read halo rows over NVLink themselves."""

    def __init__(self, torch, g, dist, rank, world, w, full_h, nsrc, seed, max_out_bytes_per_row=None, keep_host=False):
        import numpy as np
        from imageprocessingtools_b200 import ppmx as pp
        self.torch, self.g, self.dist, self.rank, self.world, self.pp = torch, g, dist, rank, world, pp
        self.w, self.full_h, self.nsrc = w, full_h, nsrc
        self.y0, self.rows = pp.band_plan(full_h, world, rank, 4)
        self.pitch = w * 3
        nbytes = max(self.rows * self.pitch, 16)
        self.srcs, self.host = [], []
        stage = np.empty((max(self.rows, 1), w, 3), np.uint8)
        for i in range(nsrc):
            p = g.device_alloc(nbytes)
            if self.rows:
                pp.synth_lcg(w, self.rows, seed ^ (i * 0x9E3779B1), y0=self.y0, out=stage[:self.rows])
                g.copy(p, stage.ctypes.data, self.rows * self.pitch, 0)
            self.srcs.append(p)
            if keep_host:
                self.host.append(stage[:self.rows].copy())
        self.seed = seed
        out_row = max_out_bytes_per_row or self.pitch
        # results: two buffers alternate; sized for the largest banded operator (resize x1.5 makes 1.5x the rows)
        self.dst_bytes = max(int(self.rows * 1.5 + 8) * out_row, 16)
        self.dsts = [g.device_alloc(self.dst_bytes) for _ in range(2)]
        self.peers, self.info = {}, None
        if world > 1:
            self.info = [None] * world
            dist.all_gather_object(self.info, (self.y0, self.rows, [g.ipc_export(p) for p in self.srcs]))
            for nb in (rank - 1, rank + 1):
                if 0 <= nb < world and self.info[nb][1] > 0:
                    self.peers[nb] = [g.ipc_open(hd) for hd in self.info[nb][2]]
            dist.barrier()

    def band(self, i, halo, **kw):
        """ppmx_band for source raster i: halo rows above / below point into the neighbours' HBM."""
        pp = self.pp
        b = pp.PpmxBand(full_h=self.full_h, y0=self.y0, halo=halo, **kw)
        if halo and self.world > 1:
            up, dn = self.rank - 1, self.rank + 1
            if up in self.peers:
                if self.info[up][1] < halo:
                    raise SystemExit("band of rank %d is thinner than the halo" % up)
                b.d_top = self.peers[up][i] + (self.info[up][1] - halo) * self.pitch
            if dn in self.peers:
                if self.info[dn][1] < halo:
                    raise SystemExit("band of rank %d is thinner than the halo" % dn)
                b.d_bottom = self.peers[dn][i]
        return b

    def calls(self, op, halo, n, layout=0, tables=0, band_kw=None, src_rows=None):
        """n ppmx_gpu_launch argument tuples cycling through the source rasters (and two result buffers)."""
        import ctypes as C
        self._keep = getattr(self, "_keep", [])
        out = []
        bands = [self.band(i, halo, **(band_kw or {})) for i in range(self.nsrc)]
        self._keep.append((op, bands))
        rows = self.rows if src_rows is None else src_rows
        for k in range(n):
            i = k % self.nsrc
            out.append((C.byref(op), C.c_void_p(self.srcs[i]), self.w, rows, layout, C.c_void_p(self.dsts[k & 1]),
                        C.byref(bands[i]), None, C.c_void_p(tables)))
        return out

    def close(self):
        for lst in self.peers.values():
            for p in lst:
                self.g.ipc_close(p)
        if self.dist is not None:
            self.dist.barrier()  # nobody frees a band a neighbour may still have mapped
        for p in self.srcs + self.dsts:
            self.g.device_free(p)
        self.srcs, self.dsts, self.peers = [], [], {}


def banded_ops(g, pp, w, full_h, world, rank):
    """(label, op, halo rows, bytes/px, band kwargs, tables, out rows of this rank, out row bytes, oracle fn)"""
    import numpy as np
    ops = []
    for label in ("blur3", "box7", "gauss7"):
        coef, div = conv_spec(label)
        k = coef.shape[0]
        ops.append(dict(label="%dx%d %s" % (k, k, label), op=g.conv_op(coef, div, 0), halo=k // 2, bpp=6.0, kind="conv",
                        coef=coef, div=div))
    ops.append(dict(label="gray", op=pp.PpmxOp(kind=pp.OP_GRAY), halo=0, bpp=4.0, kind="gray"))
    ops.append(dict(label="mono (+P4 pack)", op=pp.PpmxOp(kind=pp.OP_MONO_BITS), halo=0, bpp=3.125, kind="mono"))
    ops.append(dict(label="fliph", op=pp.PpmxOp(kind=pp.OP_FLIP, flip_direction=0), halo=0, bpp=6.0, kind="fliph"))
    # flipv / rot180 (zero-traffic form): rank r flips ITS OWN band, and the result is band N-1-r of
    # the output -- whoever consumes it (the download, the next operator) addresses it there.  No byte crosses NVLink.
    ops.append(dict(label="flipv (own band -> output band N-1-r)", op=pp.PpmxOp(kind=pp.OP_FLIP, flip_direction=1), halo=0,
                    bpp=6.0, kind="flipv", plain=True))
    ops.append(dict(label="rot180 (own band -> output band N-1-r)", op=None, halo=0, bpp=6.0, kind="rot180", plain=True))
    new_h = full_h * 3 // 2
    wt, ix = pp.calc_contributions(full_h, new_h, 1.5)
    ops.append(dict(label="imresize height pass x1.5", op=g.imresize_op(new_h, 0, wt, ix), halo=None, bpp=None, kind="resize",
                    wt=wt, ix=ix, new_h=new_h))
    return ops


def band_case(B, g, pp, spec):
    """Resolves one banded operator for this rank's band: (calls-builder kwargs, out rows, out row bytes)."""
    w = B.w
    if spec["kind"] == "rot180":
        spec = dict(spec, op=g.rotate_op(180, w, max(B.rows, 1)))
    kw = dict(op=spec["op"], halo=spec["halo"] or 0)
    out_rows, out_row_bytes, tables = B.rows, w * 3, 0
    if spec["kind"] == "gray":
        out_row_bytes = w
    elif spec["kind"] == "mono":
        out_row_bytes = (w + 7) // 8
    elif spec["kind"] == "resize":
        oy0, orows = pp.band_plan(spec["new_h"], B.world, B.rank, 1)
        need = spec["ix"][oy0:oy0 + orows]
        halo = int(max(0, B.y0 - need.min(), need.max() - (B.y0 + B.rows - 1))) if orows else 0
        tables = g.tables_upload(spec["op"])
        kw.update(halo=halo, band_kw=dict(out_y0=oy0, out_rows=orows), tables=tables)
        out_rows = orows
    if spec.get("plain"):
        kw["plain"] = True
    return spec, kw, out_rows, out_row_bytes, tables


def plain_calls(B, op, n, layout=0):
    """launches on the own band as a raster of its own (no band context: flipv / rot180 of the own band)"""
    import ctypes as C
    B._keep = getattr(B, "_keep", [])
    B._keep.append(op)
    return [(C.byref(op), C.c_void_p(B.srcs[k % B.nsrc]), B.w, B.rows, layout, C.c_void_p(B.dsts[k & 1]), None, None,
             C.c_void_p(0)) for k in range(n)]


def band_parity_check(torch, g, dist, rank, world, device):
    """Every banded operator on a 16384-wide slice: this rank's band from (a) the device-resident launch with halo
    rows read through the neighbours' IPC pointers and (b) the host call ppmx_gpu_apply_band must equal the oracle's
    rows, byte for byte.  Returns (ok, details)."""
    import numpy as np
    import oracle
    from imageprocessingtools_b200 import ppmx as pp
    orc = oracle.orc()
    w, full_h = FULL, 64 * world + 136 + 4 * (world > 1)  # uneven bands, more than the largest halo each
    B = Bands(torch, g, dist, rank, world, w, full_h, 1, 0xC0FFEE ^ 4 ^ 0x51, keep_host=True)
    # the rows the oracle needs for this band: the band plus a margin, cut from the same LCG raster
    margin = 16
    lo, hi = max(0, B.y0 - margin), min(full_h, B.y0 + B.rows + margin)
    near = pp.synth_lcg(w, hi - lo, B.seed, y0=lo)
    details, ok = [], True
    stream = torch.cuda.current_stream().cuda_stream

    def device_band(spec, kw, out_rows, out_row_bytes):
        if kw.pop("plain", False):
            calls = plain_calls(B, kw["op"], 1)
        else:
            calls = B.calls(n=1, **kw)
        import ctypes as C
        if g.L.ppmx_gpu_launch(*calls[0], C.c_void_p(stream)) != 0:
            raise SystemExit("ppmx_gpu_launch failed in band_parity")
        torch.cuda.synchronize()
        out = np.empty(max(out_rows * out_row_bytes, 1), np.uint8)
        if out_rows:
            g.copy(out.ctypes.data, B.dsts[0], out_rows * out_row_bytes, 1)
        return out[:out_rows * out_row_bytes]

    def oracle_band(spec, out_rows):
        k = spec["kind"]
        own = near[B.y0 - lo:B.y0 - lo + B.rows]
        if k == "conv":
            # mirror border at the raster's true top / bottom only: convolve the band with its margin and cut
            r = spec["coef"].shape[0] // 2
            a = max(0, B.y0 - r) if B.y0 > 0 else 0
            b = min(full_h, B.y0 + B.rows + r) if B.y0 + B.rows < full_h else full_h
            ext = near[a - lo:b - lo]
            res = orc.conv(ext, spec["coef"], spec["div"], 0)
            return res[B.y0 - a:B.y0 - a + B.rows].reshape(-1)
        if k == "gray":
            return orc.gray(own).reshape(-1)
        if k == "mono":
            full_phase = np.concatenate([np.zeros((B.y0 % 4, w, 3), np.uint8), own])  # keeps the Bayer row phase
            return orc.pack_pbm(orc.mono(full_phase))[(B.y0 % 4) * ((w + 7) // 8):]
        if k == "fliph":
            return orc.flip(own, 0).reshape(-1)
        if k == "flipv":
            return orc.flip(own, 1).reshape(-1)  # = rows of output band N-1-r
        if k == "rot180":
            return orc.rotate(own, 180).reshape(-1)
        if k == "resize":
            oy0, orows = pp.band_plan(spec["new_h"], world, rank, 1)
            ix = spec["ix"][oy0:oy0 + orows]
            if not orows:
                return np.empty(0, np.uint8)
            a, b = int(ix.min()), int(ix.max()) + 1
            if a < lo or b > hi:
                raise SystemExit("band_parity: margin too small for the resize taps")
            sub = near[a - lo:b - lo]
            return orc.imresize(sub, orows, 0, spec["wt"][oy0:oy0 + orows], (ix - a).astype(np.int32)).reshape(-1)
        raise KeyError(k)

    for spec0 in banded_ops(g, pp, w, full_h, world, rank):
        spec, kw, out_rows, out_row_bytes, tables = band_case(B, g, pp, spec0)
        got = device_band(spec, dict(kw), out_rows, out_row_bytes)
        exp = oracle_band(spec, out_rows)
        same = bool(got.size == exp.size and np.array_equal(got, exp))
        details.append({"op": spec["label"], "path": "device bands + IPC halos", "ok": same})
        ok = ok and same
        if tables:
            g.tables_free(tables)
    # (b) the host call: ppmx_gpu_apply_band on a virtual whole raster of which only this band's source rows exist
    import ctypes as C
    for label, ops_list, layout_bytes in (
            ("3x3 blur3", [g.conv_op(*conv_spec("blur3"), 0)], w * 3),
            ("7x7 box7", [g.conv_op(*conv_spec("box7"), 0)], w * 3),
            ("flipv", [pp.PpmxOp(kind=pp.OP_FLIP, flip_direction=1)], w * 3),
            ("gray", [pp.PpmxOp(kind=pp.OP_GRAY)], w)):
        oy0, orows, sy0, srows = pp.band_rows(ops_list, w, full_h, rank, world)
        src = pp.synth_lcg(w, max(srows, 1), B.seed, y0=sy0)
        out = np.zeros(max(orows * layout_bytes, 1), np.uint8)
        arr = (pp.PpmxOp * len(ops_list))(*ops_list)
        n, rw, rh, ft, by0, brows = C.c_size_t(), C.c_uint32(), C.c_uint32(), C.c_int(), C.c_uint32(), C.c_uint32()
        rc = g.L.ppmx_gpu_apply_band(g.ctx, arr, len(ops_list), C.c_void_p(src.ctypes.data - sy0 * w * 3), w, full_h, rank,
                                     world, C.c_void_p(out.ctypes.data - oy0 * layout_bytes), full_h * layout_bytes,
                                     C.byref(n), C.byref(rw), C.byref(rh), C.byref(ft), C.byref(by0), C.byref(brows))
        same = rc == 0 and (by0.value, brows.value) == (oy0, orows)
        if same and orows:
            if label == "flipv":
                exp = orc.flip(src[:srows], 1).reshape(-1)
            elif label == "gray":
                exp = orc.gray(src[:srows]).reshape(-1)
            else:
                coef, div = conv_spec(label.split()[1])
                a = oy0 - sy0
                # rows beyond the slice exist only at the raster's true edges, where the mirror rule is the oracle's too
                exp = orc.conv(src[:srows], coef, div, 0)[a:a + orows].reshape(-1)
            same = bool(np.array_equal(out[:orows * layout_bytes], exp))
        details.append({"op": label, "path": "ppmx_gpu_apply_band (host raster)", "ok": bool(same)})
        ok = ok and bool(same)
    B.close()
    if dist is not None:
        flags = [None] * world
        dist.all_gather_object(flags, ok)
        ok = all(flags)
    return ok, details


# --------------------------------------------------------------------------------------------
# per-operator extras (N = 1): distinct rasters of one size through one operator
# --------------------------------------------------------------------------------------------

class OpRunner:
    """Builds the op + device buffers of one extra workload; launches go through ppmx_gpu_launch."""

    def __init__(self, torch, g, name, device, seed):
        import ctypes as C
        import numpy as np
        from imageprocessingtools_b200 import ppmx as pp
        self.torch, self.g, self.name, self.pp = torch, g, name, pp
        w, h, batch, bpp, desc = WORKLOADS[name]
        self.w, self.h, self.batch, self.bpp, self.desc = w, h, batch, bpp, desc
        # LCG rasters; raster b = the LCG raster xor a per-raster byte, made on the device
        base = torch.from_numpy(pp.synth_lcg(w, h, seed)).to(device)
        if name == "gray_hist_const":
            base.fill_(200)
        self.src = torch.empty((batch, h, w, 3), dtype=torch.uint8, device=device)
        for b in range(batch):
            torch.bitwise_xor(base, (b * 37) & 0xFF if name != "gray_hist_const" else 0, out=self.src[b])
        del base
        self.hist = None
        self.ops, self.keep = [], []
        L = pp
        for suffix in ("_16k", "_4090", "_1080p", "_const"):
            if name.endswith(suffix):
                name = name[:-len(suffix)]
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
        elif name in ("conv3", "conv7", "sharpen3", "edge3", "gauss5", "gauss7", "sharpen7", "conv3_odd", "dense5", "dense7", "dense9", "gauss9"):
            coef, div = conv_spec(name)
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
        self.pixels_per_step = batch * w * h
        hist_ptr = C.c_void_p(self.hist.data_ptr() if self.hist is not None else 0)
        self.calls = []
        for b in range(batch):
            src, cw, chh = self.src[b].data_ptr(), w, h
            for i, (op, ow, oh, nbytes) in enumerate(self.ops):
                dst = self.dst[i][b].data_ptr()
                self.calls.append((C.byref(op), C.c_void_p(src), cw, chh, L.LAYOUT_RGB8, C.c_void_p(dst), None, hist_ptr,
                                   C.c_void_p(self.tabs[i])))
                src, cw, chh = dst, ow, oh

    def close(self):
        for t in self.tabs:
            if t:
                self.g.tables_free(t)
        for k in self.keep:
            k.close()


def time_op(torch, g, name, device, min_ms=60.0, graph=True, lanes=1):
    """One extra workload, timed over at least min_ms of launches; returns its per_op entry."""
    peak, _ = peaks()
    r = OpRunner(torch, g, name, device, seed=0xC0FFEE ^ 7)
    # chains (resize: two passes per raster) keep their launch order on one stream; independent rasters alternate
    run = StepRunner(torch, g, r.calls, 1, graph, lanes if len(r.ops) == 1 else 1)
    t, _, _ = time_runner(torch, run, 3, 3)
    nst = max(3, min(400, int(min_ms / max(t / 3, 1e-3))))
    t, _, _ = time_runner(torch, run, nst, 1)
    mp = nst * r.pixels_per_step / (t / 1e3) / 1e6
    ent = {"mpix_s": round(mp, 1), "raster": "%dx%d" % (r.w, r.h), "timed_ms": round(t, 1)}
    if r.bpp is not None:
        gbs = r.bpp * mp * 1e6 / 1e9
        ent.update({"gbs": round(gbs, 1), "frac": round(gbs / peak, 4), "static_traffic_from_capture": static_traffic(name)})
    else:
        ent["out_mpix_s"] = round(nst * r.batch * r.out_w * r.out_h / (t / 1e3) / 1e6, 1)
        # algorithmic FP64 instructions: imresize 6*K per output pixel per pass;
        # bicubic rotate ~190 per output pixel that maps inside the source (~ w*h of them)
        if name.startswith("resize"):
            dp = sum(6.0 * o[0].weights_sz * o[1] * o[2] for o in r.ops)
        else:
            dp = 190.0 * r.w * r.h
        rate = nst * r.batch * dp / (t / 1e3)
        ent.update({"bound": "fp64 issue", "dp_inst_per_s": float("%.4g" % rate), "frac_of_dp_peak": round(rate / dp_peak(), 4)})
    run.close()
    r.close()
    del r, run
    torch.cuda.empty_cache()
    return ent


# --------------------------------------------------------------------------------------------
# end to end through the reference-facing calls (pinned host rasters in, host rasters out)
# --------------------------------------------------------------------------------------------

def e2e_bands(torch, g, dist, rank, world, steps, warmup):
    """The headline operator through ppmx_gpu_apply_band: every rank uploads its band of ONE 16384^2 host raster
    (plus the halo rows, straight from the host raster), convolves, downloads its rows into the host result.  The
    host raster is per-rank pinned memory holding exactly the rows the band reads (the call never touches others)."""
    import ctypes as C
    from imageprocessingtools_b200 import ppmx as pp
    w = h = FULL
    op = g.conv_op(*conv_spec("blur3"), 0)
    ops = (pp.PpmxOp * 1)(op)
    oy0, orows, sy0, srows = pp.band_rows([op], w, h, rank, world)
    pitch = w * 3
    src = torch.empty((max(srows, 1), w, 3), dtype=torch.uint8).pin_memory()
    pp.synth_lcg(w, srows, 0xC0FFEE ^ 4, y0=sy0, out=src.numpy()[:srows])
    dst = torch.empty((max(orows, 1) * pitch,), dtype=torch.uint8).pin_memory()
    n, rw, rh, ft, by0, brows = C.c_size_t(), C.c_uint32(), C.c_uint32(), C.c_int(), C.c_uint32(), C.c_uint32()

    def one():
        rc = g.L.ppmx_gpu_apply_band(g.ctx, ops, 1, C.c_void_p(src.data_ptr() - sy0 * pitch), w, h, rank, world,
                                     C.c_void_p(dst.data_ptr() - oy0 * pitch), h * pitch, C.byref(n), C.byref(rw),
                                     C.byref(rh), C.byref(ft), C.byref(by0), C.byref(brows))
        if rc != 0:
            raise SystemExit("ppmx_gpu_apply_band failed")

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
    h2d = int(srows * pitch)
    if dist is not None:
        t = torch.tensor([h2d], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        h2d = int(t.item())
    return {"value": round(steps * w * h / dt / 1e6, 1), "unit": UNIT,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(h * pitch),
            "h2d_bytes_per_step_this_rank": int(srows * pitch), "d2h_bytes_per_step_this_rank": int(orows * pitch),
            "rasters_per_step": 1, "steps": steps,
            "api": "ppmx_gpu_apply_band (pinned host raster in, pinned host raster out; one 16384x16384 raster per step, "
                   "row band per rank, halo rows uploaded from the host raster)"}


def e2e_batch(torch, g, name, steps, warmup, dist=None):
    """An extra workload through ppmx_gpu_apply_batch: pinned host rasters in, pinned host results out."""
    import ctypes as C
    from imageprocessingtools_b200 import ppmx as pp
    w, h, batch, _, _ = WORKLOADS[name]
    batch = min(batch, 8)
    hist_host = None
    if name == "gray":
        ops = (pp.PpmxOp * 1)(pp.PpmxOp(kind=pp.OP_GRAY))
    elif name in ("gray_hist", "gray_hist_const"):
        hist_host = torch.zeros((batch, 256), dtype=torch.int64).pin_memory()
        ops = (pp.PpmxOp * 1)(pp.PpmxOp(kind=pp.OP_GRAY_HIST,
                                        hist_out=C.cast(C.c_void_p(hist_host.data_ptr()), C.POINTER(C.c_uint64))))
    elif name == "mono":
        ops = (pp.PpmxOp * 1)(pp.PpmxOp(kind=pp.OP_MONO))
    elif name in ("fliph", "flipv"):
        ops = (pp.PpmxOp * 1)(pp.PpmxOp(kind=pp.OP_FLIP, flip_direction=int(name == "flipv")))
    elif name in ("rot90", "rot180", "rot30"):
        ops = (pp.PpmxOp * 1)(g.rotate_op(int(name[3:]), w, h))
    elif name in ("conv3", "conv7", "gauss7"):
        ops = (pp.PpmxOp * 1)(g.conv_op(*conv_spec(name), 0))
    else:
        return None
    out_each = pp.chain_info([ops[0]], w, h)[3]
    src = torch.empty((batch, h, w, 3), dtype=torch.uint8).pin_memory()
    for b in range(batch):
        pp.synth_lcg(w, h, 0xC0FFEE ^ 2 ^ (b << 8), out=src.numpy()[b])
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
        one()
    dt = time.perf_counter() - t0
    # the same call on ONE raster (ppmx_gpu_apply): its row parts keep upload, kernel and download overlapped
    t1 = time.perf_counter()
    for _ in range(steps):
        rc = g.L.ppmx_gpu_apply(g.ctx, ops, 1, C.c_void_p(src.data_ptr()), w, h, C.c_void_p(dst.data_ptr()), out_each + 16,
                                C.byref(each), C.byref(ow), C.byref(oh), C.byref(ft))
        if rc != 0:
            raise SystemExit("ppmx_gpu_apply failed")
    dt1 = time.perf_counter() - t1
    if dist is not None:
        t = torch.tensor([dt, dt1], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dt1 = float(t[0].item()), float(t[1].item())
    world = dist.get_world_size() if dist is not None else 1
    return {"value": round(world * steps * batch * w * h / dt / 1e6, 1), "unit": UNIT,
            "h2d_bytes_per_step": batch * w * h * 3, "d2h_bytes_per_step": batch * int(each.value),
            "rasters_per_step": batch, "api": "ppmx_gpu_apply_batch (pinned host in/out)",
            "single_raster_apply_mpix_s": round(world * steps * w * h / dt1 / 1e6, 1)}


def batch_chain_run(torch, g, dist, world, images=128):
    """BASELINE config 5: a batch of 1920x1080 PPM rasters through full op chains, image-parallel (every rank
    takes `images` rasters), end to end through ppmx_gpu_apply_batch with pinned host buffers."""
    import ctypes as C
    from imageprocessingtools_b200 import ppmx as pp
    w, h = 1920, 1080
    src = torch.empty((images, h, w, 3), dtype=torch.uint8).pin_memory()
    pp.synth_lcg(w, h * images, 0xC0FFEE ^ 5, out=src.numpy().reshape(images * h, w, 3))
    out = []
    for label, kw in (("-w960 -r90 -gray -fv", dict(resize_w=960, angle=90, gray=True, flipv=True)),
                      ("-r90 -mono -fh", dict(angle=90, mono=True, fliph=True))):
        ph = pp._PlanHolder(w=w, h=h, **kw)
        ops = [ph.plan.ops[i] for i in range(ph.plan.nops)]
        info = pp.chain_info(ops, w, h)
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
        n0 = g.launch_count()
        t0 = time.perf_counter()
        one()
        dt = time.perf_counter() - t0
        kernels = (g.launch_count() - n0) / images
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        ent = {"chain": label, "images": images * world, "frame": "%dx%d" % (w, h),
               "out": "%dx%d type %d" % (ow.value, oh.value, ft.value),
               "kernels_per_frame": round(kernels, 2), "stages": info[5],
               "e2e_mpix_s": round(world * images * w * h / dt / 1e6, 1),
               "images_per_s": round(world * images / dt, 1)}
        del dst
        # the same chain on frames that are already in HBM (ppmx_gpu_chain_prepare / _run): kernels only
        try:
            lanes = 3
            dsrc = src.cuda()
            prepared = [g.chain_prepare(ops, w, h) for _ in range(lanes)]
            meta = prepared[0][1]
            ddst = [torch.empty(meta["out_bytes"] + 16, dtype=torch.uint8, device="cuda") for _ in range(lanes)]
            calls = [(prepared[i % lanes][0], C.c_void_p(dsrc[i].data_ptr()), C.c_void_p(ddst[i % lanes].data_ptr()))
                     for i in range(images)]
            run = StepRunner(torch, g, calls, 1, True, lanes, fn=g.L.ppmx_gpu_chain_run, kernels_per_call=meta["kernels"])
            t, _, _ = time_runner(torch, run, 3, 2, dist)
            nst = max(3, min(200, int(60.0 / max(t / 3, 1e-3))))
            t, _, _ = time_runner(torch, run, nst, 1, dist)
            fps = world * nst * images / (t / 1e3)
            ent.update({"device_resident_images_per_s": round(fps, 1),
                        "device_resident_mpix_s": round(fps * w * h / 1e6, 1),
                        "bytes_moved_per_frame": meta["bytes_moved"],
                        "device_resident_gbs": round(fps * meta["bytes_moved"] / 1e9 / world, 1),
                        "device_resident_launch": run.mode})
            run.close()
            for ch, _ in prepared:
                g.chain_free(ch)
            del dsrc, ddst
            torch.cuda.empty_cache()
        except Exception as e:
            ent["device_resident_error"] = str(e)[:160]
        out.append(ent)
        ph.close()
    return out


def cli_rows(g):
    """Wall clock of the command line on a 4096x4096 P6 (50 MB): ppmx-b200 beside the reference's own binary
    (SURVEY.md 8d(ii)); file read, header parse, operator, file write all inside."""
    import subprocess
    import tempfile
    import numpy as np
    import oracle
    from imageprocessingtools_b200 import ppmx as pp
    rows = []
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "in.ppm")
        oracle.write_p6(path, pp.synth_lcg(4096, 4096, 0xC0FFEE ^ 2))
        for flag in ("-gray", "-fh"):
            ent = {"args": flag, "file": "4096x4096 P6, 50 MB"}
            for label, exe in (("ppmx_b200_s", pp.CLI), ("reference_cli_s", oracle.REF_CLI)):
                if not os.path.exists(exe):
                    continue
                best = None
                for _ in range(2):
                    t0 = time.perf_counter()
                    p = subprocess.run([exe, flag, path], capture_output=True)
                    dt = time.perf_counter() - t0
                    if p.returncode == 0:
                        best = dt if best is None else min(best, dt)
                    if label == "ppmx_b200_s" and p.returncode == 0:
                        ent["sha_out"] = __import__("hashlib").sha256(open(path + ".out", "rb").read()).hexdigest()[:16]
                    elif p.returncode == 0:
                        same = __import__("hashlib").sha256(open(path + ".out", "rb").read()).hexdigest()[:16]
                        ent["same_bytes"] = (same == ent.get("sha_out"))
                ent[label] = round(best, 3) if best is not None else None
            ent.pop("sha_out", None)
            rows.append(ent)
        # a batch of files: one run of ppmx-b200 -batch (one device context, file I/O overlapped with the device work)
        # beside one run of the reference per file
        n = 8
        paths = []
        for i in range(n):
            p = os.path.join(td, "b%d.ppm" % i)
            oracle.write_p6(p, pp.synth_lcg(4096, 4096, 0xC0FFEE ^ 2 ^ (i + 1)))
            paths.append(p)
        ent = {"args": "-gray", "files": "%d x 4096x4096 P6 (50 MB each)" % n}
        if os.path.exists(pp.CLI):
            t0 = time.perf_counter()
            pr = subprocess.run([pp.CLI, "-batch", "-gray"] + paths, capture_output=True)
            ent["ppmx_b200_batch_s"] = round(time.perf_counter() - t0, 3) if pr.returncode == 0 else None
        if os.path.exists(oracle.REF_CLI):
            t0 = time.perf_counter()
            ok = all(subprocess.run([oracle.REF_CLI, "-gray", p], capture_output=True).returncode == 0 for p in paths)
            ent["reference_cli_one_run_per_file_s"] = round(time.perf_counter() - t0, 3) if ok else None
        rows.append(ent)
    return rows


# --------------------------------------------------------------------------------------------
# CPU arms
# --------------------------------------------------------------------------------------------

def cpu_conv_sample(threads, seconds_budget, w=FULL, rows_per_call=256):
    """The 3x3 convolution on the host cores: the plain-C oracle port (the reference has no convolution) on
    row slabs of the 16384-wide raster, one slab per call per thread, for about seconds_budget."""
    import numpy as np
    import oracle
    from imageprocessingtools_b200 import ppmx as pp
    orc = oracle.orc()
    coef, div = conv_spec("blur3")
    slabs = [pp.synth_lcg(w, rows_per_call + 2, 0xC0FFEE ^ 4, y0=i * rows_per_call) for i in range(threads)]
    count = [0] * threads
    busy = [0.0] * threads
    t_end = time.perf_counter() + seconds_budget

    def work(i):
        while True:
            t0 = time.perf_counter()
            orc.conv(slabs[i], coef, div, 0)
            busy[i] += time.perf_counter() - t0
            count[i] += 1
            if time.perf_counter() > t_end:
                break

    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    wall = time.perf_counter() - t0
    px = sum(count) * w * rows_per_call
    return {"value": round(px / wall / 1e6, 1), "unit": UNIT, "cores": threads, "kind": "port",
            "cpu": cpu_model(), "host_threads_available": os.cpu_count(),
            "sample": "%d calls of the plain-C oracle's 3x3 convolution (extension operator: the reference has none) on "
                      "%dx%d slabs of the 16384-wide LCG raster over %d thread(s), %.1f s wall" %
                      (sum(count), w, rows_per_call, threads, wall)}


def cpu_reference_sample(name, threads, seconds_budget=12.0):
    """A reference operator from the compiled reference (oracle/_ref) on this box's host cores; bounded sample."""
    import numpy as np
    import oracle
    from imageprocessingtools_b200 import ppmx as pp
    ref = oracle.ref()
    if ref is None:
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": "oracle/_ref missing"}
    w, h, _, _, _ = WORKLOADS[name]

    def call(r, img):
        if name in ("gray", "gray_hist", "gray_hist_const"):
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

    imgs = [pp.synth_lcg(w, h, 0xC0FFEE ^ 2 ^ i) for i in range(threads)]
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
    fn = {"gray_hist": "gray", "gray_hist_const": "gray", "fliph": "flip", "flipv": "flip", "rot90": "rotate",
          "rot180": "rotate", "rot30": "rotate"}.get(name, name)
    return {"value": round(total_px / max(op_time) / 1e6, 1), "unit": UNIT, "cores": threads, "kind": "reference",
            "cpu": cpu_model(), "host_threads_available": os.cpu_count(),
            "sample": "%d calls of the reference's %s() on %dx%d rasters over %d thread(s); time inside the "
                      "reference function only (its own image_buff_alloc included)%s" %
                      (sum(count), fn, w, h, threads, "; the reference has no histogram: gray() only" if "hist" in name else "")}


def cpu_reference_per_op(names):
    """One bounded, single-threaded call of the compiled reference per operator, on a 4096x4096 raster of the same
    kind; time inside the reference function only."""
    import oracle
    from imageprocessingtools_b200 import ppmx as pp
    ref = oracle.ref()
    if ref is None:
        return {}
    img = pp.synth_lcg(4096, 4096, 0xC0FFEE ^ 2)
    out, memo = {}, {}
    for name in names:
        base = name
        for suffix in ("_16k", "_4090", "_1080p", "_const"):
            if base.endswith(suffix):
                base = base[:-len(suffix)]
        try:
            if base in memo:
                out[name] = memo[base]
                continue
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
            memo[base] = out[name] = round(4096 * 4096 / t / 1e6, 1)
        except Exception:
            pass
    return out


def headline_config(world, mode=None):
    """The `config` object of the default workload -- the same dictionary in both arms."""
    cfg = {"workload": WORKLOADS["conv3_bands"][4], "op": "conv3_bands", "raster": "%dx%d" % (FULL, FULL),
           "filter": "3x3 blur (1 2 1; 2 4 2; 1 2 1)/16, mirror border, round half up",
           "rasters_per_step": RASTERS_PER_STEP,
           "l2": "inputs larger than L2: a step cycles through %d distinct resident rasters of %d MB each"
                 % (NSRC, FULL * FULL * 3 // 1000000),
           "sharding": "row bands (ppmx_band_plan), one per rank; no collective; halo rows read from peer HBM (CUDA IPC)",
           "inputs": "LCG s=s*1664525+1013904223, r=s>>24 g=s>>16 b=s>>8 (SURVEY.md 8d)"}
    return cfg


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------

def run_headline(args, torch, g, dist, rank, world, device, line):
    import numpy as np
    from imageprocessingtools_b200 import ppmx as pp
    peak, peak_src = peaks()

    ok, details = band_parity_check(torch, g, dist, rank, world, device)
    line["band_parity"] = ok
    line["band_parity_detail"] = details
    if not ok:
        if rank == 0:
            print(json.dumps(line), flush=True)
        raise SystemExit("band parity FAILED: %s" % [d for d in details if not d["ok"]])

    B = Bands(torch, g, dist, rank, world, FULL, FULL, NSRC, 0xC0FFEE ^ 4)
    op = g.conv_op(*conv_spec("blur3"), 0)
    replays = max(1, RASTERS_PER_STEP // GRAPH_LAUNCHES)
    run = StepRunner(torch, g, B.calls(op, 1, GRAPH_LAUNCHES), replays, not args.no_graph, args.lanes)
    n0 = g.launch_count()
    with ClockSampler(torch.cuda.current_device()) as clk:
        ms, ms_min, ms_med = time_runner(torch, run, args.steps, args.warmup, dist, clk)
    launches = (g.launch_count() - n0) - args.warmup * run.launches_per_step
    px = args.steps * RASTERS_PER_STEP * FULL * FULL
    value = px / (ms / 1e3) / 1e6
    band_px = B.rows * FULL
    per_launch_ms = ms / (args.steps * run.launches_per_step)
    achieved = 6.0 * band_px / (per_launch_ms / 1e3) / 1e9
    line.update({
        "value": round(value, 1), "ms_per_step": round(ms / args.steps, 4), "scaling": "strong",
        "config": headline_config(world),
        "run": {"launch": run.mode, "rows_per_rank": B.rows, "halo_rows": 1, "timed_region_s": round(ms / 1e3, 3)},
        "ranks_ms": {"max": round(ms, 3), "median": round(ms_med, 3), "min": round(ms_min, 3)},
        "gpu_launches": int(launches), "clocks": clk.summary(),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": static_traffic("conv3_bands"),
                     "traffic_note": "static, from a committed ncu --set full capture (profiles/traffic.json), per launch of a "
                                     "whole 16384^2 raster; not measured in this run",
                     "peak_source": peak_src, "kernel": "conv3_strip_kernel", "algorithmic_bytes_per_pixel": 6.0,
                     "algorithmic_bytes_per_launch": int(6.0 * band_px), "avg_launch_us": round(per_launch_ms * 1e3, 2),
                     "note": "per rank: one launch = this rank's band (%d rows x %d px)" % (B.rows, FULL)},
    })
    run.close()
    extras = []
    if not args.no_band:
        # the other banded operators on the same 16384^2 raster (strong scaling, shorter runs)
        for spec0 in banded_ops(g, pp, FULL, FULL, world, rank):
            if spec0["label"].endswith("blur3"):
                continue
            spec, kw, out_rows, out_row_bytes, tables = band_case(B, g, pp, spec0)
            n = 32
            calls = plain_calls(B, kw["op"], n) if kw.pop("plain", False) else B.calls(n=n, **kw)
            r2 = StepRunner(torch, g, calls, 1, not args.no_graph, args.lanes)
            t, _, _ = time_runner(torch, r2, 3, 2, dist)
            nst = max(3, min(200, int(80.0 / max(t / 3, 1e-3))))
            t, _, _ = time_runner(torch, r2, nst, 1, dist)
            mp = nst * n * FULL * FULL / (t / 1e3) / 1e6
            ent = {"workload": "%dx%d raster, %s, row bands over %d GPU(s)" % (FULL, FULL, spec["label"], world),
                   "scaling": "strong", "mpix_s": round(mp, 1), "ms_per_raster": round(t / (nst * n), 4),
                   "timed_ms": round(t, 1)}
            if spec["bpp"]:
                ent["gbs_total"] = round(spec["bpp"] * mp / 1e3, 1)
                ent["frac_of_n_x_peak"] = round(spec["bpp"] * mp / 1e3 / (peak * world), 4)
            else:
                ent["out_mpix_s"] = round(mp * 1.5, 1)
                ent["halo_rows"] = kw.get("halo")
            extras.append(ent)
            r2.close()
            if tables:
                g.tables_free(tables)
        line["band_split"] = extras
    B.close()
    torch.cuda.empty_cache()
    line["e2e"] = e2e_bands(torch, g, dist, rank, world, max(3, min(args.steps, 10)), 2)


def run_single_op(args, torch, g, dist, rank, world, device, line):
    """--workload <op>: distinct rasters of one size through one operator, rasters sharded over the ranks (weak)."""
    name = args.workload
    w, h, batch, bpp, desc = WORKLOADS[name]
    peak, peak_src = peaks()
    r = OpRunner(torch, g, name, device, seed=0xC0FFEE ^ (rank + 2))
    run = StepRunner(torch, g, r.calls, 1, not args.no_graph, args.lanes)
    drain = None
    if r.hist is not None and dist is not None:
        # the one real exchange of this path is the job's FINAL histogram reduction (north star: "NCCL only for the
        # final histogram reduction"): one all-reduce of the 256 accumulated bins inside the timed region
        def drain():
            dist.all_reduce(r.hist, op=dist.ReduceOp.SUM)
    n0 = g.launch_count()
    with ClockSampler(torch.cuda.current_device()) as clk:
        ms, ms_min, ms_med = time_runner(torch, run, args.steps, args.warmup, dist, clk, drain)
    launches = (g.launch_count() - n0) - args.warmup * run.launches_per_step
    value = world * args.steps * r.pixels_per_step / (ms / 1e3) / 1e6
    line.update({"value": round(value, 1), "ms_per_step": round(ms / args.steps, 4), "scaling": "weak",
                 "config": {"workload": desc, "op": name, "raster": "%dx%d" % (w, h), "rasters_per_step_per_gpu": batch,
                            "l2": "inputs larger than L2 (%d MB per step per GPU)" % (batch * w * h * 3 // 1000000),
                            "launch": run.mode, "timed_region_s": round(ms / 1e3, 3),
                            "sharding": "rasters sharded across ranks, no data-path collective" +
                                        ("; one NCCL all-reduce of the 256 accumulated histogram bins at the end of the "
                                         "timed region" if drain else "")},
                 "ranks_ms": {"max": round(ms, 3), "median": round(ms_med, 3), "min": round(ms_min, 3)},
                 "gpu_launches": int(launches), "clocks": clk.summary()})
    if bpp is not None:
        per_launch_ms = ms / (args.steps * run.launches_per_step)
        achieved = bpp * w * h / (per_launch_ms / 1e3) / 1e9
        line["roofline"] = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                            "frac": round(achieved / peak, 4), "traffic": static_traffic(name),
                            "traffic_note": "static, from a committed ncu capture; not measured in this run",
                            "peak_source": peak_src, "algorithmic_bytes_per_launch": int(bpp * w * h), "kernel": name,
                            "algorithmic_bytes_per_pixel": bpp, "avg_launch_us": round(per_launch_ms * 1e3, 2)}
    else:
        line["roofline"] = {"bound": "fp64 issue (not HBM, not tensor)", "achieved": None, "peak": None, "unit": "GB/s",
                            "frac": None, "traffic": None, "kernel": name}
    run.close()
    r.close()
    del r, run
    torch.cuda.empty_cache()
    e2e = e2e_batch(torch, g, name, max(2, min(args.steps, 8)), 1, dist)
    if e2e is not None:
        line["e2e"] = e2e


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
    # everything is launched and timed on one non-default stream (the legacy default stream can not be captured)
    torch.cuda.set_stream(torch.cuda.Stream())

    name = args.workload
    line = {"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": None, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic"}
    if name == "conv3_bands":
        run_headline(args, torch, g, dist, rank, world, device, line)
    else:
        run_single_op(args, torch, g, dist, rank, world, device, line)

    if not args.no_per_op and world == 1:
        per = {}
        for op_name in PER_OP:
            try:
                per[op_name] = time_op(torch, g, op_name, device, graph=not args.no_graph, lanes=args.lanes)
            except Exception as e:  # an extra must not take the headline down
                per[op_name] = {"error": str(e)[:120]}
        e2e_extra = {}
        for op_name in ("gray_hist", "gray", "mono", "fliph", "rot90", "conv3"):
            try:
                e2e_extra[op_name] = e2e_batch(torch, g, op_name, 3, 1)
            except Exception as e:
                e2e_extra[op_name] = {"error": str(e)[:120]}
        line["per_op"] = per
        line["per_op_e2e"] = e2e_extra
    if not args.no_band:
        try:
            line["batch_chain"] = batch_chain_run(torch, g, dist, world)
        except Exception as e:
            line["batch_chain"] = {"error": str(e)[:120]}
    if rank == 0 and world == 1 and not args.no_cpu:
        if name == "conv3_bands":
            line["cpu_baseline"] = cpu_conv_sample(1, 10.0)
        else:
            line["cpu_baseline"] = cpu_reference_sample(name, 1)
        if "per_op" in line:  # the reference's own function, 1 thread, 4096x4096, beside every operator
            for k, v in cpu_reference_per_op(list(line["per_op"])).items():
                line["per_op"][k]["cpu_reference_mpix_s_1thread"] = v
        try:
            line["cli"] = cli_rows(g)
        except Exception as e:
            line["cli"] = {"error": str(e)[:120]}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)
    g.close()


# --------------------------------------------------------------------------------------------
# reference arm: the CPU implementation of the same operator on the host cores
# --------------------------------------------------------------------------------------------

def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = min(os.cpu_count() or 1, 64)
    t0 = time.perf_counter()
    # each "step" is a bounded sample; total CPU time is capped to stay within minutes
    budget = min(60.0, max(4.0, 1.0 * (args.steps + args.warmup)))
    if name == "conv3_bands":
        cb = cpu_conv_sample(threads, budget)
        cfg = headline_config(world)
    else:
        cb = cpu_reference_sample(name, threads, seconds_budget=budget)
        w, h, batch, _, desc = WORKLOADS[name]
        cfg = {"workload": desc, "op": name, "raster": "%dx%d" % (w, h), "rasters_per_step_per_gpu": batch}
    dt = time.perf_counter() - t0
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3 / max(args.steps, 1), 3),
            "higher_is_better": True, "scaling": "strong" if name == "conv3_bands" else "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": cfg, "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-per-op", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-band", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue every launch from the host instead of replaying a CUDA graph")
    ap.add_argument("--lanes", type=int, default=3,
                    help="streams the launches of a step go round-robin over (independent rasters: ramps and tails overlap)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
