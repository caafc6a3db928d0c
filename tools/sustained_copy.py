#!/usr/bin/env python
"""tools/sustained_copy.py -- the copy bandwidth MEASURED_PEAKS.json quotes is a burst figure (best of 10 single copies).  This
times the same torch copy (1 Gi bf16 elements, read + write bytes) back to back for several seconds, the way bench.py's headline
region runs, and samples the SM clock meanwhile: what the HBM roofline is worth under the 1 kW power cap."""
import json
import subprocess
import threading
import time

import torch


def main():
    torch.cuda.set_device(0)
    a = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda")
    b = torch.empty_like(a)
    a.normal_()
    nbytes = 2 * a.numel() * a.element_size()
    out = {}
    # burst: best of 10 single copies
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        b.copy_(a)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
        time.sleep(0.05)
    out["burst_gbs"] = round(nbytes / best / 1e6, 1)
    clocks = []
    stop = threading.Event()

    def sample():
        while not stop.is_set():
            try:
                r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader,nounits",
                                    "-i", "0"], capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                clocks.append((float(r[0]), float(r[1])))
            except Exception:
                pass
            time.sleep(0.1)

    th = threading.Thread(target=sample)
    th.start()
    for secs in (0.06, 1.0, 4.0):
        n = max(1, int(secs / (best / 1e3)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            b.copy_(a)
        e1.record()
        torch.cuda.synchronize()
        out["back_to_back_%gs_gbs" % secs] = round(n * nbytes / e0.elapsed_time(e1) / 1e6, 1)
    stop.set()
    th.join()
    if clocks:
        cs = sorted(c for c, _ in clocks)
        out["sm_mhz_median_under_load"] = cs[len(cs) // 2]
        out["power_w_max"] = max(p for _, p in clocks)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
