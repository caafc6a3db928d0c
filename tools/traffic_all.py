#!/usr/bin/env python
"""tools/traffic_all.py -- DRAM bytes per launch of every per_op kernel, for profiles/traffic.json.

  step 1 (on the GPU box, under ncu, limited metrics = one or two replay passes per launch):
      ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
          --clock-control none --csv --log-file gpurun_out/traffic_launches.csv python tools/traffic_all.py run > gpurun_out/traffic_ops.jsonl
      For each workload one warm-up step runs unprofiled, then ONE raster's launches between cudaProfilerStart/Stop; the script
      prints how many launches that was, so the CSV rows map to workloads by order.
  step 2 (anywhere):  python tools/traffic_all.py merge gpurun_out/traffic_ops.jsonl gpurun_out/traffic_launches.csv
      rewrites profiles/traffic.json (entries of workloads not captured are kept).
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run():
    import torch
    import bench
    import imageprocessingtools_b200 as ip
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    g = ip.Ppmx(0)
    import ctypes as C
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def fire(call):
        if g.L.ppmx_gpu_launch(*call, stream) != 0:
            raise SystemExit("ppmx_gpu_launch failed")

    names = [n for n in sys.argv[2].split(",")] if len(sys.argv) > 2 else list(bench.PER_OP)
    for name in names:
        r = bench.OpRunner(torch, g, name, dev, seed=11)
        per_raster = len(r.ops)
        for c in r.calls:  # warm-up: every raster once (also fills the pool)
            fire(c)
        torch.cuda.synchronize()
        n0 = g.launch_count()
        torch.cuda.profiler.start()
        for c in r.calls[:per_raster]:  # one raster (ncu flushes the caches before each replay pass anyway)
            fire(c)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"op": name, "launches": g.launch_count() - n0, "raster": "%dx%d" % (r.w, r.h),
                          "algorithmic_bytes": int(r.bpp * r.w * r.h) if r.bpp else None}), flush=True)
        r.close()
        del r
        torch.cuda.empty_cache()
    g.close()


def merge(ops_path, csv_path):
    ops = [json.loads(l) for l in open(ops_path) if l.startswith("{")]
    rows = [r for r in csv.reader(l for l in open(csv_path) if not l.startswith("=="))]
    hdr = rows[0]
    i_name, i_metric, i_val, i_id = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    launches = {}
    for r in rows[1:]:
        if len(r) <= i_val:
            continue
        d = launches.setdefault(int(r[i_id]), {"kernel": r[i_name]})
        d[r[i_metric]] = float(r[i_val].replace(",", ""))
        d["unit:" + r[i_metric]] = r[hdr.index("Metric Unit")]
    seq = [launches[k] for k in sorted(launches)]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    path = os.path.join(ROOT, "profiles", "traffic.json")
    out = json.load(open(path))
    pos = 0
    for o in ops:
        mine = seq[pos:pos + o["launches"]]
        pos += o["launches"]
        if len(mine) != o["launches"]:
            raise SystemExit("launch list shorter than the workloads say")
        rd = sum(m["dram__bytes_read.sum"] * scale[m["unit:dram__bytes_read.sum"]] for m in mine)
        wr = sum(m["dram__bytes_write.sum"] * scale[m["unit:dram__bytes_write.sum"]] for m in mine)
        ent = {"kernel": " + ".join(m["kernel"].split("(")[0].replace("ppmx::", "").replace("void ", "") for m in mine),
               "raster": o["raster"], "dram_read": int(rd), "dram_write": int(wr),
               "report": "tools/traffic_all.py (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, one raster, cold-ish L2)"}
        if o.get("algorithmic_bytes"):
            ent["algorithmic_bytes"] = o["algorithmic_bytes"]
        out[o["op"]] = ent
    json.dump(out, open(path, "w"), indent=1)
    print("profiles/traffic.json: %d workloads updated" % len(ops))


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run()
    else:
        merge(sys.argv[2], sys.argv[3])
