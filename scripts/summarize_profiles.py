#!/usr/bin/env python
"""Turns gpurun_out/launches.csv and gpurun_out/prof_*.ncu-rep into small tracked summaries under profiles/.
usage: summarize_profiles.py <tag>      (e.g. r01_baseline)"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
G = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__cycles_elapsed.avg", "smsp__sass_average_data_bytes_per_sector_mem_local_op_ld.ratio",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]


def launches(fname="launches.csv", suffix="launches", cmd="python bench.py --rays 16777216 --steps 2 --warmup 3 --cpu-rays 100000 --pt-spp 32"):
    p = os.path.join(G, fname)
    if not os.path.exists(p):
        return
    txt = open(p).read().splitlines()
    start = next(i for i, l in enumerate(txt) if l.startswith('"ID"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(txt[start:]))))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v = {"us": v / 1e3, "ns": v / 1e6, "ms": v, "s": v * 1e3}.get(r["Metric Unit"], v)
        k = r["Kernel Name"].split("(")[0][:80]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(OUT, f"{tag}_{suffix}.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache serialised: compare shares)\n\n")
        f.write(f"command: `{cmd}`\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f}% |\n")
        f.write(f"\ntotal {tot:.1f} ms over {sum(v[0] for v in agg.values())} launches\n")


def full(rep, name):
    p = os.path.join(G, rep)
    if not os.path.exists(p):
        return None
    raw = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    with open(os.path.join(OUT, f"{tag}_{name}_metrics.csv"), "w") as f:
        f.write("metric,unit,value\n")
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS or h == "Kernel Name" or ("warp_issue_stalled" in h and h.endswith("per_warp_active.pct")):
                f.write(f"{h},{u},\"{v}\"\n")
                d[h] = (u, v)
    return d


os.makedirs(OUT, exist_ok=True)
launches((sys.argv[2] if len(sys.argv) > 2 else "") + "launches.csv")
launches((sys.argv[2] if len(sys.argv) > 2 else "") + "launches_default.csv", "launches_default_cmd", "python bench.py --steps 2 --warmup 1   (first 600 launches)")
PREFIX = sys.argv[2] if len(sys.argv) > 2 else ""
t = full(PREFIX + "prof_trace.ncu-rep", "trace_kernel")
full(PREFIX + "prof_extend.ncu-rep", "k_extend")
if t:
    def num(k):
        u, v = t[k]
        v = float(v.replace(",", ""))
        mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
        return v * mult
    traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    sys.path.insert(0, ROOT)
    import bench
    json.dump({"tag": tag, "trace_kernel_dram_bytes_per_launch": traffic, "rays_per_launch": 16777216,
               "dram_bytes_per_ray": traffic / 16777216, "source_hash": bench.source_hash(),
               "source_hash_of": "sha256 of csrc/{traverse.cuh,accel.cu,bvh.h,triaccel.h,bvh_build.cpp} at capture time (bench.py source_hash())",
               "layout_hash": bench.layout_hash(),
               "layout_hash_of": "sha256 of csrc/{bvh.h,bvh_dev.h,triaccel.h,bvh_build.cpp,bvh_build_gpu.cu}: unit layout and builders (bench.py layout_hash())",
               "mem_signature": bench.mem_signature(),
               "mem_signature_of": "sha256 of the ordered memory instructions (opcode + modifiers) in the SASS of trace_kernel<false,false> (bench.py mem_signature())",
               "note": "ncu --set full capture of lmb200::trace_kernel<false,false> at 16 Mi rays / 4 M triangles; bench.py scales it per ray"},
              open(os.path.join(OUT, "traffic.json"), "w"), indent=1)
print("profiles written for", tag)
