#!/usr/bin/env python
"""bench.py — headline benchmark of the lmb200 hot path (see DESIGN.md "Measurement").

Workload (BASELINE.json configs[3]): incoherent closest-hit ray casting, 64 Mi random rays against a
4 M-triangle synthetic soup BVH, per GPU (weak scaling: every rank traces its own batch, BVH replicated,
no data-path collective). One "step" = one pass of lmb200_trace_closest over the whole batch.

  value     Mrays/s, device-resident rays/hits, CUDA events on the launching stream, max over ranks
  e2e       Mrays/s through the host-buffer C-ABI call (pinned host rays in, hits out, copies inside)
  roofline  algorithmic bytes per ray (48 + nodes/ray*64 + tris/ray*48, counted by the instrumented
            kernel) * rays / kernel time, against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline  the CPU oracle on a bounded ray sample of the same scene (rank 0, N=1 only)
  path_tracing  secondary figure: wavefront ptdirect Msamples/s on configs[2] at full size (1 M-triangle
                scene, 1920x1080, 1024 spp = 2.1 G samples), sample range sharded over ranks
                (strong scaling), films summed with one NCCL reduce

`--impl reference` times the reference's own CPU implementation (oracle/_ref: accel::qbvh through the real
Accel3::Intersect and the vendored nanort, all host threads; the faster one is the reported value) on a bounded
sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lightmetrica-v2_b200"))

NODE_BYTES = 64.0     # csrc/bvh.h Node64 (round 1: 80); a triangle record is 48 algorithmic bytes (stored in a 64-byte unit)
METRIC = "Mrays/s incoherent closest-hit (64Mi random rays vs 4M-tri synthetic BVH, per-GPU batch)"
UNIT = "Mrays/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="lmb200", choices=["lmb200", "reference"])
    ap.add_argument("--tris", type=int, default=4_000_000)
    ap.add_argument("--rays", type=int, default=64 * 1024 * 1024)
    ap.add_argument("--cpu-rays", type=int, default=1_000_000, help="bounded CPU sample")
    ap.add_argument("--no-pt", action="store_true", help="skip the secondary path-tracing figure")
    ap.add_argument("--pt-tris", type=int, default=1_000_000)
    ap.add_argument("--pt-spp", type=int, default=1024)
    ap.add_argument("--pt-pool", type=int, default=0, help="wavefront pool size (0 = library default)")
    ap.add_argument("--pt-cpu-spp", type=int, default=2, help="bounded CPU sample of the path-tracing workload (samples per pixel)")
    ap.add_argument("--no-one", action="store_true", help="skip the per-ray Accel3::Intersect figure (its persistent service kernel distorts an ncu launch list)")
    ap.add_argument("--no-c4", action="store_true", help="skip the configs[4] leg (10M triangles, 4K film, NCCL film reduce)")
    ap.add_argument("--c4-spp", type=int, default=256, help="samples per pixel of the configs[4] leg (the config names 4096; throughput is per sample)")
    return ap.parse_args()


def workload_config(a):
    return {"workload": "configs[3] incoherent ray-cast microbench", "triangles": a.tris, "rays_per_gpu": a.rays,
            "scene": "soup: centres U[0,100]^3, edge 0.2, seed 42", "ray_seed": 7,
            "l2": "inputs (rays 32 B + hits 16 B per ray = %.1f GB) exceed the 126 MB L2; no explicit flush" % (a.rays * 48 / 1e9)}


def gen_rays_device(torch, n, lo, hi, seed, device):
    """Same recipe as scenes.random_rays (origins uniform in the AABB, directions uniform on the sphere), on the GPU."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    rays = torch.empty((n, 8), dtype=torch.float32, device=device)
    lo_t = torch.tensor(lo, device=device)
    hi_t = torch.tensor(hi, device=device)
    chunk = 1 << 24
    for b in range(0, n, chunk):
        m = min(chunk, n - b)
        u = torch.rand((m, 5), generator=g, device=device, dtype=torch.float32)
        r = rays[b:b + m]
        r[:, 0:3] = lo_t + u[:, 0:3] * (hi_t - lo_t)
        z = 1 - 2 * u[:, 3]
        rad = torch.sqrt(torch.clamp(1 - z * z, min=0))
        phi = 2 * np.pi * u[:, 4]
        r[:, 4] = rad * torch.cos(phi)
        r[:, 5] = rad * torch.sin(phi)
        r[:, 6] = z
        r[:, 3] = 1e-4
        r[:, 7] = 3.4028234663852886e38
    return rays


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.strip().split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for nme, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_baseline(verts, rays_host, n):
    """The C oracle (kind "port") on a bounded ray sample, all host threads (OpenMP)."""
    from oracle import bindings as ob
    P = ob.PortScene(verts)
    sample = np.ascontiguousarray(rays_host[:n])
    t0 = time.perf_counter()
    tuv, tri = P.closest(sample)
    dt = time.perf_counter() - t0
    return {"value": n / dt / 1e6, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": f"first {n} rays of the batch on the same 4M-tri scene, oracle/lm_oracle.c (OpenMP), {dt:.1f} s"}, tuv, tri


def run_reference(a):
    """--impl reference: the reference's own accel::qbvh through Accel3::Intersect on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import bindings as ob
    from lmb200py import scenes
    cfg = workload_config(a)
    cores = os.cpu_count() or 1
    verts = scenes.soup(a.tris, seed=42, extent=100.0, edge=0.2)
    lo, hi = scenes.bounds(verts)
    n = max(50_000, a.cpu_rays // 4)
    if ob.have_ref():
        t0 = time.perf_counter()
        R = ob.RefSoup(verts, "qbvh")
        build_s = time.perf_counter() - t0
        kind = "reference"

        # Two CPU tracers of the reference are timed on the same rays every step: accel::qbvh (the fastest in-tree accel whose
        # results are the parity oracle) through the real Accel3::Intersect, and nanort (the vendored third-party tracer the north
        # star names next to Embree; called as accel::nanort calls it — a SPEED figure only, its wrapper is broken in the
        # reference and its hits differ from TriAccel's). The headline `value` is the FASTER of the two. Embree 2.8.0 is an
        # un-vendored dependency and cannot be installed offline.
        try:
            NR = ob.RefNanort(verts)
        except Exception as e:      # noqa: BLE001
            NR = None
            print(f"nanort baseline unavailable: {e}", file=sys.stderr)
        nano_times = []

        def step(k):
            rays = scenes.random_rays(n, lo, hi, seed=7 + k)
            r = R.intersect(rays, threads=cores)
            if NR is not None:
                nano_times.append(NR.trace(rays, threads=cores, want_hits=False)["seconds"])
            return r["seconds"]
        sample = f"{n} rays per step, accel::qbvh via Accel3::Intersect and nanort::BVHAccel::Traverse (oracle/_ref), builds ({build_s:.0f} s / {NR.build_seconds if NR else 0:.0f} s) not timed"
        others = {"embree": "n/a (not installable offline)"}
    else:
        P = ob.PortScene(verts)
        kind = "port"

        def step(k):
            rays = scenes.random_rays(n, lo, hi, seed=7 + k)
            t0 = time.perf_counter()
            P.closest(rays)
            return time.perf_counter() - t0
        others = {}
        sample = f"{n} rays per step, oracle/lm_oracle.c (OpenMP); oracle/_ref not present"
    for k in range(a.warmup):
        step(k)
    times = [step(a.warmup + k) for k in range(a.steps)]
    total = sum(times)
    value = n * a.steps / total / 1e6
    if kind == "reference":
        others["qbvh_mrays_s"] = value
        others["timed"] = "accel::qbvh"
        if len(nano_times) >= a.steps:
            nt = sum(nano_times[-a.steps:])
            others["nanort_mrays_s"] = n * a.steps / nt / 1e6
            if others["nanort_mrays_s"] > value:      # report the stronger CPU baseline
                value, total = others["nanort_mrays_s"], nt
                others["timed"] = "nanort (faster than accel::qbvh on this box)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg,
            "cpu_baseline": dict({"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}, **others),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if not a.no_pt:
        # the other half of the metric on the reference's side: renderer::ptdirect on the configs[2] scene, bounded sample
        from lmb200py import scenedesc
        pt = pt_cpu_baseline(scenedesc.config2_scene(a.pt_tris, 1920, 1080), 1920, 1080, a.pt_cpu_spp)
        line["path_tracing"] = dict({"metric": "Msamples/s ptdirect (NEE), 1920x1080, 1M-tri synthetic scene"}, **pt)
        line["pt_msamples_s"] = pt["value"]
    emit(line)


_JSON_OUT = None


def emit(line):
    """The one JSON line, on the process's ORIGINAL stdout."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # stdout carries exactly one JSON line: keep the real stdout aside and point fd 1 at stderr, so that whatever
    # libraries write to stdout (NCCL prints its version banner there) cannot end up next to it
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    a = parse()
    if a.impl == "reference":
        run_reference(a)
        return
    import torch
    import torch.distributed as dist
    from lmb200py import capi, scenes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: lmb200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(torch, local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = capi.lib()
    launches0 = L.lmb200_launch_count()

    # ---- scene + BVH (replicated per rank) ----
    verts = scenes.soup(a.tris, seed=42, extent=100.0, edge=0.2)
    lo, hi = scenes.bounds(verts)
    accel = capi.Accel(local)
    bst = accel.build(verts)

    # ---- rays: generated on the device, mirrored once into pinned host memory for the e2e leg ----
    d_rays = gen_rays_device(torch, a.rays, lo.tolist(), hi.tolist(), 7 + rank, dev)
    d_hits = torch.empty((a.rays, 4), dtype=torch.float32, device=dev)
    h_rays = torch.empty((a.rays, 8), dtype=torch.float32, pin_memory=True)
    h_rays.copy_(d_rays)
    h_hits = torch.empty((a.rays, 4), dtype=torch.float32, pin_memory=True)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream

    def step_dev():
        capi.check(L.lmb200_trace_closest_dev(accel.h, d_rays.data_ptr(), d_hits.data_ptr(), a.rays, stream))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        step_dev()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    ev[0].record()
    for k in range(a.steps):
        step_dev()
        ev[k + 1].record()
    barrier()
    clock_info = clocks.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(a.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    value = a.rays * world * a.steps / (max_ms * 1e-3) / 1e6

    # ---- e2e: host-buffer C-ABI call, pinned rays in / hits out, copies inside the timed region ----
    # The workload's rays all carry the range Scene3::Intersect passes, (1e-4, FLT_MAX) (scene3.cpp:461), so the call a user
    # makes is the compact wire form: 24-byte rays + one [tmin, tmax] (Accel3::Intersect's own argument shape). The 32-byte
    # per-ray-range form is timed next to it.
    h_rays24 = torch.empty((a.rays, 6), dtype=torch.float32, pin_memory=True)
    h_rays24.copy_(h_rays[:, [0, 1, 2, 4, 5, 6]])

    def e2e_leg(fn):
        fn()
        barrier()
        t0 = time.perf_counter()
        steps = max(1, min(a.steps, 3))
        for _ in range(steps):
            fn()
        barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return a.rays * world * steps / float(t.item()) / 1e6
    e2e_full = e2e_leg(lambda: capi.check(L.lmb200_trace_closest(accel.h, h_rays.data_ptr(), h_hits.data_ptr(), a.rays)))
    full_hits = h_hits[:4096].clone()
    e2e_value = e2e_leg(lambda: capi.check(L.lmb200_trace_closest_compact(accel.h, h_rays24.data_ptr(), 1e-4, 3.4028234663852886e38, h_hits.data_ptr(), a.rays)))
    assert torch.equal(full_hits.view(torch.int32), h_hits[:4096].view(torch.int32)), "compact and full wire forms disagree"
    # every hit of the host-buffer call against the device-resident launch of the same rays, bit for bit (d_hits holds the last timed step)
    e2e_same = True
    for lo_i in range(0, a.rays, 1 << 24):
        hi_i = min(a.rays, lo_i + (1 << 24))
        e2e_same = e2e_same and bool(torch.equal(h_hits[lo_i:hi_i].to(dev, non_blocking=False).view(torch.int32), d_hits[lo_i:hi_i].view(torch.int32)))

    # ---- roofline of the traversal kernel ----
    npr, tpr = C.c_double(), C.c_double()
    ncount = min(a.rays, 1 << 22)
    capi.check(L.lmb200_trace_count_dev(accel.h, d_rays.data_ptr(), ncount, C.byref(npr), C.byref(tpr)))
    b_ray = 48.0 + npr.value * NODE_BYTES + tpr.value * 48.0
    mean_kernel_s = float(np.mean(kernel_ms)) * 1e-3
    achieved = b_ray * a.rays / mean_kernel_s / 1e9
    peak, peak_src = hbm_peak()
    # DRAM traffic per launch comes from one kept ncu capture (profiles/traffic.json); kept_traffic_per_ray says when it
    # still describes the shipped kernel
    per_ray_kept, traffic_note = kept_traffic_per_ray()
    traffic = per_ray_kept * a.rays if per_ray_kept is not None else None      # per launch, like `achieved`
    ceiling = pcie_ceiling(torch, dist, world, dev, h_rays, d_rays, h_hits, d_hits)

    # ---- parity spot check + CPU baseline (rank 0, N=1) ----
    cpu = None
    parity = None
    if rank == 0:
        os.sched_setaffinity(0, all_cpus)      # the CPU baseline gets every host core (its OpenMP team starts here)
        n_chk = min(a.cpu_rays, a.rays) if world == 1 else min(100_000, a.rays)
        step_dev()
        torch.cuda.synchronize()
        hits = d_hits[:n_chk].cpu().numpy()
        base, tuv, tri = cpu_port_baseline(verts, h_rays.numpy(), n_chk)
        gtri = hits[:, 3].view(np.uint32).astype(np.int64)
        gtri[gtri == capi.MISS] = -1
        parity = {"rays_checked": int(n_chk), "index_mismatches": int(np.count_nonzero(gtri != tri)),
                  "tuv_bit_mismatches": int(np.count_nonzero(hits[:, :3].view(np.uint32) != tuv.view(np.uint32)))}
        if world == 1:
            cpu = base

    # ---- the per-ray drop-in: Accel3::Intersect-shaped calls from 16 host threads (persistent service kernel) ----
    one = None
    if rank == 0 and world == 1 and not a.no_one:
        one = bench_intersect_one(a, capi, accel, h_rays)

    # ---- secondary: path-traced samples/s on configs[2] (full size by default), NCCL film reduce ----
    pt = None
    if not a.no_pt:
        pt = bench_pt(a, torch, dist, capi, world, rank, local, dev)
    c4 = None
    if not a.no_c4:
        del d_rays, d_hits, h_rays, h_hits, h_rays24
        torch.cuda.empty_cache()
        c4 = bench_config4(a, torch, dist, capi, world, rank, local, dev)

    launches = int(L.lmb200_launch_count() - launches0)
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
                "ms_per_step": max_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": workload_config(a),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": a.rays * 24, "d2h_bytes_per_step": a.rays * 16,
                        "api": "lmb200_trace_closest_compact(host pinned 24-byte rays + shared [tmin, tmax] -> host pinned hits)", "host_affinity": numa,
                        "gbs": e2e_value * 1e6 * 40 / 1e9, "pcie_ceiling_gbs": ceiling, "frac_of_pcie_ceiling": e2e_value * 1e6 * 40 / 1e9 / ceiling,
                        "frac_of_device_rate": e2e_value / value,
                        "all_hits_equal_device_launch": e2e_same,
                        "launch": "one persistent launch per call, rays taken chunk by chunk as their uploads land (LMB200_E2E_STREAM=0: one launch per chunk)",
                        "pcie_ceiling_how": "concurrent pinned H2D + D2H torch copies (2:1 bytes) of the bench's own buffers, all ranks at once; the call is bound by min(this ceiling, the kernel rate)",
                        "full_ray_form": {"value": e2e_full, "unit": UNIT, "h2d_bytes_per_step": a.rays * 32, "d2h_bytes_per_step": a.rays * 16,
                                          "api": "lmb200_trace_closest(32-byte rays with per-ray range)", "gbs": e2e_full * 1e6 * 48 / 1e9}},
                "gpu_launches": launches, "clocks": clock_info,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
                             # the north star quotes ~8 TB/s (HBM3e spec); the measured copy peak above is the binding denominator
                             "frac_of_spec_8tbs": achieved / 8000.0,
                             "frac_dram_actual": (traffic / mean_kernel_s / 1e9 / peak) if traffic else None,
                             "peak_source": peak_src, "bytes_per_ray": b_ray, "nodes_per_ray": npr.value, "tris_per_ray": tpr.value,
                             "kernel": "lmb200::trace_kernel<false,false>", "kernel_ms": mean_kernel_s * 1e3,
                             # second ceiling, informational: node fetches/s against the rate at which a fetch-only kernel reads random
                             # 64-byte aligned records with two 256-bit loads on a B200 (112.8 G/s, scripts/micro/l1_wavefront.cu pattern E,
                             # profiles/r02_sweep.md; the 80-byte node of round 1: 81.4 G/s)
                             "node_fetches_per_s": npr.value * a.rays / mean_kernel_s, "record_fetch_ceiling_per_s": 112.8e9,
                             "frac_of_fetch_ceiling": npr.value * a.rays / mean_kernel_s / 112.8e9},
                "cpu_baseline": cpu, "parity": parity,
                "bvh": {"builder": "device LBVH (library default)", "nodes": bst["num_nodes"], "node_bytes": bst["node_bytes"], "tri_bytes": bst["tri_bytes"], "build_s": bst["build_seconds"]},
                "intersect_one": one, "intersect_one_mrays_s": one["value"] if one else None,
                "path_tracing": pt, "config4": c4,
                # the two secondary figures once more as flat keys (nested objects may be dropped by a summariser)
                "pt_msamples_s": pt["value"] if pt else None, "pt_e2e_msamples_s": pt["e2e"]["value"] if pt else None,
                "pt_roofline_frac": pt["roofline"]["frac"] if pt else None,
                "pt_cpu_msamples_s": pt["cpu_baseline"]["value"] if pt and pt.get("cpu_baseline") else None,
                "incoherent_1m_mesh_mrays_s": pt["incoherent_1m_tri_mesh"]["value"] if pt else None,
                "config4_msamples_s": c4["value"] if c4 else None, "film_reduce_ms": c4["film_reduce_ms"] if c4 else None}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_intersect_one(a, capi, accel, h_rays, threads=16, n=200_000):
    """lmb200_trace_closest_one from `threads` host threads (each a loop of synchronous single-ray calls: how Scheduler_::Process
    drives Accel3::Intersect, scheduler.cpp:146-175) on the bench scene; the CPU figure beside it is the C port of accel::qbvh's
    traversal on the same rays and thread count (the reference's own accel::qbvh is timed by --impl reference)."""
    L = capi.lib()
    rays = np.ascontiguousarray(h_rays.numpy()[:n])
    hits = np.zeros(n, capi.HIT_DTYPE)
    sec = C.c_double()
    capi.check(L.lmb200_trace_closest_one_mt(accel.h, rays.ctypes.data, hits.ctypes.data, min(n, 20000), threads, C.byref(sec)))   # warm-up, starts the service
    capi.check(L.lmb200_trace_closest_one_mt(accel.h, rays.ctypes.data, hits.ctypes.data, n, threads, C.byref(sec)))
    capi.check(L.lmb200_trace_closest_one_mt(accel.h, rays.ctypes.data, hits.ctypes.data, 2000, 1, C.byref(s1 := C.c_double())))
    return {"value": n / sec.value / 1e6, "unit": "Mrays/s", "host_threads": threads, "rays": n,
            "latency_us_one_thread": s1.value / 2000 * 1e6, "api": "lmb200_trace_closest_one (persistent service kernel, mapped pinned mailboxes)",
            "scene": "the bench scene (4M-triangle soup, incoherent rays)"}


def bind_to_gpu_numa_node(torch, index):
    """Run this rank (and first-touch its pinned host buffers) on the CPUs local to its GPU's PCIe root: with 8 ranks on a
    two-socket box the host side of the e2e copies otherwise crosses the socket interconnect. Best effort; returns a note."""
    try:
        pr = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        txt = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        cpus = set()
        for part in txt.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "unchanged (no local CPUs in the allowed set)"
        os.sched_setaffinity(0, cpus)
        return f"cpus {txt} (local to GPU {bdf})"
    except Exception as e:      # noqa: BLE001
        return f"unchanged ({type(e).__name__})"


def hbm_peak():
    """(GB/s, source): the measured HBM copy bandwidth of this pool's B200s if the driver wrote it, else the recipe's fallback."""
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def incoherent_on_scene(a, torch, dist, capi, S, verts, world, dev, stream):
    """The north star's target figure: incoherent closest-hit Mrays/s per GPU on the 1M-triangle (mesh) scene — 16 Mi random
    rays (origins uniform in the scene AABB, directions uniform on the sphere) against the BVH the path tracer just used."""
    from lmb200py import scenes
    L = capi.lib()
    acc = L.lmb200_scene_accel(S.h_)
    lo, hi = scenes.bounds(verts)
    n = 1 << 24
    rays = gen_rays_device(torch, n, [float(x) for x in lo], [float(x) for x in hi], 7, dev)
    hits = torch.empty((n, 4), dtype=torch.float32, device=dev)
    for _ in range(3):
        capi.check(L.lmb200_trace_closest_dev(acc, rays.data_ptr(), hits.data_ptr(), n, stream))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        capi.check(L.lmb200_trace_closest_dev(acc, rays.data_ptr(), hits.data_ptr(), n, stream))
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    npr, tpr = C.c_double(), C.c_double()
    capi.check(L.lmb200_trace_count_dev(acc, rays.data_ptr(), 1 << 22, C.byref(npr), C.byref(tpr)))
    bpr = 48 + NODE_BYTES * npr.value + 48 * tpr.value
    rate = n / (float(ms.item()) * 1e-3)
    peak = hbm_peak()[0]
    return {"value": world * rate / 1e6, "unit": "Mrays/s", "per_gpu": rate / 1e6, "rays_per_gpu": n, "triangles": int(len(verts)),
            "nodes_per_ray": npr.value, "tris_per_ray": tpr.value, "bytes_per_ray": bpr,
            "roofline_frac": rate * bpr / 1e9 / peak, "hit_fraction": float((hits[:, 3].view(torch.int32) != -1).float().mean().item()),
            "target": ">= 1000 Mrays/s per GPU"}


S_STATE_BYTES = 113      # per-slot path state a vertex step reads and writes back (render.cu Pool: sample 8, nverts 4, thr 16,
                         # ray_o 16, ray_d 16, hit 16, traced 1, vtx_p 16, vtx_wi 16, vtx_v 4), ptdirect


def pt_roofline(st, samples, seconds, world):
    """B_sample = sum over extend rays of B_ray + sum over shadow rays of B_ray(any-hit) + V (2 S_state) + splats 16,
    with nodes / records per ray counted by the instrumented kernels over the actual extend and shadow queues."""
    er, sr = max(st.extend_rays, 1), max(st.shadow_rays, 1)
    b_ext = 32 + 16 + NODE_BYTES * st.extend_nodes / er + 48.0 * st.extend_tris / er
    b_sh = 32 + 16 + NODE_BYTES * st.shadow_nodes / sr + 48.0 * st.shadow_tris / sr      # ray + (contribution, pixel)
    n = max(st.samples, 1)
    b_sample = st.extend_rays / n * b_ext + st.shadow_rays / n * b_sh + st.vertices / n * 2 * S_STATE_BYTES + st.shadow_rays / n * 16
    peak, src = hbm_peak()
    achieved = b_sample * samples / seconds / 1e9 / world        # per GPU
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "frac_of_spec_8tbs": achieved / 8000.0, "traffic": None, "peak_source": src,
            "bytes_per_sample": b_sample, "vertices_per_sample": st.vertices / n, "extend_rays_per_sample": st.extend_rays / n,
            "shadow_rays_per_sample": st.shadow_rays / n, "extend": {"nodes_per_ray": st.extend_nodes / er, "tris_per_ray": st.extend_tris / er, "bytes_per_ray": b_ext},
            "shadow": {"nodes_per_ray": st.shadow_nodes / sr, "tris_per_ray": st.shadow_tris / sr, "bytes_per_ray": b_sh},
            "state_bytes_per_vertex": 2 * S_STATE_BYTES, "kernel": "lmb200::k_extend (dominant, ~60 % of the frame) + k_shadow + shading kernels",
            "counted_on": "a %d-sample instrumented run of the same scene (count_work)" % st.samples}


def pt_cpu_baseline(sc, W, H, spp, with_drop_in=False):
    """The reference's own renderer::ptdirect (oracle/_ref: real Scene3 + accel::qbvh + Scheduler shim on all host threads) on a
    bounded sample of the same scene; the C port when the compiled reference is not present."""
    from oracle import bindings as ob
    cores = os.cpu_count() or 1
    N = W * H * spp
    if ob.have_ref():
        t0 = time.perf_counter()
        R = ob.RefScene(sc, accel="qbvh")
        build_s = time.perf_counter() - t0
        img, sec = R.render("ptdirect", N, seed=1, threads=cores)
        out = {"value": N / sec / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "reference",
               "sample": f"{spp} spp of the same {W}x{H} scene ({N} samples, {sec:.1f} s), renderer::ptdirect + accel::qbvh from oracle/_ref, scene build {build_s:.0f} s not timed",
               "mean_rgb": [float(x) for x in img.mean(axis=(0, 1))]}
        del R
        if with_drop_in:
            # the accel drop-in alone: the reference's OWN renderer::ptdirect (its CPU threads) with every Accel3::Intersect
            # answered by accel::lmb200's per-ray service on the GPU
            try:
                plug = os.path.join(ROOT, "lightmetrica-v2_b200", "plugin", "accel_lmb200")
                if os.path.exists(plug + ".so") and ob.ref().ref_load_plugin(plug.encode()) == 1:
                    R2 = ob.RefScene(sc, accel="lmb200")
                    img2, sec2 = R2.render("ptdirect", N, seed=1, threads=cores)
                    out["reference_renderer_on_accel_lmb200"] = {"value": N / sec2 / 1e6, "unit": "Msamples/s", "threads": cores,
                                                                  "what": "renderer::ptdirect (reference, CPU threads) + accel::lmb200 (per-ray GPU service), same samples",
                                                                  "mean_rgb": [float(x) for x in img2.mean(axis=(0, 1))]}
            except Exception as e:      # noqa: BLE001
                out["reference_renderer_on_accel_lmb200"] = {"unavailable": str(e)[:200]}
        return out
    P = ob.PortPT(sc)
    N = W * H // 4
    t0 = time.perf_counter()
    img, _ = P.render(1, N, seed=1)
    sec = time.perf_counter() - t0
    return {"value": N / sec / 1e6, "unit": "Msamples/s", "cores": 1, "kind": "port",
            "sample": f"{N} samples of the same scene, oracle/lm_oracle_pt.c (scalar), {sec:.1f} s"}


def render_sharded(torch, dist, capi, S, film, W, H, N, rank, world, stream, pool=0, count=False, time_reduce=False):
    """One frame through the multi-process path: every rank renders its share of the sample range into its own unscaled film,
    one NCCL reduce to rank 0 (replaces contexts.combine_each(film->Accumulate), scheduler.cpp:280-285), rescale by W H / N."""
    from lmb200py import distributed
    L = capi.lib()
    b, e = distributed.shard_range(N, rank, world)
    st = capi.RenderStats()
    film.zero_()
    p = S.params(capi.MODE_PTDIRECT, N, seed=1, begin=b, end=e, pool=pool)
    p.count_work = 1 if count else 0
    capi.check(L.lmb200_render_dev(S.h_, C.byref(p), film.data_ptr(), stream, C.byref(st)))
    r0 = r1 = None
    if time_reduce and world > 1:
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
    distributed.reduce_film(film, dist if world > 1 else None)
    if r0 is not None:
        r1.record()
    capi.check(L.lmb200_film_rescale_dev(film.data_ptr(), W * H, distributed.film_scale(W, H, N), stream))
    return st, (r0, r1)


def bench_pt(a, torch, dist, capi, world, rank, local, dev):
    """ptdirect on the 1M-triangle mesh scene (BASELINE configs[2] geometry/materials/lights) at pt_spp samples per pixel (1024 = the config),
    sample range sharded contiguously over ranks, films summed with one NCCL reduce, then rescaled."""
    from lmb200py import scenedesc
    W, H = 1920, 1080
    sc = scenedesc.config2_scene(a.pt_tris, W, H)
    S = capi.Scene(sc, device=local)
    _k = S.keep
    N = W * H * a.pt_spp
    film = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
    h_film = torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True)
    L = capi.lib()
    stream = torch.cuda.current_stream().cuda_stream

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    st, _ = render_sharded(torch, dist, capi, S, film, W, H, N, rank, world, stream, a.pt_pool)      # warm-up frame
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 2
    for _ in range(reps):
        st, _ = render_sharded(torch, dist, capi, S, film, W, H, N, rank, world, stream, a.pt_pool)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    rays = torch.tensor([st.extend_rays + st.shadow_rays], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(rays, op=dist.ReduceOp.SUM)
    sec = float(ms.item()) * 1e-3
    mean_rgb = [float(x) for x in film[..., :3].mean(dim=(0, 1)).tolist()] if rank == 0 else None

    # ---- end to end through the public call: host film out. N = 1: lmb200_render (the call renderer::lmb200pt makes);
    # N > 1: per-rank lmb200_render_dev + NCCL reduce + rank 0 copies the film to pinned host memory. Wall clock.
    sync_all()
    t0 = time.perf_counter()
    if world == 1:
        p = S.params(capi.MODE_PTDIRECT, N, seed=1, pool=a.pt_pool)
        st1 = capi.RenderStats()
        capi.check(L.lmb200_render(S.h_, C.byref(p), h_film.data_ptr(), C.byref(st1)))
    else:
        render_sharded(torch, dist, capi, S, film, W, H, N, rank, world, stream, a.pt_pool)
        if rank == 0:
            h_film.copy_(film, non_blocking=True)
    sync_all()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e = {"value": N / float(e2e_s.item()) / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": C.sizeof(capi.RenderParams),
           "d2h_bytes_per_step": W * H * 16,
           "api": "lmb200_render(params -> host film)" if world == 1 else "lmb200_render_dev per rank + NCCL film reduce + rank-0 film to pinned host memory"}

    # ---- roofline: algorithmic bytes per sample from an instrumented (untimed) run at 16 spp
    Nc = W * H * min(16, a.pt_spp)
    stc, _ = render_sharded(torch, dist, capi, S, film, W, H, Nc, 0, 1, stream, a.pt_pool, count=True)
    torch.cuda.synchronize()
    roof = pt_roofline(stc, N, sec, world)

    target = incoherent_on_scene(a, torch, dist, capi, S, _k["verts"], world, dev, stream)
    S.close()
    cpu = None
    if rank == 0 and world == 1:
        cpu = pt_cpu_baseline(sc, W, H, a.pt_cpu_spp, with_drop_in=True)
    return {"incoherent_1m_tri_mesh": target,
            "metric": "Msamples/s ptdirect (NEE), 1920x1080, 1M-tri synthetic scene", "value": N / sec / 1e6,
            "unit": "Msamples/s", "spp": a.pt_spp, "samples": N, "ms": float(ms.item()), "rays_per_sample": float(rays.item()) / N,
            "mrays_per_s": float(rays.item()) / sec / 1e6, "film_reduce": "torch.distributed NCCL reduce" if world > 1 else "none (1 GPU)", "scaling": "strong (fixed image and spp, sample range sharded over ranks)",
            "mean_rgb": mean_rgb, "e2e": e2e, "roofline": roof, "cpu_baseline": cpu}


def bench_config4(a, torch, dist, capi, world, rank, local, dev):
    """BASELINE configs[4]: 10M-triangle instanced scene, 3840x2160 film, ptdirect at c4_spp samples per pixel (the config names 4096;
    throughput is per sample), sample range sharded over ranks, the 133 MB film summed with ONE NCCL reduce timed on its own."""
    from lmb200py import scenedesc
    W, H = 3840, 2160
    sc, verts = scenedesc.config4_scene(W, H)
    t0 = time.perf_counter()
    S = capi.Scene(sc, device=local, builder=capi.BUILD_GPU_LBVH)
    create_s = time.perf_counter() - t0
    ast = capi.AccelStats()
    L = capi.lib()
    capi.check(L.lmb200_accel_get_stats(L.lmb200_scene_accel(S.h_), C.byref(ast)))
    N = W * H * a.c4_spp
    film = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    render_sharded(torch, dist, capi, S, film, W, H, N // 8, rank, world, stream)      # warm-up (also warms the NCCL reduce)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st, (r0, r1) = render_sharded(torch, dist, capi, S, film, W, H, N, rank, world, stream, time_reduce=True)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1), r0.elapsed_time(r1) if r0 is not None else 0.0, st.seconds * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, red_ms, render_ms = [float(x) for x in t.tolist()]
    ok = bool(torch.isfinite(film).all().item()) if rank == 0 else None
    mean_rgb = [float(x) for x in film[..., :3].mean(dim=(0, 1)).tolist()] if rank == 0 else None
    S.close()
    return {"metric": "Msamples/s ptdirect, 10M-tri instanced scene, 3840x2160", "value": N / (ms * 1e-3) / 1e6, "unit": "Msamples/s",
            "triangles": int(len(verts)) + 2, "spp": a.c4_spp, "samples": N, "ms": ms, "render_ms": render_ms,
            "film_reduce_ms": red_ms if world > 1 else None, "film_bytes": W * H * 16,
            "film_reduce_gbs": (W * H * 16 / (red_ms * 1e-3) / 1e9) if (world > 1 and red_ms > 0) else None,
            "note": "film_reduce_ms is the max over ranks of the device time of the one NCCL reduce; it includes waiting for the slowest rank's render",
            "bvh": {"builder": "device LBVH", "nodes": ast.num_nodes, "node_bytes": ast.node_bytes, "tri_bytes": ast.tri_bytes,
                    "build_s": ast.build_seconds, "scene_create_s": create_s},
            "scaling": "strong", "finite": ok, "mean_rgb": mean_rgb}


def pcie_ceiling(torch, dist, world, dev, h_rays, d_rays, h_hits, d_hits):
    """What the box's host<->device path can move at all: the e2e leg's own pinned buffers copied with plain torch copies,
    H2D (rays) and D2H (hits) concurrently on two streams, all ranks at once. GB/s summed over ranks."""
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    n = min(h_rays.shape[0], 1 << 25)

    def once():
        with torch.cuda.stream(s1):
            d_rays[:n].copy_(h_rays[:n], non_blocking=True)
        with torch.cuda.stream(s2):
            h_hits[:n].copy_(d_hits[:n], non_blocking=True)
    once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return 3 * n * 48 * world / float(dt.item()) / 1e9


def source_hash():
    """Identifies the traversal kernel a kept ncu figure (profiles/traffic.json) was captured on."""
    import hashlib
    h = hashlib.sha256()
    for f in ("traverse.cuh", "accel.cu", "bvh.h", "triaccel.h", "bvh_build.cpp"):
        h.update(open(os.path.join(ROOT, "lightmetrica-v2_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def layout_hash():
    """Hash of what decides WHICH bytes the traversal kernel fetches: the unit layout and every builder."""
    import hashlib
    h = hashlib.sha256()
    for f in ("bvh.h", "bvh_dev.h", "triaccel.h", "bvh_build.cpp", "bvh_build_gpu.cu"):
        h.update(open(os.path.join(ROOT, "lightmetrica-v2_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


TRACE_KERNEL_SYMBOL = "_ZN6lmb20012trace_kernelILb0ELb0ELb0EEEvNS_6BvhDevEPK6float4PvmPKjPyS8_ff"
_MEM_OP = re.compile(r"\b(?:LDG|STG|LDS|STS|LDL|STL|ATOMG|ATOMS|ATOM|RED|LDC|LDCU|LD|ST|LDSM|MEMBAR|CCTL)(?:\.[A-Za-z0-9_]+)*")


def mem_signature(binary=None):
    """Hash of the ordered list of memory instructions (opcode and modifiers: width, cache operator, scope) in the shipped
    machine code of lmb200::trace_kernel<false,false>, read with cuobjdump. Two builds with the same signature issue the same
    loads, stores and atomics in the same order; None when cuobjdump or the library is not there."""
    import hashlib
    import shutil
    import subprocess
    binary = binary or os.path.join(ROOT, "lightmetrica-v2_b200", "lib", "liblmb200.so")
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool) or not os.path.exists(binary):
        return None
    try:
        out = subprocess.run([tool, "-sass", "-fun", TRACE_KERNEL_SYMBOL, binary], capture_output=True, text=True, timeout=120).stdout
    except Exception:
        return None
    ops = []
    for line in out.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m:
            ops += _MEM_OP.findall(m.group(1))
    if not ops:
        return None
    return hashlib.sha256(("\n".join(ops) + "\n").encode()).hexdigest()[:16]


def kept_traffic_per_ray():
    """(DRAM bytes per ray, note) from the kept ncu capture profiles/traffic.json, or (None, why not). The figure is reported
    while the traversal sources hash to what it was captured on, or - after a change that left trace_kernel<false,false>
    alone, like the streaming gate that is compiled out of it - while the unit layout and builders are unchanged AND the
    kernel's machine code still issues the same memory instructions in the same order."""
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tpath):
        return None, "no profiles/traffic.json"
    try:
        tj = json.load(open(tpath))
        per_ray = float(tj["dram_bytes_per_ray"])
    except Exception as ex:
        return None, "profiles/traffic.json unreadable: %s" % ex
    what = "ncu dram__bytes_read.sum + dram__bytes_write.sum per ray x rays per launch"
    if tj.get("source_hash") == source_hash():
        return per_ray, what + ", captured on source hash " + tj["source_hash"]
    sig = mem_signature()
    if tj.get("layout_hash") == layout_hash() and sig is not None and tj.get("mem_signature") == sig:
        return per_ray, (what + ", captured on source hash %s; the sources now hash to %s, but unit layout and builders are unchanged "
                         "(layout hash %s) and the shipped trace_kernel<false,false> issues the same memory instructions in the same "
                         "order as the captured build (signature %s)" % (tj.get("source_hash"), source_hash(), layout_hash(), sig))
    return None, ("stale: profiles/traffic.json (%.0f B/ray) was captured on source hash %s / layout %s / memory signature %s; now %s / %s / %s"
                  % (per_ray, tj.get("source_hash"), tj.get("layout_hash"), tj.get("mem_signature"), source_hash(), layout_hash(), sig))


if __name__ == "__main__":
    main()
