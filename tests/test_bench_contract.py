"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm prints exactly one JSON line on
stdout with the keys the driver reads, and non-zero ranks of a multi-rank launch stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--tris", "3000", "--steps", "1", "--warmup", "1",
                        "--cpu-rays", "40000", "--pt-tris", "20000", "--pt-cpu-spp", "1"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def test_reference_arm_prints_one_json_line():
    out = run()
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "Mrays/s"
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("configs[3]") and d["vs_baseline"] is None
    # the path-tracing half of the metric on the reference's side
    pt = d["path_tracing"]
    assert pt["unit"] == "Msamples/s" and pt["value"] > 0 and pt["kind"] in ("reference", "port") and d["pt_msamples_s"] == pt["value"]


def test_reference_arm_non_zero_rank_is_silent():
    assert run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}).strip() == ""


import pytest  # noqa: E402


@pytest.mark.gpu
def test_lmb200_arm_json_contract_small():
    """The product arm on a reduced workload: one JSON line with every key of the bench contract, parity spot check clean."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--steps", "2", "--warmup", "3", "--rays", "2097152",
                        "--tris", "200000", "--cpu-rays", "50000", "--pt-tris", "50000", "--pt-spp", "2", "--pt-cpu-spp", "1", "--c4-spp", "1"],
                       capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout[:2000]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "parity", "path_tracing"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] >= 3 and d["scaling"] == "weak" and d["dtype"] == "f32"
    assert d["value"] > 0 and 0 < d["e2e"]["value"] <= d["value"] * 1.05
    assert d["e2e"]["h2d_bytes_per_step"] == 2097152 * 24 and d["e2e"]["d2h_bytes_per_step"] == 2097152 * 16
    assert d["e2e"]["full_ray_form"]["h2d_bytes_per_step"] == 2097152 * 32 and d["e2e"]["full_ray_form"]["value"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["gpu_launches"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["parity"]["index_mismatches"] == 0 and d["parity"]["tuv_bit_mismatches"] == 0
    assert d["path_tracing"]["value"] > 0 and d["path_tracing"]["incoherent_1m_tri_mesh"]["value"] > 0
    pt = d["path_tracing"]
    assert 0 < pt["e2e"]["value"] <= pt["value"] * 1.05 and pt["e2e"]["d2h_bytes_per_step"] == 1920 * 1080 * 16
    assert pt["roofline"]["bound"] == "hbm" and 0 < pt["roofline"]["frac"] < 1.5 and pt["roofline"]["vertices_per_sample"] > 1
    assert pt["cpu_baseline"]["value"] > 0 and pt["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["pcie_ceiling_gbs"] > 0 and 0 < d["e2e"]["frac_of_pcie_ceiling"] < 1.3
    c4 = d["config4"]
    assert c4["value"] > 0 and c4["triangles"] > 9_500_000 and c4["film_bytes"] == 3840 * 2160 * 16 and c4["finite"] is True
    assert d["pt_msamples_s"] == pt["value"] and d["config4_msamples_s"] == c4["value"]


def test_kept_dram_traffic_rule():
    """roofline.traffic comes from a kept ncu capture: it is reported only while the capture still describes the shipped
    kernel (same sources, or same unit layout + builders and the same memory instructions in the kernel's machine code)."""
    import shutil
    sys.path.insert(0, ROOT)
    import bench
    tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for k in ("dram_bytes_per_ray", "source_hash", "layout_hash", "mem_signature"):
        assert k in tj, k
    per_ray, note = bench.kept_traffic_per_ray()
    lib = os.path.join(ROOT, "lightmetrica-v2_b200", "lib", "liblmb200.so")
    have_tool = os.path.exists(lib) and (shutil.which("cuobjdump") or os.path.exists("/usr/local/cuda/bin/cuobjdump"))
    fresh = tj["source_hash"] == bench.source_hash()
    same_kernel = have_tool and tj["layout_hash"] == bench.layout_hash() and tj["mem_signature"] == bench.mem_signature()
    if fresh or same_kernel:
        assert per_ray == tj["dram_bytes_per_ray"] and tj["source_hash"] in note
    else:
        assert per_ray is None and note.startswith("stale")
    if have_tool:
        sig = bench.mem_signature()
        assert sig is not None and len(sig) == 16 and sig == bench.mem_signature()      # deterministic
    assert bench.mem_signature("/nonexistent/lib.so") is None
