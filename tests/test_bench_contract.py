"""CPU checks of bench.py's output contract (no GPU): the roofline / cpu_baseline objects carry the keys the driver
reads, the arithmetic behind `achieved` and `traffic` is what DESIGN.md section 5 states, and the reference arm's command
line exists."""
import json
import os
import subprocess
import sys

import util

sys.path.insert(0, util.ROOT)


def fake_stats(steps=3):
    names = ["orient", "index_count", "index_scan", "index_fill", "index_sort", "seed", "walk", "merge", "extend", "finalize"]
    km = dict.fromkeys(names, 0.0)
    km.update(index_count=7.5 * steps, index_scan=0.2 * steps, index_fill=30.0 * steps, index_sort=7.8 * steps, seed=118.0 * steps,
              walk=13.0 * steps, extend=417.0 * steps)
    kl = dict.fromkeys(names, 0)
    kl.update(index_count=4 * steps, index_scan=3 * steps, index_fill=3 * steps, index_sort=steps, seed=13 * steps, walk=13 * steps, extend=8 * steps)
    return {"kernel_ms": km, "kernel_launches": kl, "index_bases": 1_580_000_000 * steps, "index_kmers": 1_578_000_000 * steps,
            "num_hits": 7_671_585_280 * steps, "num_candidates": 986_447 * steps, "num_extend_blocks": 21_136_000 * steps,
            "num_extend_cells": 93_000_000_000 * steps}


def test_roofline_object_contract():
    import bench
    steps = 3
    r = bench.roofline_for(fake_stats(steps), {"hbm_gbs": 6513.8}, steps)
    for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "traffic", "ms_per_launch", "launches", "all_kernels"):
        assert k in r, k
    assert r["kernel"] == "extend" and r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] == 6513.8
    # achieved = algorithmic bytes of the dominant kernel / its summed event time; per launch: bytes and ms both divided by launches
    alg_per_step = 986_447 * (2 * 15000 / 4 + 52 + 32)
    assert abs(r["achieved"] - alg_per_step / 0.417 / 1e9) < 1e-6 * r["achieved"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["launches"] == 8 * steps and abs(r["ms_per_launch"] - 417.0 / 8) < 1e-9
    assert abs(r["algorithmic_bytes_per_launch"] - alg_per_step / 8) < 1e-3
    t = json.load(open(os.path.join(util.ROOT, "profiles", "ncu_counters.json")))["extend"]
    assert abs(r["traffic"] - t["dram_bytes_per_step"] / 8) < 1.0          # committed ncu capture, scaled to one launch
    for name, k in r["all_kernels"].items():
        assert set(k) == {"ms_per_step", "algorithmic_gb_per_step", "gbps", "frac_of_hbm_peak"}, name
    # second entry: the dominant kernel against the instruction-issue roof, from the committed ncu instruction count
    iss = r["issue"]
    c = json.load(open(os.path.join(util.ROOT, "profiles", "ncu_counters.json")))["extend"]
    assert iss["bound"] == "issue" and iss["kernel"] == "extend"
    assert abs(iss["achieved"] - c["warp_inst_per_block"] * 21_136_000 / 0.417) < 1e-6 * iss["achieved"]
    assert abs(iss["peak"] - 4 * 148 * 1965e6) < 1.0 and 0 < iss["frac"] < 1.0
    assert abs(iss["cells_per_s"] - 93e9 / 0.417) < 1e-6 * iss["cells_per_s"]
    assert r["hbm_kernel"]["kernel"] == "seed" and 0 < r["hbm_kernel"]["frac"] < 1
    # a rank of a strong-scaling run: the index kernels are charged with their slice, no entry exceeds the roof
    st = fake_stats(steps)
    for k in ("index_count", "index_fill", "index_sort", "index_scan"):
        st["kernel_ms"][k] /= 8
    r8 = bench.roofline_for(st, {"hbm_gbs": 6513.8}, steps, world=8)
    assert all(v["frac_of_hbm_peak"] < 1.0 for v in r8["all_kernels"].values())
    # without a measured peak the fallback of the profiling guide is used and said so
    r2 = bench.roofline_for(fake_stats(steps), {}, steps)
    assert "fallback" in r2["peak_source"] and r2["peak"] > 0


def test_command_line_defaults():
    p = subprocess.run([sys.executable, os.path.join(util.ROOT, "bench.py"), "--help"], capture_output=True, text=True)
    assert p.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--mode", "--workload", "--volumes"):
        assert flag in p.stdout
