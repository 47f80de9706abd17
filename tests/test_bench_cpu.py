"""CPU checks of bench.py's host logic: the work formulas behind `roofline`, how a workload is shared out over N GPUs, the
`scaling` label (the same at every N -- it once said "weak" at N = 1 and "strong" at N > 1 for the same workload), the
CPU arm's sub-grid, and the committed DRAM-traffic table."""

import argparse
import json
import os

import numpy as np
import pytest

import bench
import workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def args(**kw):
    base = dict(sharding="auto", frames=None)
    base.update(kw)
    return argparse.Namespace(**base)


def test_work_formulas_match_the_survey_values():
    # F(N) = (N-1)(5N+6)+30 per grid point
    assert [workloads.flops_per_point(n) for n in (16, 32, 64, 256, 1024)] == [1320, 5176, 20568, 327960, 5243928]
    # one recurrence per unit (pair of grid lines = 8 points): 10 flops per inner step, 72 per diagonal, ~40 epilogue
    for n in (2, 16, 64, 1024):
        assert workloads.flops_per_unit_shared(n) == (n - 1) * (5 * (n - 2) + 72) + 40
    assert workloads.flops_per_unit_shared(1024) == 5301226  # what bench reports as flops_per_unit_executed for c5
    # units of an S x S mirror grid: U (U + 1) / 2 with U = ceil(S / 2)
    assert workloads.shared_units(4096) == 2048 * 2049 // 2 == 2098176
    assert workloads.shared_units(400) == 200 * 201 // 2
    assert workloads.shared_units(401) == 201 * 202 // 2


@pytest.mark.parametrize("name,frames,expect", [("c5", None, "grid"), ("c2", None, "grid"), ("c3", 512, "frames"), ("c4", None, "replicas")])
def test_sharding_mode_and_scaling_label(name, frames, expect):
    wl = bench.make_workload(name, frames) if name != "c5" else dict(workloads.config5(steps=64, N=16), name="c5")
    assert bench.sharding_mode(wl, 1, args()) == "none"
    labels = set()
    for world in (1, 2, 4, 8):
        if world > 1:
            assert bench.sharding_mode(wl, world, args()) == expect
        cfg = bench.describe(wl, world, args())
        labels.add(cfg["_scaling"])
        assert cfg["workload"].startswith(name)
    assert labels == {"weak" if expect == "replicas" else "strong"}


def test_sharding_override():
    wl = bench.make_workload("c3", 64)
    assert bench.sharding_mode(wl, 4, args(sharding="grid")) == "replicas"  # frames cannot be shared out as one grid
    one = bench.make_workload("c2")
    assert bench.sharding_mode(one, 4, args(sharding="frames")) == "replicas"  # a single frame cannot be split by frame


def test_cpu_sub_grid_is_inside_the_grid_and_bounded():
    for S, N, sec in ((200, 16, 1.0), (1024, 64, 5.0), (4096, 1024, 15.0)):
        rows, cols = bench.sub_grid(S, N, sec)
        rows, cols = np.asarray(rows), np.asarray(cols)
        assert rows.min() >= 0 and rows.max() < S and cols.min() >= 0 and cols.max() < S
        assert len(set(rows.tolist())) == len(rows) and len(set(cols.tolist())) == len(cols)
        pts = len(rows) * len(cols)
        assert 16 <= pts <= S * S
        # about `sec` seconds at the cost model's 0.9 algorithmic GFLOP/s per core (never more than 2x over)
        assert pts * workloads.flops_per_point(N) <= 2.0 * max(sec * 0.9e9, 16 * workloads.flops_per_point(N))


def test_traffic_table_schema():
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
        t = json.load(f)
    for name in ("c2", "c3", "c5"):
        e = t[name]
        assert e["dram_bytes_read"] > 0 and e["dram_bytes_write"] > 0
        assert os.path.exists(os.path.join(ROOT, e["source"].split(" ")[0])), e["source"]
