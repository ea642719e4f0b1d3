"""bench.py is the measurement contract: its host-side helpers (workload table, the PAML reader of C1, the
exact-sum checker of the in-run spot check) are pinned here on CPU."""
import json
import math
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def test_c1_reader_equals_oracle_reader(orc, gold):
    assert np.array_equal(bench.read_paml_lnf(gold / "mam15.lnf"), orc.read_lnf(gold / "mam15.lnf"))


def test_workloads_follow_baseline_json():
    base = json.loads((ROOT / "BASELINE.json").read_text())
    assert len(base["configs"]) == 5 and bench.DEFAULT_CONFIG == "C3"
    shapes = {"C1": (15, 3414, 10000), "C2": (105, 3414, 100000), "C3": (1000, 100000, 10000),
              "C4": (100, 10000000, 10000), "C5": (10000, 50000, 10000)}
    for name, (K, n, B) in shapes.items():
        c = bench.CONFIGS[name]
        assert (c["K"], c["n"], c["B"]) == (K, n, B)
    assert bench.workload("C2", 4)["B_total"] == 400000           # weak: per-GPU replicates fixed
    assert bench.workload("C3", 8)["B_total"] == 10000            # strong: total fixed
    assert bench.workload("C4", 8)["B_total"] == 10000 and bench.workload("C4", 2)["B_total"] == 2500
    cfg = bench.config_json("C3", bench.workload("C3", 2), 2)
    assert cfg["replicates_per_scale_per_gpu"] == 5000 and "workload" in cfg and "l2" in cfg


def test_exact_sum_checker_is_the_big_integer_sum(orc):
    rng = np.random.default_rng(2)
    L = -rng.exponential(4.0, size=(3, 50)) - 1.0
    counts = rng.poisson(1.2, size=(4, 50))
    got = bench.exact_sums(L, counts, 1.2)
    want = orc.exact_rounded_contract(L, counts, 1.2)
    assert np.array_equal(got, want)
    assert got[0, 0] == math.fsum(np.repeat(L[0], counts[0])) / np.float64(1.2)


def test_synthetic_matrix_is_reproducible_and_row_blocked():
    a = bench.synth_matrix(37, 501)
    b = bench.synth_matrix(37, 501)
    assert a.shape == (37, 501) and np.array_equal(a, b)
    assert -60.0 < a.min() < a.max() < 0.0


def test_cpu_baseline_drives_the_reference_scaleboot_like_genrmt(orc, monkeypatch):
    """bench.cpu_baseline_leg calls the unmodified scaleboot() scale after scale on ONE seeded MT19937 stream
    (src/makermt.c:196-207): the replicate sums are the oracle's, i.e. the .rmt payload of `makermt -s 7`"""
    import numpy as np
    import pytest

    import bench
    so = orc.REF_DIR / "libmakermt_ref.so"
    if not so.exists():
        pytest.skip("oracle/_ref/libmakermt_ref.so not built")
    rng = np.random.default_rng(3)
    K, n = 4, 500
    L = -np.exp(rng.uniform(-1, 3, size=(K, n)))
    rr, bb = [0.5, 1.0, 1.4], [5, 5, 5]
    _, bn, _ = orc.scale_setup(L, rr)
    want = orc.makermt(L, bn, bb, 7)
    bench._REF_L = np.ascontiguousarray(L)
    try:
        for s in range(len(rr)):
            _, out = bench._ref_worker((str(so), K, n, int(bn[s]), bb[s], 7), seed_it=(s == 0))
            assert np.array_equal(out, want[s])
    finally:
        bench._REF_L = None
    _, bn10, _ = orc.scale_setup(L, bench.RR)
    monkeypatch.setattr(bench, "CPU_BASELINE_ACC", 4 * 500 * 10 * 200.0)
    line = bench.cpu_baseline_leg(L, bn10, K, n)
    assert line.get("kind") == "reference" and line["cores"] == 1 and line["value"] > 0, line
