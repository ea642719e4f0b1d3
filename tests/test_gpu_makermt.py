"""GPU parity tests of the makermt side: device MT19937 jump-ahead stream, site
histograms and the strict-order contraction, through the C ABI, against the CPU
oracle and the reference's golden vectors (tests/golden, oracle/gen_golden.py)."""
import hashlib
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RR10 = [0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4]


@pytest.fixture(scope="module")
def cg():
    import consel_b200 as cg
    cg.init()
    cg.set_shard(0, 1)
    return cg


def synth(K, n, seed=20261019):
    rng = np.random.Generator(np.random.PCG64(seed))
    base = -2.3 - 4.07 * rng.exponential(size=n)
    return base[None, :] + 0.106 * rng.standard_normal((K, n)) - 0.5 * np.arange(K)[:, None] / n


@pytest.mark.parametrize("seed,skip,count", [(1234, 0, 5000), (1234, 623, 3), (7, 10_000_019, 4097),
                                              (99, 3_000_000_000, 2000), (4357, 1 << 33, 700)])
def test_device_mt19937_stream_bit_exact(cg, orc, seed, skip, count):
    got = cg.mt19937_stream(seed, skip, count)
    if skip <= 20_000_000:
        want = orc.mt_stream(seed, skip, count)
    else:
        # the CPU cannot walk 8e9 draws in a test: use the host polynomial jump (itself
        # pinned against sequential generation in test_mt_jump.py) to place a window
        import ctypes as C
        lib = cg.load()
        w = np.empty(624, np.uint32)
        lib.cgpu_test_mt_window_host(C.c_uint(seed), C.c_ulonglong(skip), w.ctypes.data_as(C.POINTER(C.c_uint)))
        want = _draws_from_window(w, count)
    assert np.array_equal(got, want)


def _draws_from_window(w, count):
    x = [int(v) for v in w]
    out = []
    for t in range(count):
        y = (x[t] & 0x80000000) | (x[t + 1] & 0x7FFFFFFF)
        v = x[t + 397] ^ (y >> 1) ^ (0x9908B0DF if (x[t + 1] & 1) else 0)
        x.append(v)
        v ^= v >> 11
        v ^= (v << 7) & 0x9D2C5680
        v ^= (v << 15) & 0xEFC60000
        v ^= v >> 18
        out.append(v & 0xFFFFFFFF)
    return np.array(out, dtype=np.uint32)


def test_counts_bit_exact_small_and_ragged(cg, orc, gold):
    rr, bb = orc.read_pa(gold / "tiny.pa")
    n = 40
    bn = (rr * n).astype(np.int32)
    want = orc.scaleboot_counts(n, bn, bb, 7)
    for s in range(len(bn)):
        got = cg.resample_counts(n, bn, bb, 7, s)
        assert np.array_equal(got.astype(np.int32), want[s]), s


def test_counts_bit_exact_mam15_shape(cg, orc):
    n = 3414
    bn = (np.array(RR10) * n).astype(np.int32)
    bb = np.array([40, 8, 9, 64, 65, 100, 3, 1, 17, 33], dtype=np.int32)
    want = orc.scaleboot_counts(n, bn, bb, 1234)
    for s in (0, 4, 5, 9):
        got = cg.resample_counts(n, bn, bb, 1234, s)
        assert np.array_equal(got.astype(np.int32), want[s]), s


def test_counts_large_n_global_histogram(cg, orc):
    n = 40_000  # beyond the shared-memory histogram: atomic path
    bn = np.array([20_000, 56_000], dtype=np.int32)
    bb = np.array([9, 5], dtype=np.int32)
    want = orc.scaleboot_counts(n, bn, bb, 11)
    for s in range(2):
        assert np.array_equal(cg.resample_counts(n, bn, bb, 11, s).astype(np.int32), want[s])


def test_counts_fp64_index_path(cg, orc):
    n = (1 << 21) + 12345  # rv*n is no longer exact: the fp64 multiply+truncate path
    bn = np.array([5000], dtype=np.int32)
    bb = np.array([3], dtype=np.int32)
    want = orc.scaleboot_counts(n, bn, bb, 5)
    assert np.array_equal(cg.resample_counts(n, bn, bb, 5, 0).astype(np.int32), want[0])


@pytest.mark.parametrize("name,seed", [("g100", 1234), ("tiny", 7)])
def test_rmt_bit_exact_vs_reference_golden(cg, orc, gold, name, seed):
    ref = orc.read_rmt(gold / f"{name}.rmt")
    if name == "g100":
        L, rr = orc.read_lnf(gold / "mam15.lnf"), RR10
    else:
        L = orc.read_mt(gold / "tiny.mt")
        rr, _ = orc.read_pa(gold / "tiny.pa")
    out = cg.makermt(L, rr, ref["bb"], seed)
    assert np.array_equal(out["datvec"], ref["datvec"])
    assert np.array_equal(out["rr"], ref["rr"])
    for s in range(len(rr)):
        assert np.array_equal(out["reps"][s], ref["reps"][s]), s


def test_full_mam15_rmt_md5_equals_reference(cg, orc, gold, tmp_path):
    """config C1 in full (10 x 10000 replicates): the .rmt FILE is byte-identical."""
    man = json.loads((gold / "MANIFEST.json").read_text())
    L = orc.read_lnf(gold / "mam15.lnf")
    out = cg.makermt(L, RR10, [10000] * 10, 1234)
    orc.write_rmt(tmp_path / "full.rmt", out["datvec"], out["rr"], out["bb"], out["reps"])
    assert hashlib.md5((tmp_path / "full.rmt").read_bytes()).hexdigest() == man["full_size"]["mam15_s1234.rmt"]
    spot = man["full_size"]["spot"]["rep_s0_k0_b0_3"]
    assert out["reps"][0][0, :4].tolist() == spot


@pytest.mark.parametrize("K,n", [(1, 130), (33, 257), (64, 128), (105, 700), (129, 300), (300, 200)])
def test_strict_contraction_all_tree_tiles(cg, orc, K, n):
    L = synth(K, n, seed=K * 1000 + n)
    rr = [0.5, 1.0, 1.4]
    bb = [13, 70, 129]
    _, bn, rr1 = orc.scale_setup(L, rr)
    want = orc.makermt(L, bn, bb, 31)
    got = cg.scaleboot_all(L, bn, bb, 31)
    for s in range(3):
        assert np.array_equal(got[s], want[s]), (K, n, s)


def test_fast_mode_within_tolerance(cg, orc):
    L = synth(105, 3414)
    _, bn, _ = orc.scale_setup(L, RR10)
    bb = [50] * 10
    want = orc.makermt(L, bn, bb, 9)
    got = cg.scaleboot_all(L, bn, bb, 9, sum_mode=cg.SUM_FAST)
    for s in range(10):
        rel = np.abs(got[s] - want[s]) / np.abs(want[s])
        assert rel.max() < 1e-12  # north star: <= 1e-9 relative; fused terms differ by <= 1 ulp each


def test_non_finite_input_propagates_like_reference(cg, orc):
    L = synth(5, 64)
    L[2, 10] = -np.inf
    _, bn, _ = orc.scale_setup(L, [1.0])
    want = orc.makermt(L, bn, [40], 3)
    got = cg.scaleboot_all(L, bn, [40], 3)
    assert np.array_equal(got[0], want[0], equal_nan=True)
    assert np.isnan(got[0][2]).all() or np.isinf(got[0][2]).any()


def test_stream_offset_continues_the_stream(cg, orc):
    # makermt -g: the second file continues the first file's stream (src/makermt.c:303-318)
    L = synth(9, 150)
    _, bn, _ = orc.scale_setup(L, [0.7, 1.0])
    bb = [20, 31]
    both = orc.makermt(L, np.concatenate([bn, bn]), bb + bb, 77)
    consumed = int(bn[0]) * bb[0] + int(bn[1]) * bb[1]
    second = cg.scaleboot_all(L, bn, bb, 77, stream_offset=consumed)
    assert np.array_equal(second[0], both[2]) and np.array_equal(second[1], both[3])


def test_sharding_does_not_change_results(cg, orc):
    L = synth(40, 500)
    _, bn, _ = orc.scale_setup(L, [0.6, 1.0, 1.3])
    bb = [30, 45, 17]
    for mode in (cg.RNG_MT19937, cg.RNG_PHILOX):
        cg.set_shard(0, 1)
        whole = cg.scaleboot_all(L, bn, bb, 5, rng_mode=mode)
        parts = [np.zeros_like(w) for w in whole]
        for rank in range(3):
            cg.set_shard(rank, 3)
            loc = cg.scaleboot_all(L, bn, bb, 5, rng_mode=mode)  # rank-local compact blocks
            for s in range(3):
                b0, b1 = cg.shard_range(bb[s])
                assert loc[s].shape == (40, b1 - b0)
                parts[s][:, b0:b1] = loc[s]
        cg.set_shard(0, 1)
        for s in range(3):
            assert np.array_equal(parts[s], whole[s]), (mode, s)
    want = orc.makermt(L, bn, bb, 5)
    cg.set_shard(0, 1)
    got = cg.scaleboot_all(L, bn, bb, 5)
    assert all(np.array_equal(a, b) for a, b in zip(got, want))


def test_philox_mode_is_a_valid_multinomial(cg):
    n = 3414
    bn = (np.array([0.5, 1.0, 1.4]) * n).astype(np.int32)
    bb = np.array([400, 400, 400], dtype=np.int32)
    for s in range(3):
        c = cg.resample_counts(n, bn, bb, 42, s, rng_mode=cg.RNG_PHILOX).astype(np.int64)
        assert (c.sum(1) == bn[s]).all()
        # every site is hit bn/n times on average; variance of a multinomial cell ~ mean
        mean = c.mean(0)
        assert abs(mean.mean() - bn[s] / n) < 1e-12
        z = (mean - bn[s] / n) / np.sqrt(bn[s] / n / 400)
        assert np.abs(z).max() < 6 and abs(z.std() - 1) < 0.1
        again = cg.resample_counts(n, bn, bb, 42, s, rng_mode=cg.RNG_PHILOX)
        assert np.array_equal(c, again)
        other = cg.resample_counts(n, bn, bb, 43, s, rng_mode=cg.RNG_PHILOX)
        assert not np.array_equal(c, other)


def test_tiny_n_overflow_is_refused(cg):
    with pytest.raises(cg.CgpuError) as e:
        cg.resample_counts(1, [300], [2], 1, 0)
    assert e.value.code == -5


def test_bad_arguments_fail_loudly(cg):
    L = synth(3, 10)
    with pytest.raises(cg.CgpuError):
        cg.scaleboot_all(L, [5], [4], 1, sum_mode=9)
    with pytest.raises(cg.CgpuError):
        cg.scaleboot_all(L, [0], [4], 1)
