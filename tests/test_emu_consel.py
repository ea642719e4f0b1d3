"""The consel side of the PRODUCT library on the CPU: tests/emu/build_emu_lib.py compiles the C ABI and the
CUDA sources of the statistics / count-query / KH-SH kernels and of the histogram kernels for the host
(kernel launches become emu::launch, one OS thread per CUDA thread) and the usual ctypes binding
(consel_b200.api) drives it.  Everything is compared bit for bit with oracle/oracle.c, exactly like the
`-m gpu` tests do on the device -- so the kernel arithmetic and the host glue of capi_stats.cu are pinned
even where no GPU is available."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def cg(emu_cg):
    return emu_cg


def _ordered(orc, datvec, ass=None):
    ptr, idx = orc.identity_assoc(len(datvec)) if ass is None else ass
    _, _, sptr, sidx = orc.order_items(datvec, ptr, idx)
    return sptr, sidx


def test_statistics_counts_order_statistics_bit_exact(cg, orc):
    rng = np.random.default_rng(1)
    K, bbs = 9, [40, 64, 33]
    reps_host = [rng.normal(size=(K, b)) for b in bbs]
    reps_host[1][3] = reps_host[1][5]  # two identical trees: ties in the max over the complement
    datvec = rng.normal(size=K) * 3
    sptr, sidx = _ordered(orc, datvec)
    reps = cg.Reps.upload(reps_host)
    reps.stats_build(sptr, sidx)
    got = reps.stats_download()
    for s, r in enumerate(reps_host):
        assert np.array_equal(got[s], orc.repminmaxs(r, sptr, sidx)), s
    srt = reps.stats_download(sorted_rows=True)
    for s in range(len(bbs)):
        assert np.array_equal(srt[s], np.sort(got[s], axis=1))
    thr = np.concatenate([rng.normal(size=(K, 2)), np.zeros((K, 1)), got[1][:, :1]], axis=1)  # incl. an exact statistic value
    q = reps.count_query(thr)
    for s in range(len(bbs)):
        v = np.sort(got[s], axis=1)
        for i in range(K):
            for k in range(thr.shape[1]):
                t = thr[i, k]
                c = int(np.searchsorted(v[i], t, side="right"))
                assert q["cnt"][i, s, k] == c
                assert q["lo"][i, s, k] == (v[i][c - 1] if c > 0 else -1e30) and q["hi"][i, s, k] == (v[i][c] if c < len(v[i]) else 1e30)
                assert (q["min1"][i, s, k], q["min2"][i, s, k], q["max1"][i, s, k], q["max2"][i, s, k]) == (v[i][0], v[i][1], v[i][-1], v[i][-2])
                # the smoothed count the AU fit uses is cntdist3 of the sorted row
                assert cg.smooth_count(q[i, s, k], len(v[i]), float(t)) == orc.cntdist(v[i], float(t), 3)
    part = reps.count_query_rows(2, thr[2:6])
    assert part.tobytes() == q[2:6].tobytes()
    rank = np.array([0, 1, bbs[1] // 2, bbs[1] - 1, 5, 7, 11, 13, 17], dtype=np.int32)
    assert np.array_equal(reps.order_stat(1, rank), np.sort(got[1], axis=1)[np.arange(K), rank])
    reps.free()


def test_kh_sh_tests_bit_exact_identity_and_grouped_associations(cg, orc):
    rng = np.random.default_rng(2)
    K, B = 8, 50
    x = rng.normal(size=(K, B)) + np.linspace(0, 1.5, K)[:, None]
    datvec = -np.linspace(0, 6, K) + rng.normal(size=K)
    for ass in (None, (np.array([0, 2, 3, 6, 8], dtype=np.int32), np.array([0, 1, 2, 3, 4, 5, 6, 7], dtype=np.int32))):
        sptr, sidx = _ordered(orc, datvec, ass)
        reps = cg.Reps.upload([x])
        got = reps.mctest(0, datvec, sptr, sidx)
        want = orc.mctest(x, datvec, sptr, sidx)
        for g, w in zip(got, want):
            assert np.array_equal(g, w)
        reps.free()


def test_philox_histogram_paths_agree_and_are_multinomial(cg):
    # byte-row paths are internal to the int8 contraction; the BJ8 seam (cgpu_resample_counts) runs the group kernel
    n, bn, bb = 700, [350, 980], [9, 17]
    a = cg.resample_counts(n, bn, bb, 5, 1, rng_mode=cg.RNG_PHILOX)
    assert a.shape == (17, n) and (a.sum(axis=1) == 980).all()
    b = cg.resample_counts(n, bn, bb, 5, 1, rng_mode=cg.RNG_PHILOX)
    c = cg.resample_counts(n, bn, bb, 6, 1, rng_mode=cg.RNG_PHILOX)
    assert np.array_equal(a, b) and not np.array_equal(a, c)


@pytest.mark.parametrize("name", ["tiny", "g100"])
def test_consel_half_end_to_end_reproduces_the_reference_pv(cg, orc, gold, name):
    """foo.rmt -> p-value table through the product's own code path on the CPU: replicate blocks uploaded,
    KH/SH/wKH/wSH tests, statistics, BP, the lockstep AU driver with its count queries and fits
    (chost_consel_run), compared with the files the unmodified reference wrote (tests/golden): items and
    counts exactly, every p-value, standard error and fit diagnostic to the nine printed digits."""
    from consel_b200 import pipeline
    ref_pv = orc.read_pv(gold / f"{name}.pv")
    rmt = orc.read_rmt(gold / f"{name}.rmt")
    reps = cg.Reps.upload(rmt["reps"])
    res = pipeline.consel(reps, rmt["datvec"], rmt["rr"], rmt["bb"])
    reps.free()
    assert np.array_equal(res["item"], ref_pv["item"])
    assert np.allclose(res["stat"], ref_pv["stat"], rtol=6e-9)
    assert np.allclose(res["pv"], ref_pv["pv"], rtol=6e-9, atol=1e-12), (res["pv"], ref_pv["pv"])
    assert np.allclose(res["se"], ref_pv["se"], rtol=6e-9, atol=1e-12)
    assert np.allclose(res["ax"], ref_pv["ax"], rtol=6e-9, atol=1e-12)
    assert np.array_equal(res["cnt"], orc.read_cnt(gold / f"{name}.cnt")["cnt"])
