"""Parity at the SHAPES of BASELINE.json's large configurations (K, n intact, few replicates so the
CPU oracle finishes in seconds): C3 (K=1000 x n=10^5: uint32 site indices, one-replicate-per-CTA
histograms, 32 x 40 GEMM tile grid), C5 (K=10^4 x n=5*10^4: 313 digit-row tiles, consel kernels at
K=10^4), a genome-length alignment (n=1.6*10^6: global-atomic histogram, n' >= 2^21 so the
recombination leaves the GEMM epilogue), several GEMM batches per scale, and the one-CTA GEMM kept
as a measurement variant.  Counts and KH/SH/AU inputs are bit-exact against oracle/oracle.c, the
int8 sums bit-equal to the correctly rounded exact sums (math.fsum over the drawn multiset)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cg():
    import consel_b200 as cg
    cg.init()
    cg.set_shard(0, 1)
    yield cg
    for opt in ("batch", "i8_2cta", "hist_wide"):
        cg.set_option(opt, -1)


def synth(K, n, seed=20261019):
    rng = np.random.Generator(np.random.PCG64(seed))
    base = -2.3 - 4.07 * rng.exponential(size=n)
    L = np.empty((K, n))
    step = max(1, (1 << 23) // n)
    for k0 in range(0, K, step):
        k1 = min(K, k0 + step)
        L[k0:k1] = base[None, :] + 0.106 * rng.standard_normal((k1 - k0, n)) - 0.5 * np.arange(k0, k1)[:, None] / n
    return L


def exact_sums(Lrows, counts, r):
    out = np.empty((Lrows.shape[0], counts.shape[0]))
    for b in range(counts.shape[0]):
        c = counts[b].astype(np.int64)
        for k in range(Lrows.shape[0]):
            out[k, b] = math.fsum(np.repeat(Lrows[k], c)) / np.float64(r)
    return out


def check_makermt(cg, orc, L, rr, bb, seed, trees, ref=None):
    """device int8 replicate sums: sampled trees bit-equal to the exact rounded sums of the ORACLE's
    counts (so the device histogram must be right too), all trees within 1e-12 of the oracle's
    sequentially rounded sums"""
    K, n = L.shape
    _, bn, rr1 = orc.scale_setup(L, rr)
    out = cg.scaleboot_all(L, bn, bb, seed, sum_mode=cg.SUM_INT8)
    counts = orc.scaleboot_counts(n, bn, bb, seed)
    for s in range(len(rr)):
        want = exact_sums(L[trees], counts[s], rr1[s])
        assert np.array_equal(out[s][trees], want), (s, np.abs(out[s][trees] - want).max())
    if ref is None:
        ref = orc.makermt(L, bn, bb, seed)
    for s in range(len(rr)):
        rel = np.abs(out[s] - ref[s]) / np.abs(ref[s])
        assert rel.max() < 1e-12, (s, rel.max())
    return out, ref, bn, rr1, counts


def check_consel(cg, orc, reps_host, datvec, rr1):
    """statistics, threshold counts and the KH/SH/wKH/wSH tests against oracle/oracle.c, bit for bit"""
    K = reps_host[0].shape[0]
    ptr, idx = orc.identity_assoc(K)
    _, _, sptr, sidx = orc.order_items(datvec, ptr, idx)
    reps = cg.Reps.upload(reps_host)
    reps.stats_build(sptr, sidx)
    got = reps.stats_download()
    for s, r in enumerate(reps_host):
        assert np.array_equal(got[s], orc.repminmaxs(r, sptr, sidx)), s
    i0 = orc.find_i0(rr1)
    B0 = reps_host[i0].shape[1]
    med = reps.order_stat(i0, np.full(K, B0 // 2, dtype=np.int32))
    assert np.array_equal(med, np.sort(got[i0], axis=1)[:, B0 // 2])
    thr = np.stack([np.zeros(K), med, got[i0].min(axis=1) - 1.0], axis=1)
    q = reps.count_query(thr)
    for s in range(len(reps_host)):
        srt = np.sort(got[s], axis=1)
        for k in range(thr.shape[1]):
            want = np.array([np.searchsorted(srt[i], thr[i, k], side="right") for i in range(K)])
            assert np.array_equal(q["cnt"][:, s, k], want), (s, k)
    kh, sh, wkh, wsh = reps.mctest(i0, datvec, sptr, sidx)
    okh, osh, owkh, owsh = orc.mctest(reps_host[i0], datvec, sptr, sidx)
    assert np.array_equal(kh, okh) and np.array_equal(sh, osh)
    assert np.array_equal(wkh, owkh) and np.array_equal(wsh, owsh)
    reps.free()


def test_c3_shape(cg, orc):
    K, n = 1000, 100000
    L = synth(K, n)
    rr, bb = [0.5, 1.0, 1.4], [24, 40, 24]
    trees = np.array([0, 1, 499, 998, 999])
    out, ref, bn, rr1, counts = check_makermt(cg, orc, L, rr, bb, 7, trees)
    # the resampler's own count seam (BJ8 layout, uint32 indices) at this n
    got = cg.resample_counts(n, bn, bb, 7, 2)
    assert np.array_equal(got, counts[2])
    datvec, _, _ = orc.scale_setup(L, rr)
    check_consel(cg, orc, ref, datvec, rr1)
    # the one-CTA GEMM (measurement variant) and the group histogram return the same bits
    cg.set_option("i8_2cta", 0)
    cg.set_option("hist_wide", 0)
    _, bnx, _ = orc.scale_setup(L, rr)
    alt = cg.scaleboot_all(L, bnx, bb, 7, sum_mode=cg.SUM_INT8)
    cg.set_option("i8_2cta", -1)
    cg.set_option("hist_wide", -1)
    for s in range(len(rr)):
        assert np.array_equal(alt[s], out[s]), s


def test_c5_shape(cg, orc):
    K, n = 10000, 50000
    L = synth(K, n, seed=5)
    rr, bb = [0.7, 1.0], [12, 20]
    trees = np.array([0, 5000, 9999])
    out, ref, bn, rr1, counts = check_makermt(cg, orc, L, rr, bb, 11, trees)
    datvec, _, _ = orc.scale_setup(L, rr)
    check_consel(cg, orc, ref, datvec, rr1)


def test_genome_length_alignment(cg, orc):
    """n = 1.6e6: rows longer than any shared-memory histogram (global atomics), and n' = 1.4 n >= 2^21 draws
    per replicate, where the digit sums no longer fold exactly in fp64 and the recombination runs as
    fx_combine_kernel on the int32 partial sums"""
    K, n = 6, 1600000
    L = synth(K, n, seed=9)
    rr, bb = [0.6, 1.4], [5, 9]
    check_makermt(cg, orc, L, rr, bb, 3, np.arange(K))


def test_several_gemm_batches_per_scale(cg, orc):
    K, n = 40, 3000
    L = synth(K, n, seed=2)
    rr, bb = [0.5, 1.3], [700, 515]
    _, bn, _ = orc.scale_setup(L, rr)
    one = cg.scaleboot_all(L, bn, bb, 21, sum_mode=cg.SUM_INT8)
    cg.set_option("i8_2cta", 1)  # short tiles default to the one-CTA kernel: force the pair kernel too
    pair = cg.scaleboot_all(L, bn, bb, 21, sum_mode=cg.SUM_INT8)
    cg.set_option("i8_2cta", -1)
    for s in range(2):
        assert np.array_equal(one[s], pair[s])
    cg.set_option("batch", 256)
    many = cg.scaleboot_all(L, bn, bb, 21, sum_mode=cg.SUM_INT8)
    strict = cg.scaleboot_all(L, bn, bb, 21, sum_mode=cg.SUM_STRICT)
    cg.set_option("batch", -1)
    ref = orc.makermt(L, bn, bb, 21)
    for s in range(2):
        assert np.array_equal(one[s], many[s])
        assert np.array_equal(strict[s], ref[s])  # the strict fp64 mode with small batches is still the reference's bits
        assert (np.abs(one[s] - ref[s]) / np.abs(ref[s])).max() < 1e-12


def test_philox_stream_offset_gives_independent_replicates(cg):
    """makermt -g continues one logical stream over the gene files: under the counter-based generator the
    consumed-draw offset is folded into the counter, so consecutive files draw different replicates"""
    n = 500
    bn, bb = [400, 600], [16, 16]
    a = cg.resample_counts(n, bn, bb, 9, 1, rng_mode=cg.RNG_PHILOX, stream_offset=0)
    b = cg.resample_counts(n, bn, bb, 9, 1, rng_mode=cg.RNG_PHILOX, stream_offset=16 * 1000)
    a2 = cg.resample_counts(n, bn, bb, 9, 1, rng_mode=cg.RNG_PHILOX, stream_offset=0)
    assert np.array_equal(a, a2) and not np.array_equal(a, b)
    assert (a.sum(axis=1) == 600).all() and (b.sum(axis=1) == 600).all()


def test_more_than_32_scales_row_range_query(cg, orc):
    """33 scales and 16+ items: the AU driver's two-half schedule queries row ranges, which must work without
    the one-launch path (S <= 32) and its edge cache"""
    rng = np.random.default_rng(4)
    K, S, B = 20, 33, 50
    reps_host = [rng.normal(size=(K, B)) for _ in range(S)]
    reps = cg.Reps.upload(reps_host)
    ptr, idx = orc.identity_assoc(K)
    reps.stats_build(ptr, idx)
    thr = rng.normal(size=(K, 2))
    whole = reps.count_query(thr)
    part = reps.count_query_rows(7, thr[7:15])
    assert part.tobytes() == whole[7:15].tobytes()
    reps.free()


def test_split_k_pair_gemm(cg, orc):
    """few tiles, long rows (the C4 shape in small): the site axis of a pair-GEMM tile is cut into parts whose
    int32 partial sums are added with red.global; same bits as the unsplit kernel and as the exact sums"""
    K, n = 12, 150000  # 96 digit rows -> 1 digit tile; 1172 k-blocks
    L = synth(K, n, seed=4)
    rr, bb = [0.5, 1.2], [40, 300]
    _, bn, rr1 = orc.scale_setup(L, rr)
    base = cg.scaleboot_all(L, bn, bb, 13, sum_mode=cg.SUM_INT8)
    outs = {}
    for ks in (1, 3, 8):
        cg.set_option("i8_ksplit", ks if ks > 1 else 0)
        outs[ks] = cg.scaleboot_all(L, bn, bb, 13, sum_mode=cg.SUM_INT8)
    cg.set_option("i8_ksplit", -1)
    for s in range(2):
        for ks in (1, 3, 8):
            assert np.array_equal(outs[ks][s], base[s]), (ks, s)
    counts = orc.scaleboot_counts(n, bn, [4, 4], 13)
    want = exact_sums(L[[0, K - 1]], counts[0], rr1[0])
    assert np.array_equal(base[0][[0, K - 1]][:, :4], want)
