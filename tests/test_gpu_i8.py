"""CGPU_SUM_INT8: exact fixed-point replicate sums on the tcgen05 int8 tensor cores.
Bit-equal to a big-integer restatement (oracle.exact_rounded_contract), within 1e-9
relative of the reference's sequentially rounded sums, independent of batch/sharding."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("K,n,bb", [(3, 100, [5]), (15, 300, [40, 9]), (40, 700, [64, 257]), (130, 260, [300])])
def test_int8_path_is_exactly_the_rounded_exact_sum(cg, orc, K, n, bb):
    L = synth(K, n, seed=K + n)
    rr = [0.6, 1.3][:len(bb)]
    _, bn, rr1 = orc.scale_setup(L, rr)
    got = cg.scaleboot_all(L, bn, bb, 17, sum_mode=cg.SUM_INT8)
    ref = orc.makermt(L, bn, bb, 17)
    for s in range(len(bb)):
        counts = cg.resample_counts(n, bn, bb, 17, s)
        want = orc.exact_rounded_contract(L, counts, rr1[s])
        assert np.array_equal(got[s], want), (K, n, s, np.abs(got[s] - want).max())
        rel = np.abs(got[s] - ref[s]) / np.abs(ref[s])
        assert rel.max() < 1e-9  # north star tolerance; in practice ~n*2^-53


def test_int8_mam15_against_reference_rmt(cg, orc, gold):
    L = orc.read_lnf(gold / "mam15.lnf")
    ref = orc.read_rmt(gold / "g100.rmt")
    out = cg.makermt(L, [0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4], ref["bb"], 1234, sum_mode=cg.SUM_INT8)
    for s in range(10):
        rel = np.abs(out["reps"][s] - ref["reps"][s]) / np.abs(ref["reps"][s])
        assert rel.max() < 1e-12, (s, rel.max())
    # the fast fp64 mode and the int8 mode bracket the reference the same way
    fast = cg.makermt(L, [1.0], [64], 5, sum_mode=cg.SUM_FAST)["reps"][0]
    i8 = cg.makermt(L, [1.0], [64], 5, sum_mode=cg.SUM_INT8)["reps"][0]
    assert np.abs(fast - i8).max() / np.abs(i8).max() < 1e-12


def test_int8_sharding_and_philox(cg, orc):
    L = synth(20, 500)
    _, bn, _ = orc.scale_setup(L, [0.7, 1.2])
    bb = [70, 33]
    for mode in (cg.RNG_MT19937, cg.RNG_PHILOX):
        cg.set_shard(0, 1)
        whole = cg.scaleboot_all(L, bn, bb, 3, rng_mode=mode, sum_mode=cg.SUM_INT8)
        for rank in range(2):
            cg.set_shard(rank, 2)
            loc = cg.scaleboot_all(L, bn, bb, 3, rng_mode=mode, sum_mode=cg.SUM_INT8)
            for s in range(2):
                b0, b1 = cg.shard_range(bb[s])
                assert np.array_equal(loc[s], whole[s][:, b0:b1])
        cg.set_shard(0, 1)


def test_int8_wide_dynamic_range(cg, orc):
    rng = np.random.default_rng(2)
    L = -np.exp(rng.uniform(-12, 5, size=(4, 130)))  # magnitudes 6e-6 .. 150: ~80 bits of fixed point
    _, bn, rr1 = orc.scale_setup(L, [1.0])
    got = cg.scaleboot_all(L, bn, [9], 1, sum_mode=cg.SUM_INT8)
    counts = cg.resample_counts(130, bn, [9], 1, 0)
    assert np.array_equal(got[0], orc.exact_rounded_contract(L, counts, rr1[0]))
    Lbad = L.copy()
    Lbad[0, 0] = -1e-30  # > 96 bits: must be refused, not truncated
    with pytest.raises(cg.CgpuError) as e:
        cg.scaleboot_all(Lbad, bn, [9], 1, sum_mode=cg.SUM_INT8)
    assert e.value.code == -5


@pytest.mark.parametrize("bits", [3, 11, 27, 40, 59, 75])
def test_int8_slice_counts_fused_and_two_kernel_forms(cg, orc, bits):
    """Matrices whose fixed-point span needs S = 1, 2, 4, 6, 8, 10 digits: the GEMM epilogue
    recombination (S divides 32) and the two-kernel form must both equal the exact sum."""
    rng = np.random.default_rng(bits)
    K, n, bb = 37, 300, [70]
    low = min(bits, 40)  # mantissas of up to 40 bits, spread over the span by a random exponent
    mag = rng.integers(1, 2 ** low, size=(K, n)).astype(np.float64) * 2.0 ** rng.integers(0, bits - low + 1, size=(K, n))
    L = np.where(rng.random((K, n)) < 0.2, 1.0, -1.0) * mag * 2.0 ** -20
    _, bn, rr1 = orc.scale_setup(L, [0.9])
    counts = cg.resample_counts(n, bn, bb, 5, 0)
    want = orc.exact_rounded_contract(L, counts, rr1[0])
    try:
        for fuse in (1, 0):
            cg.set_option("i8_fuse", fuse)
            got = cg.scaleboot_all(L, bn, bb, 5, sum_mode=cg.SUM_INT8)
            assert np.array_equal(got[0], want), (bits, fuse, np.abs(got[0] - want).max())
    finally:
        cg.set_option("i8_fuse", -1)


@pytest.mark.parametrize("n,rr,bb", [(2600, [0.5, 1.0, 1.4], [300, 77, 130]), (3414, [0.5, 0.9, 1.4], [513, 256, 40]),
                                     (4500, [0.7], [1000])])
def test_fused_draw_histogram_path(cg, orc, n, rr, bb):
    """n' >= 1248 draws per replicate and n <= 4608 sites: draws and histograms run in one
    kernel (no index stream).  Same counts, hence bit-equal sums, as the two-kernel path,
    for the whole run and for every shard of it."""
    K = 21
    L = synth(K, n, seed=n)
    _, bn, rr1 = orc.scale_setup(L, rr)
    try:
        cg.set_option("draw_fuse", 0)
        plain = cg.scaleboot_all(L, bn, bb, 23, sum_mode=cg.SUM_INT8)
        cg.set_option("draw_fuse", 1)
        fused = cg.scaleboot_all(L, bn, bb, 23, sum_mode=cg.SUM_INT8)
        for s in range(len(bb)):
            counts = cg.resample_counts(n, bn, bb, 23, s)
            want = orc.exact_rounded_contract(L, counts, rr1[s])
            assert np.array_equal(plain[s], want)
            assert np.array_equal(fused[s], want), (s, np.abs(fused[s] - want).max())
        for world in (2, 3):
            for rank in range(world):
                cg.set_shard(rank, world)
                loc = cg.scaleboot_all(L, bn, bb, 23, sum_mode=cg.SUM_INT8)
                for s in range(len(bb)):
                    b0, b1 = cg.shard_range(bb[s])
                    assert np.array_equal(loc[s], fused[s][:, b0:b1]), (world, rank, s)
    finally:
        cg.set_shard(0, 1)
        cg.set_option("draw_fuse", -1)


def test_fused_draw_histogram_small_chunks(cg, orc):
    """chunks shorter than a replicate: every histogram row is assembled from several chunks"""
    n, K, bb = 3000, 9, [60, 33]
    L = synth(K, n, seed=5)
    _, bn, rr1 = orc.scale_setup(L, [0.8, 1.2])
    try:
        cg.set_option("draw_fuse", 0)
        plain = cg.scaleboot_all(L, bn, bb, 99, sum_mode=cg.SUM_INT8)
        cg.set_option("draw_fuse", 1)
        for log2 in (11, 12, 14, 17):
            cg.set_option("chunk_log2", log2)
            fused = cg.scaleboot_all(L, bn, bb, 99, sum_mode=cg.SUM_INT8)
            for s in range(len(bb)):
                assert np.array_equal(fused[s], plain[s]), (log2, s)
    finally:
        cg.set_option("chunk_log2", -1)
        cg.set_option("draw_fuse", -1)


def test_staged_index_production_same_result(cg, orc):
    """cgpu_scaleboot_all in int8 mode produces the index stream in three stages on two
    streams for large runs ("stage" 1 forces it at any size): identical blocks."""
    K, n = 11, 900
    L = synth(K, n, seed=3)
    rr = [0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4]
    bb = [37, 50, 64, 21, 80, 33, 10, 90, 5, 41]
    _, bn, _ = orc.scale_setup(L, rr)
    try:
        cg.set_option("stage", 0)
        plain = cg.scaleboot_all(L, bn, bb, 77, sum_mode=cg.SUM_INT8)
        cg.set_option("stage", 1)
        staged = cg.scaleboot_all(L, bn, bb, 77, sum_mode=cg.SUM_INT8)
        for s in range(len(bb)):
            assert np.array_equal(plain[s], staged[s]), s
        for rank in range(2):
            cg.set_shard(rank, 2)
            loc = cg.scaleboot_all(L, bn, bb, 77, sum_mode=cg.SUM_INT8)
            for s in range(len(bb)):
                b0, b1 = cg.shard_range(bb[s])
                assert np.array_equal(loc[s], plain[s][:, b0:b1]), (rank, s)
    finally:
        cg.set_shard(0, 1)
        cg.set_option("stage", -1)


@pytest.mark.parametrize("n", [30000, 60000, 120000])
def test_int8_long_alignments_shared_histograms(cg, orc, n):
    """28160 < n <= 225280: byte-row histograms with 4, 2, 1 replicates per CTA (the fp64 paths
    count such alignments with global atomics, cgpu_resample_counts is that path)."""
    K, bb = 3, [5, 2]
    L = synth(K, n, seed=n)
    _, bn, rr1 = orc.scale_setup(L, [0.6, 1.3])
    for mode in (cg.RNG_MT19937,):
        got = cg.scaleboot_all(L, bn, bb, 31, rng_mode=mode, sum_mode=cg.SUM_INT8)
        for s in range(len(bb)):
            counts = cg.resample_counts(n, bn, bb, 31, s)
            assert counts.sum(axis=1).tolist() == [int(bn[s])] * bb[s]
            want = orc.exact_rounded_contract(L, counts, rr1[s])
            assert np.array_equal(got[s], want), (n, s)
    ph = cg.scaleboot_all(L, bn, bb, 31, rng_mode=cg.RNG_PHILOX, sum_mode=cg.SUM_INT8)
    ref = cg.scaleboot_all(L, bn, bb, 31, rng_mode=cg.RNG_PHILOX, sum_mode=cg.SUM_FAST)
    for s in range(len(bb)):
        assert np.abs(ph[s] - ref[s]).max() <= 1e-9 * np.abs(ref[s]).max()


def test_equal_fill_chunks_same_result(cg, orc):
    """Chunk lengths that are not powers of two (equal-fill, jump polynomials x^(len*2^r) built on
    first use): the same blocks as with tabulated power-of-two chunks, strict mode == oracle,
    shards == slices, staged == unstaged."""
    import ctypes as C
    from consel_b200.api import load
    lib = load()
    lib.cgpu_last_chunk_len.restype = C.c_ulonglong
    K, n = 7, 1300
    L = synth(K, n, seed=21)
    rr = [0.5, 0.8, 1.0, 1.3]
    bb = [300, 211, 150, 97]
    _, bn, _ = orc.scale_setup(L, rr)
    try:
        cg.set_option("chunk_fill", 0)
        ref8 = cg.scaleboot_all(L, bn, bb, 5, sum_mode=cg.SUM_INT8)
        assert lib.cgpu_last_chunk_len() & (lib.cgpu_last_chunk_len() - 1) == 0
        cg.set_option("chunk_fill", 2)
        got8 = cg.scaleboot_all(L, bn, bb, 5, sum_mode=cg.SUM_INT8)
        clen = lib.cgpu_last_chunk_len()
        assert clen & (clen - 1) != 0, clen  # really not a power of two
        for s in range(len(bb)):
            assert np.array_equal(got8[s], ref8[s]), s
        strict = cg.scaleboot_all(L, bn, bb, 5, sum_mode=cg.SUM_STRICT)
        want = orc.makermt(L, bn, bb, 5)
        for s in range(len(bb)):
            assert np.array_equal(strict[s], want[s]), s
        cg.set_option("stage", 1)
        staged = cg.scaleboot_all(L, bn, bb, 5, sum_mode=cg.SUM_INT8)
        cg.set_option("stage", -1)
        for s in range(len(bb)):
            assert np.array_equal(staged[s], ref8[s]), s
        for rank in range(3):
            cg.set_shard(rank, 3)
            loc = cg.scaleboot_all(L, bn, bb, 5, sum_mode=cg.SUM_INT8)
            for s in range(len(bb)):
                b0, b1 = cg.shard_range(bb[s])
                assert np.array_equal(loc[s], ref8[s][:, b0:b1]), (rank, s)
        cg.set_shard(0, 1)
        # a non-zero stream offset (makermt -g continues the stream across files)
        a = cg.scaleboot_all(L, bn, bb, 5, sum_mode=cg.SUM_INT8, stream_offset=123457)
        cg.set_option("chunk_fill", 0)
        b = cg.scaleboot_all(L, bn, bb, 5, sum_mode=cg.SUM_INT8, stream_offset=123457)
        for s in range(len(bb)):
            assert np.array_equal(a[s], b[s]), s
    finally:
        cg.set_shard(0, 1)
        cg.set_option("stage", -1)
        cg.set_option("chunk_fill", -1)


def test_site_pattern_compression_is_bit_identical(cg, orc, gold):
    """mam15 (PAML input) has 1054 distinct columns among its 3414 sites: the int8 path contracts over pattern
    slots (count rows and GEMM inner dimension 3x shorter) and must return the same bits as the
    uncompressed contraction, for both generators, for a sharded run, and when a later call with
    r > 4 cannot use the slots (site counts of a slot could pass 255) and falls back."""
    L = orc.read_lnf(gold / "mam15.lnf")
    K, n = L.shape
    rr, bb = [0.5, 1.0, 1.4], [300, 257, 64]
    _, bn, _ = orc.scale_setup(L, rr)
    m = cg.Matrix(L)
    try:
        for mode in (cg.RNG_MT19937, cg.RNG_PHILOX):
            cg.set_option("i8_patterns", 0)
            plain = m.scaleboot(bn, bb, 5, rng_mode=mode, sum_mode=cg.SUM_INT8)
            want = plain.download()
            plain.free()
            assert m.info()["patterns"] == 0
            m.free()
            m = cg.Matrix(L)
            cg.set_option("i8_patterns", -1)
            comp = m.scaleboot(bn, bb, 5, rng_mode=mode, sum_mode=cg.SUM_INT8)
            got = comp.download()
            comp.free()
            npat = m.info()["patterns"]
            assert 1054 <= npat <= 1054 + n // 32 + 8, npat  # distinct columns, plus the extra slots of patterns with > 32 sites
            for s in range(3):
                assert np.array_equal(got[s], want[s]), (mode, s)
        # sharded over two ranks: the same columns
        for rank in range(2):
            cg.set_shard(rank, 2)
            part = m.scaleboot(bn, bb, 5, rng_mode=cg.RNG_PHILOX, sum_mode=cg.SUM_INT8)
            blocks = part.download()
            part.free()
            for s in range(3):
                b0, b1 = cg.shard_range(bb[s])
                assert np.array_equal(blocks[s], want[s][:, b0:b1])
        cg.set_shard(0, 1)
        # r = 5: the slots are dropped for this matrix, results equal the exact sums of the oracle's counts
        _, bn5, rr5 = orc.scale_setup(L, [5.0])
        big = m.scaleboot(bn5, [40], 9, sum_mode=cg.SUM_INT8)
        got5 = big.download()[0]
        big.free()
        assert m.info()["patterns"] == 0
        counts = cg.resample_counts(n, bn5, [40], 9, 0)
        assert np.array_equal(got5, orc.exact_rounded_contract(L, counts, rr5[0]))
    finally:
        cg.set_option("i8_patterns", -1)
        cg.set_shard(0, 1)
        m.free()
