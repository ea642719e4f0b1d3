"""GPU parity tests of the consel side: replicate statistics (repminmaxs), sort-free
count queries (cntdist1/2/3), order statistics and the KH/SH/wKH/wSH tests, through
the C ABI, against the CPU oracle and the reference's golden .rep/.cnt/.pv files."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cg():
    import consel_b200 as cg
    cg.init()
    cg.set_shard(0, 1)
    return cg


def _case(orc, gold, name):
    rmt = orc.read_rmt(gold / ("g100.rmt" if name == "g100e" else f"{name}.rmt"))
    K = len(rmt["datvec"])
    ptr, idx = orc.read_ass(gold / "mam15.ass") if name == "g100e" else orc.identity_assoc(K)
    order, obs, sptr, sidx = orc.order_items(rmt["datvec"], ptr, idx)
    return rmt, sptr, sidx


@pytest.mark.parametrize("name", ["g100", "tiny", "g100e"])
def test_stat_blocks_bit_exact_and_rep_file_payload(cg, orc, gold, name):
    rmt, sptr, sidx = _case(orc, gold, name)
    reps = cg.Reps.upload(rmt["reps"])
    reps.stats_build(sptr, sidx)
    got = reps.stats_download()
    for s, r in enumerate(rmt["reps"]):
        assert np.array_equal(got[s], orc.repminmaxs(r, sptr, sidx)), s
    ref = orc.read_rep(gold / f"{name}.rep")
    got_sorted = reps.stats_download(sorted_rows=True)
    for s in range(len(got_sorted)):
        assert np.array_equal(got_sorted[s], ref["stats"][s]), s


@pytest.mark.parametrize("name", ["g100", "tiny", "g100e"])
def test_count_queries_match_cnt_file_and_smoothing(cg, orc, gold, name):
    rmt, sptr, sidx = _case(orc, gold, name)
    cnt = orc.read_cnt(gold / f"{name}.cnt")
    rep = orc.read_rep(gold / f"{name}.rep")
    reps = cg.Reps.upload(rmt["reps"])
    reps.stats_build(sptr, sidx)
    cm, S = reps.cm, reps.S
    q = reps.count_query(np.zeros((cm, 1)))
    assert np.array_equal(q["cnt"][:, :, 0], cnt["cnt"])  # integer counts of the .cnt file (mode 2)
    # data-dependent thresholds, several per item, incl. outside the range and exact ties
    rng = np.random.default_rng(3)
    i0 = orc.find_i0(rmt["rr"])
    row0 = rep["stats"][i0]
    thr = np.stack([row0[:, 0] - 1.0, row0[:, 0], row0[:, row0.shape[1] // 2], row0[:, -1], row0[:, -1] + 2.0,
                    rng.uniform(row0[:, 0], row0[:, -1]), np.zeros(cm)], axis=1)
    q = reps.count_query(thr)
    for i in range(cm):
        for s in range(S):
            B = int(rmt["bb"][s])
            for k in range(thr.shape[1]):
                t = float(thr[i, k])
                assert q["cnt"][i, s, k] == orc.cntdist(rep["stats"][s][i], t, 2)
                assert cg.smooth_count(q[i, s, k], B, t) == orc.cntdist(rep["stats"][s][i], t, 3), (i, s, k)


def test_order_statistics(cg, orc, gold):
    rmt, sptr, sidx = _case(orc, gold, "g100")
    rep = orc.read_rep(gold / "g100.rep")
    reps = cg.Reps.upload(rmt["reps"])
    reps.stats_build(sptr, sidx)
    cm = reps.cm
    for s in (0, 5):
        B = int(rmt["bb"][s])
        for rank in (np.zeros(cm, int), np.full(cm, B // 2), np.full(cm, B - 1), np.arange(cm) * 5 % B):
            got = reps.order_stat(s, rank)
            want = rep["stats"][s][np.arange(cm), rank]
            assert np.array_equal(got, want)


@pytest.mark.parametrize("name", ["g100", "tiny", "g100e"])
def test_mctest_matches_oracle_and_pv_file(cg, orc, gold, name):
    rmt, sptr, sidx = _case(orc, gold, name)
    pv = orc.read_pv(gold / f"{name}.pv")
    i0 = orc.find_i0(rmt["rr"])
    reps = cg.Reps.upload(rmt["reps"])
    got = reps.mctest(i0, rmt["datvec"], sptr, sidx)
    want = orc.mctest(rmt["reps"][i0], rmt["datvec"], sptr, sidx)
    for g, w, col in zip(got, want, (2, 3, 4, 5)):
        assert np.array_equal(g, w), (col, g, w)                     # same arithmetic, same bits
        assert np.allclose(g, pv["pv"][:, col], rtol=6e-9, atol=1e-15)  # .pv prints %14.9g


def test_bp_from_count_query_matches_pv_file(cg, orc, gold):
    rmt, sptr, sidx = _case(orc, gold, "g100")
    pv = orc.read_pv(gold / "g100.pv")
    i0 = orc.find_i0(rmt["rr"])
    reps = cg.Reps.upload(rmt["reps"])
    reps.stats_build(sptr, sidx)
    B = int(rmt["bb"][i0])
    q = reps.count_query(np.zeros((reps.cm, 1)))
    bp = np.array([1.0 - (1.0 - cg.smooth_count(q[i, i0, 0], B, 0.0) / B) for i in range(reps.cm)])
    assert np.allclose(bp, pv["pv"][:, 0], rtol=6e-9, atol=1e-15)


def test_synthetic_mixed_associations(cg, orc):
    rng = np.random.default_rng(11)
    K, B = 37, 333
    rep = [rng.standard_normal((K, B)) * 3 - np.arange(K)[:, None] * 0.1 for _ in range(3)]
    rep[1][5] = rep[1][6]  # exact ties between trees
    datvec = -np.sort(rng.exponential(size=K)) * 10
    # singleton, pair, large and overlapping sets
    sets = [[0], [1, 2], [3], list(range(4, 20)), [2, 7, 30], [36], [35, 0]]
    ptr = np.cumsum([0] + [len(s) for s in sets]).astype(np.int32)
    idx = np.concatenate(sets).astype(np.int32)
    reps = cg.Reps.upload(rep)
    reps.stats_build(ptr, idx)
    got = reps.stats_download()
    for s in range(3):
        assert np.array_equal(got[s], orc.repminmaxs(rep[s], ptr, idx))
    g = reps.mctest(1, datvec, ptr, idx, varadd=1.0)
    w = orc.mctest(rep[1], datvec, ptr, idx, 1.0)
    for a, b in zip(g, w):
        assert np.array_equal(a, b)
    # identity association on the same data takes the top-2 fast path
    iptr, iidx = orc.identity_assoc(K)
    g = reps.mctest(0, datvec, iptr, iidx)
    w = orc.mctest(rep[0], datvec, iptr, iidx)
    for a, b in zip(g, w):
        assert np.array_equal(a, b)


def test_consel_side_rejects_bad_input(cg, orc):
    rep = [np.zeros((4, 10))]
    reps = cg.Reps.upload(rep)
    with pytest.raises(cg.CgpuError):
        reps.count_query(np.zeros((4, 1)))  # statistics not built
    with pytest.raises(cg.CgpuError):
        reps.stats_build(np.array([0, 1]), np.array([9]))  # tree id out of range
    with pytest.raises(cg.CgpuError):
        reps.mctest(0, np.zeros(4), np.array([0, 2]), np.array([1, 1]))  # duplicate association


def test_count_query_rows_equals_slices_of_the_whole_query(cg, orc, gold):
    """cgpu_count_query_rows (the AU driver's half-range query) == the same rows of cgpu_count_query,
    for 1, 2, 5 and 10 thresholds per item, and it needs one whole-range query first (edge cache)."""
    rmt, sptr, sidx = _case(orc, gold, "g100")
    reps = cg.Reps.upload(rmt["reps"])
    reps.stats_build(sptr, sidx)
    cm = reps.cm
    rng = np.random.default_rng(9)
    with pytest.raises(cg.CgpuError):
        reps.count_query_rows(1, np.zeros((2, 1)))
    for nq in (1, 2, 5, 10):
        thr = rng.normal(0.0, 3.0, size=(cm, nq))
        thr[:, 0] = 0.0
        whole = reps.count_query(thr)
        for row0, rows in ((0, cm // 2), (cm // 2, cm - cm // 2), (1, 1)):
            part = reps.count_query_rows(row0, thr[row0:row0 + rows])
            assert part.tobytes() == whole[row0:row0 + rows].tobytes(), (nq, row0)
    with pytest.raises(cg.CgpuError):
        reps.count_query_rows(cm - 1, np.zeros((2, 1)))


@pytest.mark.parametrize("name", ["g100e", "tiny"])
def test_uploaded_statistics_serve_the_same_queries(cg, orc, gold, name):
    """cgpu_stats_upload (rep mode, src/consel.c:1195-1212): a handle over the rows of a .rep file
    answers count queries / order statistics exactly like one built from the replicates."""
    rmt, sptr, sidx = _case(orc, gold, name)
    rep = orc.read_rep(gold / f"{name}.rep")
    built = cg.Reps.upload(rmt["reps"])
    built.stats_build(sptr, sidx)
    up = cg.Reps.upload_stats(rep["stats"])
    assert up.cm == built.cm and up.S == built.S
    rng = np.random.default_rng(11)
    i0 = orc.find_i0(rmt["rr"])
    row0 = rep["stats"][i0]
    thr = np.stack([np.zeros(up.cm), row0[:, 0], row0[:, -1], rng.uniform(row0[:, 0], row0[:, -1])], axis=1)
    a, b = built.count_query(thr), up.count_query(thr)
    for f in a.dtype.names:
        assert np.array_equal(a[f], b[f]), f
    B = int(rmt["bb"][i0])
    rank = np.full(up.cm, B // 2)
    assert np.array_equal(up.order_stat(i0, rank), row0[:, B // 2])
    got = up.stats_download(sorted_rows=True)
    for s in range(up.S):
        assert np.array_equal(got[s], rep["stats"][s])
    with pytest.raises(cg.CgpuError):
        up.stats_build(sptr, sidx)


@pytest.mark.parametrize("name", ["g100", "g100e", "tiny"])
def test_rescaled_statistics_bit_exact(cg, orc, gold, name):
    """cgpu_stats_build_rescaled (consel -f, repminmaxsrs src/consel.c:1961-1984) against the oracle:
    sequential row mean, three rounded operations per element, then the ordinary statistics."""
    rmt, sptr, sidx = _case(orc, gold, name)
    i0 = orc.find_i0(rmt["rr"])
    rs = np.sqrt(rmt["rr"][i0] / np.array([0.3, 0.75, 1.0, 1.9, 4.0]))
    reps = cg.Reps.upload(rmt["reps"])
    scaled = reps.stats_build_rescaled(i0, rs, sptr, sidx)
    assert scaled.S == len(rs) and scaled.cm == len(sptr) - 1
    got = scaled.stats_download()
    for s in range(len(rs)):
        assert np.array_equal(got[s], orc.repminmaxs_rescaled(rmt["reps"][i0], sptr, sidx, float(rs[s]))), s
    # the source handle is untouched and still serves the ordinary build
    reps.stats_build(sptr, sidx)
    assert np.array_equal(reps.stats_download()[i0], orc.repminmaxs(rmt["reps"][i0], sptr, sidx))


def test_rescaled_statistics_long_rows(cg, orc):
    """Row means over more than one staging chunk per warp and a ragged tail (B = 1000 + 37)."""
    rng = np.random.default_rng(5)
    K, B = 23, 1037
    rep = -3000.0 + 40.0 * rng.standard_normal((K, B))
    ptr, idx = orc.identity_assoc(K)
    reps = cg.Reps.upload([rep])
    scaled = reps.stats_build_rescaled(0, np.array([1.25, 0.8]), ptr, idx)
    got = scaled.stats_download()
    for s, r in enumerate((1.25, 0.8)):
        assert np.array_equal(got[s], orc.repminmaxs_rescaled(rep, ptr, idx, r)), s
