/*
 * consel_gpu.h — C ABI of libconselgpu.so: the B200 (sm_100a) implementation of
 * CONSEL's multiscale RELL bootstrap hot path.
 *
 * The reference (shimo-lab/consel) has no plugin/FFI interface; its drop-in
 * surface is the makermt/consel command lines + file formats, and, at source
 * level, the call sites named beside each entry point below (paths relative to
 * the reference tree).  A reference maintainer replaces exactly those loops with
 * these calls (INTEGRATION.md shows the patch); our own host programs
 * (consel_b200/host/) do the same.
 *
 * Conventions
 *   - plain C, no C++/torch types; every size is a 32-bit int like the
 *     reference's, matrices are caller-owned HOST memory unless the name says
 *     "dev"; the library owns all device memory behind opaque handles.
 *   - return 0 on success, a negative CGPU_E* code on failure;
 *     cgpu_last_error() then returns the process-wide message of the most
 *     recent failure (whichever thread it happened on).  The host
 *     programs turn a failure into the reference's `#! Terminated: ...` + exit(1)
 *     (src/misc.c:40-49).  There is NO CPU fallback: without a CUDA device every
 *     compute entry point fails with CGPU_ENODEV.
 *   - calls on one handle are serialised by the caller (the reference is
 *     single-threaded).
 */
#ifndef CONSEL_GPU_H
#define CONSEL_GPU_H

#ifdef __cplusplus
extern "C" {
#endif

#define CGPU_OK 0
#define CGPU_EINVAL (-1)  /* bad argument                                   */
#define CGPU_ENODEV (-2)  /* no usable CUDA device / wrong architecture      */
#define CGPU_ECUDA (-3)   /* CUDA runtime error (message has the details)    */
#define CGPU_ENOMEM (-4)  /* device or host allocation failed                */
#define CGPU_ERANGE (-5)  /* a site count exceeded the int8 range (tiny n)   */

/* rng_mode */
#define CGPU_RNG_MT19937 0 /* replays the reference's seeded stream (src/mt19937.c, src/rand.c:25-37) */
#define CGPU_RNG_PHILOX 1  /* counter-based Philox4x32-10 keyed by (seed; scale, replicate)           */
/* sum_mode */
#define CGPU_SUM_STRICT 0 /* reference order of operations: bit-identical to src/makermt.c:65-70     */
#define CGPU_SUM_FAST 1   /* fp64 FMA, same site order (<= 1 ulp per term from the reference)        */
#define CGPU_SUM_INT8 2   /* fixed-point int8 slices on tcgen05 tensor cores, exact integer sum      */

/* ------------------------------------------------------------------ lifecycle */
/* Select the device (device < 0: honour LOCAL_RANK, else device 0), create the
 * streams and build the MT19937 jump table.  Idempotent.                        */
int cgpu_init(int device);
void cgpu_shutdown(void);
const char *cgpu_last_error(void);
/* Library/ABI version (major*100+minor) — lets a host binary check what it loaded. */
int cgpu_version(void);

/* Replicate sharding for one-process-per-GPU runs: this process computes, for
 * every scale s, the replicate columns [ bb[s]*rank/world , bb[s]*(rank+1)/world ).
 * Default (0,1) = everything.  Results do not depend on the partition.           */
int cgpu_set_shard(int rank, int world);
int cgpu_shard_range(int bb, int *b_begin, int *b_end);

/* Host helper: datvec[i] = sum_j L[i][j] in the reference's left-to-right order
 * (src/makermt.c:159-162), rows spread over host threads.  No device involved.      */
void cgpu_host_rowsums(const double *L, int K, int n, double *datvec);

/* --------------------------------------------------------------- makermt side */
/* Replaces the loop `for(i<kk){repmat=new_lmat; scaleboot(datmat,repmat,mm,nn,
 * bn[i],bb[i]); fwrite_bmat;}` of genrmt (src/makermt.c:201-211) together with
 * scaleboot itself (src/makermt.c:35-75) and the process-global MT19937 stream
 * it consumes (src/mt19937.c:57-60, src/rand.c:25-29).
 *
 *   L          K*n doubles, row-major == datmat[0] of new_mat (src/misc.c:115-126)
 *   bn[s]      replicate sample size (int)(rr[s]*n), computed by the HOST exactly
 *              as src/makermt.c:165-168 does; must be positive (CGPU_EINVAL otherwise: a scale
 *              that draws nothing never reaches the device -- the caller stores the reference's
 *              0.0/0.0 for it, as consel_b200/host/makermt.cpp does)
 *   bb[s]      number of replicates of scale s
 *   seed       already time()-substituted (src/rand.c:33-37)
 *   stream_offset  uniform draws already consumed from the stream (makermt -g
 *              continues one stream across files); 0 otherwise
 *   rep_out[s] K*bb[s] doubles, row-major (tree-major, replicate fastest): the
 *              exact payload fwrite_bmat (src/misc.c:637-645) writes after its
 *              two-int header.  With sharding (cgpu_set_shard) rep_out[s] is the
 *              rank-local compact block K x (b_end-b_begin) of cgpu_shard_range(bb[s]).
 * With a communicator (cgpu_comm_init) the call is COLLECTIVE: every rank passes the same L,
 * copies only its share of the rows from the host, and one NCCL all-gather over NVLink
 * completes the matrix on every device ("dist_upload" 0 restores the plain per-rank upload).
 */
int cgpu_scaleboot_all(const double *L, int K, int n, int S, const int *bn, const int *bb,
                       unsigned long seed, unsigned long long stream_offset, int rng_mode,
                       int sum_mode, double *const *rep_out);

/* The same path in three steps, so a caller can keep data resident on the device
 * (bench `value`, and makermt|consel fusion a la src/makerep.c:304-376).          */
typedef struct cgpu_matrix cgpu_matrix; /* site-log-likelihood matrix on the device */
typedef struct cgpu_reps cgpu_reps;     /* replicate blocks (+ derived statistics)  */

int cgpu_matrix_upload(const double *L, int K, int n, cgpu_matrix **out);
void cgpu_matrix_free(cgpu_matrix *m);
int cgpu_scaleboot_dev(const cgpu_matrix *m, int S, const int *bn, const int *bb, unsigned long seed,
                       unsigned long long stream_offset, int rng_mode, int sum_mode, cgpu_reps **out);
/* copy this rank's replicate columns into rep_out[s] (compact K x nb[s], see above) */
int cgpu_reps_download(const cgpu_reps *r, double *const *rep_out);
/* columns [col0, col0+ncols) of this rank's block of scale s -> out (K x ncols, row-major):
 * spot checks of a large resident result without downloading all of it                      */
int cgpu_reps_download_cols(const cgpu_reps *r, int s, int col0, int ncols, double *out);
void cgpu_reps_free(cgpu_reps *r);

/* Test/diagnostic seam: site-count matrix of scale s as the resampler produced it,
 * counts[b*n + j] for this rank's replicates (bit-exact contract of the north star:
 * equals the histogram of src/makermt.c:53-62 under CGPU_RNG_MT19937).            */
int cgpu_resample_counts(int n, int S, const int *bn, const int *bb, unsigned long seed,
                         unsigned long long stream_offset, int rng_mode, int s, unsigned char *counts);
/* first `count` tempered outputs of the seeded MT19937 stream after `skip` draws,
 * produced by the device jump-ahead + block regeneration (src/mt19937.c:108-127) */
int cgpu_mt19937_stream(unsigned long seed, unsigned long long skip, unsigned long long count,
                        unsigned int *out);

/* ---------------------------------------------------------------- consel side */
/* Replaces fread_blmat'ed repmats (src/consel.c:808-818): upload K x bb[s] blocks. */
int cgpu_reps_upload(int K, int S, const int *bb, const double *const *rep, cgpu_reps **out);

/* Replaces genrep -> repminmaxs -> calcmaxs/calcassmins (src/consel.c:751-787,
 * 1894-1959): statistic blocks stat[s] = cm x bb[s], row e =
 * min_{i in A_e} ( max_{j != i} rep[j][b] - rep[i][b] ).  Associations in CSR form,
 * already permuted into the output item order (orderv, src/consel.c:514-530).     */
int cgpu_stats_build(cgpu_reps *r, int cm, const int *ass_ptr, const int *ass_idx);
/* Replaces the fread_blmat'ed statmats of rep mode (consel -R, do_repmode
 * src/consel.c:1195-1212): a handle over statistic blocks that already exist
 * (stat[s] = cm x bb[s] doubles row-major, sorted or not).  It has no replicate blocks:
 * count queries, order statistics and downloads work, cgpu_stats_build and cgpu_mctest
 * return CGPU_EINVAL.                                                                */
int cgpu_stats_upload(int cm, int S, const int *bb, const double *const *stat, cgpu_reps **out);
/* Replaces the rescaling branch of genrep (consel -f, src/consel.c:757-771) with
 * repminmaxsrs (:1961-1984): statistic blocks of S1 target scales, all from the replicates
 * of scale i0 of `src`: x' = rs[s]*(x - mean) + mean per tree row (mean = sequential sum / B,
 * three rounded operations), then the statistics of cgpu_stats_build.  rs[s] =
 * sqrt(rr[i0]/rr1[s]) is computed by the caller.  `out` is a statistics-only handle (like
 * cgpu_stats_upload) with S1 scales of bb[i0] replicates; `src` stays valid for cgpu_mctest. */
int cgpu_stats_build_rescaled(const cgpu_reps *src, int i0, int S1, const double *rs, int cm,
                              const int *ass_ptr, const int *ass_idx, cgpu_reps **out);
/* stat_out[s] = cm x nb[s] doubles row-major (nb[s] = bb[s] unless sharded); sorted != 0
 * sorts every row ascending like sort_vec (src/consel.c:1466-1471) — the payload of
 * `.rep` (unsharded only).                                                           */
int cgpu_stats_download(const cgpu_reps *r, int sorted, double *const *stat_out);

/* Summary of one row of replicates against a threshold t: everything cntdist1/2/3
 * (src/consel.c:1796-1860) need, without sorting.  For the sorted row v[0..B):
 *   cnt = #{v <= t};  lo = max{v <= t} (or -1e30);  hi = min{v > t} (or +1e30);
 *   min1,min2 = v[0],v[1];  max1,max2 = v[B-1],v[B-2].                             */
typedef struct cgpu_rowsum {
    double cnt, lo, hi, min1, min2, max1, max2;
} cgpu_rowsum;
/* cntdist(v,B,t,3) (src/consel.c:1842-1860) from the summary of row v at threshold t:
 * pure host helper (cntdist(...,2) is q->cnt itself).                               */
double cgpu_smooth_count(const cgpu_rowsum *q, int B, double t);
/* the same for `count` summaries of rows of equal length B at one threshold t */
void cgpu_smooth_count_many(const cgpu_rowsum *q, int count, int B, double t, double *out);

/* Replaces cntdist(statmats[s][i],bb[s],x,mode) inside vcalpval/invtpval/do_bptest/
 * do_bootrep (src/consel.c:2775-2778, 2845-2848, 1503-1507): nq thresholds per item,
 * thr[i*nq+q]; out[(i*S+s)*nq+q].                                                   */
int cgpu_count_query(const cgpu_reps *r, int nq, const double *thr, cgpu_rowsum *out);
/* The same for rows [row0, row0+nrows) only (thr: nrows x nq, out: nrows x S x nq), so a host
 * driver can fit one half of the items while the device counts for the other half.  The row
 * edges (min1..max2) come from a cache that the first whole-range query fills.          */
int cgpu_count_query_rows(const cgpu_reps *r, int row0, int nrows, int nq, const double *thr, cgpu_rowsum *out);
/* k-th smallest (0-based rank[i]) of stat[s][i][:] for every item — the median and
 * the min/max that vcalpval/invtpval start from (src/consel.c:2765, 2837).         */
int cgpu_order_stat(const cgpu_reps *r, int s, const int *rank, double *val);

/* Replaces do_mctest (src/consel.c:1352-1439): centring, calckhpval (:2111-2149),
 * calcmcpval/calcmaxass (:2065-2102, 2010-2033), pairwise variances (:1404-1415)
 * and their weighted versions, on scale i0.  Outputs cm-vectors of p-values.      */
int cgpu_mctest(const cgpu_reps *r, int i0, const double *datvec, double varadd, int cm,
                const int *ass_ptr, const int *ass_idx, double *kh, double *sh, double *wkh,
                double *wsh);

/* ------------------------------------------------------ multi-GPU */
/* One process per GPU.  The reference's only parallel recipe is one makermt process per scale
 * plus catrep (src/program.tex:1089-1109); here replicates (columns) are sharded over the
 * ranks and the consel side combines per-rank results over the library's own NCCL
 * communicator.  cgpu_comm_init also does cgpu_set_shard(rank, world).  The 128-byte id comes
 * from rank 0's cgpu_comm_unique_id and reaches the other ranks through the launcher, or
 * cgpu_comm_init_env does it all from RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (TCP on
 * MASTER_PORT+17, or CGPU_COMM_PORT).  With no communicator every call below is local.      */
int cgpu_comm_unique_id(char *id128);
int cgpu_comm_init(int rank, int world, const char *id128);
int cgpu_comm_init_env(void);
void cgpu_comm_destroy(void);
int cgpu_comm_info(int *rank, int *world);
/* items (rows of the statistic blocks) are owned in equal contiguous ranges of
 * ceil(cm/world) rows: this rank fits rows [row0, row0+nrows)                               */
int cgpu_comm_item_range(int cm, int *row0, int *nrows);
/* small host-buffer collectives for the results (staged through device memory):
 * recv = world x bytes; op 0 sum, 1 max, 2 min, in place                                    */
int cgpu_comm_allgather_host(const void *send, void *recv, long long bytes);
int cgpu_comm_allreduce_host(double *buf, long long count, int op);
int cgpu_comm_barrier(void);

/* COLLECTIVE.  All-gather of the rank-local column blocks of scale s into one K x bb[s] block on
 * every rank (NCCL all-gather + strided device copies): a new single-scale, unsharded handle.
 * The KH/SH tests and the median of the AU start value need that scale whole
 * (src/consel.c:1374-1415: sequential row sums over all replicates).                        */
int cgpu_reps_gather_scale(const cgpu_reps *r, int s, cgpu_reps **whole);

/* cgpu_mctest for the items [e0, e1) only (the other entries of kh/sh/wkh/wsh are left
 * untouched): the O(K^2 B) loops shard over items when every rank holds the whole scale.
 * cgpu_mctest itself, called on a column-sharded handle with a communicator, does
 * gather + item range + all-reduce of the four vectors (COLLECTIVE in that case).            */
int cgpu_mctest_rows(const cgpu_reps *r, int i0, const double *datvec, double varadd, int cm,
                     const int *ass_ptr, const int *ass_idx, int e0, int e1, double *kh, double *sh,
                     double *wkh, double *wsh);

/* COLLECTIVE count query over column-sharded statistic blocks (one fit round of the AU test):
 *   all_rows != 0: thr is cm x nq, identical on every rank; out = cm x S x nq on every rank.
 *   all_rows == 0: thr is nrows x nq for THIS rank's items (cgpu_comm_item_range); out =
 *                  nrows x S x nq: the summaries of its items over ALL replicates.
 * nq must be the same on every rank.  Each rank counts every item over its own columns
 * (thresholds all-gathered on the device), the integer counts add, lo is a max, hi a min
 * (one grouped NCCL reduce[-scatter]): exact and order independent, hence bit-identical to
 * the single-device result.  `active` is this rank's "I still iterate" flag; *any_active
 * returns the OR over all ranks from the same exchange, so converged ranks know how long to
 * keep serving rounds.  Without a communicator this is cgpu_count_query.                     */
int cgpu_count_query_dist(const cgpu_reps *r, int all_rows, int nq, const double *thr, cgpu_rowsum *out,
                          int active, int *any_active);

/* -------------------------------------------------------------- introspection */
/* Device time (ms, CUDA events on the library's streams) of the last cgpu_scaleboot_*
 * call: [0] upload, [1] rng/jump, [2] resample, [3] contraction (all kernels),
 * [4] download, [5] whole call (first to last event, includes launch gaps),
 * [6] tensor-core GEMM only, [7] fixed-point recombination only (CGPU_SUM_INT8).   */
int cgpu_last_timing(double *ms8);
/* the same slots followed by finer ones: [8] MT19937 jump tree, [9] MT19937 draw kernel
 * ([1] is their sum), [10] / [11] the number of tensor-core GEMM / histogram launches; slots past
 * the last defined one read 0                          */
int cgpu_last_timing_ex(double *ms, int n);
/* Tuning switches (development / measurement): "overlap" 0|1 (two-stream pipeline of
 * resampling and contraction), "hist_algo" 0|1|2, "contract_algo" -1|0|1|2,
 * "i8_oneshot" 0|1, "i8_fuse" 0|1 (recombination in the GEMM epilogue), "draw_nch" 1|2|3,
 * "draw_fuse" 0|1 (draws + histograms in one kernel, int8 mode, opt-in), "stage" 0|1 (staged
 * index production of host-output int8 calls: 0 off, 1 always, default by size),
 * "hist_wide" 0|1 (one replicate per CTA
 * on long alignments), "hist_bucket" 0|1|2 (byte rows longer than 225280 sites: 1 = partition the draws by
 * 65536-site bucket and count in shared memory instead of global atomics, opt-in; 2 = that path for every
 * uint32-index row, tests), "i8_2cta" 0|1 (cta_group::2 GEMM, default on), "i8_patterns" 0|1 (site-pattern compression of the int8 path,
 * default on), "i8_panel" N (tile order of the pair GEMM), "dist_upload" 0|1, "batch" N (replicates per
 * GEMM batch, tests), "ktime" 0|1 (per-kernel events, see cgpu_kernel_times), "chunk_log2" 10..30 (force a power-of-two chunk of the MT19937 jump tree), "chunk_fill" 0|1|2
 * (equal-fill chunk lengths that are not powers of two: off, by cost model (default), forced), "i8_dbg" 0..3 (switches parts of the GEMM epilogue off: timing only,
 * results are wrong), "mc_smooth" 0|1 (0 = consel --quick_mctest: the KH/SH p-values of
 * cgpu_mctest use the plain count #{v<=t} instead of the smoothed one, sw_pvaldist
 * src/consel.c:1875-1884).  A negative value restores the default.  Returns CGPU_EINVAL for an
 * unknown name.  Environment (development): CGPU_TRACE=1 prints every timed span of a
 * resampling call, CGPU_STAGE_FRACS=a,b,c overrides the stage sizes.                    */
int cgpu_set_option(const char *name, int value);
/* With cgpu_set_option("ktime", 1) every kernel launch of the consel side (statistics, count
 * queries, order statistics, KH/SH tests) is bracketed by CUDA events; this call synchronises the
 * device, writes one line "kernel_name launches total_ms algorithmic_work" (bytes for the streaming kernels,
 * fp64 instructions for the FP64-bound ones, summed over the launches) per kernel into buf and clears the
 * record (bench.py's per-kernel rooflines).                                              */
int cgpu_kernel_times(char *buf, int len);
/* shape of an uploaded matrix; slices/frac_bits/mpad describe the fixed-point split of
 * CGPU_SUM_INT8 (0 before the first int8 call): L = M * 2^-frac_bits, M in `slices` signed
 * base-256 digits, mpad = GEMM rows (slices*K rounded up to 128), patterns = number of column slots
 * when identical columns of L share one (site-pattern compression, 0 = none).  Any pointer may be NULL. */
int cgpu_matrix_info(const cgpu_matrix *m, int *K, int *n, int *NP, int *slices, int *frac_bits, long long *mpad,
                     int *patterns);
/* measured roofline denominators (MEASURED_PEAKS.json carries HBM and bf16 only): dense int8
 * tcgen05.mma issue rate from shared-memory-resident operands (TOP/s), FP64 DFMA rate (TFLOP/s) */
int cgpu_bench_i8_peak(double *tops);
/* the same loop back to back for `seconds` (<= 30), rate of the second half: the clock a long
 * tensor-bound step settles at under the board's power limit                                */
int cgpu_bench_i8_peak_sustained(double seconds, double *tops);
int cgpu_bench_fp64_peak(double *tflops);
/* draws per chunk of the most recent MT19937 jump plan (diagnostics / tests) */
unsigned long long cgpu_last_chunk_len(void);
/* number of kernels launched by this library since cgpu_init (bench `gpu_launches`) */
long long cgpu_launch_count(void);
/* raw device pointer + shape of a replicate block, so a host framework can wrap it
 * (e.g. for an NCCL all-gather) without a copy: K x ld doubles, this rank's columns
 * [b_begin,b_end) at column offset 0.                                              */
int cgpu_reps_devptr(const cgpu_reps *r, int s, void **ptr, int *rows, int *cols, int *ld);

#ifdef __cplusplus
}
#endif
#endif /* CONSEL_GPU_H */
