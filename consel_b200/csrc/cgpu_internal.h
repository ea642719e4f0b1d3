// Internal declarations shared by the translation units of libconselgpu.so.
// Nothing here is part of the C ABI (include/consel_gpu.h is).
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/consel_gpu.h"

namespace cgpu {

// ---------------------------------------------------------------- errors
void set_error(int code, const char *fmt, ...);
extern std::atomic<long long> g_launches;

#define CGPU_CUDA(call)                                                                      \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            cgpu::set_error(CGPU_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,           \
                            cudaGetErrorString(e__));                                        \
            return CGPU_ECUDA;                                                               \
        }                                                                                    \
    } while (0)

#define CGPU_LAUNCH_CHECK()                                                                  \
    do {                                                                                     \
        cgpu::g_launches.fetch_add(1, std::memory_order_relaxed);                            \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess) {                                                            \
            cgpu::set_error(CGPU_ECUDA, "%s:%d kernel launch: %s", __FILE__, __LINE__,       \
                            cudaGetErrorString(e__));                                        \
            return CGPU_ECUDA;                                                               \
        }                                                                                    \
    } while (0)

#define CGPU_TRY(expr)                                                                       \
    do {                                                                                     \
        int rc__ = (expr);                                                                   \
        if (rc__ != CGPU_OK) return rc__;                                                    \
    } while (0)

// ---------------------------------------------------------------- MT19937
constexpr int MT_N = 624;
constexpr int MT_M = 397;
constexpr int MT_DEG = 19937;
constexpr int MT_POLY_W64 = 312;  // 64-bit words holding a residue mod phi (19968 bits)
constexpr int MT_POLY_W32 = 624;
constexpr int MT_MAXLOG = 48;     // jump table covers x^(2^i), i = 0..MT_MAXLOG

struct Gf2Poly {
    uint64_t w[MT_POLY_W64];
};

// x^(2^i) mod phi for i = 0..MT_MAXLOG (host table, built once)
const Gf2Poly *mt_jump_table();
// x^J mod phi for arbitrary J (square and multiply-by-x)
void mt_poly_xpow(uint64_t J, Gf2Poly *out);
// seeded state array of init_genrand (window W_0 = x_0..x_623)
void mt_seed_window(uint32_t seed, uint32_t *w624);
// host application of a jump polynomial to a window (tests / stream_offset set-up)
void mt_jump_host(const uint32_t *in624, const Gf2Poly *g, uint32_t *out624);

// ---------------------------------------------------------------- globals
struct Globals {
    bool inited = false;
    int device = -1;
    int sm_count = 0;
    int shard_rank = 0, shard_world = 1;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // aux stream of the resample/contract software pipeline
    cudaStream_t stream3 = nullptr;  // copy stream: D2H of finished scales behind the compute
    uint32_t *d_jump_table = nullptr;  // (MT_MAXLOG+1) x MT_POS_STRIDE byte offsets of the set bits of x^(2^i) mod phi
    int jump_nquads[64] = {0};
    cudaEvent_t ev[8] = {};
    double last_ms[12] = {0};
    // grow-only scratch arena for the large per-call workspaces of the resampling path (index
    // stream, count buffers, int32 partial sums): they keep their addresses between calls, so
    // interleaved smaller allocations cannot fragment the pool and force a re-map of GBs
    void *scratch[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t scratch_cap[4] = {0, 0, 0, 0};
    // tuning switches (cgpu_set_option); -1 = built-in default
    int opt_overlap = -1, opt_hist_algo = -1, opt_contract_algo = -1, opt_i8_oneshot = -1, opt_draw_nch = -1, opt_chunk_log2 = -1, opt_i8_fuse = -1, opt_draw_fuse = -1, opt_stage = -1, opt_i8_dbg = -1, opt_mc_smooth = -1, opt_chunk_fill = -1, opt_ktime = -1, opt_hist_wide = -1, opt_i8_2cta = -1, opt_batch = -1, opt_dist_upload = -1, opt_i8_panel = -1, opt_i8_patterns = -1, opt_i8_ksplit = -1, opt_hist_bucket = -1;
};
Globals &G();
int ensure_init();

// stream-ordered device allocations from the default pool (recycled across calls)
template <typename T>
inline cudaError_t dev_alloc(T **p, size_t bytes) {
    return cudaMallocAsync(reinterpret_cast<void **>(p), bytes ? bytes : 16, G().stream);
}
inline void dev_free(void *p) {
    if (p) cudaFreeAsync(p, G().stream);
}

// ---------------------------------------------------------------- layouts
// Site axis padded to a multiple of 128 so every pipeline chunk is full.
inline int pad_sites(int n) { return (n + 127) / 128 * 128; }
// trees per lane for the contraction kernels (tile of KT = 32*TK trees)
inline int pick_tk(int K) { return K <= 32 ? 1 : (K <= 64 ? 2 : (K <= 128 ? 4 : 8)); }

struct DevMatrix {  // cgpu_matrix
    int K = 0, n = 0, NP = 0, TK = 0, KT = 0, ktiles = 0;
    bool finite = true;
    double *d_L = nullptr;   // K x n row-major (as uploaded)
    double *d_LT = nullptr;  // [ktiles][NP][KT] site-major tiles, zero padded (fp64 paths; built on first use)
    cudaEvent_t ready = nullptr;  // recorded behind the upload of d_L: streams that read it wait on this first
    // CGPU_SUM_INT8 (built lazily by fx_prepare): L = M * 2^-fxF, M = sum_s digit_s * 256^s
    signed char *d_A8 = nullptr;  // [fxMpad][NP] int8 digits, row k*fxS+s
    int fxS = 0, fxF = 0;
    size_t fxMpad = 0;
    // site-pattern compression of the int8 path: identical columns of L (PAML-style inputs:
    // mam15 has 1054 distinct columns of 3414) share ONE column of A8, the histogram counts into pattern
    // slots.  The integer sums are exact, so the result is bit-identical to the uncompressed contraction.
    int npat = 0;                  // pattern slots (0: no compression)
    int NPc = 0;                   // inner dimension of the GEMM: pad_sites(npat), or NP
    unsigned int *d_pid = nullptr; // [n] site -> slot
};

struct DevReps {  // cgpu_reps
    int K = 0, S = 0;
    int *bb = nullptr;       // host: full replicate counts per scale
    int *b0 = nullptr;       // host: first column owned by this rank
    int *nb = nullptr;       // host: number of columns owned by this rank
    double **d_rep = nullptr;   // host array of device pointers, K x nb[s]
    // statistics (cgpu_stats_build)
    int cm = 0;
    double **d_stat = nullptr;  // cm x nb[s]
    void *d_edges = nullptr;    // (unused)
    // host cache of the row edges (min1, min2, max1, max2 per (item, scale)); filled by the first
    // count query after a statistics build, so later queries skip the edge tracking
    mutable std::vector<double> h_edges;
    int *d_ass_ptr = nullptr, *d_ass_idx = nullptr;
};

// ---------------------------------------------------------------- kernels (host launchers)
// mt_kernels.cu
constexpr int MT_POS_STRIDE = 9984;   // u32 words reserved per position list (two positions per word, MT_DEG rounded up to 8)
int mt_poly_positions(const Gf2Poly *g, uint32_t *out_stride);  // returns the number of 8-position groups
int launch_mt_jump(uint32_t *d_states, const int2 *d_pairs, const uint32_t *d_pos, int nocts, int count,
                   cudaStream_t st);
// site indices are stored as uint16 when they fit (halves the index traffic); n == 0 asks
// for raw tempered 32-bit words
inline bool idx_is16(uint32_t n) { return n > 0 && n <= 65536; }
inline size_t idx_bytes(uint32_t n) { return idx_is16(n) ? 2 : 4; }
struct DrawPlan {
    int nranges, nintervals;
    unsigned long long lo[32], hi[32], off[32];  // stream ranges to keep -> index-buffer offsets
    unsigned long long first[32];                // chunk intervals to run
    int cum[33];
    int slot[32];                                // state slot of the first chunk of each interval
    unsigned long long stream_end;               // draws beyond this are never generated
};
// chunk start states of one MT19937 stream: only the chunks that cover the ranges a rank
// keeps (plus the inner nodes of the jump tree that leads to them) are materialised
struct MtPlan {
    uint32_t *d_states = nullptr;  // nslots x 624 words
    int2 *d_pairs = nullptr;       // (src slot, dst slot) of every jump, level by level
    unsigned long long chunk_len = 0;  // draws per chunk: a power of two (tabulated jump polynomials) or an
                                       // equal-fill length with its own cached polynomials x^(chunk_len*2^r)
    const uint32_t *d_pos = nullptr;   // device position lists of the jump polynomials, level r at r*MT_POS_STRIDE
    const int *nquads = nullptr;       // host: 8-position groups per level
    int nintervals = 0;
    unsigned long long first[32];
    int count[32], slot[32];
    std::vector<uint32_t> h_stage;  // host staging of the seed window and the pair lists
    std::vector<int> level_cnt;     // jumps per tree level (index = bit of the chunk number)
    int nslots = 0, root = 0;
    size_t npairs = 0;
    long long slot_of(unsigned long long chunk) const {
        for (int i = 0; i < nintervals; i++)
            if (chunk >= first[i] && chunk < first[i] + (unsigned long long)count[i]) return slot[i] + (long long)(chunk - first[i]);
        return -1;
    }
};
int make_draw_plan(int nranges, const unsigned long long *lo, const unsigned long long *hi,
                   const unsigned long long *off, unsigned long long chunk_len, unsigned long long stream_end, DrawPlan *plan);
int resolve_draw_slots(const MtPlan &mt, DrawPlan *plan);
// fused draws + site histograms (byte rows of NP sites; needs bn >= 1248 and NP small enough for shared memory)
bool mt_draw_hist_fits(int NP);
int launch_mt_draw_hist(const MtPlan &mt, const DrawPlan &plan, const unsigned int *bn_per_range,
                        unsigned char *const *row0_per_range, int NP, uint32_t n, int *d_err, cudaStream_t st);
int launch_mt_draw(const MtPlan &mt, const DrawPlan &plan, uint32_t n, void *d_idx, cudaStream_t st);
int launch_mt_draw_range(const MtPlan &mt, unsigned long long lo, unsigned long long hi, uint32_t n, void *d_idx,
                         cudaStream_t st);
// resample_kernels.cu
// counts layout: rows == false -> "BJ8" words (8 replicates interleaved per site);
//                rows == true  -> plain uint8 rows [replicate][NP] (tensor-core operand)
// pid (optional, byte rows only): site -> pattern slot table; NP is then the padded slot count
int launch_hist_from_idx(const void *d_idx, int n, int NP, int bn, int nb,
                         unsigned long long *d_counts, int *d_err, cudaStream_t st, bool rows = false,
                         const unsigned int *pid = nullptr);
int launch_hist_philox(unsigned long long seed, int scale, int b_first, int n, int NP, int bn, int nb,
                       unsigned long long *d_counts, int *d_err, cudaStream_t st, bool rows = false,
                       unsigned long long stream_offset = 0, const unsigned int *pid = nullptr);
int launch_counts_unpack(const unsigned long long *d_counts, int n, int NP, int nb,
                         unsigned char *d_out, cudaStream_t st);
// contract_kernels.cu
int launch_transpose_L(const double *d_L, int K, int n, int NP, int TK, double *d_LT, int *d_nonfinite,
                       cudaStream_t st);
int launch_contract(const DevMatrix &m, const unsigned long long *d_counts, int nb, double r,
                    double *d_rep, int ldrep, int col0, int sum_mode, cudaStream_t st);
// i8_kernels.cu
// allow_patterns: look for identical columns and compress (a matrix prepared compressed is re-sliced
// uncompressed when a later call cannot use the compression)
int fx_prepare(DevMatrix &m, cudaStream_t st, bool allow_patterns = false);
// does the tensor-core path need the int32 partial-sum buffer (recombination outside the GEMM epilogue)?
bool fx_needs_partial(const DevMatrix &m, int max_draws, int max_nb_pad);
struct PhaseTimer;
int launch_contract_i8(const DevMatrix &m, const unsigned char *d_counts_rows, int nb, int nb_pad, int *d_partial,
                       double r, double *d_rep, int ldrep, int col0, cudaStream_t st, PhaseTimer *pt = nullptr,
                       int max_draws = 1 << 30 /* draws per replicate (bounds the digit sums) */);

// accumulates device time per phase over many (start, stop) event pairs
struct PhaseTimer {
    struct Span {
        int phase;
        cudaEvent_t a, b;
        bool ended;
    };
    static constexpr int NPHASE = 12;  // 0..7 as cgpu_last_timing documents, 8 mt_jump, 9 mt_draw (their sum is slot 1)
    Span *spans = nullptr;
    int nspans = 0, cap = 0;
    cudaStream_t st;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    cudaStream_t cur = nullptr;
    explicit PhaseTimer(cudaStream_t s);
    ~PhaseTimer();
    void begin(int phase, cudaStream_t on = nullptr);
    void end();
    void finish(double *ms);  // call after the streams have been synchronised
};

// comm.cu: the library's NCCL communicator (one process per GPU); every helper degenerates to a
// local copy when no communicator exists
int comm_world();
int comm_rank();
int comm_group_start();
int comm_group_end();
int comm_allgather_f64(const double *send, double *recv, size_t count_per_rank, cudaStream_t st);
int comm_allreduce_f64(const double *send, double *recv, size_t count, int op /*0 sum, 1 max, 2 min*/, cudaStream_t st);
int comm_reduce_scatter_f64(const double *send, double *recv, size_t recvcount, int op, cudaStream_t st);

// Per-kernel device times of the consel side (cgpu_set_option("ktime", 1); read and cleared by
// cgpu_kernel_times): a launcher brackets its launch with a KT object.  Off by default: no
// events are created and nothing is recorded.
struct KT {
    int slot = -1;
    cudaStream_t st;
    // work: the launch's ALGORITHMIC bytes (HBM-bound kernels) or fp64 instructions (FP64-bound ones)
    KT(const char *name, cudaStream_t s, double work = 0.0);
    ~KT();
};

// stats_kernels.cu
// Row definitions for the row-summary kernel:
//  mode 0: plain row            v = x[row*ld + b]
//  mode 1: KH difference        v = x[ja[row]*ld + b] - x[ia[row]*ld + b]
//  mode 2: singleton SH/stat    v = (i == a1[b] ? m2[b] : m1[b]) - x[i*ld + b], i = ia[row]
struct RowSrc {
    const double *x;
    int ld;
    const int *ia, *ja;
    const double *m1, *m2;
    const int *a1;
};
int launch_rowsum(int mode, const RowSrc &src, int rows, int B, int nq, const double *d_thr, cgpu_rowsum *d_out,
                  int out_stride, int out_off, cudaStream_t st);
int launch_rowsum_all_scales(const double *const *d_stat, const int *B, int S, int cm, int nq, const double *d_thr,
                             cgpu_rowsum *d_part, cgpu_rowsum *d_out, int nsplit, cudaStream_t st, bool edges = true,
                             int row0 = 0 /* cm = number of rows queried, starting at this row of the blocks */);
int launch_rowsum_planes(const cgpu_rowsum *d_q, size_t n, size_t padded, double *d_planes, cudaStream_t st);
int launch_rowsum_edges(const cgpu_rowsum *d_q, size_t nrs, int nq, double *d_edges, cudaStream_t st);
int launch_order_stat(const double *d_stat, int rows, int B, int ld, const int *d_rank, double *d_val, cudaStream_t st);
int launch_col_top2(const double *x, int K, int B, int ld, double *m1, double *m2, int *a1, cudaStream_t st);
int launch_stat_build(const double *x, int K, int B, int ld, const double *m1, const double *m2, const int *a1,
                      int cm, const int *ass_ptr, const int *ass_idx, double *stat, cudaStream_t st);
// xc = x - rowmean (rescale == false) or rs*(x - rowmean) + rowmean (consel -f); rowmean = sequential sum / B
int launch_center(const double *x, int K, int B, int ld, double *xc, cudaStream_t st, bool rescale = false,
                  double rs = 1.0);
int launch_pair_weight(const double *xc, int K, int B, int ld, double varadd, double *wt, int *nsmall, cudaStream_t st,
                       const int *d_rows = nullptr /* only these weight rows (device list) */, int nrows = 0);
int launch_wsh_single(const double *x, int K, int B, int ld, const double *wt, const int *mem_tree, int nmem, double *out,
                      cudaStream_t st);
int launch_sh_rows(bool weighted, const double *x, int K, int B, int ld, const double *wt, int nmem,
                   const int *mem_item, const int *mem_tree, const int *cass_ptr, const int *cass_idx, double *out,
                   cudaStream_t st);
int launch_weighted_obs(const double *dat, int K, const double *wt, int nmem, const int *mem_item,
                        const int *mem_tree, const int *cass_ptr, const int *cass_idx, double *obs, int *argj,
                        cudaStream_t st);

}  // namespace cgpu
