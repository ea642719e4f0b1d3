// makerep — drop-in command line of the reference's makerep (src/makerep.c): makermt's
// resampling and consel's replicate statistics in one process, `.rep` out, no `.rmt` in
// between.  The reference does it one replicate at a time (do_makerep :304-331:
// scaleboot1 -> calcmaxs -> calcassmins); here the replicate blocks are produced on the
// device (cgpu_scaleboot_dev), stay there, and the statistics kernels of consel
// (cgpu_stats_build) read them in place.  The MT19937 stream is consumed in the same
// scale-major, replicate-major order, so the statistic rows equal the reference's
// bit for bit in the default strict mode.
//
// Extensions (not in the reference): --rng mt|philox, --sum strict|fast|int8, --gpu N.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "../../include/consel_gpu.h"
#include "assoc.h"
#include "refio.h"

using namespace refio;

enum SeqMode { SEQ_MT, SEQ_MOLPHY, SEQ_PAML, SEQ_PAUP, SEQ_PUZZLE };

static void byebye() { die("error in command line"); }
#define CG(call)                                             \
    do {                                                     \
        if ((call) != CGPU_OK) die("%s", cgpu_last_error()); \
    } while (0)

int main(int argc, char **argv) {
    const char *fname_pa = nullptr, *fname_mt = nullptr, *fname_rep = nullptr, *fname_ass = nullptr;
    unsigned long seed = 0;
    SeqMode seqmode = SEQ_MT;
    int rng_mode = CGPU_RNG_MT19937, sum_mode = CGPU_SUM_STRICT, gpu = -1;

    printf("# consel_b200 makerep (sm_100a); command line of makerep.c,v 1.5");

    for (int i = 1, j = 1; i < argc; i++) {  // src/makerep.c:99-131
        const std::string a = argv[i];
        auto need = [&]() { if (i + 1 >= argc) byebye(); };
        if (argv[i][0] != '-') {
            if (j == 1) fname_mt = argv[i];
            else if (j == 2) fname_rep = argv[i];
            else byebye();
            j++;
        } else if (a == "-s") {
            need();
            if (sscanf(argv[++i], "%lu", &seed) != 1) byebye();
        } else if (a == "-p") { need(); fname_pa = argv[++i]; }
        else if (a == "-a") { need(); fname_ass = argv[++i]; }
        else if (a == "--molphy") seqmode = SEQ_MOLPHY;
        else if (a == "--paml") seqmode = SEQ_PAML;
        else if (a == "--paup") seqmode = SEQ_PAUP;
        else if (a == "--puzzle") seqmode = SEQ_PUZZLE;
        else if (a == "--rng") {
            need();
            const std::string v = argv[++i];
            if (v == "mt" || v == "mt19937") rng_mode = CGPU_RNG_MT19937;
            else if (v == "philox") rng_mode = CGPU_RNG_PHILOX;
            else byebye();
        } else if (a == "--sum") {
            need();
            const std::string v = argv[++i];
            if (v == "strict") sum_mode = CGPU_SUM_STRICT;
            else if (v == "fast") sum_mode = CGPU_SUM_FAST;
            else if (v == "int8") sum_mode = CGPU_SUM_INT8;
            else byebye();
        } else if (a == "--gpu") {
            need();
            if (sscanf(argv[++i], "%d", &gpu) != 1) byebye();
        } else byebye();
    }

    // smrand (src/rand.c:33-37): seed 0 means "seed from the clock"
    const unsigned long seed_used = seed ? seed : (unsigned long)time(nullptr);

    std::string used;
    std::vector<double> rr = {0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4};  // src/makerep.c:43-45
    std::vector<int> bb(10, 10000);
    if (fname_pa) {  // :137-146
        FILE *fp = open_file(fname_pa, ".pa", "r", &used);
        printf("\n# reading %s", used.c_str());
        rr = read_vec(fp);
        putchar('.');
        bb = read_ivec(fp, (int)rr.size());
        putchar('.');
        fclose(fp);
    }
    const int kk = (int)rr.size();
    printf("\n# seed:%lu", seed);
    printf("\n# K:%d", kk);
    printf("\n# R:");
    for (int i = 0; i < kk; i++) printf("%g ", rr[i]);
    printf("\n# B:");
    for (int i = 0; i < kk; i++) printf("%d ", bb[i]);

    const char *fext = ".mt";
    switch (seqmode) {  // src/freadmat.h:13-18
        case SEQ_MOLPHY: fext = ".lls"; break;
        case SEQ_PAML: fext = ".lnf"; break;
        case SEQ_PAUP: fext = ".txt"; break;
        case SEQ_PUZZLE: fext = ".sitelh"; break;
        default: break;
    }
    FILE *fp = stdin;
    if (fname_mt) {
        fp = open_file(fname_mt, fext, "r", &used);
        printf("\n# reading %s", used.c_str());
    } else {
        printf("\n# reading from stdin");
    }
    int mm = 0, nn = 0;
    std::vector<double> L;
    switch (seqmode) {
        case SEQ_MOLPHY: L = read_mat_molphy(fp, &mm, &nn); break;
        case SEQ_PAML: L = read_mat_paml(fp, &mm, &nn); break;
        case SEQ_PAUP: L = read_mat_paup(fp, &mm, &nn); break;
        case SEQ_PUZZLE: L = read_mat_puzzle(fp, &mm, &nn); break;
        default: L = read_mat(fp, &mm, &nn); break;
    }
    if (fname_mt) fclose(fp);
    printf("\n# M:%d N:%d", mm, nn);
    if (mm <= 0 || nn <= 0) die("empty matrix");

    std::vector<int> bn(kk);
    for (int i = 0; i < kk; i++) {  // :184-188
        bn[i] = (int)(rr[i] * nn);
        rr[i] = (double)bn[i] / nn;
    }
    std::vector<double> datvec(mm);
    for (int i = 0; i < mm; i++) {  // :191-195
        double x = 0;
        for (int j = 0; j < nn; j++) x += L[(size_t)i * nn + j];
        datvec[i] = x;
    }

    Assoc ass;
    std::vector<double> obsvec;
    std::vector<int> orderv;
    const int cm = read_assoc_and_order(fname_ass, datvec, 0, false, &ass, &obsvec, &orderv);

    // ---- header of the .rep (:275-291)
    std::string outbase;
    if (fname_mt && !fname_rep) {
        outbase = strip_extension(fname_mt);
        fname_rep = outbase.c_str();
    }
    FILE *out = nullptr;
    if (fname_rep) {
        out = open_file(fname_rep, ".rep", "wb", &used);
        printf("\n# writing %s", used.c_str());
        write_bivec(out, orderv.data(), cm);
        write_bvec(out, obsvec.data(), cm);
        write_bvec(out, rr.data(), kk);
        write_bivec(out, bb.data(), kk);
        write_bint(out, kk);
    } else {
        printf("\n# writing to stdout\n");
        write_ivec(stdout, orderv.data(), cm);
        write_vec(stdout, obsvec.data(), cm);
        write_vec(stdout, rr.data(), kk);
        write_ivec(stdout, bb.data(), kk);
        printf("%d\n", kk);
    }
    fflush(stdout);
    const double t0 = cpu_seconds();

    // ---- do_makerep on the device: replicate blocks -> statistic blocks, nothing in between on the host
    CG(cgpu_init(gpu));
    cgpu_matrix *mat = nullptr;
    cgpu_reps *reps = nullptr;
    CG(cgpu_matrix_upload(L.data(), mm, nn, &mat));
    // A scale with n' = (int)(r*n) = 0 draws nothing: every replicate sum of the reference is 0.0/0.0, every
    // comparison of calcmaxs / calcassmins with it is false, and the statistic stays at the 1e30 the minimum over an
    // association starts from (src/consel.c:1923-1937 via makerep.c:304-331).  Such scales consume no draws and
    // never reach the device.
    std::vector<int> bn_live, bb_live;
    std::vector<std::vector<double>> st(kk);
    std::vector<double *> ptrs;
    for (int k = 0; k < kk; k++) {
        st[k].assign((size_t)cm * (bb[k] > 0 ? bb[k] : 0) + 1, 1.0e30);
        if (bn[k] > 0) {
            bn_live.push_back(bn[k]);
            bb_live.push_back(bb[k]);
            ptrs.push_back(st[k].data());
        }
    }
    if (!bn_live.empty()) {
        CG(cgpu_scaleboot_dev(mat, (int)bn_live.size(), bn_live.data(), bb_live.data(), seed_used, 0, rng_mode, sum_mode, &reps));
        CG(cgpu_stats_build(reps, cm, ass.ptr.data(), ass.idx.data()));
        CG(cgpu_stats_download(reps, /*sorted=*/0, ptrs.data()));  // makerep writes the rows unsorted
    }
    for (int k = 0; k < kk; k++) {
        if (out) {
            write_bmat(out, st[k].data(), cm, bb[k]);
            putchar('.');
            fflush(stdout);
        } else {
            write_mat(stdout, st[k].data(), cm, bb[k]);
        }
    }
    const double t1 = cpu_seconds();
    if (out) fclose(out);
    printf("\n# time elapsed for bootstrap t=%g sec", t1 - t0);
    cgpu_reps_free(reps);
    cgpu_matrix_free(mat);
    printf("\n# exit normally\n");
    cgpu_shutdown();
    return 0;
}
