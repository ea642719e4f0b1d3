// makermt — drop-in command line of the reference's makermt (src/makermt.c) whose
// resampling loop (genrmt's `for scale: scaleboot(); fwrite_bmat();`, :201-211) runs on
// the GPU through the C ABI (cgpu_scaleboot_all).  Everything else — argument parsing,
// .pa/.mt/.lnf/... readers, scale set-up, .vt/.rmt writers, stdout lines — is host code
// that follows the reference's observable behaviour line by line (citations inline).
//
// Extensions (not in the reference): --rng mt|philox, --sum strict|fast|int8, --gpu N.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "../../include/consel_gpu.h"
#include "refio.h"

using namespace refio;

enum SeqMode { SEQ_MT, SEQ_MOLPHY, SEQ_PAML, SEQ_PAUP, SEQ_PUZZLE, SEQ_PHYML };

static void byebye() { die("error in command line"); }

struct Params {
    unsigned long seed = 0;
    unsigned long seed_used = 0;
    std::vector<double> rr;
    std::vector<int> bb;
    SeqMode seqmode = SEQ_MT;
    int rng_mode = CGPU_RNG_MT19937, sum_mode = CGPU_SUM_STRICT;
    unsigned long long consumed = 0;  // draws taken from the stream so far (-g continues it)
};

static int genrmt(Params &P, const char *infile, const char *outfile) {
    const char *fext = ".mt";
    switch (P.seqmode) {  // default extensions: src/freadmat.h:13-18
        case SEQ_MOLPHY: fext = ".lls"; break;
        case SEQ_PAML: fext = ".lnf"; break;
        case SEQ_PAUP: fext = ".txt"; break;
        case SEQ_PUZZLE: fext = ".sitelh"; break;
        case SEQ_PHYML: fext = ".txt"; break;
        default: break;
    }
    FILE *fp;
    std::string used;
    if (infile) {
        fp = open_file(infile, fext, "r", &used);
        printf("\n# reading %s", used.c_str());
    } else {
        fp = stdin;
        printf("\n# reading from stdin");
    }
    int mm = 0, nn = 0;
    std::vector<double> L;
    switch (P.seqmode) {
        case SEQ_MOLPHY: L = read_mat_molphy(fp, &mm, &nn); break;
        case SEQ_PAML: L = read_mat_paml(fp, &mm, &nn); break;
        case SEQ_PAUP: L = read_mat_paup(fp, &mm, &nn); break;
        case SEQ_PUZZLE: L = read_mat_puzzle(fp, &mm, &nn); break;
        case SEQ_PHYML: L = read_mat_phyml(fp, &mm, &nn); break;
        default: L = read_mat(fp, &mm, &nn); break;
    }
    if (infile) fclose(fp);
    printf("\n# M:%d N:%d", mm, nn);
    if (mm <= 0 || nn <= 0) die("empty matrix");

    const int kk = (int)P.rr.size();
    std::vector<double> datvec(mm), rr1(kk);
    std::vector<int> bn(kk);
    for (int i = 0; i < mm; i++) {  // src/makermt.c:159-162
        double x = 0;
        for (int j = 0; j < nn; j++) x += L[(size_t)i * nn + j];
        datvec[i] = x;
    }
    for (int i = 0; i < kk; i++) {  // :165-168
        bn[i] = (int)(P.rr[i] * nn);
        rr1[i] = (double)bn[i] / nn;
    }

    // The device keeps site counts in bytes (a site drawn more than 255 times in one replicate is refused
    // with CGPU_ERANGE).  That can only happen for toy alignments with large scales (n = 10, r = 30); it is
    // recognised here, before any output file is opened, instead of after the whole run.
    for (int i = 0; i < kk; i++) {
        const double mu = (double)bn[i] / (double)nn;  // mean count per site
        if (bn[i] < 0) die("scale %g gives a negative replicate size (n'=%d) for %d sites", P.rr[i], bn[i], nn);
        if (mu + 9.0 * sqrt(mu) + 10.0 > 255.0)
            die("scale %g on %d sites draws a site about %g times per replicate: beyond the byte counts of the device path", P.rr[i], nn, mu);
    }

    // one process per GPU (RANK / WORLD_SIZE, cgpu_comm_init_env): every rank resamples its share of the
    // replicate columns, the blocks are all-gathered over NVLink and rank 0 alone writes the files
    int g_rank = 0, g_world = 1;
    cgpu_comm_info(&g_rank, &g_world);
    const bool writer = g_rank == 0;
    if (!writer) outfile = nullptr;  // (stdout of the other ranks is /dev/null)
    FILE *out = nullptr;
    if (outfile) {
        fp = open_file(outfile, ".vt", "w", &used);
        printf("\n# writing %s", used.c_str());
        write_vec(fp, datvec.data(), mm);
        fclose(fp);
        out = open_file(outfile, ".rmt", "wb", &used);
        printf("\n# writing %s", used.c_str());
        write_bvec(out, datvec.data(), mm);
        write_bvec(out, rr1.data(), kk);
        write_bivec(out, P.bb.data(), kk);
        write_bint(out, kk);
    } else {
        printf("\n# writing to stdout");
        printf("\n# OBS:\n"); write_vec(stdout, datvec.data(), mm);
        printf("\n# R:\n"); write_vec(stdout, rr1.data(), kk);
        printf("\n# B:\n"); write_ivec(stdout, P.bb.data(), kk);
        printf("\n# RMAT:\n");
        printf("%d\n", kk);
    }

    int total = 0;
    for (int i = 0; i < kk; i++) total += P.bb[i];
    printf("\n# start generating total %d replicates for %d items", total, mm);
    fflush(stdout);
    const double t0 = cpu_seconds();

    // ---- the hot loop, on the device
    std::vector<std::vector<double>> rep(kk);
    std::vector<double *> rep_ptr(kk);
    for (int i = 0; i < kk; i++) {
        rep[i].resize((size_t)mm * (P.bb[i] > 0 ? P.bb[i] : 0) + 1);
        rep_ptr[i] = rep[i].data();
    }
    // A scale with n' = (int)(r*n) = 0 (a toy alignment under a small scale) draws nothing: the reference's loop
    // (src/makermt.c:52-70) leaves x = 0 and stores x / r with r = 0/n -- the NaN of 0.0/0.0, formed here by the
    // same division at run time.  Such scales consume no draws of the stream and never reach the device.
    std::vector<int> live, bn_live, bb_live;
    std::vector<double *> ptr_live;
    for (int i = 0; i < kk; i++) {
        if (bn[i] > 0) {
            live.push_back(i);
            bn_live.push_back(bn[i]);
            bb_live.push_back(P.bb[i]);
            ptr_live.push_back(rep_ptr[i]);
        } else {
            volatile double zero = 0.0;
            const double v = zero / (zero / (double)nn);
            std::fill(rep[i].begin(), rep[i].end(), v);
        }
    }
    const int kl = (int)live.size();
    if (kl == 0) {
        // nothing to resample
    } else if (g_world == 1) {
        int rc = cgpu_scaleboot_all(L.data(), mm, nn, kl, bn_live.data(), bb_live.data(), P.seed_used, P.consumed, P.rng_mode,
                                    P.sum_mode, ptr_live.data());
        if (rc != CGPU_OK) die("%s", cgpu_last_error());
    } else {
        cgpu_matrix *dm = nullptr;
        cgpu_reps *dr = nullptr;
        if (cgpu_matrix_upload(L.data(), mm, nn, &dm) != CGPU_OK) die("%s", cgpu_last_error());
        if (cgpu_scaleboot_dev(dm, kl, bn_live.data(), bb_live.data(), P.seed_used, P.consumed, P.rng_mode, P.sum_mode, &dr) != CGPU_OK)
            die("%s", cgpu_last_error());
        for (int q = 0; q < kl; q++) {
            cgpu_reps *whole = nullptr;
            if (cgpu_reps_gather_scale(dr, q, &whole) != CGPU_OK) die("%s", cgpu_last_error());
            double *one[1] = {ptr_live[q]};
            if (writer && bb_live[q] > 0 && cgpu_reps_download(whole, one) != CGPU_OK) die("%s", cgpu_last_error());
            cgpu_reps_free(whole);
        }
        cgpu_reps_free(dr);
        cgpu_matrix_free(dm);
    }
    for (int i = 0; i < kk; i++) P.consumed += (unsigned long long)P.bb[i] * (unsigned long long)bn[i];

    for (int i = 0; i < kk; i++) {
        if (bn[i] == 0) warn("zero size in malloc");  // scaleboot's new_vec(bn) of an empty draw list (src/makermt.c:48, misc.c:98)
        if (outfile) {
            write_bmat(out, rep[i].data(), mm, P.bb[i]);
            putchar('.');
            fflush(stdout);
        } else {
            printf("\n## RMAT[%d]:\n", i);
            write_mat(stdout, rep[i].data(), mm, P.bb[i]);
        }
    }
    const double t1 = cpu_seconds();
    printf("\n# time elapsed for bootstrap t=%g sec", t1 - t0);
    if (outfile) fclose(out);
    return 0;
}

int main(int argc, char **argv) {
    Params P;
    const char *fname_pa = nullptr, *fname_mt = nullptr, *fname_rmt = nullptr;
    double bbfact = 1.0;
    bool fastrep = false, multi = false;
    int gpu = -1;

    if (getenv("RANK") && atoi(getenv("RANK")) > 0 && getenv("WORLD_SIZE") && atoi(getenv("WORLD_SIZE")) > 1)
        if (!freopen("/dev/null", "w", stdout)) return 1;  // rank 0 speaks for the job
    printf("# consel_b200 makermt (sm_100a); command line of makermt.c,v 1.16");

    for (int i = 1, j = 1; i < argc; i++) {  // src/makermt.c:237-279
        const std::string a = argv[i];
        auto need = [&](int more) { if (i + more >= argc) byebye(); };
        if (argv[i][0] != '-') {
            if (j == 1) fname_mt = argv[i];
            else if (j == 2) fname_rmt = argv[i];
            else byebye();
            j++;
        } else if (a == "-s") {
            need(1);
            if (sscanf(argv[++i], "%lu", &P.seed) != 1) byebye();
        } else if (a == "-p") {
            need(1);
            fname_pa = argv[++i];
        } else if (a == "-b") {
            need(1);
            if (sscanf(argv[++i], "%lf", &bbfact) != 1) byebye();
        } else if (a == "-f") fastrep = true;
        else if (a == "-g") multi = true;
        else if (a == "--molphy") P.seqmode = SEQ_MOLPHY;
        else if (a == "--paml") P.seqmode = SEQ_PAML;
        else if (a == "--paup") P.seqmode = SEQ_PAUP;
        else if (a == "--puzzle") P.seqmode = SEQ_PUZZLE;
        else if (a == "--phyml") P.seqmode = SEQ_PHYML;
        else if (a == "-d") {
            need(1);
            if (sscanf(argv[++i], "%d", &debugmode) != 1) byebye();
        } else if (a == "--rng") {
            need(1);
            std::string v = argv[++i];
            if (v == "mt" || v == "mt19937") P.rng_mode = CGPU_RNG_MT19937;
            else if (v == "philox") P.rng_mode = CGPU_RNG_PHILOX;
            else byebye();
        } else if (a == "--sum") {
            need(1);
            std::string v = argv[++i];
            if (v == "strict") P.sum_mode = CGPU_SUM_STRICT;
            else if (v == "fast") P.sum_mode = CGPU_SUM_FAST;
            else if (v == "int8") P.sum_mode = CGPU_SUM_INT8;
            else byebye();
        } else if (a == "--gpu") {
            need(1);
            if (sscanf(argv[++i], "%d", &gpu) != 1) byebye();
        } else byebye();
    }

    // smrand (src/rand.c:33-37): seed 0 means "seed from the clock"
    P.seed_used = P.seed ? P.seed : (unsigned long)time(nullptr);

    if (fname_pa) {  // src/makermt.c:285-291
        std::string used;
        FILE *fp = open_file(fname_pa, ".pa", "r", &used);
        printf("\n# reading %s", used.c_str());
        P.rr = read_vec(fp);
        P.bb = read_ivec(fp, (int)P.rr.size());
        fclose(fp);
    } else if (fastrep) {
        P.rr = {1.0};
        P.bb = {10000};
    } else {
        P.rr = {0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4};
        P.bb.assign(10, 10000);
    }
    const int kk = (int)P.rr.size();
    for (int i = 0; i < kk; i++) P.bb[i] = (int)(P.bb[i] * bbfact);  // `bb[i] *= bbfact`, :297
    printf("\n# seed:%lu (MT19937 generator)", P.seed);
    printf("\n# K:%d", kk);
    printf("\n# R:");
    for (int i = 0; i < kk; i++) printf("%g ", P.rr[i]);
    printf("\n# B:");
    for (int i = 0; i < kk; i++) printf("%d ", P.bb[i]);

    if (cgpu_init(gpu) != CGPU_OK) die("%s", cgpu_last_error());
    if (cgpu_comm_init_env() != CGPU_OK) die("%s", cgpu_last_error());

    if (multi) {  // src/makermt.c:303-318
        std::string used;
        FILE *fp = fname_mt ? open_file(fname_mt, ".svt", "r", &used) : stdin;
        std::vector<std::string> infiles = read_svec(fp), outfiles;
        if (fname_mt) fclose(fp);
        if (fname_rmt) {
            fp = open_file(fname_rmt, ".svt", "r", &used);
            outfiles = read_svec(fp);
            fclose(fp);
            if (outfiles.size() != infiles.size()) die("size mismatch in svec");
        } else {
            for (auto &s : infiles) outfiles.push_back(strip_extension(s));
        }
        for (size_t i = 0; i < infiles.size(); i++) {
            printf("\n# %d/%d %s %s", (int)i + 1, (int)infiles.size(), infiles[i].c_str(), outfiles[i].c_str());
            genrmt(P, infiles[i].c_str(), outfiles[i].c_str());
        }
    } else {
        std::string outname;
        if (fname_mt && !fname_rmt) {
            outname = strip_extension(fname_mt);
            fname_rmt = outname.c_str();
        }
        genrmt(P, fname_mt, fname_rmt);
    }
    printf("\n# exit normally\n");
    cgpu_shutdown();  // (also tears the communicator down, on every rank)
    return 0;
}
