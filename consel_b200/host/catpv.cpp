// catpv — prints the p-value table of one or more .pv files exactly like the
// reference's catpv (src/catpv.c:105-330): `# rank item obs au np | bp pp kh sh wkh wsh |`.
// Pure host code, no computation; it defines the "printed precision" of the parity
// contract.  The aggregate modes (-o, -c) are not carried by this build.
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "refio.h"

using namespace refio;

static bool sw_printse = false;
static int sw_sort = 0;
static void byebye() { die("error in command line"); }

static void print_real(double x) {  // src/catpv.c:40-46
    if (x < 100.0 && x > -10.0 && fabs(x) > 0.0005) printf(" %6.3f", x);
    else if (x == 0.0) printf(" %6s", "0");
    else if (fabs(x) < 0.0005 || x > 999999 || x < -99999) printf(" %6.0e", x);
    else printf(" %6.0f", x);
}
static void print_pvname(const char *name) {
    printf(" %6s", name);
    if (sw_printse) printf(" %7s", "(se)");
}
static char indicate_sort(int col) { return col == sw_sort ? '+' : (col == -sw_sort ? '-' : ' '); }
static void print_pvcol(int col) {
    printf(" %5d%c", col, indicate_sort(col));
    if (sw_printse) printf(" %7s", "");
}
static void print_pval(double pv, double se) {
    print_real(pv);
    if (sw_printse) printf(" (%5.3f)", se);
}

int main(int argc, char **argv) {
    std::vector<std::string> files;
    bool verbose = false, pba = true, pau = true, pbp = true, pmc = true, prt = true, prank = false;
    int labeladd = 1;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto integer = [&](int *x) { if (i + 1 >= argc || sscanf(argv[++i], "%d", x) != 1) byebye(); };
        if (argv[i][0] != '-') files.push_back(a);
        else if (a == "-d") integer(&debugmode);
        else if (a == "-l") integer(&labeladd);
        else if (a == "-s") integer(&sw_sort);
        else if (a == "-r") prank = true;
        else if (a == "-e") sw_printse = true;
        else if (a == "-v") verbose = true;
        else if (a == "--no_au") pau = false;
        else if (a == "--no_bp") pbp = false;
        else if (a == "--no_pp") pba = false;
        else if (a == "--no_sh") pmc = false;
        else if (a == "--no_print") prt = false;
        else if (a == "-o" || a == "-c" || a == "-t" || a == "-i" || a == "-h")
            die("switch %s is not supported by this build of catpv", a.c_str());
        else byebye();
    }
    if (files.empty()) die("must specify input-file");

    for (const std::string &fname : files) {
        std::string used;
        FILE *fp = open_file(fname, ".pv", "r", &used);
        if (prt) printf("\n# reading %s", used.c_str());
        if (read_int(fp) != 1) die("wrong file type");
        std::vector<int> orderv = read_ivec(fp);
        const int cm = (int)orderv.size();
        std::vector<double> obsvec = read_vec(fp, cm);
        const int outbit = read_int(fp);
        int m = cm, pvnum = 0, auxnum = 0;
        std::vector<double> pvmat = read_mat(fp, &m, &pvnum);
        m = cm;
        std::vector<double> semat = read_mat(fp, &m, &pvnum);
        m = cm;
        std::vector<double> auxmat = read_mat(fp, &m, &auxnum);
        fclose(fp);

        std::vector<int> rev(cm);
        for (int i = 0; i < cm; i++) rev[i] = i;
        if (sw_sort) {
            std::vector<double> key(cm);
            int j = sw_sort < 0 ? -sw_sort : sw_sort;
            if (j == 1) for (int i = 0; i < cm; i++) key[i] = (double)orderv[i];
            else if (j == 2) for (int i = 0; i < cm; i++) key[i] = obsvec[i];
            else if (j < pvnum + 3) for (int i = 0; i < cm; i++) key[i] = -pvmat[(size_t)i * pvnum + j - 3];
            else die("-s # <= %d\n", pvnum + 2);
            if (sw_sort < 0) for (int i = 0; i < cm; i++) key[i] = -key[i];
            sort_with_order(key, &rev);
        }
        int j = 0, jbp = 0, jba = 0, jmc = 0, jau = 0;
        bool sbp = false, sba = false, smc = false, sau = false;
        if (outbit & 1) { sbp = true; jbp = j; j += 1; }
        if (outbit & 2) { sba = true; jba = j; j += 1; }
        if (outbit & 4) { smc = true; jmc = j; j += 4; }
        if (outbit & 8) { sau = true; jau = j; j += 2; }
        sbp = sbp && pbp && prt; sba = sba && pba && prt; sau = sau && pau && prt; smc = smc && pmc && prt;

        if (prt && (sw_sort || verbose)) {
            printf("\n# %3d%c %3d%c %5d%c", 0, indicate_sort(0), 1, indicate_sort(1), 2, indicate_sort(2));
            if (sau) for (int i = jau; i < jau + 2; i++) print_pvcol(3 + i);
            printf(" |");
            if (sbp) print_pvcol(3 + jbp);
            if (sba) print_pvcol(3 + jba);
            if (smc) for (int i = jmc; i < jmc + 4; i++) print_pvcol(3 + i);
        }
        if (prt) printf("\n# %4s %4s %6s", "rank", "item", "obs");
        if (sau) { print_pvname("au"); print_pvname("np"); }
        if (prt) printf(" |");
        if (sbp) print_pvname("bp");
        if (sba) print_pvname("pp");
        if (smc) { print_pvname("kh"); print_pvname("sh"); print_pvname("wkh"); print_pvname("wsh"); }
        if (prt) printf(" |");
        if (verbose && sau) printf(" %6s %6s %2s %6s %6s %6s %6s %6s", "pf", "rss", "df", "x", "c", "th", "a", "dim");

        for (int i = 0; i < cm; i++) {
            const int ir = rev[i];
            const double *pv = &pvmat[(size_t)ir * pvnum], *se = &semat[(size_t)ir * pvnum];
            if (prt) printf("\n# %4d %4d %6.1f", i + 1, orderv[ir] + labeladd, obsvec[ir]);
            if (sau) { print_pval(pv[jau], se[jau]); print_pval(pv[jau + 1], se[jau + 1]); }
            if (prt) printf(" |");
            if (sbp) print_pval(pv[jbp], se[jbp]);
            if (sba) print_pval(pv[jba], se[jba]);
            if (smc) for (int c = 0; c < 4; c++) print_pval(pv[jmc + c], se[jmc + c]);
            if (prt) printf(" |");
            if (verbose && sau) {
                const double *ax = &auxmat[(size_t)ir * auxnum];
                print_real(ax[0]); print_real(ax[1]);
                printf(" %2.0f", ax[2]);
                print_real(ax[3]); print_real(ax[4]); print_real(ax[5]);
                if (auxnum > 6) {
                    double x = (ax[6] != 0.0 && ax[4] != 0.0) ? 0.5 * (ax[6] - 1.0) / ax[4] : 0.0;
                    print_real(x);
                    print_real(ax[6]);
                }
            }
        }
        if (prank) {
            printf("\n\n# %6s %6s %6s", "rank", "ordr", "item");
            for (int i = 0; i < cm; i++) printf("\n  %6d %6d %6d", i + 1, rev[i] + 1, orderv[rev[i]] + labeladd);
        }
        if (prt) printf("\n");
    }
    printf("\n");
    return 0;
}
