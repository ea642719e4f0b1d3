// catci — prints the confidence-limit table of one or more .ci files (consel -c) exactly like
// the reference's catci (src/catci.c:27-129): AU limits | bootstrap (np) limits per alpha,
// -v adds the standard errors and error indices.  Pure host code, no computation.
#include <cstdio>
#include <string>
#include <vector>

#include "refio.h"

using namespace refio;

static void byebye() { die("error in command line"); }
static void repchar(char c, int n) { for (int i = 0; i < n; i++) putchar(c); }

int main(int argc, char **argv) {
    std::vector<std::string> names;
    bool verbose = false, sw_au = true, sw_np = true;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a[0] != '-') names.push_back(a);
        else if (a == "-d") { if (i + 1 >= argc || sscanf(argv[i + 1], "%d", &debugmode) != 1) byebye(); i++; }
        else if (a == "-v") verbose = true;
        else if (a == "--no_au") sw_au = false;
        else if (a == "--no_np") sw_np = false;
        else byebye();
    }
    for (const std::string &name : names) {
        std::string used;
        FILE *fp = open_file(name, ".ci", "r", &used);
        printf("\n# reading %s", used.c_str());
        std::vector<int> order = read_ivec(fp);
        int cm = (int)order.size();
        std::vector<double> obs = read_vec(fp, cm);
        std::vector<double> alpha = read_vec(fp);
        int na = (int)alpha.size();
        std::vector<double> t[6];  // ci, se, ei of the AU limits, then of the np limits
        for (auto &x : t) {
            int m = cm, n = na;
            x = read_mat(fp, &m, &n);
            if (m != cm) die("size of rows mismatch in mat");
            if (n != na) die("size of columns mismatch in mat");
        }
        fclose(fp);

        auto alphas = [&] {
            for (int j = 0; j < na; j++) {
                printf(" %6.3f", alpha[j]);
                if (verbose) printf(" %4s %4s", "se", "ei");
            }
        };
        auto limits = [&](int i, int k) {
            for (int j = 0; j < na; j++) {
                size_t p = (size_t)i * na + j;
                printf(" %6.1f", t[k][p]);
                if (verbose) printf(" %4.1f %4.1f", t[k + 1][p], t[k + 2][p]);
            }
        };
        printf("\n#");
        repchar(' ', 17);
        int w = na * (verbose ? 17 : 7), w2 = w / 2 - 2;
        if (sw_au) { repchar('-', w2); printf(" au "); repchar('-', w - w2 - 4); }
        printf(" |");
        if (sw_np) { repchar('-', w2); printf(" np "); repchar('-', w - w2 - 4); }
        printf("\n# %4s %4s", "rank", "item");
        printf(" %6s", "obs");
        if (sw_au) alphas();
        printf(" |");
        if (sw_np) alphas();
        for (int i = 0; i < cm; i++) {
            printf("\n# %4d %4d", i + 1, order[i] + 1);
            printf(" %6.1f", obs[i]);
            if (sw_au) limits(i, 0);
            printf(" |");
            if (sw_np) limits(i, 3);
        }
        printf("\n");
    }
    return 0;
}
