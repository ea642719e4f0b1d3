// catmt — concatenates .mt files site-wise (the trees must agree in number) and, with
// --kht alpha, prints the KH test of every tree against the best one and keeps the trees
// with p >= alpha, writing their ids to <out>.vt.  Same command line, stdout protocol and
// output files as the reference's catmt (src/catmt.c:63-201).  Pure host code on the
// input side of makermt.
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "refio.h"
#include "specfun.h"

using namespace refio;

static void byebye() { die("error in command line"); }

int main(int argc, char **argv) {
    std::vector<std::string> names;
    double khtalpha = 0.0;
    bool sw_kht = false;
    const double varadd = 1.0, varaddwarn = 100.0;

    printf("# %s", "catmt (consel_b200 host tool; reference: src/catmt.c)");
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a[0] != '-') names.push_back(a);
        else if (a == "-d") { if (i + 1 >= argc || sscanf(argv[i + 1], "%d", &debugmode) != 1) byebye(); i++; }
        else if (a == "--kht") { if (i + 1 >= argc || sscanf(argv[i + 1], "%lf", &khtalpha) != 1) byebye(); sw_kht = true; i++; }
        else byebye();
    }
    if (names.size() < 2) die("must specify input-file and output-file");
    std::string fname_out = names.back();
    names.pop_back();

    // ---- read and append columns (src/catmt.c:28-50, :112-119)
    int m = 0, n = 0;
    std::vector<std::vector<double>> rows;  // m rows, growing site-wise
    for (const std::string &name : names) {
        std::string used;
        FILE *fp = open_file(name, ".mt", "r", &used);
        printf("\n# reading %s", used.c_str());
        int mi = 0, ni = 0;
        std::vector<double> x = read_mat(fp, &mi, &ni);
        fclose(fp);
        if (rows.empty()) { m = mi; rows.resize(m); }
        else if (mi != m) die("matrix row size mismatch");
        for (int i = 0; i < m; i++) rows[i].insert(rows[i].end(), x.begin() + (size_t)i * ni, x.begin() + (size_t)(i + 1) * ni);
        n += ni;
        printf(" %d", ni);
    }
    printf("\n# M:%d N:%d", m, n);

    std::vector<int> keep(m);
    for (int i = 0; i < m; i++) keep[i] = i;
    if (sw_kht) {  // src/catmt.c:138-199
        std::vector<double> tot(m), pv(m);
        double best = -1.0e30;
        int i0 = 0;
        for (int i = 0; i < m; i++) {
            double x = 0.0;
            for (int j = 0; j < n; j++) x += rows[i][j];
            tot[i] = x;
            if (x > best) { best = x; i0 = i; }
        }
        printf("\n# %4s %6s %6s %5s %6s", "item", "obs", "se", "T", "kh");
        double s = 0.0, x = 0.0;
        for (int i = 0; i < m; i++) {
            if (i == i0) {
                pv[i] = 1.0; s = x = 0.0;
            } else {
                double mn = (tot[i0] - tot[i]) / n;
                x = 0.0;
                for (int j = 0; j < n; j++) { double y = rows[i0][j] - rows[i][j] - mn; x += y * y; }
                x = (double)n * x;
                if (x < varadd * varaddwarn)
                    warn("small variance [%d,%d]: %g<%g*%g", i0 + 1, i + 1, x, varadd, varaddwarn);
                double v = (x + varadd) / (n - 1);
                s = sqrt(v);
                x = (tot[i0] - tot[i]) / s;
                pv[i] = 1.0 - specfun::poz(x);
            }
            printf("\n# %4d %6.2f %6.2f %5.2f %6.3f", i + 1, tot[i0] - tot[i], s, x, pv[i]);
        }
        if (khtalpha > 0.0) {
            keep.clear();
            for (int i = 0; i < m; i++) if (pv[i] >= khtalpha) keep.push_back(i);
            printf("\n# SCREENING BY KH > %6.3f\n# M:%d -> %d", khtalpha, m, (int)keep.size());
            std::string used;
            FILE *fp = open_file(strip_extension(fname_out), ".vt", "w", &used);
            printf("\n# writing %s", used.c_str());
            write_ivec(fp, keep.data(), (int)keep.size());
            fclose(fp);
        }
    }

    std::vector<double> out;
    out.reserve(keep.size() * (size_t)n);
    for (int i : keep) out.insert(out.end(), rows[i].begin(), rows[i].end());
    std::string used;
    FILE *fp = open_file(fname_out, ".mt", "w", &used);
    printf("\n# writing %s", used.c_str());
    write_mat(fp, out.data(), (int)keep.size(), n);
    fclose(fp);
    printf("\n# exit normally\n");
    return 0;
}
