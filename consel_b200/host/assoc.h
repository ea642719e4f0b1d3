// Associations and item order shared by consel and makerep: read_assoc of the reference
// (src/consel.c:459-541; the same code in src/makerep.c:196-272).
#pragma once
#include <cstdio>
#include <string>
#include <vector>

#include "refio.h"

static const double kHuge = 1.0e30;  // src/consel.c:155

struct Assoc {
    std::vector<int> ptr{0}, idx;
    int size() const { return (int)ptr.size() - 1; }
    int len(int e) const { return ptr[e + 1] - ptr[e]; }
    const int *at(int e) const { return idx.data() + ptr[e]; }
    void push(const std::vector<int> &v) {
        idx.insert(idx.end(), v.begin(), v.end());
        ptr.push_back((int)idx.size());
    }
};

// Reads the association file (or builds the identity association), computes the observed
// statistic vectors, sorts the items (unless nosort) and truncates to cm_arg items.
// Prints the reference's stdout lines.  Returns cm.
inline int read_assoc_and_order(const char *fname_ass, const std::vector<double> &datvec, int cm_arg, bool nosort,
                                Assoc *ass_out, std::vector<double> *obsvec_out, std::vector<int> *orderv_out) {
    using namespace refio;
    const int mm = (int)datvec.size();
    std::string used;
    Assoc &ass = *ass_out;
    std::vector<double> &obsvec = *obsvec_out;
    std::vector<int> &orderv = *orderv_out;
    int cm = 0;
    // ---------------- associations and item order (read_assoc, src/consel.c:459-541)
    Assoc ass0;
    if (fname_ass) {
        FILE *fp = open_file(fname_ass, ".ass", "r", &used);
        printf("\n# reading %s", used.c_str());
        const int cm0 = read_int(fp);
        for (int i = 0; i < cm0; i++) {
            std::vector<int> v = read_ivec(fp);
            for (size_t j = 0; j < v.size(); j++)
                if (v[j] < 0 || v[j] >= mm) die("ass out of range: ass[%d][%d]=%d", i + 1, (int)j + 1, v[j] + 1);
            ass0.push(v);
        }
        fclose(fp);
    } else {
        printf("\n# generate the identity association");
        for (int i = 0; i < mm; i++) ass0.push({i});
    }
    const int cm0 = ass0.size();
    // observed statistics: max over the complement minus each member, sorted ascending
    std::vector<std::vector<double>> obsv(cm0);
    {
        std::vector<char> mark(mm);
        for (int e = 0; e < cm0; e++) {
            std::fill(mark.begin(), mark.end(), 0);
            int nset = 0;
            for (int q = 0; q < ass0.len(e); q++) {
                if (!mark[ass0.at(e)[q]]) nset++;
                mark[ass0.at(e)[q]] = 1;
            }
            if (nset != ass0.len(e)) die("\n# duplicate associations");
            double mx = -kHuge;
            for (int jx = 0; jx < mm; jx++)
                if (!mark[jx] && datvec[jx] > mx) mx = datvec[jx];
            std::vector<double> v(ass0.len(e));
            for (int q = 0; q < ass0.len(e); q++) v[q] = mx - datvec[ass0.at(e)[q]];
            sort_with_order(v, nullptr);
            obsv[e] = v;
        }
    }
    printf("\n# CM:%d", cm0);
    cm = cm_arg;
    if (cm <= 0 || cm > cm0) cm = cm0;
    if (cm < cm0) printf(" -> %d", cm);
    orderv.resize(cm0);
    if (nosort) {
        for (int i = 0; i < cm0; i++) orderv[i] = i;
    } else {
        psort_order(cm0, orderv, [&](int a, int b) {  // compdvec, src/consel.c:1986-2000
            const std::vector<double> &x = obsv[a], &y = obsv[b];
            int out, len;
            if (x.size() < y.size()) { out = -1; len = (int)x.size(); }
            else if (x.size() > y.size()) { out = 1; len = (int)y.size(); }
            else { out = 0; len = (int)y.size(); }
            for (int i = 0; i < len; i++) {
                if (x[i] < y[i]) { out = -1; break; }
                else if (x[i] > y[i]) { out = 1; break; }
            }
            return out;
        });
    }
    dprint(1, "\n# observed statistics for the associations");
    dprint(1, "\n# %4s %4s ", "rank", "item");
    for (int i = 0; i < cm0; i++) {
        dprint(1, "\n# %4d %4d ", i + 1, orderv[i] + 1);
        for (double v : obsv[orderv[i]]) dprint(1, " %6.2f", v);
    }
    obsvec.resize(cm);
    for (int i = 0; i < cm; i++) {
        obsvec[i] = obsv[orderv[i]][0];
        ass.push(std::vector<int>(ass0.at(orderv[i]), ass0.at(orderv[i]) + ass0.len(orderv[i])));
    }
    return cm;
}
