// Host-side runtime shared by our makermt / consel / catpv: the reference's text and
// binary file conventions, its error/warning protocol and file-name rules, restated
// from the behaviour of src/misc.c (cited per function in refio.cpp).  Plain host
// code: the I/O and CLI surface stays on the CPU by design (BASELINE.json north_star).
#pragma once
#include <cstdio>
#include <string>
#include <vector>

namespace refio {

extern int debugmode;

// `\n#! Terminated: <msg>.\n` on stdout, exit(1)          (src/misc.c:40-49)
[[noreturn]] void die(const char *fmt, ...);
// `\n#! Warning: <msg>. ` on stdout, continue            (src/misc.c:51-59)
void warn(const char *fmt, ...);
// printf when debugmode >= level                         (src/misc.c:29-38)
void dprint(int level, const char *fmt, ...);
double cpu_seconds();  // clock()/CLOCKS_PER_SEC          (src/misc.c:65-73)

// extension rules of openfp/chkaxt/rmvaxt                (src/misc.c:235-276)
bool has_extension(const std::string &name);
std::string strip_extension(const std::string &name);
FILE *open_file(const std::string &name, const char *ext, const char *mode, std::string *used);

// ---- ascii readers: whitespace separated tokens, `#...\n` comments (src/misc.c:174-193)
void skip_junk(FILE *fp);
void skip_line(FILE *fp);  // src/freadmat.c:13-39 (handles \r\n and \r)
void skip_word(FILE *fp);  // src/freadmat.c:41-53
int read_int(FILE *fp);
double read_double(FILE *fp);
std::string read_word(FILE *fp);
std::vector<double> read_vec(FILE *fp, int expect = 0);
std::vector<int> read_ivec(FILE *fp, int expect = 0);
std::vector<std::string> read_svec(FILE *fp);
// K x n matrix, row-major; m/n in: expected (0 = any), out: actual
std::vector<double> read_mat(FILE *fp, int *m, int *n);

// ---- ascii writers: "%14.9g " / "%10d ", five per line   (src/misc.c:447-519)
void write_vec(FILE *fp, const double *a, int m);
void write_ivec(FILE *fp, const int *a, int m);
void write_mat(FILE *fp, const double *a, int m, int n);  // row-major

// ---- binary: native int32 / double, no padding           (src/misc.c:525-645)
int read_bint(FILE *fp);
std::vector<double> read_bvec(FILE *fp, int expect = 0);
std::vector<int> read_bivec(FILE *fp, int expect = 0);
std::vector<double> read_bmat(FILE *fp, int *m, int *n);
void write_bint(FILE *fp, int x);
void write_bvec(FILE *fp, const double *a, int m);
void write_bivec(FILE *fp, const int *a, int m);
void write_bmat(FILE *fp, const double *a, int m, int n);

// ---- foreign site-log-likelihood formats                  (src/freadmat.c)
std::vector<double> read_mat_paml(FILE *fp, int *m, int *n);    // .lnf  :110-154
std::vector<double> read_mat_molphy(FILE *fp, int *m, int *n);  // .lls  :55-75
std::vector<double> read_mat_puzzle(FILE *fp, int *m, int *n);  // .sitelh :77-106
std::vector<double> read_mat_paup(FILE *fp, int *m, int *n);    // .txt  :205-300
std::vector<double> read_mat_phyml(FILE *fp, int *m, int *n);   // .txt  :321-360

// the reference's in-place quicksort with permutation output (src/misc.c:657-700): ties
// are ordered by THIS partition scheme, so the item order of equal statistics matches.
void sort_with_order(std::vector<double> &x, std::vector<int> *order);
template <typename Less3>  // cmp(a,b) -> <0, 0, >0 on indices into a caller array
void psort_order(int dim, std::vector<int> &order, Less3 cmp);

}  // namespace refio

// ---------------------------------------------------------------- template body
namespace refio {
template <typename Cmp>
void psort_order(int dim, std::vector<int> &order, Cmp cmp) {
    // src/misc.c:745-789 (mypsort): median-position pivot, Hoare-style scans with the
    // reference's boundary handling; `order` ends as the permutation of sorted items.
    order.resize(dim);
    for (int i = 0; i < dim; i++) order[i] = i;
    if (dim <= 1) return;
    std::vector<int> item(order);  // item[p] = original index currently at position p
    int stack[64], sp = 0, l = 0, r = dim - 1;
    for (;;) {
        while (r > l) {
            int k = (r + l) / 2;
            const int v = item[k];
            int i = l - 1, j = r + 1;
            for (;;) {
                while (i < r && cmp(item[++i], v) <= 0) {}
                while (j > l && cmp(item[--j], v) >= 0) {}
                if (i > j) break;
                if (i == r) j = k; else if (j == l) i = k;
                std::swap(item[i], item[j]);
                if (i == r) { j = r - 1; break; } else if (j == l) { i = l + 1; break; }
            }
            if (j - l > r - i) { stack[sp++] = l; stack[sp++] = j; l = i; }
            else { stack[sp++] = i; stack[sp++] = r; r = j; }
        }
        if (sp == 0) break;
        r = stack[--sp];
        l = stack[--sp];
    }
    order = item;
}
}  // namespace refio
