// catrep — merges .rep (or, with -m, .rmt) files: blocks whose scales agree within
// -e (default 0.005) are concatenated replicate-wise in file order, -p keeps only the
// scales of a .pa file (in its order), rows are sorted ascending unless --no_sort.
// Same command line, stdout protocol and output bytes as the reference's catrep
// (src/catrep.c:61-285).  Pure host code: the tool only moves replicate blocks that
// makermt / consel produced on the device.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "refio.h"

using namespace refio;

namespace {

struct Block {               // one scale: cm rows x b replicates, row-major
    double r = 0.0;
    int b = 0;
    std::vector<double> x;
};
struct RepFile {
    int cm = 0;
    std::vector<int> ord;    // item ids (.rep only)
    std::vector<double> obs;
    std::vector<Block> blocks;
};

bool mode_rmt = false;

void byebye() { die("error in command line"); }
void putdot() { putchar('.'); fflush(stdout); }

void print_header(int cm, const std::vector<Block> &bl) {  // src/catrep.c:31-39
    printf("\n# K:%d", (int)bl.size());
    printf("\n# R:");
    for (const Block &k : bl) printf("%g ", k.r);
    printf("\n# B:");
    for (const Block &k : bl) printf("%d ", k.b);
    printf(mode_rmt ? "\n# MM:%d" : "\n# CM:%d", cm);
}

int find_scale(const std::vector<Block> &bl, double r, double eps) {
    for (size_t j = 0; j < bl.size(); j++)
        if (fabs(bl[j].r - r) < eps) return (int)j;
    return -1;
}

}  // namespace

int main(int argc, char **argv) {
    std::vector<std::string> names;
    const char *fname_pa = nullptr;
    bool mode_nop = false, sw_sort = true, sw_ascii = false;
    double eps = 0.5e-2;

    printf("# %s", "catrep (consel_b200 host tool; reference: src/catrep.c)");
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a[0] != '-') names.push_back(a);
        else if (a == "-p") { if (i + 1 >= argc) byebye(); fname_pa = argv[++i]; }
        else if (a == "-d") { if (i + 1 >= argc || sscanf(argv[i + 1], "%d", &debugmode) != 1) byebye(); i++; }
        else if (a == "-e") { if (i + 1 >= argc || sscanf(argv[i + 1], "%lf", &eps) != 1) byebye(); i++; }
        else if (a == "-n") mode_nop = true;
        else if (a == "-m") mode_rmt = true;
        else if (a == "-a") sw_ascii = true;
        else if (a == "-L") {}  // the reference's "large matrix" allocator; same file format
        else if (a == "--no_sort") sw_sort = false;
        else byebye();
    }
    if (names.size() < 2) { die("must specify input-file and output-file"); }
    std::string fname_out = names.back();
    names.pop_back();
    const char *ext = mode_rmt ? ".rmt" : ".rep";

    std::vector<double> want;  // scales of the .pa file
    if (fname_pa) {
        std::string used;
        FILE *fp = open_file(fname_pa, ".pa", "r", &used);
        printf("\n# reading %s", used.c_str());
        want = read_vec(fp);
        fclose(fp);
    }

    // ---- read every input (src/catrep.c:139-163)
    std::vector<RepFile> in(names.size());
    for (size_t f = 0; f < names.size(); f++) {
        std::string used;
        FILE *fp = open_file(names[f], ext, "rb", &used);
        printf("\n# reading %s", used.c_str());
        RepFile &rp = in[f];
        if (!mode_rmt) { rp.ord = read_bivec(fp); rp.cm = (int)rp.ord.size(); }
        rp.obs = read_bvec(fp, rp.cm);
        rp.cm = (int)rp.obs.size();
        std::vector<double> rr = read_bvec(fp);
        std::vector<int> bb = read_bivec(fp, (int)rr.size());
        if (read_bint(fp) != (int)rr.size()) die("wrong number of matrices");
        rp.blocks.resize(rr.size());
        for (size_t i = 0; i < rr.size(); i++) { rp.blocks[i].r = rr[i]; rp.blocks[i].b = bb[i]; }
        if (!mode_nop) {
            for (Block &k : rp.blocks) {
                int m = rp.cm, n = k.b;
                k.x = read_bmat(fp, &m, &n);
                if (m != rp.cm) die("size of rows mismatch in bmat");
                if (n != k.b) die("size of columns mismatch in bmat");
                putdot();
            }
        } else {
            printf(" --- skipped.");
        }
        print_header(rp.cm, rp.blocks);
        fclose(fp);
    }

    // ---- compatibility and the union of scales (:166-200)
    const int cm = in[0].cm;
    for (const RepFile &rp : in) {
        if (rp.cm != cm) die("cm mismatch");
        for (int i = 0; i < cm; i++)
            if (fabs(rp.obs[i] - in[0].obs[i]) > 1e-6) die("obs mismatch");
    }
    std::vector<Block> merged;
    for (const RepFile &rp : in)
        for (const Block &k : rp.blocks) {
            int j = find_scale(merged, k.r, eps);
            if (j < 0) { merged.emplace_back(); merged.back().r = k.r; merged.back().b = k.b; }
            else merged[j].b += k.b;
        }

    // ---- concatenate replicate-wise, file order (:203-229).  With -n nothing was read:
    // the blocks stay zero (the reference dereferences the unread blocks there and crashes);
    // the printed headers are the purpose of -n.
    std::vector<int> filled(merged.size(), 0);
    for (Block &k : merged) k.x.assign((size_t)cm * k.b, 0.0);
    for (const RepFile &rp : in)
        for (const Block &k : rp.blocks) {
            int j = find_scale(merged, k.r, eps);
            if (j < 0) die("internal error");
            Block &dst = merged[j];
            if (!k.x.empty())
                for (int row = 0; row < cm; row++)
                    std::copy(k.x.begin() + (size_t)row * k.b, k.x.begin() + (size_t)(row + 1) * k.b,
                              dst.x.begin() + (size_t)row * dst.b + filled[j]);
            filled[j] += k.b;
        }
    if (in.size() > 1) { printf("\n# merged input"); print_header(cm, merged); }
    for (size_t j = 0; j < merged.size(); j++)
        if (filled[j] != merged[j].b) die("internal error");

    // ---- selection by the .pa scales (:232-259)
    std::vector<Block> out;
    if (fname_pa) {
        for (double r : want) {
            int j = find_scale(merged, r, eps);
            if (j >= 0) out.push_back(merged[j]);
            else warn("scale not found for rr=%g", r);
        }
        printf("\n# selected output");
        print_header(cm, out);
    } else {
        out.swap(merged);
    }

    if (sw_sort) {
        printf("\n# sorting");
        for (Block &k : out) {
            for (int row = 0; row < cm; row++)
                std::sort(k.x.begin() + (size_t)row * k.b, k.x.begin() + (size_t)(row + 1) * k.b);
            putdot();
        }
    }

    // ---- output (:272-297)
    std::string used;
    FILE *fp = open_file(fname_out, ext, sw_ascii ? "w" : "wb", &used);
    printf("\n# writing %s", used.c_str());
    std::vector<double> rr;
    std::vector<int> bb;
    for (const Block &k : out) { rr.push_back(k.r); bb.push_back(k.b); }
    const int kk = (int)out.size();
    if (sw_ascii) {
        if (!mode_rmt) write_ivec(fp, in[0].ord.data(), cm);
        write_vec(fp, in[0].obs.data(), cm);
        write_vec(fp, rr.data(), kk);
        write_ivec(fp, bb.data(), kk);
        fprintf(fp, "%d\n", kk);
        for (const Block &k : out) { write_mat(fp, k.x.data(), cm, k.b); putdot(); }
    } else {
        if (!mode_rmt) write_bivec(fp, in[0].ord.data(), cm);
        write_bvec(fp, in[0].obs.data(), cm);
        write_bvec(fp, rr.data(), kk);
        write_bivec(fp, bb.data(), kk);
        write_bint(fp, kk);
        for (const Block &k : out) { write_bmat(fp, k.x.data(), cm, k.b); putdot(); }
    }
    fclose(fp);
    printf("\n# exit normally\n");
    return 0;
}
