// seqmt — converts the site-wise log-likelihood output of MOLPHY / PAML / PAUP / TREE-PUZZLE
// / PhyML (or an .mt file) into the .mt text matrix makermt reads.  Same command line and
// output as the reference's seqmt (src/seqmt.c:17-104); the readers are the ones our makermt
// uses (refio.cpp, restating src/freadmat.c).  Pure host code.
#include <cstdio>
#include <string>
#include <vector>

#include "refio.h"

using namespace refio;

static void byebye() { die("error in command line"); }

int main(int argc, char **argv) {
    enum { MT, MOLPHY, PAML, PAUP, PUZZLE, PHYML } mode = MT;
    const char *fname_in = nullptr, *fname_out = nullptr;

    printf("# %s", "seqmt (consel_b200 host tool; reference: src/seqmt.c)");
    for (int i = 1, j = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a[0] != '-') {
            if (j == 1) fname_in = argv[i];
            else if (j == 2) fname_out = argv[i];
            else byebye();
            j++;
        } else if (a == "--molphy") mode = MOLPHY;
        else if (a == "--paml") mode = PAML;
        else if (a == "--paup") mode = PAUP;
        else if (a == "--puzzle") mode = PUZZLE;
        else if (a == "--phyml") mode = PHYML;
        else if (a == "-d") { if (i + 1 >= argc || sscanf(argv[i + 1], "%d", &debugmode) != 1) byebye(); i++; }
        else byebye();
    }
    static const char *exts[] = {".mt", ".lls", ".lnf", ".txt", ".sitelh", ".txt"};  // src/freadmat.h:13-18

    FILE *fp = stdin;
    if (fname_in) {
        std::string used;
        fp = open_file(fname_in, exts[mode], "r", &used);
        printf("\n# reading %s", used.c_str());
    } else {
        printf("\n# reading from stdin");
    }
    int m = 0, n = 0;
    std::vector<double> x;
    switch (mode) {
        case MOLPHY: x = read_mat_molphy(fp, &m, &n); break;
        case PAML: x = read_mat_paml(fp, &m, &n); break;
        case PAUP: x = read_mat_paup(fp, &m, &n); break;
        case PUZZLE: x = read_mat_puzzle(fp, &m, &n); break;
        case PHYML: x = read_mat_phyml(fp, &m, &n); break;
        default: x = read_mat(fp, &m, &n); break;
    }
    if (fname_in) fclose(fp);
    printf("\n# M:%d N:%d", m, n);

    std::string out_name;
    if (fname_out) out_name = fname_out;
    else if (fname_in) out_name = strip_extension(fname_in);
    if (!out_name.empty()) {
        std::string used;
        fp = open_file(out_name, ".mt", "w", &used);
        printf("\n# writing %s", used.c_str());
    } else {
        fp = stdout;
        printf("\n# writing to stdout\n");
    }
    write_mat(fp, x.data(), m, n);
    if (!out_name.empty()) fclose(fp);
    putchar('\n');
    return 0;
}
