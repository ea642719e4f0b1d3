// Development tool (not shipped to users): times the INPUT PATH of makermt -- refio::read_mat (.mt) or
// refio::read_mat_paml (PAML .lnf), the readers behind src/misc.c:385-404 and src/freadmat.c:110-154 --
// on a file, and prints a checksum of every value read (tools/parse_time.py compares it with the values
// it wrote and times the unmodified reference on the same file).
//
//   parse_bench mt|paml FILE [repeat=1]
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "refio.h"

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: parse_bench mt|paml FILE [repeat]\n");
        return 2;
    }
    const bool paml = !strcmp(argv[1], "paml");
    const int repeat = argc > 3 ? atoi(argv[3]) : 1;
    double best = 1e30;
    for (int it = 0; it < repeat; it++) {
        FILE *fp = fopen(argv[2], "r");
        if (!fp) {
            perror(argv[2]);
            return 1;
        }
        fseek(fp, 0, SEEK_END);
        const long bytes = ftell(fp);
        fseek(fp, 0, SEEK_SET);
        int m = 0, n = 0;
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<double> a = paml ? refio::read_mat_paml(fp, &m, &n) : refio::read_mat(fp, &m, &n);
        const double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        fclose(fp);
        uint64_t h = 1469598103934665603ull;
        for (double x : a) {
            uint64_t b;
            memcpy(&b, &x, 8);
            h = (h ^ b) * 1099511628211ull;
        }
        best = t < best ? t : best;
        printf("{\"format\": \"%s\", \"trees\": %d, \"sites\": %d, \"file_bytes\": %ld, \"seconds\": %.4f, \"MB_per_s\": %.1f, "
               "\"values_per_s\": %.4g, \"fnv1a_of_values\": \"%016llx\"}\n",
               argv[1], m, n, bytes, t, bytes / t / 1e6, (double)a.size() / t, (unsigned long long)h);
    }
    return 0;
}
