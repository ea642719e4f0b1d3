// Drives the library's OWN histogram launchers (cgpu::launch_hist_from_idx / launch_hist_philox of
// consel_b200/csrc/resample_kernels.cu, byte-row layout = the operand of the int8 contraction) inside the
// emulated library and compares the count rows with a direct histogram of the same draws.  With
// cgpu_set_option("hist_bucket", 1|2) this is hist_bucket_launch -- its workspace sizes, table offsets, grids
// and launch order -- which the C ABI can only reach through the (GPU-only) tensor-core contraction.
//
//   emu_hist_rows n bn nb src(0 index stream | 1 philox) hist_bucket_option
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cgpu_internal.h"
#include "philox.cuh"

using namespace cgpu;

int main(int argc, char **argv) {
    if (argc < 6) {
        fprintf(stderr, "usage: %s n bn nb src opt\n", argv[0]);
        return 2;
    }
    const int n = atoi(argv[1]), bn = atoi(argv[2]), nb = atoi(argv[3]), src = atoi(argv[4]), opt = atoi(argv[5]);
    if (cgpu_init(0) != CGPU_OK || cgpu_set_option("hist_bucket", opt) != CGPU_OK) {
        fprintf(stderr, "init: %s\n", cgpu_last_error());
        return 2;
    }
    const int NP = pad_sites(n);
    const bool is16 = idx_is16((uint32_t)n);
    cudaStream_t st = G().stream;
    const unsigned long long seed = 0x123456789abcdefull;
    const int scale = 4, b_first = 77;
    void *d_idx = nullptr;
    unsigned long long *d_counts = nullptr;
    int *d_err = nullptr;
    const size_t rows_alloc = (size_t)((nb + 7) / 8) * 8 * NP;  // what scaleboot_core reserves per batch (whole groups of 8)
    cudaMalloc(&d_counts, rows_alloc);
    cudaMalloc(&d_err, sizeof(int));
    memset(d_counts, 0xee, rows_alloc);
    *d_err = 0;
    std::vector<unsigned> want((size_t)nb * NP, 0);
    if (src == 0) {
        cudaMalloc(&d_idx, (size_t)nb * bn * (is16 ? 2 : 4));
        uint64_t s = 0x9E3779B97F4A7C15ull ^ (uint64_t)n;
        for (size_t e = 0; e < (size_t)nb * bn; e++) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            const uint32_t k = (uint32_t)(((s >> 32) * (uint64_t)n) >> 32);
            if (is16) static_cast<uint16_t *>(d_idx)[e] = (uint16_t)k; else static_cast<uint32_t *>(d_idx)[e] = k;
            want[(e / bn) * NP + k]++;
        }
        if (launch_hist_from_idx(d_idx, n, NP, bn, nb, d_counts, d_err, st, true, nullptr) != CGPU_OK) {
            fprintf(stderr, "launch: %s\n", cgpu_last_error());
            return 2;
        }
    } else {
        for (int r = 0; r < nb; r++)
            for (int i = 0; i < bn; i++) {
                const Philox4 p = philox4x32_10((uint32_t)(i >> 2), (uint32_t)(b_first + r), (uint32_t)scale, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
                want[(size_t)r * NP + __umulhi(p.v[i & 3], (uint32_t)n)]++;
            }
        if (launch_hist_philox(seed, scale, b_first, n, NP, bn, nb, d_counts, d_err, st, true, 0, nullptr) != CGPU_OK) {
            fprintf(stderr, "launch: %s\n", cgpu_last_error());
            return 2;
        }
    }
    cudaStreamSynchronize(st);
    const unsigned char *rows = reinterpret_cast<const unsigned char *>(d_counts);
    long long bad = 0;
    for (size_t e = 0; e < (size_t)nb * NP; e++)
        if (rows[e] != (unsigned char)want[e]) bad++;
    printf("n=%d NP=%d bn=%d nb=%d src=%d hist_bucket=%d: %lld mismatches, err=%d, %lld launches\n", n, NP, bn, nb, src, opt, bad, *d_err,
           cgpu_launch_count());
    return (bad == 0 && *d_err == 0) ? 0 : 1;
}
