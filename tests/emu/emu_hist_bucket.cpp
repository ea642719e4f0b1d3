// Runs the bucketed site-histogram kernels of consel_b200/csrc/hist_bucket.cuh on the CPU
// (tests/emu/cuda_emu.h) and compares the count rows byte for byte with a direct histogram of
// the same draws (the loop of src/makermt.c:53-62).
//
//   emu_hist_bucket n bn nb src(0 index stream | 1 philox) wbl [hot]
//
// hot > 0: `hot` extra draws of one site are planted in replicate 0 (a count above 255 must raise
// the error flag).  Exit status 0 = all rows equal and the flag is as expected.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cuda_emu.h"
#include "hist_bucket.cuh"

using namespace cgpu;

template <int SRC, int WBL>
static int run(uint32_t n, int bn, int nb, int hot) {
    const int NP = (int)((n + 127) / 128 * 128);
    const int nbk = (NP + (1 << WBL) - 1) >> WBL;
    if (nbk > BK_MAXB) {
        fprintf(stderr, "too many buckets (%d)\n", nbk);
        return 2;
    }
    const uint32_t key0 = 0x1234567u, key1 = 0x89abcdefu, scale = 3, b_first = 1000, ctr3 = 77;
    std::vector<uint32_t> idx;
    if (SRC == 0) {
        idx.resize((size_t)nb * bn);
        uint64_t st = 0x9E3779B97F4A7C15ull ^ n;
        for (auto &k : idx) {
            st ^= st << 13; st ^= st >> 7; st ^= st << 17;
            k = (uint32_t)(((st >> 32) * (uint64_t)n) >> 32);
        }
        for (int h = 0; h < hot && h < bn; h++) idx[h] = n / 3;
    }
    // ---- direct histogram
    std::vector<unsigned> want((size_t)nb * NP, 0);
    for (int r = 0; r < nb; r++)
        for (int i = 0; i < bn; i++) {
            uint32_t k;
            if (SRC == 0) {
                k = idx[(size_t)r * bn + i];
            } else {
                const Philox4 p = philox4x32_10((uint32_t)(i >> 2), b_first + (uint32_t)r, scale, ctr3, key0, key1);
                k = __umulhi(p.v[i & 3], n);
            }
            want[(size_t)r * NP + k]++;
        }
    // ---- the kernels
    const size_t pstride = (size_t)((bn + 7) / 8 * 8) + (size_t)8 * nbk;
    std::vector<unsigned> cnt((size_t)nb * nbk, 0), off((size_t)nb * nbk, 0xdeadbeefu), cur((size_t)nb * nbk, 0xdeadbeefu);
    // "device" buffers: exactly sized and 16-byte aligned (cudaMalloc gives 256), so that the sanitizer builds see
    // any overrun or misaligned 16-byte access
    unsigned short *part = static_cast<unsigned short *>(aligned_alloc(16, ((size_t)nb * pstride * 2 + 15) / 16 * 16));
    unsigned char *rows_al = static_cast<unsigned char *>(aligned_alloc(16, (size_t)nb * NP));
    memset(part, 0xff, (size_t)nb * pstride * 2);
    memset(rows_al, 0xee, (size_t)nb * NP);
    int err = 0;
    const int tiles = (bn + BK_TILE - 1) / BK_TILE;
    const uint32_t *d_idx = SRC == 0 ? idx.data() : nullptr;
    emu::launch(dim3(tiles, nb), dim3(BK_THREADS), 0, bucket_count_kernel<SRC, WBL>, d_idx, n, bn, nbk, cnt.data(), key0, key1,
                scale, b_first, ctr3);
    emu::launch(dim3((nb + 127) / 128), dim3(128), 0, bucket_scan_kernel, (const unsigned *)cnt.data(), nb, nbk, off.data(),
                cur.data());
    emu::launch(dim3(tiles, nb), dim3(BK_THREADS), bk_scatter_smem(), bucket_scatter_kernel<SRC, WBL>, d_idx, n, bn, nbk,
                (const unsigned *)off.data(), cur.data(), part, pstride, key0, key1, scale, b_first, ctr3);
    emu::launch(dim3(nbk, nb), dim3(BKH_THREADS), (size_t)1 << WBL, bucket_hist_kernel<WBL>, (const unsigned short *)part,
                pstride, (const unsigned *)cnt.data(), (const unsigned *)off.data(), nbk, NP, rows_al, &err);
    // ---- compare
    long long bad = 0;
    bool over = false;
    for (size_t e = 0; e < (size_t)nb * NP; e++) {
        if (want[e] > 255) over = true;
        if (rows_al[e] != (unsigned char)want[e]) bad++;
    }
    if (over) bad = 0;  // a wrapped byte carries into its neighbour: the rows are void, the flag is what counts
    for (int r = 0; r < nb; r++) {  // the cursors must have handed out exactly the bucket totals
        unsigned tot = 0;
        for (int t = 0; t < nbk; t++) {
            if (cur[(size_t)r * nbk + t] != cnt[(size_t)r * nbk + t]) bad++;
            tot += cnt[(size_t)r * nbk + t];
        }
        if (tot != (unsigned)bn) bad++;
    }
    // ---- the row checksum of the global-atomic path on the same rows: silent on exact rows, raised by a wrap
    {
        unsigned char *ex = static_cast<unsigned char *>(aligned_alloc(16, (size_t)nb * NP));
        for (size_t e = 0; e < (size_t)nb * NP; e++) ex[e] = (unsigned char)want[e];  // a count above 255 loses 256 here
        int err2 = 0;
        emu::launch(dim3(nb), dim3(RCS_THREADS), 0, row_checksum_kernel, (const unsigned char *)ex, NP, bn, &err2);
        if (err2 != (over ? 1 : 0)) bad++;
        free(ex);
    }
    printf("n=%u NP=%d bn=%d nb=%d src=%d wbl=%d buckets=%d tiles=%d: %lld mismatches, err=%d (expected %d)\n", n, NP, bn, nb, SRC,
           WBL, nbk, tiles, bad, err, over ? 1 : 0);
    free(part);
    free(rows_al);
    return (bad == 0 && err == (over ? 1 : 0)) ? 0 : 1;
}

int main(int argc, char **argv) {
    if (argc < 6) {
        fprintf(stderr, "usage: %s n bn nb src wbl [hot]\n", argv[0]);
        return 2;
    }
    const uint32_t n = (uint32_t)atoll(argv[1]);
    const int bn = atoi(argv[2]), nb = atoi(argv[3]), src = atoi(argv[4]), wbl = atoi(argv[5]), hot = argc > 6 ? atoi(argv[6]) : 0;
    if (wbl == 16) return src ? run<1, 16>(n, bn, nb, hot) : run<0, 16>(n, bn, nb, hot);
    if (wbl == 8) return src ? run<1, 8>(n, bn, nb, hot) : run<0, 8>(n, bn, nb, hot);
    fprintf(stderr, "wbl must be 16 or 8\n");
    return 2;
}
