// Stand-in for the tcgen05 tensor-core GEMMs of consel_b200/csrc/i8_kernels.cu in the emulated build (spliced in
// by tests/emu/build_emu_lib.py in place of launch_contract_i8; it sits inside namespace cgpu of that file).
//
// ONLY the matrix-multiply instruction is replaced: T[(k*S+s)][b] = sum_j A8[(k*S+s)][j] * counts[b][j] in int32, by
// plain loops.  Everything around it is the product's own code: the digit matrix A8 comes from fx_scan_kernel /
// fx_slice_kernel (pattern slots from the fx_find_patterns kernels), the count rows from the byte-row histogram
// kernels, the decision fused / unfused is launch_contract_i8's, the fused recombination is the REAL
// fx_finish_block<S> called on 32 x 32 blocks of digit sums laid out as the GEMM epilogue lays them out (transposed,
// stride 33), the unfused one the REAL fx_combine_kernel on the int32 partial sums.
template <int S>
static void emu_fused_blocks(const std::vector<int> &T, size_t ldT, const DevMatrix &m, int nb, const FxEpilogue &ep) {
    constexpr int TREES = 32 / S;
    int tile[32 * 33];
    for (int k0 = 0; k0 < m.K; k0 += TREES)      // 32 digit rows = TREES trees
        for (int b0 = 0; b0 < nb; b0 += 32) {    // 32 replicates
            for (int row = 0; row < 32; row++)
                for (int l = 0; l < 32; l++) {
                    const size_t mrow = (size_t)k0 * S + row;
                    tile[row * 33 + l] = (mrow < m.fxMpad && b0 + l < (int)ldT) ? T[mrow * ldT + b0 + l] : 0;
                }
            for (int lane = 0; lane < 32; lane++) fx_finish_block<S>(tile, lane, k0, b0 + lane, ep);
        }
}

int launch_contract_i8(const DevMatrix &m, const unsigned char *d_counts_rows, int nb, int nb_pad, int *d_partial,
                       double r, double *d_rep, int ldrep, int col0, cudaStream_t st, PhaseTimer *pt, int max_draws) {
    if (nb <= 0) return CGPU_OK;
    static const bool env_one_shot = getenv("CGPU_I8_ONESHOT") != nullptr;
    const bool one_shot = G().opt_i8_oneshot >= 0 ? G().opt_i8_oneshot != 0 : env_one_shot;
    const bool two_cta = !one_shot && (G().opt_i8_2cta == 1 || (G().opt_i8_2cta != 0 && m.NPc / G8_BK > 48)) &&
                         (G().sm_count % 2 == 0 || G().sm_count <= 0);
    const int ksplit = (two_cta && d_partial) ? g2_ksplit(m, nb_pad) : 1;
    const bool fused = !one_shot && fx_can_fuse(m, max_draws) && ksplit == 1;
    if (!fused && !d_partial) {
        set_error(CGPU_EINVAL, "launch_contract_i8: no partial-sum buffer for the unfused recombination");
        return CGPU_EINVAL;
    }
    if (pt) pt->begin(6, st);
    // ---- the "tensor core": int8 x uint8 -> int32, rows of A8 against the count rows of the batch
    const size_t ldT = (size_t)nb_pad;
    std::vector<int> T(m.fxMpad * ldT, 0);
    const size_t rows_used = std::min(m.fxMpad, (size_t)m.K * m.fxS);
    {
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        std::vector<std::thread> th;
        for (unsigned w = 0; w < hw; w++)
            th.emplace_back([&, w] {
                for (size_t mr = w; mr < rows_used; mr += hw) {
                    const signed char *a = m.d_A8 + mr * (size_t)m.NPc;
                    for (int b = 0; b < nb; b++) {
                        const unsigned char *c = d_counts_rows + (size_t)b * m.NPc;
                        int acc = 0;
                        for (int j = 0; j < m.NPc; j++) acc += (int)a[j] * (int)c[j];
                        T[mr * ldT + b] = acc;
                    }
                }
            });
        for (auto &t : th) t.join();
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (pt) pt->end();
    FxEpilogue ep{fused ? d_rep : nullptr, r, {(m.fxF > -1000 && m.fxF < 1000) ? std::ldexp(1.0, -m.fxF) : 0.0, m.fxF},
                  m.K, m.fxS, nb, ldrep, col0, 0};
    if (fused) {
        switch (m.fxS) {
            case 1: emu_fused_blocks<1>(T, ldT, m, nb, ep); break;
            case 2: emu_fused_blocks<2>(T, ldT, m, nb, ep); break;
            case 4: emu_fused_blocks<4>(T, ldT, m, nb, ep); break;
            case 8: emu_fused_blocks<8>(T, ldT, m, nb, ep); break;
            default: set_error(CGPU_EINVAL, "fused recombination with S=%d", m.fxS); return CGPU_EINVAL;
        }
        return CGPU_OK;
    }
    memcpy(d_partial, T.data(), T.size() * sizeof(int));
    if (pt) pt->begin(7, st);
    dim3 cgrid((nb + 255) / 256, m.K);
    emu::launch(dim3(cgrid), dim3(256), (size_t)0, fx_combine_kernel, (const int *)d_partial, nb_pad, m.K, m.fxS, m.fxF, nb, r, d_rep, ldrep, col0);
    CGPU_LAUNCH_CHECK();
    if (pt) pt->end();
    return CGPU_OK;
}
