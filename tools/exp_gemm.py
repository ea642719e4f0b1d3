"""GEMM time vs inner dimension (development aid): per-tile overhead of the persistent kernel."""
import sys, numpy as np
sys.path.insert(0, ".")
import consel_b200 as cg
cg.init()
K, B = 105, 100000
for n in (1707, 3414, 6828, 13656):
    rng = np.random.Generator(np.random.PCG64(1))
    L = -2.3 - 4.07 * rng.exponential(size=n)[None, :] + 0.1 * rng.standard_normal((K, n))
    m = cg.Matrix(L)
    _, bn, _ = cg.scale_setup(L, [1.0] * 4)
    for fuse in (0,):
        cg.set_option("i8_dbg", fuse)
        for it in range(3):
            r = m.scaleboot(bn, [B] * 4, 7, rng_mode=1, sum_mode=2); t = cg.last_timing(); r.free()
        NP = (n + 127) // 128 * 128
        tiles = 7 * 391
        per_tile_us = t["gemm_i8"] / 4 * 1e3 / (tiles / 148)
        ideal = 2 * 128 * 256 * NP / (2 * 1644.9e12 / 148) * 1e6
        print(f"n={n} kblocks={NP//128} fuse={fuse}: gemm {t['gemm_i8']/4*1e3:.1f} us/launch, {per_tile_us:.2f} us per tile-slot (ideal {ideal:.2f}), combine {t['fx_combine']/4*1e3:.1f} us", flush=True)
    m.free()
