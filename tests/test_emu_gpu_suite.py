"""`-m gpu` parity tests run UNCHANGED against the emulated build of the product library
(tests/emu/build_emu_lib.py), i.e. on the CPU: the same assertions, the same goldens, the same C ABI --
only the `cg` fixture differs.

  tests/test_gpu_makermt.py  all of it: the device MT19937 stream (incl. jumps of 3e9 and 2^33 draws), site counts
                             on every histogram path, `.rmt` payloads bit-equal to the reference's (tiny, 1 % and
                             FULL-SIZE mam15: the md5 of SURVEY Appendix C), the strict contraction over all tree
                             tiles, the fast mode's tolerance, NaN/inf propagation, stream offsets, sharding, Philox
  tests/test_gpu_consel.py   all of it: statistics and .rep payloads, count queries vs .cnt files and cntdist
                             smoothing, order statistics, KH/SH/wKH/wSH vs oracle and .pv, BP, mixed associations,
                             input validation, row-range queries, uploaded statistics, rescaled statistics
  tests/test_gpu_i8.py       all of it, WITH THE MMA INSTRUCTION REPLACED by a plain int32 product (tests/emu/shim/
                             i8_gemm_emu.h): the fixed-point set-up (exponent scan, digit slicing), site-pattern
                             compression, byte-row histograms, fused draw+histogram kernel, staged index production, both
                             recombinations (the GEMM epilogue's fx_finish_block and fx_combine_kernel) are the product's
  tests/test_gpu_shapes.py   the tests whose shapes the CPU finishes in seconds (several GEMM batches per scale, split-K,
                             33 scales, Philox stream offsets); the C3 / C5 / genome-length shapes stay GPU-only
  tests/test_gpu_dist.py     the single-rank behaviour of the collective entry points

What stays GPU-only is the tcgen05 / TMEM / TMA GEMM itself (its tiling, swizzles, pipelines) and everything about
speed."""
import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))

import test_gpu_consel as _consel  # noqa: E402
import test_gpu_dist as _dist  # noqa: E402
import test_gpu_i8 as _i8  # noqa: E402
import test_gpu_makermt as _makermt  # noqa: E402
import test_gpu_shapes as _shapes  # noqa: E402

# re-collected here without the source modules' `gpu` mark
for _mod in (_consel, _makermt, _i8):
    for _name in dir(_mod):
        if _name.startswith("test_"):
            globals()[_name] = getattr(_mod, _name)
for _name in ("test_philox_stream_offset_gives_independent_replicates", "test_more_than_32_scales_row_range_query",
              "test_several_gemm_batches_per_scale", "test_split_k_pair_gemm"):
    globals()[_name] = getattr(_shapes, _name)
# the single-rank behaviour of the collective entry points (the multi-rank runs are tests/test_emu_dist.py)
for _name in ("test_collective_entry_points_without_communicator", "test_mctest_item_ranges_compose"):
    globals()[_name] = getattr(_dist, _name)


@pytest.fixture(scope="module")
def cg(emu_cg):
    return emu_cg
