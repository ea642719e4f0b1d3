"""`-m gpu` parity tests run UNCHANGED against the emulated build of the product library
(tests/emu/build_emu_lib.py), i.e. on the CPU: the same assertions, the same goldens, the same C ABI --
only the `cg` fixture differs.  All of tests/test_gpu_consel.py (statistics and .rep payloads, count queries
vs .cnt files and cntdist smoothing, order statistics, KH/SH/wKH/wSH vs oracle and .pv, BP, mixed
associations, input validation, row-range queries, uploaded statistics, rescaled statistics) and the tests
of the other GPU modules that need neither the MT19937 stream nor a contraction (those stay GPU-only:
tensor cores, TMA and inline PTX are not emulated)."""
import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))

import test_gpu_consel as _consel  # noqa: E402
import test_gpu_makermt as _makermt  # noqa: E402
import test_gpu_shapes as _shapes  # noqa: E402

# re-collected here without the source modules' `gpu` mark
for _name in dir(_consel):
    if _name.startswith("test_"):
        globals()[_name] = getattr(_consel, _name)
for _mod, _names in ((_makermt, ("test_philox_mode_is_a_valid_multinomial", "test_bad_arguments_fail_loudly")),
                     (_shapes, ("test_philox_stream_offset_gives_independent_replicates",
                                "test_more_than_32_scales_row_range_query"))):
    for _name in _names:
        globals()[_name] = getattr(_mod, _name)


@pytest.fixture(scope="module")
def cg(emu_cg):
    return emu_cg
