"""consel_b200 — B200-native hot path of CONSEL's multiscale RELL bootstrap.

This package is a thin host-side mirror (ctypes) of the C ABI in
``include/consel_gpu.h``; all compute happens in ``libconselgpu.so`` (hand-written
sm_100a CUDA, built in-tree by ``__graft_entry__.build()``).  There is no CPU
fallback: if the shared library is missing or no B200 is visible, every compute
call raises :class:`CgpuError`.

Names follow the reference (shimo-lab/consel): ``scaleboot`` is
``src/makermt.c:35-75``, ``repminmaxs`` is ``src/consel.c:1939-1959`` and so on.
"""
from .api import (  # noqa: F401
    CgpuError,
    Matrix,
    Reps,
    RNG_MT19937,
    RNG_PHILOX,
    SUM_FAST,
    SUM_INT8,
    SUM_STRICT,
    bench_peaks,
    exported_symbols,
    kernel_times,
    init,
    last_timing,
    launch_count,
    lib_path,
    load,
    makermt,
    mt19937_stream,
    resample_counts,
    scale_setup,
    scaleboot_all,
    set_option,
    set_shard,
    shard_range,
    smooth_count,
)

__all__ = [n for n in dir() if not n.startswith("_")]
