"""The bucketed site-histogram kernels (consel_b200/csrc/hist_bucket.cuh, the C4 / genome-length path)
executed ON THE CPU: the kernel source is compiled as C++ against tests/emu/cuda_emu.h (one OS thread per
CUDA thread, real barriers, real atomics) and compared byte for byte with a direct histogram of the same
draws -- index-stream and Philox sources, ragged tiles, 1 to 1024 buckets, the >255 overflow flag -- and
once more under ThreadSanitizer, where a missing __syncthreads shows up as a data race."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
EMU = ROOT / "tests" / "emu"
CSRC = ROOT / "consel_b200" / "csrc"


def _build(tmp: Path, name: str, extra) -> Path:
    exe = tmp / name
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-pthread", "-DCGPU_EMU", f"-I{EMU}", f"-I{CSRC}", *extra, "-o", str(exe),
           str(EMU / "emu_hist_bucket.cpp")]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    return _build(tmp_path_factory.mktemp("emu"), "emu_hist_bucket", ["-DCGPU_EMU_FIBERS"])  # coroutine mode: fast


CASES = [
    # n, draws per replicate, replicates, source (0 index stream, 1 Philox), log2(bucket width), planted draws
    (150000, 120000, 2, 0, 16, 0),   # 3 buckets of 65536 sites, 15 tiles per replicate
    (150000, 120000, 2, 1, 16, 0),
    (40000, 30001, 2, 0, 8, 0),      # 157 buckets (C4 has 153), ragged last tile
    (40000, 30001, 2, 1, 8, 0),
    (200000, 70001, 3, 0, 8, 0),     # 782 buckets
    (200000, 70001, 3, 1, 8, 0),
    (262144, 9000, 3, 0, 8, 0),      # 1024 buckets (the most a replicate may have), sparse
    (1000000, 1400000, 2, 0, 16, 0), # 16 buckets of 65536 sites, 171 tiles per replicate (a C4 row in small)
    (1000000, 1400000, 2, 1, 16, 0),
    (1000, 5, 2, 0, 8, 0),           # fewer draws than threads
    (1000, 5, 2, 1, 8, 0),
    (70000, 8192, 1, 0, 16, 0),      # exactly one tile
    (70000, 8193, 1, 1, 16, 0),      # one tile + one draw
    (65536, 20000, 2, 1, 16, 0),     # a single bucket
    (300, 3000, 2, 0, 8, 300),       # a site drawn more than 255 times: the error flag must rise
    (150000, 100000, 1, 0, 16, 400),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "n%d_bn%d_nb%d_src%d_w%d_hot%d" % c)
def test_bucket_histogram_equals_direct_histogram(exe, case):
    r = subprocess.run([str(exe), *map(str, case)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 mismatches" in r.stdout


def test_bucket_kernels_are_race_free_under_thread_sanitizer(tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    try:
        exe = _build(tmp_path, "emu_hist_bucket_tsan", ["-fsanitize=thread"])
    except subprocess.CalledProcessError as e:  # no libtsan on this host
        pytest.skip("ThreadSanitizer build unavailable: " + e.stderr[-200:])
    for case in ((20000, 20000, 2, 0, 8), (20000, 9000, 1, 1, 8)):
        r = subprocess.run([str(exe), *map(str, case)], capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "ThreadSanitizer" not in r.stderr and " 0 mismatches" in r.stdout


def test_bucket_kernels_stay_in_bounds_and_aligned(tmp_path):
    """AddressSanitizer + UndefinedBehaviorSanitizer: the "device" buffers are exactly sized and 16-byte aligned,
    uint4 is 16-byte aligned as on the device, so an overrun or a misaligned vector access aborts the run
    (what would be an illegal-address / misaligned-address fault on the GPU)."""
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    try:
        exe = _build(tmp_path, "emu_hist_bucket_asan", ["-fsanitize=address,undefined", "-fno-sanitize-recover=all"])
    except subprocess.CalledProcessError as e:
        pytest.skip("sanitizer build unavailable: " + e.stderr[-200:])
    for case in ((150000, 120001, 2, 0, 16), (150000, 120001, 1, 1, 16), (40000, 30001, 1, 0, 8), (1000, 5, 2, 1, 8),
                 (70000, 8193, 1, 0, 16), (300, 3000, 2, 0, 8, 300)):
        r = subprocess.run([str(exe), *map(str, case)], capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr and " 0 mismatches" in r.stdout


@pytest.mark.parametrize("case", [
    # n, draws per replicate, replicates, source, hist_bucket option (0 = default kernels: the test's own control)
    (300000, 200001, 3, 0, 1), (300000, 200001, 3, 1, 1),    # above 225280 sites: the opt-in takes over from global atomics
    (300000, 200001, 3, 0, 0), (300000, 200001, 3, 1, 0),
    (100000, 140000, 9, 0, 2), (100000, 140000, 9, 1, 2),    # forced below that (tests): instead of hist_wide_kernel
    (100000, 140000, 9, 0, 0),
    (70000, 3, 2, 0, 2),                                     # fewer draws than threads
], ids=lambda c: "n%d_bn%d_nb%d_src%d_opt%d" % c)
def test_library_histogram_launchers_byte_rows(emu_lib, tmp_path_factory, case):
    """the library's own launchers (hist_bucket_launch behind cgpu_set_option("hist_bucket", ...)) inside the
    emulated library: workspace sizes, table offsets, grids, launch order"""
    exe = emu_lib._path.parent / "emu_hist_rows"
    if not exe.exists():
        cmd = ["g++", "-O1", "-g", "-std=c++17", "-pthread", "-DCGPU_EMU", "-DCGPU_EMU_FIBERS", "-w", f"-I{EMU / 'shim'}", f"-I{EMU}",
               f"-I{CSRC}", f"-I{ROOT / 'include'}", "-o", str(exe), str(EMU / "emu_hist_rows.cpp"), f"-L{emu_lib._path.parent}",
               "-lconselgpu_emu", f"-Wl,-rpath,{emu_lib._path.parent}"]
        subprocess.run(cmd, check=True, capture_output=True, text=True)
    r = subprocess.run([str(exe), *map(str, case)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 mismatches, err=0" in r.stdout
    launches = int(r.stdout.split(",")[-1].split()[0])
    if case[4] in (1, 2):
        assert launches == 4  # count, scan, scatter, histogram


def test_bucketed_path_through_the_c_abi_on_the_emulated_library(emu_cg, orc):
    """what tests/test_zz_gpu_hist_bucket.py runs on the device (tests/gpu_hist_bucket_check.py: the int8 replicate
    sums with cgpu_set_option("hist_bucket", 1|2) against the default histogram kernels and against the exact sums
    of the oracle's counts), at sizes the CPU finishes in seconds"""
    import sys
    sys.path.insert(0, str(ROOT / "tests"))
    import gpu_hist_bucket_check as chk
    launches = emu_cg.launch_count()
    bad = chk.check(emu_cg, orc, cases=((3, 300000, [0.6, 1.3], [3, 2], 1), (3, 100000, [0.5, 1.2], [20, 3], 2)))
    assert bad == [] and emu_cg.launch_count() > launches
