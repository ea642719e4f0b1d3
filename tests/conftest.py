import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLD = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gold():
    return GOLD


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def emu_lib(tmp_path_factory):
    """libconselgpu_emu.so: the consel side of the product library compiled for the CPU (tests/emu/build_emu_lib.py),
    built once per session."""
    import ctypes as C
    import shutil
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    sys.path.insert(0, str(ROOT / "tests" / "emu"))
    import build_emu_lib
    path = build_emu_lib.build(tmp_path_factory.mktemp("emulib"))
    lib = C.CDLL(str(path))
    lib._path = path
    lib.cgpu_last_error.restype = C.c_char_p
    lib.cgpu_launch_count.restype = C.c_longlong
    lib.cgpu_smooth_count.restype = C.c_double
    return lib


@pytest.fixture(scope="module")
def emu_cg(emu_lib):
    """consel_b200 bound to the emulated library for the duration of a test module (the binding of the real
    library, with its refusal to compute without a GPU, is restored afterwards)"""
    import consel_b200 as cg
    from consel_b200 import api, pipeline
    saved, saved_host = api._lib, pipeline._host
    api._lib, pipeline._host = emu_lib, emu_lib  # the emulated library also carries libconselhost's sources
    try:
        cg.init()
        cg.set_shard(0, 1)
        yield cg
    finally:
        api._lib, pipeline._host = saved, saved_host


@pytest.fixture(scope="session")
def emu_bin(emu_lib):
    """a bin/ directory whose makermt / makerep / consel are the product's host programs (the sources of the real
    binaries) linked against the EMULATED device library"""
    import subprocess
    host = ROOT / "consel_b200" / "host"
    out = emu_lib._path.parent / "bin"
    out.mkdir(exist_ok=True)
    programs = {  # as consel_b200/host/Makefile links them, with the emulated library in place of libconselgpu.so
        "consel": ["consel.cpp", "pipeline.cpp", "audriver.cpp", "aufit.cpp", "specfun.cpp", "refio.cpp"],
        "makermt": ["makermt.cpp", "refio.cpp"],
        "makerep": ["makerep.cpp", "refio.cpp"],
    }
    procs = []
    for prog, srcs in programs.items():
        cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-pthread", "-w", "-o", str(out / prog), *[str(host / f) for f in srcs],
               f"-L{emu_lib._path.parent}", "-lconselgpu_emu", f"-Wl,-rpath,{emu_lib._path.parent}", "-lm"]
        procs.append((prog, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for prog, p in procs:
        _, err = p.communicate()
        assert p.returncode == 0, f"{prog}: {err[-3000:]}"
    for prog in ("catpv", "catrep", "catmt", "seqmt", "catci"):  # never touch the device: the real binaries
        dst = out / prog
        if not dst.exists():
            dst.symlink_to(ROOT / "consel_b200" / "bin" / prog)
    return out
