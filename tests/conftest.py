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
    lib = C.CDLL(str(build_emu_lib.build(tmp_path_factory.mktemp("emulib"))))
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
