"""Builds libconselgpu_emu.so: libconselgpu.so WITHOUT its tensor-core path -- the C ABI (capi.cu, capi_stats.cu),
the MT19937 jump-ahead and draw kernels (mt_kernels.cu), the histogram kernels (resample_kernels.cu), the
fp64 contractions with their mbarrier / bulk-copy pipeline (contract_kernels.cu), the statistics, count-query
and KH/SH kernels (stats_kernels.cu) and the native pipeline / AU driver (consel_b200/host) -- compiled FOR
THE CPU, so that the `-m gpu` parity tests of the strict (bit-exact) makermt path and of the whole consel side
run in the CPU test tier against the same oracle and goldens as on the device.  Test infrastructure only: the
product library is built by nvcc from the very same files (consel_b200/csrc/Makefile), nothing of the product
loads this one, and the product still refuses to compute without an sm_100 device.

How: the sources are taken as they are, with mechanical rewrites --
    kernel<<<grid, block, smem, stream>>>(args);     ->  emu::launch(dim3(grid), dim3(block), smem, kernel, args);
    extern __shared__ [__align__(N)] T name[];       ->  T *name = reinterpret_cast<T *>(emu::dyn_smem());
    three one-instruction inline-PTX statements of mt_kernels.cu -> the same instruction in C++ (PATCHES)
    the mbarrier / cp.async.bulk helper block of contract_kernels.cu -> tests/emu/shim/mbarrier_emu.h (BLOCKS)
-- and compiled by g++ against tests/emu/shim/cuda_runtime.h (host memory as device memory, synchronous
streams) and tests/emu/cuda_emu.h (CUDA threads as coroutines or OS threads, real barriers, real atomics).
What is NOT emulated: the tcgen05 / TMEM / TMA GEMM of the int8 mode (i8_kernels.cu), the peak-measurement
kernels (bench_kernels.cu); emu_stubs.cpp answers for them (CGPU_ENODEV).  comm.cu is compiled as it is; the
libnccl.so.2 it binds at run time is tests/emu/fake_nccl.cpp (collectives between processes over shared memory).
"""
from __future__ import annotations

import re
import subprocess
import sys
from pathlib import Path

EMU = Path(__file__).resolve().parent
ROOT = EMU.parent.parent
CSRC = ROOT / "consel_b200" / "csrc"
SOURCES = ["capi.cu", "capi_stats.cu", "stats_kernels.cu", "resample_kernels.cu", "mt_kernels.cu", "contract_kernels.cu", "comm.cu", "i8_kernels.cu"]
# Inline PTX cannot be compiled for the host: the few helper functions that contain it are replaced, by exact text,
# with C++ statements of the same instruction (every patch must apply exactly once, or the build fails).
PATCHES = {
    "mt_kernels.cu": [
        ('asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(pow2_32_minus_s));', "r = __umulhi(v, pow2_32_minus_s);"),
        ('asm("lop3.b32 %0, %1, %2, 0x80000000, 0xe4;" : "=r"(y) : "r"(hi), "r"(lo));', "y = (hi & 0x80000000u) | (lo & 0x7fffffffu);"),
        ('asm("mul.lo.u32 %0, %1, %2;" : "=r"(mag) : "r"(lo & 1u), "r"(k.mag));', "mag = (lo & 1u) * k.mag;"),
    ],
}
# contract_kernels.cu: the block of mbarrier / cp.async.bulk helpers -> tests/emu/shim/mbarrier_emu.h
# i8_kernels.cu: the tcgen05 / TMEM / TMA GEMM kernels with their tensor-map helpers are cut, and the launcher that
# drives them is replaced by tests/emu/shim/i8_gemm_emu.h (a plain int32 matrix product IN PLACE OF THE MMA; the
# fixed-point set-up, the pattern kernels and both recombinations stay the product's own code)
BLOCKS = {
    "contract_kernels.cu": [("// ------------------------------------------------------------------ PTX helpers",
                             "// ------------------------------------------------------------------ L -> site-major tiles",
                             '}  // namespace cgpu\n#include "mbarrier_emu.h"\nnamespace cgpu {\n')],
    "i8_kernels.cu": [("// ------------------------------------------------------------------ tcgen05 GEMM tiles",
                       "// ---- site patterns: identical columns of L",
                       "constexpr int G8_BM = 128, G8_BN = 256, G8_BK = 128;\n"),
                      ("// counts: plain rows [nb_pad][NP] uint8 (nb_pad multiple of 256, padding rows zero)",
                       "}  // namespace cgpu",
                       '#include <thread>\n#include "i8_gemm_emu.h"\n')],
}
HOST = ROOT / "consel_b200" / "host"
# the native makermt->consel pipeline and AU driver (libconselhost.so) go INTO the emulated library, so that their
# cgpu_* calls bind to the emulated entry points whatever else the process has loaded
HOST_SOURCES = ["pipeline.cpp", "audriver.cpp", "aufit.cpp", "specfun.cpp", "refio.cpp"]

LAUNCH = re.compile(r"(?P<kern>[A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*<<<(?P<cfg>[^;]*?)>>>\s*\((?P<args>[^;]*?)\);")
DYN = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?(?P<type>[A-Za-z_][\w ]*?)\s+(?P<name>\w+)\[\];")


def split_top(s: str) -> list[str]:
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":  # (angle brackets are not tracked: `1 << 16` is a shift; a template argument list with a
            depth += 1   #  comma inside a launch configuration would yield more than four parts and be refused)
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite(text: str) -> tuple[str, int]:
    n = 0

    def launch(m):
        nonlocal n
        cfg = split_top(m.group("cfg"))
        if not 2 <= len(cfg) <= 4:
            raise ValueError(f"launch configuration not understood: {m.group(0)}")
        smem = cfg[2] if len(cfg) > 2 else "0"
        n += 1
        return f"emu::launch(dim3({cfg[0]}), dim3({cfg[1]}), (size_t)({smem}), {m.group('kern')}, {m.group('args')});"

    text = LAUNCH.sub(launch, text)
    text = DYN.sub(lambda m: f"{m.group('type')} *{m.group('name')} = reinterpret_cast<{m.group('type')} *>(emu::dyn_smem());", text)
    if "<<<" in text:
        raise ValueError("a kernel launch survived the rewrite")
    return text, n


def build(out_dir: Path, extra_flags=(), fibers: bool = True) -> Path:
    """fibers: coroutine execution (fast; several CTAs at once).  False: one OS thread per CUDA thread, CTAs in
    order -- the mode the sanitizers understand."""
    out_dir.mkdir(parents=True, exist_ok=True)
    cpps = []
    for name in SOURCES:
        src = (CSRC / name).read_text()
        for old, new in PATCHES.get(name, []):
            if src.count(old) != 1:
                raise ValueError(f"{name}: patch target not found exactly once: {old[:50]}")
            src = src.replace(old, new)
        for a, b, repl in BLOCKS.get(name, []):
            if src.count(a) != 1 or src.count(b) < 1:
                raise ValueError(f"{name}: block markers not found")
            ia = src.index(a)
            src = src[:ia] + repl + src[src.index(b, ia):]
        text, n = rewrite(src)
        dst = out_dir / (Path(name).stem + "_emu.cpp")
        dst.write_text(f'#line 1 "{CSRC / name}"\n' + text)
        cpps.append(dst)
    lib = out_dir / "libconselgpu_emu.so"
    # -Bsymbolic: the library's own calls bind to ITSELF even when the real libconselgpu.so (same symbol names) is
    # already loaded into the process with global visibility (libconselhost.so pulls it in that way)
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-DCGPU_EMU", "-w", "-Wl,-Bsymbolic",
           *(["-DCGPU_EMU_FIBERS"] if fibers else []),
           f"-I{EMU / 'shim'}", f"-I{EMU}", f"-I{CSRC}", f"-I{ROOT / 'include'}", f"-I{HOST}", *extra_flags, "-o", str(lib),
           *map(str, cpps), str(EMU / "emu_stubs.cpp"), str(CSRC / "gf2_mt.cpp"), *[str(HOST / h) for h in HOST_SOURCES]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulated library failed to build:\n" + r.stderr[-6000:])
    # the libnccl.so.2 that comm.cu dlopens in a multi-rank run: collectives between processes over shared memory
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-pthread", f"-I{EMU}", "-o", str(out_dir / "libnccl.so.2"),
                        str(EMU / "fake_nccl.cpp"), "-lrt"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("stand-in NCCL failed to build:\n" + r.stderr[-3000:])
    return lib


if __name__ == "__main__":
    print(build(Path(sys.argv[1]) if len(sys.argv) > 1 else EMU / "_build"))
