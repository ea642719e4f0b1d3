// Entry points of the translation units that are NOT part of the emulated library (see
// build_emu_lib.py): none of the compute paths any more.  (comm.cu is compiled in as it is; the libnccl.so.2 it dlopens is tests/emu/fake_nccl.cpp.)
#include <string.h>

#include "cgpu_internal.h"

namespace cgpu {
static int gpu_only(const char *what) {
    set_error(CGPU_ENODEV, "%s needs the GPU (the emulated library has everything but the tensor-core path)", what);
    return CGPU_ENODEV;
}
// (nothing left to stub inside namespace cgpu: the int8 path is compiled in with a stand-in for the MMA,
//  tests/emu/shim/i8_gemm_emu.h; gpu_only() stays for future use)
}  // namespace cgpu
