// Entry points of the translation units that are NOT part of the emulated library (see
// build_emu_lib.py): the tcgen05 tensor-core path (i8_kernels.cu) needs the GPU and fails with
// CGPU_ENODEV here.  (comm.cu is compiled in as it is; the libnccl.so.2 it dlopens is tests/emu/fake_nccl.cpp.)
#include <string.h>

#include "cgpu_internal.h"

namespace cgpu {
static int gpu_only(const char *what) {
    set_error(CGPU_ENODEV, "%s needs the GPU (the emulated library has everything but the tensor-core path)", what);
    return CGPU_ENODEV;
}
int fx_prepare(DevMatrix &, cudaStream_t, bool) { return gpu_only("fx_prepare"); }
bool fx_needs_partial(const DevMatrix &, int, int) { return false; }
int launch_contract_i8(const DevMatrix &, const unsigned char *, int, int, int *, double, double *, int, int, cudaStream_t, PhaseTimer *, int) {
    return gpu_only("the tcgen05 GEMM");
}

}  // namespace cgpu
