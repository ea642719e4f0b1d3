// Entry points of the translation units that are NOT part of the emulated library (see
// build_emu_lib.py): the tcgen05 tensor-core path (i8_kernels.cu) needs the GPU and fails with
// CGPU_ENODEV here; the communicator (comm.cu, NCCL) degenerates to a single rank.
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

int comm_world() { return 1; }
int comm_rank() { return 0; }
int comm_group_start() { return CGPU_OK; }
int comm_group_end() { return CGPU_OK; }
int comm_allgather_f64(const double *send, double *recv, size_t n, cudaStream_t) {
    if (send != recv) memmove(recv, send, n * sizeof(double));
    return CGPU_OK;
}
int comm_allreduce_f64(const double *send, double *recv, size_t n, int, cudaStream_t) {
    if (send != recv) memmove(recv, send, n * sizeof(double));
    return CGPU_OK;
}
int comm_reduce_scatter_f64(const double *send, double *recv, size_t n, int, cudaStream_t) {
    if (send != recv) memmove(recv, send, n * sizeof(double));
    return CGPU_OK;
}
}  // namespace cgpu

extern "C" {
int cgpu_comm_allreduce_host(double *, long long, int) { return CGPU_OK; }
int cgpu_comm_item_range(int cm, int *row0, int *nrows) {
    if (row0) *row0 = 0;
    if (nrows) *nrows = cm;
    return CGPU_OK;
}
int cgpu_comm_info(int *rank, int *world) {
    if (rank) *rank = 0;
    if (world) *world = 1;
    return CGPU_OK;
}
void cgpu_comm_destroy(void) {}
int cgpu_comm_init_env(void) {  // single rank: what the real entry point does when WORLD_SIZE <= 1
    if (int rc = cgpu::ensure_init()) return rc;
    return cgpu_set_shard(0, 1);
}
int cgpu_comm_allgather_host(const void *send, void *recv, long long bytes) {
    if (send != recv && bytes > 0) memmove(recv, send, (size_t)bytes);
    return CGPU_OK;
}
}
