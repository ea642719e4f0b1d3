// A stand-in libnccl.so.2 for the emulated library (test infrastructure): the nine entry points
// consel_b200/csrc/comm.cu binds, as collectives between PROCESSES of one host over a POSIX shared-memory
// segment named by the unique id.  Buffers are host pointers (device memory of the emulation), data type is
// always fp64.  Every collective: copy the send buffer into the rank's slot, barrier, combine all slots into
// the receive buffer, barrier -- so in-place calls are safe and results do not depend on timing.
#include <fcntl.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>

#include "shim/nccl.h"

namespace {
constexpr size_t kSlot = (size_t)16 << 20;  // bytes per rank and collective
struct Header {
    volatile int ready;
    int world;
    pthread_barrier_t bar;
};
}  // namespace
struct ncclComm {
    char name[64];
    int rank, world;
    Header *hdr;
    unsigned char *slots;
    size_t bytes;
};

static double combine(double a, double b, ncclRedOp_t op) { return op == ncclMax ? std::max(a, b) : op == ncclMin ? std::min(a, b) : a + b; }

extern "C" {
ncclResult_t ncclGetUniqueId(ncclUniqueId *id) {
    memset(id->internal, 0, NCCL_UNIQUE_ID_BYTES);
    timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    snprintf(id->internal, 64, "/cgpu_emu_%d_%ld", (int)getpid(), (long)(ts.tv_nsec ^ ts.tv_sec));
    return ncclSuccess;
}
ncclResult_t ncclCommInitRank(ncclComm_t *out, int world, ncclUniqueId id, int rank) {
    ncclComm *c = new ncclComm();
    memcpy(c->name, id.internal, 63);
    c->name[63] = 0;
    c->rank = rank;
    c->world = world;
    c->bytes = 4096 + (size_t)world * kSlot;
    int fd = -1;
    if (rank == 0) {
        fd = shm_open(c->name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)c->bytes) != 0) return ncclSystemError;
    } else {
        for (int tries = 0; tries < 6000 && fd < 0; tries++) {  // up to ~60 s for rank 0 to create the segment
            fd = shm_open(c->name, O_RDWR, 0600);
            struct stat sb;
            if (fd >= 0 && (fstat(fd, &sb) != 0 || (size_t)sb.st_size < c->bytes)) {
                close(fd);
                fd = -1;
            }
            if (fd < 0) usleep(10000);
        }
        if (fd < 0) return ncclSystemError;
    }
    void *p = mmap(nullptr, c->bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return ncclSystemError;
    c->hdr = static_cast<Header *>(p);
    c->slots = static_cast<unsigned char *>(p) + 4096;
    if (rank == 0) {
        pthread_barrierattr_t a;
        pthread_barrierattr_init(&a);
        pthread_barrierattr_setpshared(&a, PTHREAD_PROCESS_SHARED);
        pthread_barrier_init(&c->hdr->bar, &a, (unsigned)world);
        c->hdr->world = world;
        __sync_synchronize();
        c->hdr->ready = 1;
    } else {
        while (!c->hdr->ready) usleep(1000);
        __sync_synchronize();
    }
    pthread_barrier_wait(&c->hdr->bar);
    *out = c;
    return ncclSuccess;
}
ncclResult_t ncclCommDestroy(ncclComm_t c) {
    if (!c) return ncclSuccess;
    pthread_barrier_wait(&c->hdr->bar);
    munmap(c->hdr, c->bytes);
    if (c->rank == 0) shm_unlink(c->name);
    delete c;
    return ncclSuccess;
}
static ncclResult_t exchange(ncclComm_t c, const void *send, size_t bytes) {
    if (bytes > kSlot) {
        fprintf(stderr, "fake nccl: a %zu-byte message exceeds the %zu-byte slot\n", bytes, kSlot);
        return ncclInvalidArgument;
    }
    memcpy(c->slots + (size_t)c->rank * kSlot, send, bytes);
    pthread_barrier_wait(&c->hdr->bar);
    return ncclSuccess;
}
ncclResult_t ncclAllGather(const void *send, void *recv, size_t count, ncclDataType_t, ncclComm_t c, void *) {
    if (ncclResult_t r = exchange(c, send, count * 8)) return r;
    for (int p = 0; p < c->world; p++) memcpy(static_cast<char *>(recv) + (size_t)p * count * 8, c->slots + (size_t)p * kSlot, count * 8);
    pthread_barrier_wait(&c->hdr->bar);
    return ncclSuccess;
}
ncclResult_t ncclAllReduce(const void *send, void *recv, size_t count, ncclDataType_t, ncclRedOp_t op, ncclComm_t c, void *) {
    if (ncclResult_t r = exchange(c, send, count * 8)) return r;
    double *out = static_cast<double *>(recv);
    for (size_t i = 0; i < count; i++) {
        double v = reinterpret_cast<const double *>(c->slots)[i];
        for (int p = 1; p < c->world; p++) v = combine(v, reinterpret_cast<const double *>(c->slots + (size_t)p * kSlot)[i], op);
        out[i] = v;
    }
    pthread_barrier_wait(&c->hdr->bar);
    return ncclSuccess;
}
ncclResult_t ncclReduceScatter(const void *send, void *recv, size_t recvcount, ncclDataType_t, ncclRedOp_t op, ncclComm_t c, void *) {
    if (ncclResult_t r = exchange(c, send, recvcount * 8 * (size_t)c->world)) return r;
    double *out = static_cast<double *>(recv);
    const size_t base = (size_t)c->rank * recvcount;
    for (size_t i = 0; i < recvcount; i++) {
        double v = reinterpret_cast<const double *>(c->slots)[base + i];
        for (int p = 1; p < c->world; p++) v = combine(v, reinterpret_cast<const double *>(c->slots + (size_t)p * kSlot)[base + i], op);
        out[i] = v;
    }
    pthread_barrier_wait(&c->hdr->bar);
    return ncclSuccess;
}
ncclResult_t ncclGroupStart() { return ncclSuccess; }
ncclResult_t ncclGroupEnd() { return ncclSuccess; }
const char *ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "fake nccl error"; }
}
