// One process per GPU: the library's own NCCL communicator.
//
// The makermt side needs no data-path collective (every rank jumps the MT19937 stream to its
// own replicate columns).  The consel side combines per-rank row summaries and gathers the one
// replicate block the KH/SH tests need whole; those collectives are issued from C++ on the
// library stream, so a fit round costs one H2D, one all-gather, one counting kernel, one grouped
// reduce-scatter and one D2H -- no host framework in the loop.
//
// libnccl.so.2 is bound at run time (dlopen + dlsym): a single-GPU run never loads it, and a
// process that already carries a copy (a host framework's bundled NCCL has the same soname)
// shares it.  The 128-byte unique id travels either through the caller (cgpu_comm_unique_id /
// cgpu_comm_init: any launcher can broadcast it) or through a minimal TCP rendezvous on
// MASTER_ADDR:MASTER_PORT+17 (cgpu_comm_init_env: the torchrun environment, no Python needed).
#include <arpa/inet.h>
#include <dlfcn.h>
#include <nccl.h>  // types and enums only; every function is bound with dlsym
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <unistd.h>

#include <chrono>
#include <thread>

#include "cgpu_internal.h"

namespace cgpu {

namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // the copy this process already has
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error(CGPU_ENODEV, "libnccl.so.2 cannot be loaded: %s", dlerror());
        return nullptr;
    }
    bool ok = true;
    auto bind = [&](auto &fn, const char *name) {
        void *p = dlsym(h, name);
        if (!p) ok = false;
        fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(p);
    };
    bind(api.GetUniqueId, "ncclGetUniqueId");
    bind(api.CommInitRank, "ncclCommInitRank");
    bind(api.CommDestroy, "ncclCommDestroy");
    bind(api.AllReduce, "ncclAllReduce");
    bind(api.AllGather, "ncclAllGather");
    bind(api.ReduceScatter, "ncclReduceScatter");
    bind(api.GroupStart, "ncclGroupStart");
    bind(api.GroupEnd, "ncclGroupEnd");
    bind(api.GetErrorString, "ncclGetErrorString");
    if (!ok) {
        set_error(CGPU_ENODEV, "libnccl.so.2 lacks a required entry point");
        return nullptr;
    }
    api.handle = h;
    return &api;
}

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    // grow-only staging for the host-buffer helpers
    void *d_stage = nullptr;
    size_t d_cap = 0;
};
Comm &C() {
    static Comm c;
    return c;
}
}  // namespace

#define CGPU_NCCL(call)                                                                                  \
    do {                                                                                                 \
        ncclResult_t r__ = (call);                                                                       \
        if (r__ != ncclSuccess) {                                                                        \
            NcclApi *a__ = nccl_api();                                                                   \
            set_error(CGPU_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,                             \
                      a__ && a__->GetErrorString ? a__->GetErrorString(r__) : "NCCL error");             \
            return CGPU_ECUDA;                                                                           \
        }                                                                                                \
    } while (0)

int comm_world() { return C().comm ? C().world : 1; }
int comm_rank() { return C().comm ? C().rank : 0; }

static ncclRedOp_t red_op(int op) { return op == 1 ? ncclMax : op == 2 ? ncclMin : ncclSum; }

int comm_group_start() {
    if (!C().comm) return CGPU_OK;
    CGPU_NCCL(nccl_api()->GroupStart());
    return CGPU_OK;
}
int comm_group_end() {
    if (!C().comm) return CGPU_OK;
    CGPU_NCCL(nccl_api()->GroupEnd());
    return CGPU_OK;
}
int comm_allgather_f64(const double *send, double *recv, size_t count, cudaStream_t st) {
    if (!C().comm) {
        if (send != recv) CGPU_CUDA(cudaMemcpyAsync(recv, send, count * 8, cudaMemcpyDeviceToDevice, st));
        return CGPU_OK;
    }
    CGPU_NCCL(nccl_api()->AllGather(send, recv, count, ncclFloat64, C().comm, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return CGPU_OK;
}
int comm_allreduce_f64(const double *send, double *recv, size_t count, int op, cudaStream_t st) {
    if (!C().comm) {
        if (send != recv) CGPU_CUDA(cudaMemcpyAsync(recv, send, count * 8, cudaMemcpyDeviceToDevice, st));
        return CGPU_OK;
    }
    CGPU_NCCL(nccl_api()->AllReduce(send, recv, count, ncclFloat64, red_op(op), C().comm, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return CGPU_OK;
}
int comm_reduce_scatter_f64(const double *send, double *recv, size_t recvcount, int op, cudaStream_t st) {
    if (!C().comm) {
        if (send != recv) CGPU_CUDA(cudaMemcpyAsync(recv, send, recvcount * 8, cudaMemcpyDeviceToDevice, st));
        return CGPU_OK;
    }
    CGPU_NCCL(nccl_api()->ReduceScatter(send, recv, recvcount, ncclFloat64, red_op(op), C().comm, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return CGPU_OK;
}

static int stage_get(size_t bytes, void **out) {
    Comm &c = C();
    if (c.d_cap < bytes) {
        if (c.d_stage) cudaFree(c.d_stage);
        c.d_stage = nullptr;
        c.d_cap = 0;
        CGPU_CUDA(cudaMalloc(&c.d_stage, bytes + bytes / 4 + 4096));
        c.d_cap = bytes + bytes / 4 + 4096;
    }
    *out = c.d_stage;
    return CGPU_OK;
}

// ---- TCP rendezvous for the unique id (rank 0 listens, everybody else connects)
static int send_all(int fd, const char *p, size_t n) {
    while (n) {
        ssize_t w = ::send(fd, p, n, MSG_NOSIGNAL);
        if (w <= 0) return -1;
        p += w;
        n -= (size_t)w;
    }
    return 0;
}
static int recv_all(int fd, char *p, size_t n) {
    while (n) {
        ssize_t r = ::recv(fd, p, n, 0);
        if (r <= 0) return -1;
        p += r;
        n -= (size_t)r;
    }
    return 0;
}

static int rendezvous(int rank, int world, const char *addr, int port, char *id /*128 bytes, filled on rank 0*/) {
    if (rank == 0) {
        int ls = socket(AF_INET, SOCK_STREAM, 0);
        if (ls < 0) {
            set_error(CGPU_ECUDA, "rendezvous: socket() failed");
            return CGPU_ECUDA;
        }
        int one = 1;
        setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
        sockaddr_in sa;
        memset(&sa, 0, sizeof(sa));
        sa.sin_family = AF_INET;
        sa.sin_addr.s_addr = htonl(INADDR_ANY);
        sa.sin_port = htons((uint16_t)port);
        if (bind(ls, (sockaddr *)&sa, sizeof(sa)) != 0 || listen(ls, world) != 0) {
            close(ls);
            set_error(CGPU_ECUDA, "rendezvous: cannot listen on port %d", port);
            return CGPU_ECUDA;
        }
        for (int i = 1; i < world; i++) {
            int fd = accept(ls, nullptr, nullptr);
            if (fd < 0 || send_all(fd, id, NCCL_UNIQUE_ID_BYTES) != 0) {
                if (fd >= 0) close(fd);
                close(ls);
                set_error(CGPU_ECUDA, "rendezvous: lost a peer while sending the id");
                return CGPU_ECUDA;
            }
            close(fd);
        }
        close(ls);
        return CGPU_OK;
    }
    addrinfo hints, *res = nullptr;
    memset(&hints, 0, sizeof(hints));
    hints.ai_family = AF_INET;
    hints.ai_socktype = SOCK_STREAM;
    char ports[16];
    snprintf(ports, sizeof(ports), "%d", port);
    if (getaddrinfo(addr, ports, &hints, &res) != 0 || !res) {
        set_error(CGPU_ECUDA, "rendezvous: cannot resolve %s", addr);
        return CGPU_ECUDA;
    }
    int rc = CGPU_ECUDA;
    for (int tries = 0; tries < 1200 && rc != CGPU_OK; tries++) {  // up to ~60 s for rank 0 to come up
        int fd = socket(AF_INET, SOCK_STREAM, 0);
        if (fd >= 0 && connect(fd, res->ai_addr, res->ai_addrlen) == 0 && recv_all(fd, id, NCCL_UNIQUE_ID_BYTES) == 0) rc = CGPU_OK;
        if (fd >= 0) close(fd);
        if (rc != CGPU_OK) std::this_thread::sleep_for(std::chrono::milliseconds(50));
    }
    freeaddrinfo(res);
    if (rc != CGPU_OK) set_error(CGPU_ECUDA, "rendezvous: rank %d could not fetch the id from %s:%d", rank, addr, port);
    return rc;
}

}  // namespace cgpu

using namespace cgpu;

extern "C" {

int cgpu_comm_unique_id(char *id128) {
    if (!id128) return CGPU_EINVAL;
    NcclApi *api = nccl_api();
    if (!api) return CGPU_ENODEV;
    ncclUniqueId id;
    CGPU_NCCL(api->GetUniqueId(&id));
    memcpy(id128, id.internal, NCCL_UNIQUE_ID_BYTES);
    return CGPU_OK;
}

int cgpu_comm_init(int rank, int world, const char *id128) {
    CGPU_TRY(ensure_init());
    if (world < 1 || rank < 0 || rank >= world) {
        set_error(CGPU_EINVAL, "comm_init: rank %d of %d", rank, world);
        return CGPU_EINVAL;
    }
    if (C().comm) cgpu_comm_destroy();
    CGPU_TRY(cgpu_set_shard(rank, world));
    if (world == 1) return CGPU_OK;
    if (!id128) return CGPU_EINVAL;
    NcclApi *api = nccl_api();
    if (!api) return CGPU_ENODEV;
    ncclUniqueId id;
    memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
    ncclComm_t comm = nullptr;
    CGPU_NCCL(api->CommInitRank(&comm, world, id, rank));
    C().comm = comm;
    C().rank = rank;
    C().world = world;
    return CGPU_OK;
}

int cgpu_comm_init_env(void) {
    auto geti = [](const char *a, const char *b, int dflt) {
        const char *e = getenv(a);
        if (!e && b) e = getenv(b);
        return e ? atoi(e) : dflt;
    };
    const int world = geti("CGPU_WORLD", "WORLD_SIZE", 1), rank = geti("CGPU_RANK", "RANK", 0);
    if (world <= 1) {
        CGPU_TRY(ensure_init());
        return cgpu_set_shard(0, 1);
    }
    const char *addr = getenv("MASTER_ADDR");
    if (!addr || !*addr) addr = "127.0.0.1";
    const int port = getenv("CGPU_COMM_PORT") ? atoi(getenv("CGPU_COMM_PORT")) : geti("MASTER_PORT", nullptr, 29500) + 17;
    CGPU_TRY(ensure_init());
    char id[NCCL_UNIQUE_ID_BYTES];
    memset(id, 0, sizeof(id));
    if (rank == 0) CGPU_TRY(cgpu_comm_unique_id(id));
    CGPU_TRY(rendezvous(rank, world, addr, port, id));
    return cgpu_comm_init(rank, world, id);
}

void cgpu_comm_destroy(void) {
    Comm &c = C();
    if (c.comm) {
        if (G().inited) cudaStreamSynchronize(G().stream);
        NcclApi *api = nccl_api();
        if (api) api->CommDestroy(c.comm);
    }
    if (c.d_stage) cudaFree(c.d_stage);
    c = Comm();
}

int cgpu_comm_info(int *rank, int *world) {
    if (rank) *rank = comm_rank();
    if (world) *world = comm_world();
    return CGPU_OK;
}

int cgpu_comm_item_range(int cm, int *row0, int *nrows) {
    if (cm < 0) return CGPU_EINVAL;
    const int world = comm_world(), rank = comm_rank();
    const int per = (cm + world - 1) / world;
    const int r0 = std::min(cm, rank * per), r1 = std::min(cm, r0 + per);
    if (row0) *row0 = r0;
    if (nrows) *nrows = r1 - r0;
    return CGPU_OK;
}

int cgpu_comm_allgather_host(const void *send, void *recv, long long bytes) {
    CGPU_TRY(ensure_init());
    if (!send || !recv || bytes < 0) return CGPU_EINVAL;
    const int world = comm_world();
    if (world == 1) {
        if (send != recv) memcpy(recv, send, (size_t)bytes);
        return CGPU_OK;
    }
    (void)cudaSetDevice(G().device);
    const size_t cnt = ((size_t)bytes + 7) / 8;  // whole doubles
    void *stage = nullptr;
    CGPU_TRY(stage_get(cnt * 8 * (size_t)(world + 1), &stage));
    double *d_send = static_cast<double *>(stage), *d_recv = d_send + cnt;
    cudaStream_t st = G().stream;
    CGPU_CUDA(cudaMemcpyAsync(d_send, send, (size_t)bytes, cudaMemcpyHostToDevice, st));
    CGPU_TRY(comm_allgather_f64(d_send, d_recv, cnt, st));
    CGPU_CUDA(cudaMemcpy2DAsync(recv, (size_t)bytes, d_recv, cnt * 8, (size_t)bytes, (size_t)world, cudaMemcpyDeviceToHost, st));
    CGPU_CUDA(cudaStreamSynchronize(st));
    return CGPU_OK;
}

int cgpu_comm_allreduce_host(double *buf, long long count, int op) {
    CGPU_TRY(ensure_init());
    if (!buf || count < 0 || op < 0 || op > 2) return CGPU_EINVAL;
    if (comm_world() == 1 || count == 0) return CGPU_OK;
    (void)cudaSetDevice(G().device);
    void *stage = nullptr;
    CGPU_TRY(stage_get((size_t)count * 8, &stage));
    cudaStream_t st = G().stream;
    CGPU_CUDA(cudaMemcpyAsync(stage, buf, (size_t)count * 8, cudaMemcpyHostToDevice, st));
    CGPU_TRY(comm_allreduce_f64(static_cast<double *>(stage), static_cast<double *>(stage), (size_t)count, op, st));
    CGPU_CUDA(cudaMemcpyAsync(buf, stage, (size_t)count * 8, cudaMemcpyDeviceToHost, st));
    CGPU_CUDA(cudaStreamSynchronize(st));
    return CGPU_OK;
}

int cgpu_comm_barrier(void) {
    double x = 0.0;
    return cgpu_comm_allreduce_host(&x, 1, 0);
}

}  // extern "C"
