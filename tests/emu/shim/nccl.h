// Types and enums of <nccl.h> that consel_b200/csrc/comm.cu uses (it binds every function with dlsym), for the
// emulated build.  The functions themselves come from tests/emu/fake_nccl.cpp (built as libnccl.so.2 next to
// the emulated library): collectives between PROCESSES over POSIX shared memory on host pointers.
#pragma once
#include <stddef.h>
#define NCCL_UNIQUE_ID_BYTES 128
typedef struct {
    char internal[NCCL_UNIQUE_ID_BYTES];
} ncclUniqueId;
typedef struct ncclComm *ncclComm_t;
typedef enum { ncclSuccess = 0, ncclUnhandledCudaError = 1, ncclSystemError = 2, ncclInternalError = 3, ncclInvalidArgument = 4 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclFloat64 = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
