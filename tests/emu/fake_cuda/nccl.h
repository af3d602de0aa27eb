// fake_cuda/nccl.h — the handful of NCCL declarations sharded.cu needs, for the CPU build of the library
// (tests/emu; TEST INFRASTRUCTURE ONLY).  The functions themselves live in tests/emu/fake_nccl.cpp (libfake_nccl.so,
// found by sharded.cu's dlopen through COZO_GPU_NCCL_LIB): ranks are THREADS of one process, collectives are copies
// between their buffers around a barrier.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#define NCCL_UNIQUE_ID_BYTES 128
typedef struct {
  char internal[NCCL_UNIQUE_ID_BYTES];
} ncclUniqueId;
typedef struct ncclComm* ncclComm_t;
typedef enum { ncclSuccess = 0, ncclUnhandledCudaError = 1, ncclSystemError = 2, ncclInternalError = 3, ncclInvalidArgument = 4 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5, ncclFloat16 = 6, ncclFloat32 = 7, ncclFloat64 = 8 } ncclDataType_t;
