// TEST INFRASTRUCTURE ONLY: single-rank stand-in for the NCCL calls of palace_b200/csrc (host emulation build).
#pragma once
#include <cstring>

#include "cuda_runtime.h"
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;
enum
{
  ncclSuccess = 0,
  ncclInvalidUsage = 5
};
enum ncclDataType_t
{
  ncclDouble = 8
};
enum ncclRedOp_t
{
  ncclSum = 0
};
struct ncclUniqueId
{
  char internal[128];
};
inline const char *ncclGetErrorString(ncclResult_t) { return "NCCL is not available in the host emulation"; }
inline ncclResult_t ncclGetUniqueId(ncclUniqueId *id)
{
  std::memset(id, 0, sizeof(*id));
  return ncclSuccess;
}
inline ncclResult_t ncclCommInitRank(ncclComm_t *, int, ncclUniqueId, int) { return ncclInvalidUsage; }
inline ncclResult_t ncclCommDestroy(ncclComm_t) { return ncclSuccess; }
inline ncclResult_t ncclGroupStart() { return ncclSuccess; }
inline ncclResult_t ncclGroupEnd() { return ncclSuccess; }
inline ncclResult_t ncclSend(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) { return ncclInvalidUsage; }
inline ncclResult_t ncclRecv(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) { return ncclInvalidUsage; }
inline ncclResult_t ncclAllReduce(const void *s, void *r, size_t n, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t)
{
  if (s != r) std::memmove(r, s, n * sizeof(double));
  return ncclSuccess;
}
