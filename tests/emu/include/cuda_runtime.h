// TEST INFRASTRUCTURE ONLY: the sliver of the CUDA runtime API that palace_b200/csrc uses, on host memory.
// "Device memory" is malloc'd host memory, streams are synchronous, graphs and IPC are unsupported (the
// library's eager fallbacks run instead).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../cuda_emu.hpp"

typedef int cudaError_t;
enum
{
  cudaSuccess = 0,
  cudaErrorNotSupported = 801,
  cudaErrorInvalidValue = 1
};
typedef struct CUstream_emu *cudaStream_t;
typedef struct CUevent_emu *cudaEvent_t;
typedef struct CUgraph_emu *cudaGraph_t;
typedef struct CUgraphExec_emu *cudaGraphExec_t;
enum cudaMemcpyKind
{
  cudaMemcpyHostToHost,
  cudaMemcpyHostToDevice,
  cudaMemcpyDeviceToHost,
  cudaMemcpyDeviceToDevice,
  cudaMemcpyDefault
};
enum cudaFuncAttribute
{
  cudaFuncAttributeMaxDynamicSharedMemorySize = 8,
  cudaFuncAttributePreferredSharedMemoryCarveout = 9
};
enum
{
  cudaSharedmemCarveoutMaxShared = 100,
  cudaStreamNonBlocking = 1,
  cudaEventDisableTiming = 2,
  cudaStreamCaptureModeThreadLocal = 1,
  cudaIpcMemLazyEnablePeerAccess = 1
};
struct cudaDeviceProp
{
  char name[256];
  int major, minor, multiProcessorCount;
  size_t sharedMemPerBlockOptin;
};
struct cudaIpcMemHandle_t
{
  char reserved[64];
};

inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "not supported by the host emulation"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n)
{
  *n = 1;
  return cudaSuccess;
}
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
  std::memset(p, 0, sizeof(*p));
  std::snprintf(p->name, sizeof(p->name), "host SIMT emulation");
  p->major = 10;
  p->minor = 0;
  const char *e = std::getenv("B2P_EMU_SMS");
  p->multiProcessorCount = e ? std::atoi(e) : 2;
  p->sharedMemPerBlockOptin = 227 * 1024;
  return cudaSuccess;
}
inline cudaError_t cudaMalloc(void **p, size_t n)
{
  // 256-byte alignment like cudaMalloc; poisoned so reads of never-written device memory show up
  if (posix_memalign(p, 256, n ? ((n + 255) & ~(size_t)255) : 256)) return cudaErrorInvalidValue;
  std::memset(*p, 0xCD, n);
  return cudaSuccess;
}
template <typename T>
inline cudaError_t cudaMalloc(T **p, size_t n)
{
  return cudaMalloc((void **)p, n);
}
inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
template <typename T>
inline cudaError_t cudaMallocHost(T **p, size_t n)
{
  return cudaMalloc((void **)p, n);
}
inline cudaError_t cudaFree(void *p)
{
  std::free(p);
  return cudaSuccess;
}
inline cudaError_t cudaFreeHost(void *p)
{
  std::free(p);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind)
{
  std::memmove(d, s, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr)
{
  std::memmove(d, s, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemset(void *d, int v, size_t n)
{
  std::memset(d, v, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr)
{
  std::memset(d, v, n);
  return cudaSuccess;
}
template <typename F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int)
{
  return cudaSuccess;
}
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi)
{
  *lo = *hi = 0;
  return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int)
{
  *s = nullptr;
  return cudaSuccess;
}
inline cudaError_t cudaStreamCreate(cudaStream_t *s)
{
  *s = nullptr;
  return cudaSuccess;
}
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned)
{
  *e = nullptr;
  return cudaSuccess;
}
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { return cudaEventCreateWithFlags(e, 0); }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t)
{
  *ms = 0.f;
  return cudaSuccess;
}
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
enum cudaStreamCaptureStatus
{
  cudaStreamCaptureStatusNone = 0,
  cudaStreamCaptureStatusActive = 1
};
inline cudaError_t cudaStreamIsCapturing(cudaStream_t, cudaStreamCaptureStatus *st)
{
  *st = cudaStreamCaptureStatusNone;
  return cudaSuccess;
}
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, int) { return cudaErrorNotSupported; }
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t *) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t *, cudaGraph_t, unsigned long long = 0) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return cudaErrorNotSupported; }
inline cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
