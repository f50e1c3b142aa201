// cuda_runtime.h (EMULATION SHIM) -- development/test tool, never part of the shipped library.
// Lets the stage kernels in p3dfft.3_b200/csrc/*.cuh be compiled by g++ and executed on the CPU with one
// OS thread per CUDA thread, so kernel index arithmetic can be validated in the GPU-less authoring
// container before spending time on a real B200.  Used only by tools/cuda_emu/Makefile (libp3dfft_emu.so)
// and by the `not gpu` tests that load that library explicitly.
#pragma once
#define P3B_EMU 1
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <algorithm>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __restrict__
#define __shared__
#define __align__(x)

struct float2 { float x, y; };
struct double2 { double x, y; };
struct emu_uint3 { unsigned x, y, z; };
extern thread_local emu_uint3 threadIdx;
extern emu_uint3 blockIdx, blockDim, gridDim;
extern pthread_barrier_t emu_block_barrier;
inline void __syncthreads() { pthread_barrier_wait(&emu_block_barrier); }
inline void __syncwarp() { pthread_barrier_wait(&emu_block_barrier); }  // callers are uniform across the CTA
template <class T> inline T __ldg(const T *p) { return *p; }
inline int __ffs(int x) { return __builtin_ffs(x); }
using std::max;
using std::min;

typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
template <class F> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) {
  *n = 1;
  return cudaSuccess;
}

namespace p3b { extern unsigned char smem_raw[]; }

// run `grid` blocks one after another, each with `block` OS threads
template <class Kernel, class... Params> void emu_launch(Kernel k, int grid, int block, size_t smem, const Params &...P) {
  if (smem > 232448) { fprintf(stderr, "emu: shared memory request too large\n"); abort(); }
  gridDim.x = grid; gridDim.y = gridDim.z = 1;
  blockDim.x = block; blockDim.y = blockDim.z = 1;
  for (int b = 0; b < grid; b++) {
    blockIdx.x = b; blockIdx.y = blockIdx.z = 0;
    pthread_barrier_init(&emu_block_barrier, nullptr, block);
    std::vector<std::thread> th;
    th.reserve(block);
    for (int t = 0; t < block; t++)
      th.emplace_back([&, t]() {
        threadIdx.x = t; threadIdx.y = threadIdx.z = 0;
        k(P...);
      });
    for (auto &x : th) x.join();
    pthread_barrier_destroy(&emu_block_barrier);
  }
}
#define P3B_LAUNCH(kernel, grid, block, smem, stream, params) emu_launch(kernel, (grid) < 2 ? (grid) : 2, block, smem, params)
#define P3B_LAUNCH2(kernel, grid, block, smem, stream, p1, p2) emu_launch(kernel, (grid) < 2 ? (grid) : 2, block, smem, p1, p2)

// ---- host runtime stubs: "device" memory is plain host memory
struct cudaDeviceProp { char name[64]; int major, minor, multiProcessorCount; size_t sharedMemPerBlockOptin; };
struct cudaPointerAttributes { int type; };
enum { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
  strcpy(p->name, "EMULATED (CPU threads)"); p->major = 10; p->minor = 0; p->multiProcessorCount = 2;
  p->sharedMemPerBlockOptin = 232448; return cudaSuccess;
}
// "device" allocations live in POSIX shared memory so that the CUDA-IPC calls below can map a peer
// process's buffer: multi-rank exchanges (peer stores + flag barrier) run on the CPU exactly as written
inline cudaError_t cudaMemGetInfo(size_t *fr, size_t *tot) { *fr = *tot = 0; return cudaSuccess; }
cudaError_t cudaMalloc(void **p, size_t n);
cudaError_t cudaFree(void *p);
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
enum { cudaStreamNonBlocking = 1 };
inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -1; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = malloc(8); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }  // launches run in issue order
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *) { a->type = cudaMemoryTypeUnregistered; return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = malloc(8); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return cudaSuccess; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *);
cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, int);
cudaError_t cudaIpcCloseMemHandle(void *);
