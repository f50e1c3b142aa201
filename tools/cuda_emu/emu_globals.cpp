// definitions of the emulated CUDA built-ins (see cuda_runtime.h in this directory)
#include "cuda_runtime.h"
thread_local emu_uint3 threadIdx;
emu_uint3 blockIdx, blockDim, gridDim;
pthread_barrier_t emu_block_barrier;
namespace p3b { alignas(16) unsigned char smem_raw[232448]; }
