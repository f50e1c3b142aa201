// peer_store_bench.cu -- development micro-benchmark: NVLink write bandwidth of SM-issued stores into a peer GPU's memory
// as a function of the contiguous run length per row (the fused exchange stores runs of P*16 bytes per output row).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peer_store_bench peer_store_bench.cu ; needs >= 2 GPUs with P2P
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// every thread stores E double2 values; a group of W lanes covers one run of W*16 bytes; consecutive runs of a warp
// (and the E stores of a thread) are `stride` elements apart, like the rows of a transposed tile
template <int W, int E>
__global__ void __launch_bounds__(512) run_store(double2 *__restrict__ dst, long long stride, long long nrows, long long row_len) {
  const int lane_in_run = threadIdx.x % W, run_in_cta = threadIdx.x / W;
  const int runs_per_cta = blockDim.x / W;
  // tile = (column block cb of W elements, row block rb of runs_per_cta*E rows)
  const long long cbs = row_len / W, rbs = nrows / (runs_per_cta * E);
  const double2 val = make_double2(1.0, 2.0);
  for (long long t = blockIdx.x; t < cbs * rbs; t += gridDim.x) {
    const long long cb = t % cbs, rb = t / cbs;
    double2 *p = dst + (rb * runs_per_cta * E + run_in_cta) * stride + cb * W + lane_in_run;
#pragma unroll
    for (int m = 0; m < E; m++) p[(long long)m * runs_per_cta * stride] = val;
  }
}
__global__ void copy16(const double2 *__restrict__ in, double2 *__restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = in[i];
}

int main() {
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
  const long long row_len = 1024, nrows = 1024 * 128;  // 2 GiB of double2
  const size_t bytes = (size_t)row_len * nrows * 16;
  double2 *local, *remote, *src;
  CK(cudaSetDevice(1));
  CK(cudaMalloc(&remote, bytes));
  CK(cudaSetDevice(0));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  CK(cudaMalloc(&local, bytes));
  CK(cudaMalloc(&src, bytes));
  CK(cudaMemset(src, 0, bytes));
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  auto time = [&](auto f) {
    f(); CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int i = 0; i < 3; i++) f();
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / 3;
  };
  for (int pass = 0; pass < 2; pass++) {
    double2 *dst = pass ? remote : local;
    printf("---- destination: %s\n", pass ? "PEER over NVLink" : "local HBM");
    for (int ctas : {148, 74, 296}) {
      float ms;
      ms = time([&] { run_store<8, 16><<<ctas, 512>>>(dst, row_len, nrows, row_len); });
      printf("runs 128B  ctas %3d  %.3f ms  %.0f GB/s\n", ctas, ms, bytes / ms / 1e6);
      ms = time([&] { run_store<16, 16><<<ctas, 512>>>(dst, row_len, nrows, row_len); });
      printf("runs 256B  ctas %3d  %.3f ms  %.0f GB/s\n", ctas, ms, bytes / ms / 1e6);
      ms = time([&] { run_store<32, 16><<<ctas, 512>>>(dst, row_len, nrows, row_len); });
      printf("runs 512B  ctas %3d  %.3f ms  %.0f GB/s\n", ctas, ms, bytes / ms / 1e6);
      ms = time([&] { run_store<4, 16><<<ctas, 512>>>(dst, row_len, nrows, row_len); });
      printf("runs  64B  ctas %3d  %.3f ms  %.0f GB/s\n", ctas, ms, bytes / ms / 1e6);
      ms = time([&] { copy16<<<ctas * 4, 512>>>(src, dst, bytes / 16); });
      printf("copy16     ctas %3d  %.3f ms  %.0f GB/s (written)\n", ctas * 4, ms, bytes / ms / 1e6);
    }
    float ms = time([&] { CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, 0)); });
    printf("cudaMemcpyAsync (copy engine)  %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
  }
  return 0;
}
