// a2a_bench.cu -- development micro-benchmark (round 2): ceiling of the ALL-TO-ALL store pattern of the fused exchange stage
// over NVSwitch.  One process drives all GPUs; every GPU runs one kernel that scatters tiles of 1024 rows x 128 bytes to all
// GPUs (rows [q*1024/N, (q+1)*1024/N) of a tile go to GPU q, the own share to local memory), all GPUs at once -- the store
// pattern of pow2_pipe_kernel<double,1024,C2C,8,TS=1> with one segment per peer, without the FFT.  Also: the same volume by
// the copy engines (one cudaMemcpyAsync per peer block), and stores to a single peer for reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o a2a_bench a2a_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct Dst { double2 *p[8]; };

// tile t = (column block cb of 8 elements, slab sb of 1024 rows); thread: lane = tid % 8, tB = tid / 8 (0..63), 16 stores to rows
// k = tB + 64 m; row k of slab sb lands on GPU q = k / rpp at row sb * rpp + k % rpp of its [rows][row_len] array
__global__ void __launch_bounds__(512) a2a_store(Dst d, int n, long long row_len, long long nslabs, int only_peer) {
  const int lane = threadIdx.x % 8, tB = threadIdx.x / 8;
  const int rpp = 1024 / n;
  const long long cbs = row_len / 8;
  const double2 val = make_double2(1.0, 2.0);
  for (long long t = blockIdx.x; t < cbs * nslabs; t += gridDim.x) {
    const long long cb = t % cbs, sb = t / cbs;
#pragma unroll
    for (int m = 0; m < 16; m++) {
      const int k = tB + 64 * m;
      const int q = only_peer >= 0 ? only_peer : k / rpp;
      const long long row = only_peer >= 0 ? sb * 1024 + k : sb * rpp + k % rpp;
      d.p[q][row * row_len + cb * 8 + lane] = val;
    }
  }
}

int main(int argc, char **argv) {
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  if (argc > 1) n = atoi(argv[1]) < n ? atoi(argv[1]) : n;
  if (n < 2) { printf("needs >= 2 GPUs\n"); return 0; }
  if (n > 8) n = 8;
  while (1024 % n) n--;
  const long long row_len = 1024, nslabs = 64;  // per GPU: 64 slabs x 1024 rows x 1024 cols x 16 B = 1 GiB sent (own share included)
  const size_t bytes = (size_t)row_len * nslabs * 1024 * 16;
  std::vector<double2 *> recv(n), src(n);
  std::vector<cudaStream_t> st(n);
  std::vector<std::vector<cudaStream_t>> cst(n);
  std::vector<cudaEvent_t> a(n), b(n);
  for (int d = 0; d < n; d++) {
    CK(cudaSetDevice(d));
    for (int q = 0; q < n; q++)
      if (q != d) CK(cudaDeviceEnablePeerAccess(q, 0));
    CK(cudaMalloc(&recv[d], bytes));
    CK(cudaMalloc(&src[d], bytes));
    CK(cudaMemset(src[d], 0, bytes));
    CK(cudaStreamCreateWithFlags(&st[d], cudaStreamNonBlocking));
    cst[d].resize(n);
    for (int q = 0; q < n; q++) CK(cudaStreamCreateWithFlags(&cst[d][q], cudaStreamNonBlocking));
    CK(cudaEventCreate(&a[d]));
    CK(cudaEventCreate(&b[d]));
  }
  auto sync_all = [&] { for (int d = 0; d < n; d++) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); } };
  auto report = [&](const char *what, double remote_frac) {
    float worst = 0, best = 1e9;
    for (int d = 0; d < n; d++) {
      CK(cudaSetDevice(d));
      CK(cudaEventSynchronize(b[d]));
      float ms;
      CK(cudaEventElapsedTime(&ms, a[d], b[d]));
      worst = ms > worst ? ms : worst;
      best = ms < best ? ms : best;
    }
    printf("%-58s %d GPUs  %.3f..%.3f ms  %.0f GB/s per GPU over NVLink (slowest)\n", what, n, best, worst, bytes * remote_frac / worst / 1e6);
  };
  Dst dd;
  for (int q = 0; q < 8; q++) dd.p[q] = q < n ? recv[q] : nullptr;
  for (int ctas : {148, 92, 74}) {
    for (int rep = 0; rep < 3; rep++) {
      sync_all();
      for (int d = 0; d < n; d++) {
        CK(cudaSetDevice(d));
        CK(cudaEventRecord(a[d], st[d]));
        a2a_store<<<ctas, 512, 0, st[d]>>>(dd, n, row_len, nslabs, -1);
        CK(cudaEventRecord(b[d], st[d]));
      }
    }
    char nm[100];
    snprintf(nm, sizeof nm, "all-to-all SM stores, 128 B runs, %d CTAs per GPU", ctas);
    report(nm, (double)(n - 1) / n);
  }
  // every GPU stores everything into its right neighbour only
  for (int rep = 0; rep < 3; rep++) {
    sync_all();
    for (int d = 0; d < n; d++) {
      CK(cudaSetDevice(d));
      CK(cudaEventRecord(a[d], st[d]));
      a2a_store<<<148, 512, 0, st[d]>>>(dd, n, row_len, nslabs, (d + 1) % n);
      CK(cudaEventRecord(b[d], st[d]));
    }
  }
  report("ring: every GPU stores to its right neighbour only", 1.0);
  // copy engines: one copy per peer block (bytes / n each), all GPUs at once, every copy on its own stream
  for (int pieces : {1, 4, 8}) {
    for (int rep = 0; rep < 3; rep++) {
      sync_all();
      for (int d = 0; d < n; d++) {
        CK(cudaSetDevice(d));
        CK(cudaEventRecord(a[d], st[d]));
        const size_t blk = bytes / n, pc = blk / pieces;
        for (int q = 0; q < n; q++) {
          if (q == d) continue;
          CK(cudaStreamWaitEvent(cst[d][q], a[d], 0));
          for (int i = 0; i < pieces; i++)
            CK(cudaMemcpyAsync((char *)recv[q] + d * blk + i * pc, (char *)src[d] + q * blk + i * pc, pc, cudaMemcpyDeviceToDevice, cst[d][q]));
          CK(cudaEventRecord(b[d], cst[d][q]));  // (re-recorded: the last one wins; joined below)
          CK(cudaStreamWaitEvent(st[d], b[d], 0));
        }
        CK(cudaEventRecord(b[d], st[d]));
      }
    }
    char nm[100];
    snprintf(nm, sizeof nm, "all-to-all copy engines, %d copies of %.1f MB per peer", pieces, bytes / n / pieces / 1e6);
    report(nm, (double)(n - 1) / n);
  }
  return 0;
}
