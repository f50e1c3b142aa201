// stage_bench.cu -- development micro-benchmark (not part of the library): times stage-kernel variants on the three
// access patterns of the 1024^3 double R2C transform, plus plain copies that calibrate what the pattern itself allows.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I include -I p3dfft.3_b200/csrc [-DP3B_SKELETON] ...
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "common.cuh"
#include "generic_stage.cuh"
#include "pow2_stage.cuh"
#include "pow2_pipe.cuh"
using namespace p3b;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// ---- calibration kernels
__global__ void copy16(const double2 *__restrict__ in, double2 *__restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = in[i];
}
// tile copy with the y-stage pattern and no shared memory: each thread moves E elements of column u from rows t+m*TP;
// loads are W-wide runs (W*16 bytes) with row stride `rs`, stores keep the same pattern (transposed on both sides)
template <int W, int ROWS, int E>
__global__ void __launch_bounds__(W * ROWS / E) tile_copy(const double2 *__restrict__ in, double2 *__restrict__ out, long long rs, long long nu,
                                                          long long nv, long long vs) {
  constexpr int TP = ROWS / E;
  const int p = threadIdx.x % W, t = threadIdx.x / W;
  const long long tiles_u = nu / W, ntiles = tiles_u * nv;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long u0 = (tile % tiles_u) * W, v = tile / tiles_u;
    const double2 *src = in + v * vs + u0 + p;
    double2 *dst = out + v * vs + u0 + p;
    double2 r[E];
#pragma unroll
    for (int m = 0; m < E; m++) r[m] = src[(long long)(t + m * TP) * rs];
#pragma unroll
    for (int m = 0; m < E; m++) dst[(long long)(t + m * TP) * rs] = r[m];
  }
}

struct Timer {
  cudaEvent_t a, b;
  Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
  template <class F> float run(F f, int reps = 5) {
    f();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int i = 0; i < reps; i++) f();
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
  }
};

template <typename T> const void *twiddles(int n) {
  std::vector<T> h(2 * (size_t)n);
  for (int j = 0; j < n; j++) { h[2 * j] = (T)cos(-2.0 * M_PI * j / n); h[2 * j + 1] = (T)sin(-2.0 * M_PI * j / n); }
  void *d;
  CK(cudaMalloc(&d, h.size() * sizeof(T)));
  CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

int ilog2(int x) { int l = 0; while ((1 << l) < x) l++; return l; }

int main(int argc, char **argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 1024;
  const long long NX = N, NY = N, NZ = N, NXC = N / 2 + 1;
  const size_t cbytes = (size_t)NXC * NY * NZ * 16;
  void *A, *B;
  CK(cudaMalloc(&A, cbytes));
  CK(cudaMalloc(&B, cbytes));
  CK(cudaMemset(A, 0, cbytes));
  CK(cudaMemset(B, 0, cbytes));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  Timer tm;
  const double gb_c = 2.0 * cbytes / 1e9;
#ifdef P3B_SKELETON
  printf("== SKELETON build (no butterflies)\n");
#endif
  {
    long long n = cbytes / 16;
    float ms = tm.run([&] { copy16<<<sms * 8, 512>>>((const double2 *)A, (double2 *)B, n); });
    printf("copy16 contiguous                      %.3f ms  %.0f GB/s\n", ms, gb_c / ms * 1e3);
  }
  // y-stage pattern copies: rows of NXC complex, 1024 rows (y), NZ planes
  {
    float ms = tm.run([&] { tile_copy<8, 1024, 16><<<sms, 512>>>((const double2 *)A, (double2 *)B, NXC, NXC - 1, NZ, NXC * NY); });
    printf("tile_copy W=8  (128B runs) 512thr x1    %.3f ms  %.0f GB/s\n", ms, gb_c / ms * 1e3);
    ms = tm.run([&] { tile_copy<8, 1024, 16><<<sms * 2, 512>>>((const double2 *)A, (double2 *)B, NXC, NXC - 1, NZ, NXC * NY); });
    printf("tile_copy W=8  (128B runs) 512thr x2    %.3f ms  %.0f GB/s\n", ms, gb_c / ms * 1e3);
    ms = tm.run([&] { tile_copy<4, 1024, 16><<<sms * 4, 256>>>((const double2 *)A, (double2 *)B, NXC, NXC - 1, NZ, NXC * NY); });
    printf("tile_copy W=4  (64B runs)  256thr x4    %.3f ms  %.0f GB/s\n", ms, gb_c / ms * 1e3);
    ms = tm.run([&] { tile_copy<16, 1024, 16><<<sms, 1024>>>((const double2 *)A, (double2 *)B, NXC, NXC - 1, NZ, NXC * NY); });
    printf("tile_copy W=16 (256B runs) 1024thr x1   %.3f ms  %.0f GB/s\n", ms, gb_c / ms * 1e3);
    ms = tm.run([&] { tile_copy<2, 1024, 16><<<sms * 4, 128>>>((const double2 *)A, (double2 *)B, NXC, NXC - 1, NZ, NXC * NY); });
    printf("tile_copy W=2  (32B runs)  128thr x4    %.3f ms  %.0f GB/s\n", ms, gb_c / ms * 1e3);
    ms = tm.run([&] { tile_copy<2, 1024, 16><<<sms * 8, 128>>>((const double2 *)A, (double2 *)B, NXC, NXC - 1, NZ, NXC * NY); });
    printf("tile_copy W=2  (32B runs)  128thr x8    %.3f ms  %.0f GB/s\n", ms, gb_c / ms * 1e3);
    ms = tm.run([&] { tile_copy<1, 1024, 16><<<sms * 8, 64>>>((const double2 *)A, (double2 *)B, NXC, NXC - 1, NZ, NXC * NY); });
    printf("tile_copy W=1  (16B runs)  64thr x8     %.3f ms  %.0f GB/s\n", ms, gb_c / ms * 1e3);
    ms = tm.run([&] { tile_copy<1, 1024, 16><<<sms * 16, 64>>>((const double2 *)A, (double2 *)B, NXC, NXC - 1, NZ, NXC * NY); });
    printf("tile_copy W=1  (16B runs)  64thr x16    %.3f ms  %.0f GB/s\n", ms, gb_c / ms * 1e3);
    ms = tm.run([&] { tile_copy<4, 1024, 16><<<sms * 2, 256>>>((const double2 *)A, (double2 *)B, NXC, NXC - 1, NZ, NXC * NY); });
    printf("tile_copy W=4  (64B runs)  256thr x2    %.3f ms  %.0f GB/s\n", ms, gb_c / ms * 1e3);
    ms = tm.run([&] { tile_copy<4, 1024, 16><<<sms * 8, 256>>>((const double2 *)A, (double2 *)B, NXC, NXC - 1, NZ, NXC * NY); });
    printf("tile_copy W=4  (64B runs)  256thr x8    %.3f ms  %.0f GB/s\n", ms, gb_c / ms * 1e3);
    ms = tm.run([&] { tile_copy<8, 1024, 16><<<sms * 4, 512>>>((const double2 *)A, (double2 *)B, NXC, NXC - 1, NZ, NXC * NY); });
    printf("tile_copy W=8  (128B runs) 512thr x4    %.3f ms  %.0f GB/s\n", ms, gb_c / ms * 1e3);
    ms = tm.run([&] { tile_copy<8, 256, 16><<<sms * 16, 128>>>((const double2 *)A, (double2 *)B, NXC, NXC - 1, NZ * 4, NXC * NY / 4); });
    printf("tile_copy W=8 256 rows 128thr x16       %.3f ms  %.0f GB/s\n", ms, gb_c / ms * 1e3);
    ms = tm.run([&] { tile_copy<8, 1024, 8><<<sms * 2, 1024>>>((const double2 *)A, (double2 *)B, NXC, NXC - 1, NZ, NXC * NY); });
    printf("tile_copy W=8 E=8 1024thr x2            %.3f ms  %.0f GB/s\n", ms, gb_c / ms * 1e3);
  }
  // ---- stage kernels: y-stage forward (load U runs, store D contiguous): in mo 012 {NXC,NY,NZ} -> out mo 102 {NY,NXC,NZ}
  StageParams P;
  memset(&P, 0, sizeof P);
  P.in = A;
  P.kind = P3DFFTCU_K_C2C_FWD; P.dt_in = 2; P.dt_out = 2;
  P.nfft = P.n_in = P.n_out = N; P.L = N;
  P.nu = NXC; P.nv = NZ; P.is_d = NXC; P.is_u = 1; P.is_v = NXC * NY;
  P.tw = twiddles<double>(N);
  P.nseg = 1;
  P.seg[0].base = B; P.seg[0].k0 = 0; P.seg[0].k1 = N; P.seg[0].off = 0;
  P.seg[0].os_d = 1; P.seg[0].os_u = NY; P.seg[0].os_v = NY * NXC;
  auto set_tiles = [&](int tu, int tv, int lo, int so) {
    P.tile_u = tu; P.tile_v = tv; P.tu_log2 = ilog2(tu); P.load_ord = lo; P.store_ord = so;
    P.tiles_u = (P.nu + tu - 1) / tu; P.tiles_v = (P.nv + tv - 1) / tv; P.ntiles = P.tiles_u * P.tiles_v;
    P.vfast = P.seg[0].os_v < P.seg[0].os_u;
  };
  auto run_old = [&](auto kern, int threads, size_t smem, int ctas, const char *name) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    float ms = tm.run([&] { kern<<<sms * ctas, threads, smem>>>(P); });
    CK(cudaGetLastError());
    printf("%-38s %.3f ms  %.0f GB/s\n", name, ms, gb_c / ms * 1e3);
  };
  auto run_pipe = [&](const PipeInfo *info, int ctas, const char *name) {
    if (!info) { printf("%-38s (not instantiated)\n", name); return; }
    CK(cudaFuncSetAttribute(info->func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)info->smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, info->func, info->threads, info->smem));
    if (ctas <= 0 || ctas > occ) ctas = occ;
    float ms = tm.run([&] { info->launch(P, sms * ctas, 0); });
    CK(cudaGetLastError());
    printf("%-38s %.3f ms  %.0f GB/s  (occ %d, used %d, smem %zu)\n", name, ms, gb_c / ms * 1e3, occ, ctas, info->smem);
  };
  if (N == 1024) {
    constexpr int M = 1024;
    size_t pen = (size_t)Pow2Smem<M>::PENCIL * 16;
    // (1) contiguous both sides: in {NY,NXC,NZ} pencils along y unit-stride, same layout out
    P.is_d = 1; P.is_u = NY; P.is_v = NY * NXC;
    P.seg[0].os_d = 1; P.seg[0].os_u = NY; P.seg[0].os_v = NY * NXC;
    P.in = A; P.seg[0].base = B;
    set_tiles(4, 1, ORD_D, ORD_D);
    run_old(pow2_stage_kernel<double, M, 256, 2>, 256, 4 * pen, 2, "contig old P=4 256thr x2");
    for (int p : {1, 2, 4, 8}) {
      char nm[64]; snprintf(nm, sizeof nm, "contig pipe P=%d", p);
      set_tiles(p, 1, ORD_D, ORD_D);
      run_pipe(pipe_info<double, 1, 0>(M, p), 0, nm);
    }
    // (2) contiguous loads, transposed stores along u: out {NXC(u) fastest, NY(d), NZ}
    P.seg[0].os_d = NXC - 1; P.seg[0].os_u = 1; P.seg[0].os_v = (NXC - 1) * NY;  // u extent 512: aligned 128-byte runs
    P.nu = NXC - 1;
    for (int p : {2, 4, 8}) {
      char nm[64]; snprintf(nm, sizeof nm, "transposed-u pipe P=%d", p);
      set_tiles(p, 1, ORD_D, ORD_U);
      run_pipe(pipe_info<double, 1, 1>(M, p), 0, nm);
    }
    set_tiles(8, 1, ORD_D, ORD_U);
    run_old(pow2_stage_kernel<double, M, 512, 1>, 512, 8 * pen, 1, "transposed-u old P=8 512thr");
    // (3) as the planner's y stage: in {NY(d) fastest, NXC(u), NZ(v)} -> out {NZ(v) fastest, NXC(u), NY(d)}
    P.nu = NXC; P.nv = NZ;
    P.seg[0].os_d = NZ * NXC; P.seg[0].os_u = NZ; P.seg[0].os_v = 1;
    for (int p : {4, 8}) {
      char nm[64]; snprintf(nm, sizeof nm, "transposed-v pipe P=%d", p);
      set_tiles(1, p, ORD_D, ORD_V);
      run_pipe(pipe_info<double, 1, 1>(M, p), 0, nm);
    }
  }
  return 0;
}
