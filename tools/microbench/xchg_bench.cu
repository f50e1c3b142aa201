// xchg_bench.cu -- development micro-benchmark (round 2): which store path into a peer GPU's memory gets closest to the
// NVLink limit, under the conditions of the fused exchange stage (short runs across pencils, traffic in both directions,
// a local HBM-bound kernel running next to it).
//   variants: SM st.global (128 B runs) | TMA bulk store smem->peer (cp.async.bulk.global.shared::cta) by run length |
//             TMA tensor store (2-D box, 128 B inner) | copy engine, whole and in chunks | SM + copy engine together |
//             both directions at once | SM stores + a local streaming kernel on the other SMs
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xchg_bench xchg_bench.cu ; needs >= 2 GPUs with P2P
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int W, int E>
__global__ void __launch_bounds__(512) run_store(double2 *__restrict__ dst, long long stride, long long nrows, long long row_len) {
  const int lane_in_run = threadIdx.x % W, run_in_cta = threadIdx.x / W;
  const int runs_per_cta = blockDim.x / W;
  const long long cbs = row_len / W, rbs = nrows / (runs_per_cta * E);
  const double2 val = make_double2(1.0, 2.0);
  for (long long t = blockIdx.x; t < cbs * rbs; t += gridDim.x) {
    const long long cb = t % cbs, rb = t / cbs;
    double2 *p = dst + (rb * runs_per_cta * E + run_in_cta) * stride + cb * W + lane_in_run;
#pragma unroll
    for (int m = 0; m < E; m++) p[(long long)m * runs_per_cta * stride] = val;
  }
}

// local streaming copy (stands for the local FFT stage of an overlapped pair)
__global__ void __launch_bounds__(512) copy16(const double2 *__restrict__ in, double2 *__restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = in[i];
}

// TMA bulk stores: a CTA owns a 16 KiB shared-memory tile per stage (STAGES of them); each of its warps' lane 0 issues
// runs of RUN bytes to rows `stride` elements apart (the layout of a transposed tile: tile = RUN/16 columns x 16384/RUN rows)
template <int RUN>
__global__ void __launch_bounds__(256) bulk_store(double2 *__restrict__ dst, long long stride, long long nrows, long long row_len) {
  extern __shared__ __align__(128) unsigned char sm[];
  constexpr int TILE = 16384, STAGES = 4, ROWS = TILE / RUN, W = RUN / 16;
  for (int i = threadIdx.x; i < TILE * STAGES / 16; i += blockDim.x) ((double2 *)sm)[i] = make_double2(1.0, 2.0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nwarps = blockDim.x / 32;
  const long long cbs = row_len / W, rbs = nrows / ROWS;
  int stage = 0;
  for (long long t = blockIdx.x; t < cbs * rbs; t += gridDim.x) {
    const long long cb = t % cbs, rb = t / cbs;
    if (lane == 0) {
      // before re-using a stage's buffer wait until its bulk group has been READ (at most STAGES-1 groups pending)
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(STAGES - 1) : "memory");
      for (int r = warp; r < ROWS; r += nwarps) {
        double2 *g = dst + (rb * ROWS + r) * stride + cb * W;
        unsigned s = (unsigned)__cvta_generic_to_shared(sm + stage * TILE + r * RUN);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(s), "n"(RUN) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    stage = (stage + 1) % STAGES;
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// TMA tensor stores: one 2-D box (W16 elements of 16 B x ROWS rows = 16 KiB) per operation
__global__ void __launch_bounds__(128) tensor_store(const __grid_constant__ CUtensorMap tm, int w_dbl, int rows, long long nrows, long long row_len_dbl) {
  extern __shared__ __align__(128) unsigned char sm[];
  constexpr int TILE = 16384, STAGES = 4;
  for (int i = threadIdx.x; i < TILE * STAGES / 16; i += blockDim.x) ((double2 *)sm)[i] = make_double2(1.0, 2.0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const long long cbs = row_len_dbl / w_dbl, rbs = nrows / rows;
  int stage = 0;
  if (threadIdx.x == 0) {
    for (long long t = blockIdx.x; t < cbs * rbs; t += gridDim.x) {
      const int x = (int)((t % cbs) * w_dbl), y = (int)((t / cbs) * rows);
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(STAGES - 1) : "memory");
      unsigned s = (unsigned)__cvta_generic_to_shared(sm + stage * TILE);
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&tm), "r"(x), "r"(y), "r"(s) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      stage = (stage + 1) % STAGES;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  const bool have_peer = n >= 2;
  const long long row_len = 1024, nrows = 1024 * 64;  // 1 GiB of double2
  const size_t bytes = (size_t)row_len * nrows * 16;
  double2 *local, *remote = nullptr, *src, *back0 = nullptr, *src1 = nullptr, *loc2;
  cudaStream_t s1a = nullptr;
  if (have_peer) {
    CK(cudaSetDevice(1));
    CK(cudaDeviceEnablePeerAccess(0, 0));
    CK(cudaMalloc(&remote, bytes));
    CK(cudaMalloc(&src1, bytes));
    CK(cudaMemset(src1, 0, bytes));
    CK(cudaStreamCreateWithFlags(&s1a, cudaStreamNonBlocking));
  }
  CK(cudaSetDevice(0));
  if (have_peer) CK(cudaDeviceEnablePeerAccess(1, 0));
  CK(cudaMalloc(&local, bytes));
  CK(cudaMalloc(&src, bytes));
  CK(cudaMalloc(&back0, bytes));  // what device 1 writes into device 0
  CK(cudaMalloc(&loc2, bytes));
  CK(cudaMemset(src, 0, bytes));
  cudaStream_t st[4], hi;
  for (auto &s : st) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  int lo_p, hi_p;
  CK(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
  CK(cudaStreamCreateWithPriority(&hi, cudaStreamNonBlocking, hi_p));
  cudaEvent_t a, b, a1, b1, ej[4];
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (auto &e : ej) cudaEventCreate(&e);
  if (have_peer) { CK(cudaSetDevice(1)); cudaEventCreate(&a1); cudaEventCreate(&b1); CK(cudaSetDevice(0)); }
  // f enqueues work on stream st[0] (others must join st[0] through events before returning)
  auto time = [&](auto f) {
    f(); CK(cudaDeviceSynchronize());
    cudaEventRecord(a, st[0]);
    for (int i = 0; i < 3; i++) f();
    cudaEventRecord(b, st[0]);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / 3;
  };
  EncodeFn encode = nullptr;
  {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    encode = (EncodeFn)fn;
  }
  CK(cudaFuncSetAttribute(bulk_store<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(bulk_store<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(bulk_store<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(bulk_store<2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(bulk_store<16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(tensor_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));

  for (int pass = have_peer ? 0 : 0; pass < (have_peer ? 2 : 1); pass++) {
    double2 *dst = pass ? remote : local;
    printf("---- destination: %s   (1 GiB per measurement)\n", pass ? "PEER over NVLink" : "local HBM");
    float ms;
    for (int ctas : {148, 74}) {
      ms = time([&] { run_store<8, 16><<<ctas, 512, 0, st[0]>>>(dst, row_len, nrows, row_len); });
      printf("SM st.global runs 128B      ctas %3d  %.3f ms  %.0f GB/s\n", ctas, ms, bytes / ms / 1e6);
    }
    for (int ctas : {148, 296, 74}) {
      ms = time([&] { bulk_store<128><<<ctas, 256, 65536, st[0]>>>(dst, row_len, nrows, row_len); });
      printf("TMA bulk store runs 128B    ctas %3d  %.3f ms  %.0f GB/s\n", ctas, ms, bytes / ms / 1e6);
      ms = time([&] { bulk_store<256><<<ctas, 256, 65536, st[0]>>>(dst, row_len, nrows, row_len); });
      printf("TMA bulk store runs 256B    ctas %3d  %.3f ms  %.0f GB/s\n", ctas, ms, bytes / ms / 1e6);
      ms = time([&] { bulk_store<512><<<ctas, 256, 65536, st[0]>>>(dst, row_len, nrows, row_len); });
      printf("TMA bulk store runs 512B    ctas %3d  %.3f ms  %.0f GB/s\n", ctas, ms, bytes / ms / 1e6);
      ms = time([&] { bulk_store<2048><<<ctas, 256, 65536, st[0]>>>(dst, row_len, nrows, row_len); });
      printf("TMA bulk store runs 2KiB    ctas %3d  %.3f ms  %.0f GB/s\n", ctas, ms, bytes / ms / 1e6);
      ms = time([&] { bulk_store<16384><<<ctas, 256, 65536, st[0]>>>(dst, row_len, nrows, row_len); });
      printf("TMA bulk store runs 16KiB   ctas %3d  %.3f ms  %.0f GB/s\n", ctas, ms, bytes / ms / 1e6);
    }
    if (encode) {
      for (int w16 : {8, 16, 64}) {  // inner box = w16 elements of 16 bytes
        CUtensorMap tm;
        const int w_dbl = w16 * 2, rows = 16384 / (w16 * 16);
        cuuint64_t gdim[2] = {(cuuint64_t)row_len * 2, (cuuint64_t)nrows};
        cuuint64_t gstr[1] = {(cuuint64_t)row_len * 16};
        cuuint32_t box[2] = {(cuuint32_t)w_dbl, (cuuint32_t)rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, dst, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("tensor map encode failed (%d) for inner %d B\n", (int)r, w16 * 16); continue; }
        for (int ctas : {148, 296}) {
          ms = time([&] { tensor_store<<<ctas, 128, 65536, st[0]>>>(tm, w_dbl, rows, nrows, row_len * 2); });
          printf("TMA tensor store box %4dB x %3d rows  ctas %3d  %.3f ms  %.0f GB/s\n", w16 * 16, rows, ctas, ms, bytes / ms / 1e6);
        }
      }
    }
    ms = time([&] { CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st[0])); });
    printf("copy engine, one copy                 %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
    for (int nchunk : {8, 64, 448}) {
      for (int nstream : {1, 2, 4}) {
        ms = time([&] {
          const size_t cb = bytes / nchunk;
          cudaEventRecord(ej[0], st[0]);
          for (int s = 1; s < nstream; s++) cudaStreamWaitEvent(st[s], ej[0], 0);
          for (int c = 0; c < nchunk; c++)
            CK(cudaMemcpyAsync((char *)dst + c * cb, (char *)src + c * cb, cb, cudaMemcpyDeviceToDevice, st[c % nstream]));
          for (int s = 1; s < nstream; s++) { cudaEventRecord(ej[s], st[s]); cudaStreamWaitEvent(st[0], ej[s], 0); }
        });
        printf("copy engine, %3d chunks of %6.2f MB on %d streams  %.3f ms  %.0f GB/s\n", nchunk, bytes / nchunk / 1e6, nstream, ms, bytes / ms / 1e6);
      }
    }
    // SM stores (half of the data, 74 CTAs) and the copy engine (other half) at the same time
    ms = time([&] {
      cudaEventRecord(ej[0], st[0]);
      cudaStreamWaitEvent(st[1], ej[0], 0);
      run_store<8, 16><<<74, 512, 0, st[0]>>>(dst, row_len, nrows / 2, row_len);
      CK(cudaMemcpyAsync((char *)dst + bytes / 2, (char *)src + bytes / 2, bytes / 2, cudaMemcpyDeviceToDevice, st[1]));
      cudaEventRecord(ej[1], st[1]);
      cudaStreamWaitEvent(st[0], ej[1], 0);
    });
    printf("SM stores (74 CTAs, half) + copy engine (half) together  %.3f ms  %.0f GB/s aggregate\n", ms, bytes / ms / 1e6);
    // SM stores on 92 high-priority CTAs + a local streaming copy on the rest of the machine
    for (int xc : {148, 92, 74}) {
      ms = time([&] {
        cudaEventRecord(ej[0], st[0]);
        cudaStreamWaitEvent(hi, ej[0], 0);
        run_store<8, 16><<<xc, 512, 0, hi>>>(dst, row_len, nrows, row_len);
        cudaEventRecord(ej[1], hi);
        if (xc < 148)
          for (int r = 0; r < (pass ? 6 : 1); r++) copy16<<<(148 - xc), 512, 0, st[0]>>>(src, loc2, bytes / 16 / 2);
        cudaStreamWaitEvent(st[0], ej[1], 0);
      });
      printf("SM stores on %3d CTAs next to a local copy kernel on %3d CTAs  %.3f ms  %.0f GB/s (stores only)\n", xc, 148 - xc, ms, bytes / ms / 1e6);
    }
    if (pass == 1) {
      // both directions at once: device 1 stores into device 0 while device 0 stores into device 1
      for (int mode = 0; mode < 2; mode++) {
        float ms0 = 0, ms1 = 0;
        for (int rep = 0; rep < 3; rep++) {
          CK(cudaSetDevice(1)); CK(cudaDeviceSynchronize());
          CK(cudaSetDevice(0)); CK(cudaDeviceSynchronize());
          cudaEventRecord(a, st[0]);
          if (mode == 0) run_store<8, 16><<<148, 512, 0, st[0]>>>(remote, row_len, nrows, row_len);
          else CK(cudaMemcpyAsync(remote, src, bytes, cudaMemcpyDeviceToDevice, st[0]));
          cudaEventRecord(b, st[0]);
          CK(cudaSetDevice(1));
          cudaEventRecord(a1, s1a);
          if (mode == 0) run_store<8, 16><<<148, 512, 0, s1a>>>(back0, row_len, nrows, row_len);
          else CK(cudaMemcpyAsync(back0, src1, bytes, cudaMemcpyDeviceToDevice, s1a));
          cudaEventRecord(b1, s1a);
          CK(cudaEventSynchronize(b1));
          cudaEventElapsedTime(&ms1, a1, b1);
          CK(cudaSetDevice(0));
          CK(cudaEventSynchronize(b));
          cudaEventElapsedTime(&ms0, a, b);
        }
        printf("both directions at once (%s): 0->1 %.3f ms %.0f GB/s, 1->0 %.3f ms %.0f GB/s\n", mode ? "copy engine" : "SM stores", ms0,
               bytes / ms0 / 1e6, ms1, bytes / ms1 / 1e6);
      }
    }
  }
  return 0;
}
