// fastcore_inst.cu -- instantiates the fastcore stage kernels (fastcore_stage.cuh) and their launch table.
#include "fastcore_stage.cuh"

namespace p3b {

template <typename T, int M, int BLUE> void fast_launcher(const StageParams &P, int grid, int threads, size_t smem, cudaStream_t s) {
  P3B_LAUNCH((fastcore_stage_kernel<T, M, BLUE>), grid, threads, smem, s, P);
}

template <typename T, int M, int BLUE> FastInfo fast_info_one() {
  FastInfo f;
  f.launch = fast_launcher<T, M, BLUE>;
  f.func = (const void *)fastcore_stage_kernel<T, M, BLUE>;
  f.tp = FastCfg<T, M>::TP;
  f.pitch = FastCfg<T, M>::PITCH;
  f.table_elems = FastCfg<T, M>::T2N + FastCfg<T, M>::T3N;
  return f;
}

template <typename T, int BLUE> bool fast_info_m(int M, FastInfo *out) {
  switch (M) {
    case 64: *out = fast_info_one<T, 64, BLUE>(); return true;
    case 128: *out = fast_info_one<T, 128, BLUE>(); return true;
    case 256: *out = fast_info_one<T, 256, BLUE>(); return true;
    case 512: *out = fast_info_one<T, 512, BLUE>(); return true;
    case 1024: *out = fast_info_one<T, 1024, BLUE>(); return true;
    case 2048: *out = fast_info_one<T, 2048, BLUE>(); return true;
    case 4096: *out = fast_info_one<T, 4096, BLUE>(); return true;
  }
  return false;
}

bool fast_lookup(int prec, int M, int blue, FastInfo *out) {
  if (prec == 8) return blue ? fast_info_m<double, 1>(M, out) : fast_info_m<double, 0>(M, out);
  return blue ? fast_info_m<float, 1>(M, out) : fast_info_m<float, 0>(M, out);
}

}  // namespace p3b
