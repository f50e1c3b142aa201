// pow2_tload_inst.cu -- instantiates the tensor-load stage kernels (pow2_tload.cuh) of ONE (precision, kind) pair, kind in
// {C2C forward, C2C backward, R2C}; the Makefile compiles this file six times (-DPIPE_PREC=4|8 -DPIPE_KIND=1..3).
#define P3B_PIPE_TU 1
#include "pow2_tload.cuh"

#if PIPE_PREC == 8
#define PIPE_T double
#else
#define PIPE_T float
#endif
#define TL_CAT2(a, b, c) a##b##_##c
#define TL_CAT(a, b, c) TL_CAT2(a, b, c)
#define TL_FN TL_CAT(tload_lookup_p, PIPE_PREC, PIPE_KIND)

namespace p3b {

const TLoadInfo *TL_FN(int ts, int M, int P) {
  return ts ? tload_info<PIPE_T, PIPE_KIND, 1>(M, P) : tload_info<PIPE_T, PIPE_KIND, 0>(M, P);
}

}  // namespace p3b
