// mixed_pipe_inst.cu -- instantiates the smooth-length stage kernels (mixed_pipe.cuh) of ONE (precision, kind) pair; the
// Makefile compiles this file eight times (-DPIPE_PREC=4|8 -DPIPE_KIND=1..4) so the instantiations build in parallel.
#define P3B_PIPE_TU 1
#include "mixed_pipe.cuh"

#if PIPE_PREC == 8
#define PIPE_T double
#else
#define PIPE_T float
#endif
#define MIX_CAT2(a, b, c) a##b##_##c
#define MIX_CAT(a, b, c) MIX_CAT2(a, b, c)
#define MIX_FN MIX_CAT(mixed_lookup_p, PIPE_PREC, PIPE_KIND)

namespace p3b {

const PipeInfo *MIX_FN(int ts, int Q, int MC, int P) {
  return ts ? mixed_info<PIPE_T, PIPE_KIND, 1>(Q, MC, P) : mixed_info<PIPE_T, PIPE_KIND, 0>(Q, MC, P);
}

}  // namespace p3b
