// pow2_pipe_inst.cu -- instantiates the pipelined stage kernels of ONE (precision, kind) pair; the Makefile compiles this
// file eight times (-DPIPE_PREC=4|8 -DPIPE_KIND=1..4) so the instantiations build in parallel.
#define P3B_PIPE_TU 1
#include "pow2_pipe.cuh"

#if PIPE_PREC == 8
#define PIPE_T double
#else
#define PIPE_T float
#endif
#define PIPE_CAT2(a, b, c) a##b##_##c
#define PIPE_CAT(a, b, c) PIPE_CAT2(a, b, c)
#define PIPE_FN PIPE_CAT(pipe_lookup_p, PIPE_PREC, PIPE_KIND)

namespace p3b {

const PipeInfo *PIPE_FN(int ts, int M, int P) {
  return ts ? pipe_info<PIPE_T, PIPE_KIND, 1>(M, P) : pipe_info<PIPE_T, PIPE_KIND, 0>(M, P);
}

}  // namespace p3b
