// common.cuh -- shared device-side definitions for the stage kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "p3dfft_b200.h"

#ifndef P3B_LAUNCH
#define P3B_LAUNCH(kernel, grid, block, smem, stream, params) kernel<<<grid, block, smem, stream>>>(params)
#define P3B_LAUNCH2(kernel, grid, block, smem, stream, p1, p2) kernel<<<grid, block, smem, stream>>>(p1, p2)
#endif

namespace p3b {

template <typename T> struct cx;
template <> struct cx<float> { typedef float2 type; };
template <> struct cx<double> { typedef double2 type; };

template <typename T> __host__ __device__ __forceinline__ typename cx<T>::type mk(T a, T b) {
  typename cx<T>::type r;
  r.x = a;
  r.y = b;
  return r;
}
template <typename C> __device__ __forceinline__ C cadd(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
  C r;
  r.x = a.x * b.x - a.y * b.y;
  r.y = a.x * b.y + a.y * b.x;
  return r;
}
template <typename C> __device__ __forceinline__ C cconj(C a) { C r; r.x = a.x; r.y = -a.y; return r; }
template <typename C> __device__ __forceinline__ C cneg(C a) { C r; r.x = -a.x; r.y = -a.y; return r; }
// multiply by +i
template <typename C> __device__ __forceinline__ C cmuli(C a) { C r; r.x = -a.y; r.y = a.x; return r; }
// multiply by -i
template <typename C> __device__ __forceinline__ C cmulmi(C a) { C r; r.x = a.y; r.y = -a.x; return r; }

// device copy of one output segment (see p3dfftcu_seg): base already resolved to a (possibly peer) pointer
struct SegDev {
  void *base;
  long long off, os_d, os_u, os_v;
  int k0, k1;
};

// iteration orders for the load / store loops of a pencil tile
enum { ORD_D = 0, ORD_U = 1, ORD_V = 2 };

struct StageParams {
  const void *in;
  long long nu, nv, is_d, is_u, is_v;
  long long tiles_u, ntiles;
  long long tiles_v;  // pipelined kernel only: tile count along v and
  int vfast;          // whether consecutive tile numbers run along v (else along u)
  int kind, dt_in, dt_out;
  int nfft, n_in, n_out, L;  // L = length of the internal complex FFT
  int tile_u, tile_v, tu_log2;
  int load_ord, store_ord;
  int lstride;               // shared-memory pitch of one pencil (complex elements)
  int nfac;
  int fac[20];
  const void *tw;   // exp(-2 pi i j / L), j < L
  const void *tw2;  // exp(-i pi j / (2 n)), j < 2n   (r2r kinds II-IV)
  const void *tw3;  // exp(-i pi (2k+1) / (4 n)), k < n (r2r kinds IV)
  const void *tw_core;  // fastcore kernel: exp(-2 pi i j / M) of the power-of-two core
  const void *chirp;    // Bluestein: c_j = exp(-i pi j^2 / L), j < L
  const void *bhat;     // Bluestein: FFT_M of the wrapped conj chirp, divided by M
  int pipe_bytes;   // TMA kernel, r2r kinds: bytes of one pencil's bulk copy (n_in elements rounded up to 16 bytes)
  int tl_swap;      // tensor-load kernel: != 0 when the transform dimension is dim 2 of the tensor map (dims 1, 2 are ordered by stride)
  int deriv_g;      // > 0: spectral derivative epilogue with full length g
  int nseg;
  SegDev seg[P3DFFTCU_MAXSEG];
};

// ---- tile groups and flags: one persistent launch per stage of an overlapped pair (pow2_pipe.cuh, SYNC = 1)
// The pencils of a stage are cut into groups (rectangles of the (u, v) pencil plane, processed in table order).  A group may
// WAIT for a flag before its pencils are loaded (its input is produced by another kernel: the neighbouring stage on this GPU,
// or the exchange stages of the peer GPUs) and may SIGNAL a flag once all its outputs have been stored.  Flags are 8-byte
// words holding an epoch: the number of flag-synchronised execs the writer and the reader have run TOGETHER (kept per pair
// of ranks by the host, so plans on different sub-communicators cannot confuse each other); epochs only grow, so flags are
// never reset.
#define P3B_MAXGRP 24
#define P3B_MAXSRC 32
struct TileGroupDev {
  int u0, u1, v0, v1;    // pencil ranges [u0,u1) x [v0,v1)
  int tiles_u, tiles_v;  // tiles of the group along u and v
  long long tile0;       // global number of its first tile
  int wait_id;           // >= 0: flag id every wait source must have published before the group is loaded
  int signal_id;         // >= 0: flag id published to every signal target when the group is complete
  int count;             // != 0: completed pencils are counted in ctl[1 + g] (implied by signal_id >= 0)
  int after;             // the flag is published only once the groups [0, after) are complete as well (they must count)
};
struct SyncDev {
  int ngroups;
  int dynamic;                  // work comes from the counter ctl[0] instead of blockIdx striding: whole tiles (transposed-output
                                // kernels, one CTA barrier per tile anyway) or single pencils (contiguous-output kernels)
  int keep_ctas;                // dynamic only: CTAs >= keep_ctas stop taking work once the counter has reached boost_limit
  unsigned long long boost_limit;  //   (the stage starts on every SM and then leaves room for its partner kernel)
  unsigned long long *ctl;      // [0] tile counter, [1 + g] pencils of group g completed (zeroed by the host before the launch)
  unsigned long long timeout_ns;  // > 0: trap when a wait lasts longer (0 = wait for ever)
  const unsigned long long *wait_base;  // word of (source j, id) = wait_base[wait_off[j] + id]
  int wait_n, sig_n;
  int wait_off[P3B_MAXSRC];
  unsigned long long wait_epoch[P3B_MAXSRC];  // value source j writes for this exec
  unsigned long long *sig_ptr[P3B_MAXSRC];    // word of (target j, id) = sig_ptr[j][id] (peer memory over NVLink, or local)
  unsigned long long sig_epoch[P3B_MAXSRC];   // value target j expects for this exec
  TileGroupDev grp[P3B_MAXGRP];
};

// derivative multiplier of output index k for full spectral length g (reference exec.C:228-287):
// returns kappa with out = i*kappa*in; kappa = k (k < g/2), 0 (k == g/2), k-g (k > g/2)
__device__ __forceinline__ int deriv_kappa(int k, int g) {
  int mid = g / 2;
  return k < mid ? k : (k == mid ? 0 : k - g);
}

}  // namespace p3b
