// common.cuh -- shared device-side definitions for the stage kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "p3dfft_b200.h"

#ifndef P3B_LAUNCH
#define P3B_LAUNCH(kernel, grid, block, smem, stream, params) kernel<<<grid, block, smem, stream>>>(params)
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
  int deriv_g;      // > 0: spectral derivative epilogue with full length g
  int nseg;
  SegDev seg[P3DFFTCU_MAXSEG];
};

// derivative multiplier of output index k for full spectral length g (reference exec.C:228-287):
// returns kappa with out = i*kappa*in; kappa = k (k < g/2), 0 (k == g/2), k-g (k > g/2)
__device__ __forceinline__ int deriv_kappa(int k, int g) {
  int mid = g / 2;
  return k < mid ? k : (k == mid ? 0 : k - g);
}

}  // namespace p3b
