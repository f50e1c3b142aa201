// generic_stage.cuh -- the any-length, any-kind stage kernel.
//
// One CTA owns a tile of tile_u x tile_v pencils.  It (1) gathers them from global memory in the
// order that makes the input's unit-stride dimension the fastest thread index, applying the kind's
// pre-processing (Hermitian / even / odd extension, half-sample pre-twiddle) on the fly, (2) runs a
// Stockham autosort FFT of length L in shared memory with one pass per prime factor (radix 4 where
// possible, direct DFT butterflies for odd primes), (3) applies the kind's post-twiddle, the optional
// spectral-derivative factor, and scatters through the segment table in the order that makes the
// OUTPUT's unit-stride dimension the fastest thread index.  So a pencil crosses HBM once per stage
// whatever the storage permutation (replaces reference exec.C:737-2032 + init.C:1146-1607).
//
// Every r2r kind is expressed as a complex-linear map (symmetric extension, FFT, twiddle), so the
// same code serves the REAL and the COMPLEX (re/im transformed separately, init.C:1179-1188) variants.
// The power-of-two fast path lives in pow2_stage.cuh; this kernel is the reference for correctness.
#pragma once
#include "common.cuh"

namespace p3b {

template <typename T>
__device__ __forceinline__ void store_out(const StageParams &P, int k, long long u, long long v,
                                          typename cx<T>::type val) {
  typedef typename cx<T>::type C;
  if (P.deriv_g > 0) {
    T kap = (T)deriv_kappa(k, P.deriv_g);
    val = mk<T>(-kap * val.y, kap * val.x);
  }
  int s = 0;
  while (s + 1 < P.nseg && k >= P.seg[s].k1) s++;
  const SegDev &sg = P.seg[s];
  long long a = sg.off + (long long)(k - sg.k0) * sg.os_d + u * sg.os_u + v * sg.os_v;
  if (P.dt_out == 2) ((C *)sg.base)[a] = val;
  else ((T *)sg.base)[a] = val.x;
}

template <typename T>
__global__ void __launch_bounds__(256) generic_stage_kernel(const __grid_constant__ StageParams P) {
  typedef typename cx<T>::type C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NP = P.tile_u * P.tile_v;
  C *buf0 = reinterpret_cast<C *>(smem_raw);
  C *buf1 = buf0 + (size_t)NP * P.lstride;
  const C *tw = (const C *)P.tw;
  const C *tw2 = (const C *)P.tw2;
  const C *tw3 = (const C *)P.tw3;
  const int L = P.L, n = P.nfft;
  const int tid = threadIdx.x, nth = blockDim.x;

  for (long long tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
    const long long u0 = (tile % P.tiles_u) * P.tile_u;
    const long long v0 = (tile / P.tiles_u) * P.tile_v;
    const int cu = (int)min((long long)P.tile_u, P.nu - u0);
    const int cv = (int)min((long long)P.tile_v, P.nv - v0);

    // ---- zero-fill positions that no input element maps to
    if (P.kind == P3DFFTCU_K_DST1 || P.kind == P3DFFTCU_K_DCT3 || P.kind == P3DFFTCU_K_DST3) {
      for (int p = tid; p < NP; p += nth) {
        C z = mk<T>(0, 0);
        if (P.kind == P3DFFTCU_K_DST1) { buf0[p * P.lstride] = z; buf0[p * P.lstride + n + 1] = z; }
        else if (P.kind == P3DFFTCU_K_DCT3) buf0[p * P.lstride + n] = z;
        else buf0[p * P.lstride] = z;
      }
    }
    // ---- gather + pre-processing
    const int nin = P.n_in;
    const int total_in = NP * nin;
    for (int idx = tid; idx < total_in; idx += nth) {
      int j, pu, pv;
      if (P.load_ord == ORD_D) { j = idx % nin; int p = idx / nin; pu = p % P.tile_u; pv = p / P.tile_u; }
      else if (P.load_ord == ORD_U) { pu = idx % P.tile_u; int r = idx / P.tile_u; pv = r % P.tile_v; j = r / P.tile_v; }
      else { pv = idx % P.tile_v; int r = idx / P.tile_v; pu = r % P.tile_u; j = r / P.tile_u; }
      C x = mk<T>(0, 0);
      if (pu < cu && pv < cv) {
        long long a = (long long)j * P.is_d + (u0 + pu) * P.is_u + (v0 + pv) * P.is_v;
        if (P.dt_in == 2) x = ((const C *)P.in)[a];
        else x.x = ((const T *)P.in)[a];
      }
      C *b = buf0 + (pv * P.tile_u + pu) * P.lstride;
      switch (P.kind) {
        case P3DFFTCU_K_C2R:  // Hermitian extension; imag of X_0 (and X_{N/2}) drops out of the real part
          b[j] = x;
          if (j > 0 && 2 * j < n) b[n - j] = cconj(x);
          break;
        case P3DFFTCU_K_DCT1:  // even extension, L = 2(n-1)
          b[j] = x;
          if (j > 0 && j < n - 1) b[L - j] = x;
          break;
        case P3DFFTCU_K_DST1:  // odd extension, L = 2(n+1)
          b[j + 1] = x;
          b[L - 1 - j] = cneg(x);
          break;
        case P3DFFTCU_K_DCT2:  // half-sample even extension, L = 2n
          b[j] = x;
          b[L - 1 - j] = x;
          break;
        case P3DFFTCU_K_DST2:  // half-sample odd extension
          b[j] = x;
          b[L - 1 - j] = cneg(x);
          break;
        case P3DFFTCU_K_DCT3:  // z_j = x_j w_j, z_{2n-j} = -x_j w_{2n-j}, z_n = 0
          b[j] = cmul(x, tw2[j]);
          if (j > 0) b[L - j] = cneg(cmul(x, tw2[L - j]));
          break;
        case P3DFFTCU_K_DST3: {  // s_m = x_{m-1}: z_m = i s_m w_m (1<=m<=n), z_{2n-m} = i s_m w_{2n-m} (m<n), z_0 = 0
          int m = j + 1;
          b[m] = cmuli(cmul(x, tw2[m]));
          if (m < n) b[L - m] = cmuli(cmul(x, tw2[L - m]));
          break;
        }
        case P3DFFTCU_K_DCT4:  // y_j = x_j (j<n), -x_{2n-1-j} (j>=n); z_j = y_j w_j
          b[j] = cmul(x, tw2[j]);
          b[L - 1 - j] = cneg(cmul(x, tw2[L - 1 - j]));
          break;
        case P3DFFTCU_K_DST4:
          b[j] = cmul(x, tw2[j]);
          b[L - 1 - j] = cmul(x, tw2[L - 1 - j]);
          break;
        default:  // EMPTY, C2C, R2C
          b[j] = x;
      }
    }
    __syncthreads();

    // ---- Stockham passes, ping-pong buf0 <-> buf1
    C *src = buf0, *dst = buf1;
    if (P.kind != P3DFFTCU_K_EMPTY) {
      const bool bwd = (P.kind == P3DFFTCU_K_C2C_BWD || P.kind == P3DFFTCU_K_C2R);
      int Ns = 1;
      for (int f = 0; f < P.nfac; f++) {
        const int r = P.fac[f];
        const int Lr = L / r;
        const int tws = L / (Ns * r);
        const int total = NP * L;
        for (int idx = tid; idx < total; idx += nth) {
          int p = idx / L, o = idx - p * L;
          int k = o % Ns;
          int t = o / Ns;
          int q2 = t % r;
          int j = (t / r) * Ns + k;
          const C *xs = src + p * P.lstride + j;
          int step = (k * tws + q2 * Lr) % L;
          C acc = xs[0];
          int e = 0;
          for (int q = 1; q < r; q++) {
            e += step;
            if (e >= L) e -= L;
            C w = tw[e];
            if (bwd) w.y = -w.y;
            C x = xs[q * Lr];
            acc.x += x.x * w.x - x.y * w.y;
            acc.y += x.x * w.y + x.y * w.x;
          }
          dst[p * P.lstride + o] = acc;
        }
        __syncthreads();
        C *tmp = src; src = dst; dst = tmp;
        Ns *= r;
      }
    }

    // ---- post-processing + scatter
    const int nout = P.n_out;
    const int total_out = NP * nout;
    for (int idx = tid; idx < total_out; idx += nth) {
      int k, pu, pv;
      if (P.store_ord == ORD_D) { k = idx % nout; int p = idx / nout; pu = p % P.tile_u; pv = p / P.tile_u; }
      else if (P.store_ord == ORD_U) { pu = idx % P.tile_u; int r = idx / P.tile_u; pv = r % P.tile_v; k = r / P.tile_v; }
      else { pv = idx % P.tile_v; int r = idx / P.tile_v; pu = r % P.tile_u; k = r / P.tile_u; }
      if (pu >= cu || pv >= cv) continue;
      const C *b = src + (pv * P.tile_u + pu) * P.lstride;
      C y;
      switch (P.kind) {
        case P3DFFTCU_K_DST1: y = cmuli(b[k + 1]); break;
        case P3DFFTCU_K_DCT2: y = cmul(b[k], tw2[k]); break;
        case P3DFFTCU_K_DST2: y = cmuli(cmul(b[k + 1], tw2[k + 1])); break;
        case P3DFFTCU_K_DCT4: y = cmul(b[k], tw3[k]); break;
        case P3DFFTCU_K_DST4: y = cmuli(cmul(b[k], tw3[k])); break;
        default: y = b[k];
      }
      store_out<T>(P, k, u0 + pu, v0 + pv, y);
    }
    __syncthreads();
  }
}

}  // namespace p3b
