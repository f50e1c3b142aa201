// fastcore_stage.cuh -- any-kind stage kernel on the register-resident power-of-two FFT core.
//
// Serves every transform whose internal complex FFT length L is either a power of two (64...4096) that the TMA kernel
// does not take (all r2r kinds: DCT-I of 2^k+1 points, DST-I of 2^k-1, DCT/DST-II..IV of 2^(k-1); strided or unaligned
// inputs) or ANY other length with 2L-1 <= 4096, through Bluestein's chirp-z identity
//     X_k = c_k * sum_j (x_j c_j) conj(c)_{k-j},  c_j = exp(-i pi j^2 / L),
// i.e. two length-M FFTs (M = the power of two >= 2L-1) around a pointwise product with the precomputed spectrum of the
// chirp.  The kind's pre-processing (Hermitian / even / odd extension, half-sample pre-twiddle), post-twiddle, fused
// derivative, storage reorder and per-peer segment table are the generic kernel's (generic_stage.cuh); only its
// one-pass-per-prime-factor shared-memory Stockham loop is replaced by the radix-16 register passes of pow2_stage.cuh
// with smem twiddle tables (pow2_pipe.cuh), which is what makes the r2r kinds and awkward lengths (1022 = 2*7*73 for a
// 512-point Chebyshev DCT-I) run at FFT speed instead of O(L * sum of prime factors).
// Replaces reference FFTW r2r / c2c plans of init.C:1146-1607 + reorder_trans (exec.C:737-1326).
#pragma once
#include "common.cuh"
#include "generic_stage.cuh"
#include "pow2_stage.cuh"
#include "pow2_pipe.cuh"

namespace p3b {

// forward FFT core of M points on the E register values of the TP threads of one pencil; exchanges go through the
// pencil's padded buffer Bp; `sync` is a barrier over (at least) the threads of the pencil.  In: v[m] = x[t + m TP];
// out: v[m] = X[t + m TP].  The caller guarantees that nobody still reads Bp.
template <typename T, int M, typename Sync>
__device__ __forceinline__ void fast_core(typename cx<T>::type *v, int t, typename cx<T>::type *Bp, const typename cx<T>::type *T2,
                                          const typename cx<T>::type *T3, const typename cx<T>::type *tw, Sync sync) {
  typedef Pow2Cfg<M> R;
  constexpr int E = R::E, R1 = R::R1, R2 = R::R2, R3 = R::R3;
  reg_pass<T, M, E, R1, false>(v, t, 1, tw, 1);
  smem_scatter<T, M, E, R1>(v, Bp, t, 1);
  sync();
  smem_gather<T, M, E>(v, Bp, t);
  reg_pass2<T, M, E, R1, R2>(v, t, T2);
  if constexpr (R3 > 1) {
    sync();
    smem_scatter<T, M, E, R2>(v, Bp, t, R1);
    sync();
    smem_gather<T, M, E>(v, Bp, t);
    reg_pass3<T, M, E, R3>(v, t, T3);
  }
}

template <typename T, int M> struct FastCfg {
  enum { E = Pow2Cfg<M>::E, TP = M / E, PITCH = Pow2Smem<M>::PENCIL };
  enum { R1 = Pow2Cfg<M>::R1, R2 = Pow2Cfg<M>::R2, R3 = Pow2Cfg<M>::R3, T2N = R1 * R2, T3N = R3 > 1 ? R3 * TP : 0 };
  static constexpr size_t csz = 2 * sizeof(T);
  static size_t smem(int np) { return ((size_t)np * PITCH + T2N + T3N) * csz; }
};

// BLUE = 0: L == M.  BLUE = 1: L < M/2 + 1, chirp-z.  blockDim.x = NP * TP with NP = tile_u * tile_v pencils per tile
template <typename T, int M, int BLUE>
__global__ void __launch_bounds__(512) fastcore_stage_kernel(const __grid_constant__ StageParams P) {
  typedef typename cx<T>::type C;
  typedef FastCfg<T, M> Cfg;
  constexpr int E = Cfg::E, TP = Cfg::TP, PITCH = Cfg::PITCH;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NP = P.tile_u * P.tile_v;
  C *buf0 = reinterpret_cast<C *>(smem_raw);
  C *T2 = buf0 + (size_t)NP * PITCH;
  C *T3 = T2 + Cfg::T2N;
  const C *tw = (const C *)P.tw;        // exp(-2 pi i j / L): the kinds' own twiddles (C2R split is not used here)
  const C *twc = (const C *)P.tw_core;  // exp(-2 pi i j / M): the FFT core
  const C *tw2 = (const C *)P.tw2;
  const C *tw3 = (const C *)P.tw3;
  const C *chirp = (const C *)P.chirp;  // c_j, j < L
  const C *bhat = (const C *)P.bhat;    // FFT_M of the wrapped conj chirp, divided by M
  const int L = P.L, n = P.nfft;
  const int tid = threadIdx.x, nth = blockDim.x;
  (void)tw;
  for (int i = tid; i < Cfg::T2N; i += nth) T2[i] = twc[(i / Cfg::R1) * (i % Cfg::R1) * (M / (Cfg::R1 * Cfg::R2))];
  for (int i = tid; i < Cfg::T3N; i += nth) T3[i] = twc[(i / TP) * (i % TP)];
  const int slot = tid / TP, t = tid % TP;  // FFT mapping: pencil-major
  const int tv_log2 = __ffs(P.tile_v) - 1, np_log2 = P.tu_log2 + tv_log2;  // tile_u, tile_v are powers of two
  C *Bp = buf0 + slot * PITCH;
  const bool bwd = (P.kind == P3DFFTCU_K_C2C_BWD || P.kind == P3DFFTCU_K_C2R);
  __syncthreads();
  // when input and output are both unit-stride along the transform dimension every phase of a pencil is done by its own
  // TP threads: barriers then span that group only (whole warps), and the pencils of a CTA drift apart and overlap their
  // load / shared-memory / FP64 / store phases; otherwise the gather or the scatter crosses pencils and the CTA synchronises
  const bool grouped = P.load_ord == ORD_D && P.store_ord == ORD_D && TP >= 32 && NP <= 15;
  auto sync = [&]() {
    if (grouped) {
      if (TP == 32) __syncwarp();
      else group_bar(1 + slot, TP);
    } else __syncthreads();
  };

  for (long long tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
    const long long u0 = (tile % P.tiles_u) * P.tile_u;
    const long long v0 = (tile / P.tiles_u) * P.tile_v;
    const int cu = (int)min((long long)P.tile_u, P.nu - u0);
    const int cv = (int)min((long long)P.tile_v, P.nv - v0);

    // ---- zero-fill positions that no input element maps to (each pencil by its own first thread)
    if (t == 0 && (P.kind == P3DFFTCU_K_DST1 || P.kind == P3DFFTCU_K_DCT3 || P.kind == P3DFFTCU_K_DST3)) {
      const C z = mk<T>(0, 0);
      if (P.kind == P3DFFTCU_K_DST1) { Bp[0] = z; Bp[n + 1] = z; }
      else if (P.kind == P3DFFTCU_K_DCT3) Bp[n] = z;
      else Bp[0] = z;
    }
    // ---- gather + pre-processing
    const int nin = P.n_in;
    {  // thread -> (pencil, first index, step): along d when the input is unit-stride there, else lanes across pencils
      int pq, j0, jstep;
      if (P.load_ord == ORD_D) { pq = slot; j0 = t; jstep = TP; }
      else { pq = tid & (NP - 1); j0 = tid >> np_log2; jstep = nth >> np_log2; }
      int pu, pv;
      if (P.load_ord == ORD_V) { pv = pq & (P.tile_v - 1); pu = pq >> tv_log2; }
      else { pu = pq & (P.tile_u - 1); pv = pq >> P.tu_log2; }
      const bool live_in = pu < cu && pv < cv;
      const long long a0 = (u0 + pu) * P.is_u + (v0 + pv) * P.is_v;
      C *b = buf0 + (pv * P.tile_u + pu) * PITCH;
    for (int j = j0; j < nin; j += jstep) {
      C x = mk<T>(0, 0);
      if (live_in) {
        const long long a = a0 + (long long)j * P.is_d;
        if (P.dt_in == 2) x = ((const C *)P.in)[a];
        else x.x = ((const T *)P.in)[a];
      }
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
    }
    sync();


    // ---- length-L FFT of every pencil of the tile on the register core
    if (P.kind != P3DFFTCU_K_EMPTY) {
      C v[E];
#pragma unroll
      for (int m = 0; m < E; m++) {
        const int j = t + m * TP;
        C x = mk<T>(0, 0);
        if (!BLUE || j < L) {
          x = Bp[j];
          if (bwd) x.y = -x.y;
          if (BLUE) x = cmul(x, __ldg(&chirp[j]));
        }
        v[m] = x;
      }
      sync();  // everything is in registers: the buffers now carry the exchanges
      fast_core<T, M>(v, t, Bp, T2, T3, twc, sync);
      if (BLUE) {
        // pointwise product with the chirp spectrum, inverse FFT by the conjugation trick (1/M is folded into bhat)
#pragma unroll
        for (int m = 0; m < E; m++) v[m] = cconj(cmul(v[m], __ldg(&bhat[t + m * TP])));
        sync();
        fast_core<T, M>(v, t, Bp, T2, T3, twc, sync);
#pragma unroll
        for (int m = 0; m < E; m++) {
          const int k = t + m * TP;
          if (k < L) v[m] = cmul(cconj(v[m]), __ldg(&chirp[k]));
        }
      }
      sync();  // the last exchange has been read back
#pragma unroll
      for (int m = 0; m < E; m++) {
        const int k = t + m * TP;
        if (!BLUE || k < L) Bp[k] = bwd ? cconj(v[m]) : v[m];
      }
      sync();
    }

    // ---- post-processing + scatter
    const int nout = P.n_out;
    {
      int pq, k0, kstep;
      if (P.store_ord == ORD_D) { pq = slot; k0 = t; kstep = TP; }
      else { pq = tid & (NP - 1); k0 = tid >> np_log2; kstep = nth >> np_log2; }
      int pu, pv;
      if (P.store_ord == ORD_V) { pv = pq & (P.tile_v - 1); pu = pq >> tv_log2; }
      else { pu = pq & (P.tile_u - 1); pv = pq >> P.tu_log2; }
      const bool live_out = pu < cu && pv < cv;
      const C *b = buf0 + (pv * P.tile_u + pu) * PITCH;
    for (int k = live_out ? k0 : nout; k < nout; k += kstep) {
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
    }
    sync();
  }
}

// ------------------------------------------------------------------ host side
struct FastInfo {
  void (*launch)(const StageParams &, int grid, int threads, size_t smem, cudaStream_t);
  const void *func;
  int tp, pitch, table_elems;
};
// defined in fastcore_inst.cu
bool fast_lookup(int prec, int M, int blue, FastInfo *out);

}  // namespace p3b
