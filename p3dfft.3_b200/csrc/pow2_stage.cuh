// pow2_stage.cuh -- fast path for C2C / R2C / C2R stages whose FFT core length M is a power of two.
//
// A pencil is handled by TP = M/E threads, each holding E complex elements in registers.  The FFT is
// a 2- or 3-pass Stockham decomposition M = R1*R2[*R3] with radix-{2,4,8,16} butterflies done entirely
// in registers; between passes the data is exchanged through padded shared memory (one buffer, the
// padding makes the stride-R scatter and the unit-stride gather both bank-conflict free).  The first
// pass consumes values straight from global memory and the last pass stores straight to global
// memory, so per stage every element crosses HBM exactly once and shared memory at most twice
// (three times for R2C, which needs Z[M-k] next to Z[k]).
//
// Thread -> (pencil, slot) mapping is chosen separately for the load side (pass 1) and the store side
// (later passes): along the transform dimension when that is the unit-stride direction, across the
// tile's pencils when a different dimension is.  The exchange between pass 1 and pass 2 re-maps for
// free.  This is what lets one kernel replace the reference's FFTW call + reorder_trans variants +
// pack_sendbuf (exec.C:737-1326, 2792-2879) without a separate transpose pass.
//
// R2C uses the length-N/2 complex FFT of the even/odd-packed real data plus a Hermitian split;
// C2R is the mirror image.  Backward transforms use conj(F(conj(x))).
#pragma once
#include <cstdlib>
#include <string>
#include "common.cuh"

namespace p3b {

// ------------------------------------------------------------------ in-register butterflies (forward, exp(-i...))
template <typename C> __device__ __forceinline__ void bfly2(C &a, C &b) {
  C t = csub(a, b);
  a = cadd(a, b);
  b = t;
}
template <typename C> __device__ __forceinline__ void bfly4(C &v0, C &v1, C &v2, C &v3) {
  C a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3), a3 = cmulmi(csub(v1, v3));
  v0 = cadd(a0, a2);
  v2 = csub(a0, a2);
  v1 = cadd(a1, a3);
  v3 = csub(a1, a3);
}

template <typename T, int R> struct Radix;
template <typename T> struct Radix<T, 2> {
  typedef typename cx<T>::type C;
  static __device__ __forceinline__ void run(C *v) { bfly2(v[0], v[1]); }
};
template <typename T> struct Radix<T, 4> {
  typedef typename cx<T>::type C;
  static __device__ __forceinline__ void run(C *v) { bfly4(v[0], v[1], v[2], v[3]); }
};
template <typename T> struct Radix<T, 8> {
  typedef typename cx<T>::type C;
  static __device__ __forceinline__ void run(C *v) {
    // n = 2*n1 + n2 (n1<4, n2<2): radix-4 over n1 for each n2, twiddle W8^{n2*k1}, radix-2 over n2; X[k1 + 4*k2]
    const T h = (T)0.70710678118654752440084436210485;
    C e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    C o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    bfly4(e0, e1, e2, e3);
    bfly4(o0, o1, o2, o3);
    o1 = mk<T>(h * (o1.x + o1.y), h * (o1.y - o1.x));   // * W8^1 = (1-i)/sqrt2
    o2 = cmulmi(o2);                                    // * W8^2 = -i
    o3 = mk<T>(h * (o3.y - o3.x), -h * (o3.x + o3.y));  // * W8^3 = (-1-i)/sqrt2
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
    v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
  }
};
template <typename T> struct Radix<T, 16> {
  typedef typename cx<T>::type C;
  static __device__ __forceinline__ C tw16(C a, int m) {
    // a * W16^m for the exponents that occur (m = n2*k1, n2,k1 < 4)
    const T c = (T)0.92387953251128675612818318939679, s = (T)0.38268343236508977172845998403040;
    const T h = (T)0.70710678118654752440084436210485;
    switch (m) {
      case 0: return a;
      case 1: return mk<T>(a.x * c + a.y * s, a.y * c - a.x * s);
      case 2: return mk<T>(h * (a.x + a.y), h * (a.y - a.x));
      case 3: return mk<T>(a.x * s + a.y * c, a.y * s - a.x * c);
      case 4: return cmulmi(a);
      case 6: return mk<T>(h * (a.y - a.x), -h * (a.x + a.y));
      default: /* 9 */ return mk<T>(-a.x * c - a.y * s, a.x * s - a.y * c);
    }
  }
  static __device__ __forceinline__ void run(C *v) {
    // n = 4*n1 + n2: radix-4 over n1 for each n2 -> A[k1][n2]; twiddle W16^{n2*k1}; radix-4 over n2 -> X[k1 + 4*k2]
    C a[16];
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) {
      C x0 = v[n2], x1 = v[4 + n2], x2 = v[8 + n2], x3 = v[12 + n2];
      bfly4(x0, x1, x2, x3);
      a[0 * 4 + n2] = x0;
      a[1 * 4 + n2] = tw16(x1, n2);
      a[2 * 4 + n2] = tw16(x2, 2 * n2);
      a[3 * 4 + n2] = tw16(x3, 3 * n2);
    }
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) {
      C x0 = a[k1 * 4 + 0], x1 = a[k1 * 4 + 1], x2 = a[k1 * 4 + 2], x3 = a[k1 * 4 + 3];
      bfly4(x0, x1, x2, x3);
      v[k1] = x0;
      v[k1 + 4] = x1;
      v[k1 + 8] = x2;
      v[k1 + 12] = x3;
    }
  }
};

// ------------------------------------------------------------------ size configuration
template <int M> struct Pow2Cfg;  // E elements per thread, radices R1*R2*R3 = M
template <> struct Pow2Cfg<64>   { enum { E = 8,  R1 = 8,  R2 = 8,  R3 = 1 }; };
template <> struct Pow2Cfg<128>  { enum { E = 16, R1 = 16, R2 = 8,  R3 = 1 }; };
template <> struct Pow2Cfg<256>  { enum { E = 16, R1 = 16, R2 = 16, R3 = 1 }; };
#ifdef P3B_M512_E8  // experiment: 8 values per thread (64 threads per pencil, 64 registers) for twice the warps per SM.
                    // Measured (1024^3 double): R2C stage 4.23 vs 4.6-4.7 TB/s, C2R 5.27 vs 6.1-6.2 TB/s -> not the default
template <> struct Pow2Cfg<512>  { enum { E = 8,  R1 = 8,  R2 = 8,  R3 = 8 }; };
#else
template <> struct Pow2Cfg<512>  { enum { E = 16, R1 = 16, R2 = 16, R3 = 2 }; };
#endif
template <> struct Pow2Cfg<1024> { enum { E = 16, R1 = 16, R2 = 16, R3 = 4 }; };
template <> struct Pow2Cfg<2048> { enum { E = 16, R1 = 16, R2 = 16, R3 = 8 }; };
template <> struct Pow2Cfg<4096> { enum { E = 16, R1 = 16, R2 = 16, R3 = 16 }; };
// values per thread of the M-point core (host side)
inline int pow2_values_per_thread(int M) {
  switch (M) {
    case 64: return Pow2Cfg<64>::E;
    case 128: return Pow2Cfg<128>::E;
    case 256: return Pow2Cfg<256>::E;
    case 512: return Pow2Cfg<512>::E;
    case 1024: return Pow2Cfg<1024>::E;
    case 2048: return Pow2Cfg<2048>::E;
    default: return Pow2Cfg<4096>::E;
  }
}

// padded shared-memory index: one pad element after every 16
__device__ __forceinline__ int padidx(int i) { return i + (i >> 4); }
template <int M> struct Pow2Smem { enum { PENCIL = M + M / 16 + 1 }; };  // odd pitch in elements

// one radix-R pass over the E register values of a thread.
//   values v[b + q*(E/R)] (q<R) form butterfly j = t + b*TP; Ns = product of earlier radices
template <typename T, int M, int E, int R, bool TWIDDLE>
__device__ __forceinline__ void reg_pass(typename cx<T>::type *v, int t, int Ns, const typename cx<T>::type *__restrict__ tw,
                                         int twscale) {
  typedef typename cx<T>::type C;
  constexpr int TP = M / E;
  constexpr int NB = E / R;
#ifdef P3B_SKELETON  // tools/microbench only: data movement without arithmetic, to find the ceiling of the access pattern
  return;
#endif
#pragma unroll
  for (int b = 0; b < NB; b++) {
    C a[R];
#pragma unroll
    for (int q = 0; q < R; q++) a[q] = v[b + q * NB];
    if (TWIDDLE) {
      int j = t + b * TP;
      int k = j & (Ns - 1);
      int step = k * (M / (Ns * R)) * twscale;  // exponent unit of the table is 2 pi / (M * twscale)
#pragma unroll
      for (int q = 1; q < R; q++) {
        C w = __ldg(&tw[q * step]);
        a[q] = cmul(a[q], w);
      }
    }
    Radix<T, R>::run(a);
#pragma unroll
    for (int q = 0; q < R; q++) v[b + q * NB] = a[q];
  }
}

// scatter the outputs of a pass into shared memory in Stockham order
template <typename T, int M, int E, int R>
__device__ __forceinline__ void smem_scatter(const typename cx<T>::type *v, typename cx<T>::type *sp, int t, int Ns) {
  constexpr int TP = M / E;
  constexpr int NB = E / R;
#pragma unroll
  for (int b = 0; b < NB; b++) {
    int j = t + b * TP;
    int k = j & (Ns - 1);
    int base = (j - k) * R + k;
#pragma unroll
    for (int q = 0; q < R; q++) sp[padidx(base + q * Ns)] = v[b + q * NB];
  }
}
template <typename T, int M, int E>
__device__ __forceinline__ void smem_gather(typename cx<T>::type *v, const typename cx<T>::type *sp, int t) {
  constexpr int TP = M / E;
#pragma unroll
  for (int m = 0; m < E; m++) v[m] = sp[padidx(t + m * TP)];
}

// ------------------------------------------------------------------ the kernel
template <typename T, int M, int THREADS>
__device__ __forceinline__ void pow2_stage_body(const StageParams &P) {
  typedef typename cx<T>::type C;
  typedef Pow2Cfg<M> Cfg;
  constexpr int E = Cfg::E, R1 = Cfg::R1, R2 = Cfg::R2, R3 = Cfg::R3;
  constexpr int TP = M / E;
  constexpr int NPB = THREADS / TP;  // pencils per CTA iteration
  constexpr int PITCH = Pow2Smem<M>::PENCIL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C *sm = reinterpret_cast<C *>(smem_raw);
  const C *__restrict__ tw = (const C *)P.tw;  // exp(-2 pi i j / nfft)
  const int kind = P.kind;
  const bool r2c = kind == P3DFFTCU_K_R2C, c2r = kind == P3DFFTCU_K_C2R;
  const bool bwd = kind == P3DFFTCU_K_C2C_BWD || c2r;
  const int twscale = (r2c || c2r) ? 2 : 1;  // table is for nfft = 2M in the real cases
  const int tid = threadIdx.x;

  // load-side and store-side thread mappings: (pencil-in-tile, slot t)
  int pL, tL, pS, tS;
  if (P.load_ord == ORD_D) { pL = tid / TP; tL = tid % TP; } else { pL = tid % NPB; tL = tid / NPB; }
  if (P.store_ord == ORD_D) { pS = tid / TP; tS = tid % TP; } else { pS = tid % NPB; tS = tid / NPB; }
  // pencil index -> (pu, pv) inside the tile; ORD_V enumerates v fastest
  int puL, pvL, puS, pvS;
  if (P.load_ord == ORD_V) { pvL = pL % P.tile_v; puL = pL / P.tile_v; } else { puL = pL % P.tile_u; pvL = pL / P.tile_u; }
  if (P.store_ord == ORD_V) { pvS = pS % P.tile_v; puS = pS / P.tile_v; } else { puS = pS % P.tile_u; pvS = pS / P.tile_u; }
  C *smL = sm + (pvL * P.tile_u + puL) * PITCH;
  C *smS = sm + (pvS * P.tile_u + puS) * PITCH;

  for (long long tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
    const long long u0 = (tile % P.tiles_u) * P.tile_u;
    const long long v0 = (tile / P.tiles_u) * P.tile_v;
    const bool liveL = (u0 + puL < P.nu) && (v0 + pvL < P.nv);
    const bool liveS = (u0 + puS < P.nu) && (v0 + pvS < P.nv);
    C v[E];

    // ---------------- load (pass-1 mapping)
    {
      const long long base = (u0 + puL) * P.is_u + (v0 + pvL) * P.is_v;
      if (!liveL) {
#pragma unroll
        for (int m = 0; m < E; m++) v[m] = mk<T>(0, 0);
      } else if (r2c) {
        const T *in = (const T *)P.in;
        if (P.is_d == 1 && (((uintptr_t)(in + base)) & (sizeof(C) - 1)) == 0) {
          const C *inc = (const C *)(in + base);
#pragma unroll
          for (int m = 0; m < E; m++) v[m] = inc[tL + m * TP];
        } else {
#pragma unroll
          for (int m = 0; m < E; m++) {
            long long a = base + (long long)(2 * (tL + m * TP)) * P.is_d;
            v[m] = mk<T>(in[a], in[a + P.is_d]);
          }
        }
      } else if (c2r) {
        // Z[k] = (X[k] + conj X[M-k]) + i e^{+2 pi i k/N} (X[k] - conj X[M-k]); conj-trick folded in (we need conj Z)
        const C *in = (const C *)P.in + base;
#pragma unroll
        for (int m = 0; m < E; m++) {
          int k = tL + m * TP;
          C a = in[(long long)k * P.is_d];
          C b = cconj(in[(long long)(M - k) * P.is_d]);
          if (k == 0) { a.y = 0; b.y = 0; }  // FFTW's c2r ignores Im X[0] and Im X[N/2]
          C s = cadd(a, b), d = csub(a, b);
          C w = cconj(__ldg(&tw[k]));  // e^{+2 pi i k / N}
          C e = cmuli(cmul(d, w));
          v[m] = cconj(cadd(s, e));
        }
      } else {
        const C *in = (const C *)P.in + base;
#pragma unroll
        for (int m = 0; m < E; m++) {
          C x = in[(long long)(tL + m * TP) * P.is_d];
          v[m] = bwd ? cconj(x) : x;
        }
      }
    }

    // ---------------- pass 1 (no twiddles), exchange, pass 2 [, exchange, pass 3]
    reg_pass<T, M, E, R1, false>(v, tL, 1, tw, twscale);
    smem_scatter<T, M, E, R1>(v, smL, tL, 1);
    __syncthreads();
    smem_gather<T, M, E>(v, smS, tS);
    reg_pass<T, M, E, R2, true>(v, tS, R1, tw, twscale);
    if (R3 > 1) {
      __syncthreads();
      smem_scatter<T, M, E, R2>(v, smS, tS, R1);
      __syncthreads();
      smem_gather<T, M, E>(v, smS, tS);
      reg_pass<T, M, E, (R3 > 1 ? R3 : 2), true>(v, tS, R1 * R2, tw, twscale);
    }
    // now v[m] = F[tS + m*TP] (forward core) of pencil pS

    // ---------------- epilogue + store (store mapping)
    const long long uo = u0 + puS, vo = v0 + pvS;
    if (r2c) {
      // X[k] = ((Z[k] + conj Z[M-k]) - i e^{-2 pi i k/N} (Z[k] - conj Z[M-k])) / 2,  k = 0..M
      __syncthreads();
#pragma unroll
      for (int m = 0; m < E; m++) smS[padidx(tS + m * TP)] = v[m];
      __syncthreads();
      if (liveS) {
#pragma unroll
        for (int m = 0; m < E; m++) {
          int k = tS + m * TP;
          C zk = v[m];
          C zm = cconj(smS[padidx((M - k) & (M - 1))]);
          C s = cadd(zk, zm), d = csub(zk, zm);
          C e = cmulmi(cmul(d, __ldg(&tw[k])));
          C x = mk<T>((T)0.5 * (s.x + e.x), (T)0.5 * (s.y + e.y));
          store_out<T>(P, k, uo, vo, x);
          if (k == 0) store_out<T>(P, M, uo, vo, mk<T>(zk.x - zk.y, (T)0));
        }
      }
    } else if (c2r) {
      if (liveS) {
        // conj(F(conj Z)) = z[j] = x[2j] + i x[2j+1]
        int s = 0;
        const SegDev &sg = P.seg[s];
        T *out = (T *)sg.base + sg.off + uo * sg.os_u + vo * sg.os_v;
        if (sg.os_d == 1 && (((uintptr_t)out) & (sizeof(C) - 1)) == 0) {
          C *oc = (C *)out;
#pragma unroll
          for (int m = 0; m < E; m++) oc[tS + m * TP] = cconj(v[m]);
        } else {
#pragma unroll
          for (int m = 0; m < E; m++) {
            long long a = (long long)(2 * (tS + m * TP)) * sg.os_d;
            out[a] = v[m].x;
            out[a + sg.os_d] = -v[m].y;
          }
        }
      }
    } else {
      if (liveS) {
#pragma unroll
        for (int m = 0; m < E; m++) {
          C x = bwd ? cconj(v[m]) : v[m];
          store_out<T>(P, tS + m * TP, uo, vo, x);
        }
      }
    }
    __syncthreads();  // shared memory is reused by the next tile
  }
}

// MINB = resident CTAs per SM the register allocation must allow (caps registers at 65536 / (THREADS * MINB))
template <typename T, int M, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) pow2_stage_kernel(const __grid_constant__ StageParams P) {
  pow2_stage_body<T, M, THREADS>(P);
}

// register budget: 128 per thread in double (16 complex values = 64 registers + butterfly temporaries),
// 80 in single -> 512 resp. 768 resident threads per SM whatever the CTA size
template <typename T, int THREADS> struct Pow2MinBlocks {
  enum { BUDGET = sizeof(T) == 8 ? 512 : 768, VALUE = (BUDGET / THREADS) < 1 ? 1 : (BUDGET / THREADS) };
};

// ------------------------------------------------------------------ host side
struct Pow2Plan {
  int M = 0, prec = 0, threads = 0, grid = 0;
  size_t smem = 0;
  int tile_u = 1, tile_v = 1, load_ord = 0, store_ord = 0;
  long long tiles_u = 0, ntiles = 0;
  void (*launch)(const StageParams &, int grid, int threads, size_t smem, cudaStream_t) = nullptr;
  const void *func = nullptr;
};

inline bool pow2_supported(const p3dfftcu_stage_desc &d) {
  int M;
  if (d.kind == P3DFFTCU_K_C2C_FWD || d.kind == P3DFFTCU_K_C2C_BWD) M = d.nfft;
  else if (d.kind == P3DFFTCU_K_R2C || d.kind == P3DFFTCU_K_C2R) {
    if (d.nfft % 2) return false;
    M = d.nfft / 2;
    if (d.kind == P3DFFTCU_K_C2R && d.nseg != 1) return false;  // real output is never exchanged
  } else return false;
  return M == 64 || M == 128 || M == 256 || M == 512 || M == 1024 || M == 2048 || M == 4096;
}

#ifndef P3B_PIPE_TU  // (the pow2_pipe_inst.cu units only need the device code above: skip ~100 kernel instantiations each)
template <typename T, int M, int THREADS> void pow2_launcher(const StageParams &P, int grid, int threads, size_t smem, cudaStream_t s) {
  (void)threads;
  P3B_LAUNCH((pow2_stage_kernel<T, M, THREADS, Pow2MinBlocks<T, THREADS>::VALUE>), grid, THREADS, smem, s, P);
}

template <typename T, int M, int THREADS> int pow2_bind(Pow2Plan *pl, size_t smem_optin, int num_sms) {
  auto kern = pow2_stage_kernel<T, M, THREADS, Pow2MinBlocks<T, THREADS>::VALUE>;
  constexpr int TP = M / Pow2Cfg<M>::E;
  int npb = THREADS / TP;
  pl->threads = THREADS;
  pl->smem = (size_t)npb * Pow2Smem<M>::PENCIL * 2 * sizeof(T);
  if (pl->smem > smem_optin) return -1;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem) != cudaSuccess) return 1;
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, pl->smem) != cudaSuccess) return 1;
  if (occ < 1) occ = 1;
  pl->grid = num_sms * occ;
  pl->launch = pow2_launcher<T, M, THREADS>;
  pl->func = (const void *)kern;
  return 0;
}

// pick THREADS so that a CTA holds `want` pencils (power of two) where possible
template <typename T, int M> int pow2_bind_size(Pow2Plan *pl, int want_pencils, size_t smem_optin, int num_sms) {
  constexpr int TP = M / Pow2Cfg<M>::E;
  int threads = TP * want_pencils;
  if (threads < 64) threads = 64;
  if (threads > 512) threads = 512;
  if (threads < TP) threads = TP;
  int rc = -1;
  // try the wanted size, then smaller ones if shared memory does not fit
  for (; threads >= TP && threads >= 32; threads /= 2) {
    switch (threads) {
      case 32: if (TP <= 32) rc = pow2_bind<T, M, (TP <= 32 ? 32 : TP)>(pl, smem_optin, num_sms); break;
      case 64: if (TP <= 64) rc = pow2_bind<T, M, (TP <= 64 ? 64 : TP)>(pl, smem_optin, num_sms); break;
      case 128: if (TP <= 128) rc = pow2_bind<T, M, (TP <= 128 ? 128 : TP)>(pl, smem_optin, num_sms); break;
      case 256: if (TP <= 256) rc = pow2_bind<T, M, (TP <= 256 ? 256 : TP)>(pl, smem_optin, num_sms); break;
      case 512: if (TP <= 512) rc = pow2_bind<T, M, (TP <= 512 ? 512 : TP)>(pl, smem_optin, num_sms); break;
      default: rc = -1;
    }
    if (rc == 0) return 0;
    if (rc > 0) return rc;
  }
  return -1;
}

template <typename T> int pow2_bind_any(Pow2Plan *pl, int M, int want, size_t smem_optin, int num_sms) {
  switch (M) {
    case 64: return pow2_bind_size<T, 64>(pl, want, smem_optin, num_sms);
    case 128: return pow2_bind_size<T, 128>(pl, want, smem_optin, num_sms);
    case 256: return pow2_bind_size<T, 256>(pl, want, smem_optin, num_sms);
    case 512: return pow2_bind_size<T, 512>(pl, want, smem_optin, num_sms);
    case 1024: return pow2_bind_size<T, 1024>(pl, want, smem_optin, num_sms);
    case 2048: return pow2_bind_size<T, 2048>(pl, want, smem_optin, num_sms);
    case 4096: return pow2_bind_size<T, 4096>(pl, want, smem_optin, num_sms);
  }
  return -1;
}

// returns 0 ok, <0 not applicable (caller falls back to the generic kernel), >0 CUDA error
inline int pow2_setup(const p3dfftcu_stage_desc &d, int num_sms, size_t smem_optin, Pow2Plan *pl, std::string *name) {
  const bool real = d.kind == P3DFFTCU_K_R2C || d.kind == P3DFFTCU_K_C2R;
  const int M = real ? d.nfft / 2 : d.nfft;
  // unit-stride directions on both sides (0 d, 1 u, 2 v)
  auto fast = [](long long sd, long long su, long long sv, long long nd, long long nu, long long nv) {
    long long best = -1;
    int which = 0;
    long long s[3] = {sd, su, sv}, n[3] = {nd, nu, nv};
    for (int i = 0; i < 3; i++) {
      if (n[i] <= 1) continue;
      if (best < 0 || s[i] < best) { best = s[i]; which = i; }
    }
    return which;
  };
  int fin = fast(d.is_d, d.is_u, d.is_v, d.n_in, d.nu, d.nv);
  int fout = fast(d.seg[0].os_d, d.seg[0].os_u, d.seg[0].os_v, d.seg[0].k1 - d.seg[0].k0, d.nu, d.nv);
  const bool needU = fin == 1 || fout == 1, needV = fin == 2 || fout == 2;
  size_t esz = (size_t)d.prec * 2;
  int want = (needU || needV) ? (int)(128 / esz) : 4;  // 128-byte runs across pencils when transposing
  if (needU && needV) want = 16;
  if (const char *e = getenv("P3DFFT_B200_POW2_PENCILS")) {  // tuning override: pencils per CTA iteration
    int w = atoi(e);
    if (w > 0) want = w;
  }
  pl->M = M;
  pl->prec = d.prec;
  int rc = d.prec == 8 ? pow2_bind_any<double>(pl, M, want, smem_optin, num_sms) : pow2_bind_any<float>(pl, M, want, smem_optin, num_sms);
  if (rc) return rc;
  const int E = pow2_values_per_thread(M);
  const int npb = pl->threads / (M / E);
  int tu = 1, tv = 1;
  if (needU && needV) {
    tu = 1;
    while (tu * tu < npb) tu *= 2;
    tv = npb / tu;
  } else if (needV) tv = npb;
  else tu = npb;
  pl->tile_u = tu;
  pl->tile_v = tv;
  pl->load_ord = fin == 0 ? ORD_D : (fin == 1 ? ORD_U : ORD_V);
  pl->store_ord = fout == 0 ? ORD_D : (fout == 1 ? ORD_U : ORD_V);
  pl->tiles_u = (d.nu + tu - 1) / tu;
  pl->ntiles = pl->tiles_u * ((d.nv + tv - 1) / tv);
  if (pl->ntiles < pl->grid) pl->grid = (int)(pl->ntiles > 0 ? pl->ntiles : 1);
  char nm[200];
  snprintf(nm, sizeof nm, "pow2<%s,M=%d> threads=%d tile=%dx%d load=%d store=%d smem=%zu grid=%d", d.prec == 8 ? "f64" : "f32", M,
           pl->threads, tu, tv, pl->load_ord, pl->store_ord, pl->smem, pl->grid);
  *name = nm;
  return 0;
}

inline int pow2_launch(const Pow2Plan &pl, StageParams P, cudaStream_t s) {
  P.tile_u = pl.tile_u;
  P.tile_v = pl.tile_v;
  P.load_ord = pl.load_ord;
  P.store_ord = pl.store_ord;
  P.tiles_u = pl.tiles_u;
  P.ntiles = pl.ntiles;
  if (pl.ntiles == 0) return 0;
  pl.launch(P, pl.grid, pl.threads, pl.smem, s);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
#endif  // P3B_PIPE_TU

}  // namespace p3b
