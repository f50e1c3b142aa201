// pow2_pipe.cuh -- TMA-fed, software-pipelined stage kernel for power-of-two FFT cores (the production fast path).
//
// Same arithmetic as pow2_stage.cuh (register-resident radix-16/8/4/2 Stockham passes, E complex values per
// thread, padded shared-memory exchanges, R2C/C2R through the half-length complex core) but restructured so
// that HBM traffic and arithmetic overlap inside one resident CTA:
//   * a persistent CTA walks its tiles of P pencils; while tile i is being transformed, tile i+1 streams from
//     global memory into the staging buffer S through the TMA unit (cp.async.bulk -> UBLKCP) and signals an
//     mbarrier: no registers, no LSU wavefronts and no per-element address arithmetic are spent on loads.
//       LM_PENCIL  transform dimension is unit-stride: one bulk copy per pencil, S is pencil-major
//       LM_ROWS    pencils are adjacent in memory (transposed input): one bulk copy per row of P elements,
//                  S is row-major with pitch P+1 so that the column reads below are bank-conflict free
//   * at the top of an iteration the tile is read out of S into registers (one LDS per value, applying the
//     kind's pre-processing), S is released and the bulk copies of the next tile are issued at once, so a
//     full tile of loads is in flight during all the passes, exchanges and stores of the current tile;
//   * passes exchange through a second buffer X; when S + X do not fit in 227 KB (1024-point double with 8
//     pencils) two pencils share one X region and take turns; the barriers of an exchange only span the
//     threads that share a region (named barriers), so different pencils of a tile drift apart and their
//     shared-memory and FP64 phases overlap;
//   * stores go straight from registers to global (or to a peer's buffer over NVLink through the segment
//     table) in the OUTPUT's coalescing order: the thread -> (pencil, slot) mapping is chosen for the store side.
// Bulk copies need 16-byte aligned rows; the host selects this kernel only then (pow2_stage.cuh otherwise).
// Replaces reference FFTW execute + reorder_trans + pack_sendbuf_trans (exec.C:737-1326, 2792-2879).
#pragma once
#include "common.cuh"
#include "generic_stage.cuh"
#include "pow2_stage.cuh"

namespace p3b {

enum { LM_PENCIL = 0, LM_ROWS = 1 };

// ------------------------------------------------------------------ mbarrier / bulk-copy wrappers
#ifdef P3B_EMU
// CPU emulation: a bulk copy is a memcpy at issue time; waiting for the tile is a CTA barrier (every thread issues its
// copies before it waits), group barriers are CTA barriers
inline void mbar_init(unsigned long long *, int) {}
inline void mbar_expect_tx(unsigned long long *, unsigned) {}
inline void mbar_wait(unsigned long long *, unsigned) { __syncthreads(); }
inline void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *) { memcpy(dst, src, bytes); }
inline void fence_async_smem() {}
inline void group_bar(int, int) { __syncthreads(); }
#else
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(a), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *bar) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst), b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(gsrc), "r"(bytes),
               "r"(b)
               : "memory");
}
// orders earlier generic-proxy accesses to shared memory before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void group_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
#endif

constexpr size_t kPipeSmemMax = 232448 - 1024;  // 227 KB opt-in minus the static reserve

template <typename T, int M, int KIND, int P, int LM> struct PipeCfg {
  enum { E = Pow2Cfg<M>::E, TP = M / E, THREADS = P * TP, XPITCH = Pow2Smem<M>::PENCIL };
  enum { NIN = KIND == P3DFFTCU_K_C2R ? M + 1 : M };          // complex-sized elements per pencil in S
  static constexpr size_t csz = 2 * sizeof(T);
  // S pitch: every pencil (LM_PENCIL) or row (LM_ROWS) starts 16-byte aligned for the bulk copies and an odd number of
  // 16-byte units after the previous one, which spreads the strided reads of the compute mapping over the banks
  static constexpr int span_units = (int)(((LM == LM_PENCIL ? NIN : P) * csz + 15) / 16);
  enum { SPITCH = (int)(((span_units % 2 == 0 ? span_units + 1 : span_units + 2) * 16) / csz) };
  static constexpr size_t s_bytes = (LM == LM_PENCIL ? (size_t)P * SPITCH : (size_t)NIN * SPITCH) * csz;
  static constexpr size_t x_pencil = (size_t)XPITCH * csz;
  // compact twiddle tables kept in shared memory: with ~200 KB of it in use the L1 that is left cannot hold the
  // global table, and a miss on the critical path of every pass costs an L2 round trip
  // (with little shared memory in use the L1 holds the global table and plain __ldg loads are cheaper: TW_SMEM = 0)
  static constexpr bool TW_SMEM = s_bytes + (P / 2) * x_pencil > 100 * 1024;
  enum { T2N = TW_SMEM ? Pow2Cfg<M>::R1 * Pow2Cfg<M>::R2 : 0, T3N = (TW_SMEM && Pow2Cfg<M>::R3 > 1) ? TP * Pow2Cfg<M>::R3 : 0 };
  static constexpr size_t t_bytes = (size_t)(T2N + T3N) * csz;
  static constexpr bool fits1 = s_bytes + P * x_pencil + t_bytes + 64 <= kPipeSmemMax;
  static constexpr bool fits2 = (P >= 2) && s_bytes + (P / 2) * x_pencil + t_bytes + 64 <= kPipeSmemMax;
  static constexpr bool valid = (THREADS >= 32) && (THREADS <= 1024) && (fits1 || fits2);
  enum { XS = fits1 ? 1 : 2, PX = P / XS };
  static constexpr size_t smem = s_bytes + (size_t)PX * x_pencil + t_bytes + 64;
  // register budget as in pow2_stage.cuh: 128 per thread in double, 80 in single
  enum { BUDGET = sizeof(T) == 8 ? 512 : 768, MINB = (BUDGET / THREADS) < 1 ? 1 : (BUDGET / THREADS) };
  // an exchange region is shared by XS pencils; their threads form a barrier group when they are whole warps
  enum { GROUP = XS * TP, GROUPED = (TP % 32 == 0) && (P / XS <= 15) };
};

// a * exp(-2 pi i m / 16) for a compile-time m (folds to the cheapest form after unrolling)
template <typename T, typename C> __device__ __forceinline__ C mul_w16(C a, int m) {
  const T c1 = (T)0.92387953251128675612818318939679, s1 = (T)0.38268343236508977172845998403040;
  const T h = (T)0.70710678118654752440084436210485;
  switch (m & 15) {
    case 0: return a;
    case 1: return mk<T>(a.x * c1 + a.y * s1, a.y * c1 - a.x * s1);
    case 2: return mk<T>(h * (a.x + a.y), h * (a.y - a.x));
    case 3: return mk<T>(a.x * s1 + a.y * c1, a.y * s1 - a.x * c1);
    case 4: return cmulmi(a);
    case 5: return mk<T>(a.y * c1 - a.x * s1, -a.x * c1 - a.y * s1);
    case 6: return mk<T>(h * (a.y - a.x), -h * (a.x + a.y));
    case 7: return mk<T>(a.y * s1 - a.x * c1, -a.x * s1 - a.y * c1);
    case 8: return cneg(a);
    case 9: return mk<T>(-a.x * c1 - a.y * s1, a.x * s1 - a.y * c1);
    case 10: return mk<T>(-h * (a.x + a.y), h * (a.x - a.y));
    case 11: return mk<T>(-a.x * s1 - a.y * c1, a.x * c1 - a.y * s1);
    case 12: return cmuli(a);
    case 13: return mk<T>(a.x * s1 - a.y * c1, a.x * c1 + a.y * s1);
    case 14: return mk<T>(h * (a.x - a.y), h * (a.x + a.y));
    default: return mk<T>(a.x * c1 - a.y * s1, a.x * s1 + a.y * c1);
  }
}

// second pass (Ns = R1): twiddle of butterfly input q is T2[k][q], k = j mod R1
template <typename T, int M, int E, int R1, int R>
__device__ __forceinline__ void reg_pass2(typename cx<T>::type *v, int t, const typename cx<T>::type *T2) {
  typedef typename cx<T>::type C;
  constexpr int TP = M / E, NB = E / R;
#ifdef P3B_SKELETON
  return;
#endif
#pragma unroll
  for (int b = 0; b < NB; b++) {
    C a[R];
    const C *row = T2 + ((t + b * TP) & (R1 - 1)) * R;
#pragma unroll
    for (int q = 0; q < R; q++) a[q] = v[b + q * NB];
#pragma unroll
    for (int q = 1; q < R; q++) a[q] = cmul(a[q], row[q]);
    Radix<T, R>::run(a);
#pragma unroll
    for (int q = 0; q < R; q++) v[b + q * NB] = a[q];
  }
}

// third pass (Ns * R = M): the twiddle w_M^{q (t + b TP)} factors into T3[t][q] and the 16th root w_16^{q b}
// (TP = M / 16), so one small table serves all NB butterflies of a thread
template <typename T, int M, int E, int R>
__device__ __forceinline__ void reg_pass3(typename cx<T>::type *v, int t, const typename cx<T>::type *T3) {
  typedef typename cx<T>::type C;
  constexpr int NB = E / R;
  static_assert(E == 16, "third pass assumes 16 values per thread");
#ifdef P3B_SKELETON
  return;
#endif
  C w[R];
#pragma unroll
  for (int q = 1; q < R; q++) w[q] = T3[t * R + q];
#pragma unroll
  for (int b = 0; b < NB; b++) {
    C a[R];
#pragma unroll
    for (int q = 0; q < R; q++) a[q] = v[b + q * NB];
#pragma unroll
    for (int q = 1; q < R; q++) a[q] = cmul(mul_w16<T>(a[q], q * b), w[q]);
    Radix<T, R>::run(a);
#pragma unroll
    for (int q = 0; q < R; q++) v[b + q * NB] = a[q];
  }
}

// ------------------------------------------------------------------ the kernel
template <typename T, int M, int KIND, int P, int LM>
__global__ void __launch_bounds__(PipeCfg<T, M, KIND, P, LM>::THREADS, PipeCfg<T, M, KIND, P, LM>::MINB)
pow2_pipe_kernel(const __grid_constant__ StageParams Q) {
  typedef typename cx<T>::type C;
  typedef PipeCfg<T, M, KIND, P, LM> Cfg;
  typedef Pow2Cfg<M> R;
  constexpr int E = R::E, R1 = R::R1, R2 = R::R2, R3 = R::R3;
  constexpr int TP = Cfg::TP, THREADS = Cfg::THREADS, XPITCH = Cfg::XPITCH, SPITCH = Cfg::SPITCH, XS = Cfg::XS, PX = Cfg::PX;
  constexpr int NIN = Cfg::NIN;
  constexpr bool r2c = KIND == P3DFFTCU_K_R2C, c2r = KIND == P3DFFTCU_K_C2R;
  constexpr bool bwd = KIND == P3DFFTCU_K_C2C_BWD || c2r;
  constexpr int twscale = (r2c || c2r) ? 2 : 1;  // the table is exp(-2 pi i j / nfft), nfft = 2M in the real cases
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long *full = reinterpret_cast<unsigned long long *>(smem_raw);  // mbarrier: tile landed in S
  C *S = reinterpret_cast<C *>(smem_raw + 64);
  C *X = reinterpret_cast<C *>(smem_raw + 64 + Cfg::s_bytes);
  C *T2 = X + PX * XPITCH;    // [R1][R2]
  C *T3 = T2 + Cfg::T2N;      // [TP][R3]
  const C *__restrict__ tw = (const C *)Q.tw;
  const int tid = threadIdx.x;
  const int tile_u = Q.tile_u, tu_log2 = Q.tu_log2;
  for (int i = tid; i < Cfg::T2N; i += THREADS) T2[i] = tw[(i % R2) * (i / R2) * (M / (R1 * R2)) * twscale];
  for (int i = tid; i < Cfg::T3N; i += THREADS) T3[i] = tw[(i % R3) * (i / R3) * twscale];

  // compute/store mapping: thread -> (pencil slot in tile, FFT slot t).  slot = pv * tile_u + pu
  int slot, t;
  const bool pencil_major = Q.store_ord == ORD_D;
  if (pencil_major) { slot = tid / TP; t = tid % TP; }
  else { slot = tid % P; t = tid / P; }  // 1-D tiles: the lanes run across the tile's pencils
  const int pu = slot & (tile_u - 1), pv = slot >> tu_log2;
  C *Xp = X + (slot % PX) * XPITCH;
  const int xh = slot / PX;  // turn of this pencil in an exchange region shared by XS pencils
  // barrier scope of an exchange: the threads sharing an X region when they are whole warps, else the CTA
  const bool grouped = Cfg::GROUPED && pencil_major;
  const int bar_id = 1 + (slot % PX);
  auto xsync = [&]() {
    if (grouped) {
      if (Cfg::GROUP <= 32) __syncwarp();
      else group_bar(bar_id, Cfg::GROUP);
    } else __syncthreads();
  };

  if (tid == 0) mbar_init(full, 1);
  __syncthreads();

  // bulk copies of one tile into S (TMA); every thread issues its share, thread 0 announces the byte count
  auto prefetch = [&](long long tile) {
    const long long u0 = (tile % Q.tiles_u) * tile_u, v0 = (tile / Q.tiles_u) * Q.tile_v;
    const int cu = (int)min((long long)tile_u, Q.nu - u0), cv = (int)min((long long)Q.tile_v, Q.nv - v0);
    fence_async_smem();
    if (LM == LM_PENCIL) {
      constexpr unsigned bytes = (unsigned)(NIN * Cfg::csz);  // R2C: 2M reals = M complex-sized elements
      if (tid == 0) mbar_expect_tx(full, bytes * (unsigned)(cu * cv));
      if (tid < P) {
        const int lu = tid & (tile_u - 1), lv = tid >> tu_log2;
        if (lu < cu && lv < cv) {
          const long long base = (u0 + lu) * Q.is_u + (v0 + lv) * Q.is_v;
          const void *src = r2c ? (const void *)((const T *)Q.in + base) : (const void *)((const C *)Q.in + base);
          bulk_g2s(S + tid * SPITCH, src, bytes, full);
        }
      }
    } else {
      // rows of `w` adjacent pencils; the tile is 1-D (tile_u == P or tile_v == P)
      const int w = cu * cv;
      const unsigned bytes = (unsigned)(w * Cfg::csz);
      if (tid == 0) mbar_expect_tx(full, bytes * (unsigned)NIN);
      const C *src = (const C *)Q.in + u0 * Q.is_u + v0 * Q.is_v;
      for (int j = tid; j < NIN; j += THREADS) bulk_g2s(S + j * SPITCH, src + (long long)j * Q.is_d, bytes, full);
    }
  };
  // element j of this thread's pencil in S
  auto s_at = [&](int j) -> C { return LM == LM_PENCIL ? S[slot * SPITCH + j] : S[j * SPITCH + slot]; };

  // one exchange through X: scatter in Stockham order, gather in slot order; pencils sharing a region take turns
  auto exchange = [&](C *v, auto scatter, bool lead_sync) {
#pragma unroll
    for (int h = 0; h < XS; h++) {
      if (h > 0 || lead_sync) xsync();  // the region is free again
      if (XS == 1 || xh == h) scatter(v);
      xsync();
      if (XS == 1 || xh == h) smem_gather<T, M, E>(v, Xp, t);
    }
  };

  long long tile = blockIdx.x;
  unsigned parity = 0;
  if (tile < Q.ntiles) prefetch(tile);
  for (; tile < Q.ntiles; tile += gridDim.x) {
    const long long u0 = (tile % Q.tiles_u) * tile_u, v0 = (tile / Q.tiles_u) * Q.tile_v;
    const long long uo = u0 + pu, vo = v0 + pv;
    const bool live = uo < Q.nu && vo < Q.nv;
    C v[E];
    mbar_wait(full, parity);  // the whole tile has landed in S
    parity ^= 1;
    // ---------------- S -> registers (+ pre-processing)
    if (c2r) {
      // Z[k] = (X[k] + conj X[M-k]) + i e^{+2 pi i k/N} (X[k] - conj X[M-k]); we need conj Z for the conj-trick inverse
#pragma unroll
      for (int m = 0; m < E; m++) {
        const int k = t + m * TP;
        C a = s_at(k);
        C b = cconj(s_at(M - k));
        if (k == 0) { a.y = 0; b.y = 0; }  // FFTW's c2r ignores Im X[0] and Im X[N/2]
        C s = cadd(a, b), d = csub(a, b);
        C w = cconj(__ldg(&tw[k]));
        C e = cmuli(cmul(d, w));
        v[m] = cconj(cadd(s, e));
      }
    } else {
#pragma unroll
      for (int m = 0; m < E; m++) {
        C x = s_at(t + m * TP);
        v[m] = bwd ? cconj(x) : x;
      }
    }
    __syncthreads();  // everyone has read S: release it to the next tile
    if (tile + gridDim.x < Q.ntiles) prefetch(tile + gridDim.x);

    // ---------------- passes
    reg_pass<T, M, E, R1, false>(v, t, 1, tw, twscale);
    exchange(v, [&](C *w) { smem_scatter<T, M, E, R1>(w, Xp, t, 1); }, false);  // X idle since the CTA barrier above
    if constexpr (Cfg::TW_SMEM) reg_pass2<T, M, E, R1, R2>(v, t, T2);
    else reg_pass<T, M, E, R2, true>(v, t, R1, tw, twscale);
    if constexpr (R3 > 1) {
      exchange(v, [&](C *w) { smem_scatter<T, M, E, R2>(w, Xp, t, R1); }, true);
      if constexpr (Cfg::TW_SMEM) reg_pass3<T, M, E, R3>(v, t, T3);
      else reg_pass<T, M, E, R3, true>(v, t, R1 * R2, tw, twscale);
    }
    // v[m] = forward core output F[t + m*TP] of this thread's pencil

    // ---------------- epilogue + stores
    if (r2c) {
      // X[k] = ((Z[k] + conj Z[M-k]) - i e^{-2 pi i k/N} (Z[k] - conj Z[M-k])) / 2, k = 0..M
#pragma unroll
      for (int h = 0; h < XS; h++) {
        xsync();  // the region is free
        if (XS == 1 || xh == h) {
#pragma unroll
          for (int m = 0; m < E; m++) Xp[padidx(t + m * TP)] = v[m];
        }
        xsync();
        if ((XS == 1 || xh == h) && live) {
#pragma unroll
          for (int m = 0; m < E; m++) {
            const int k = t + m * TP;
            const C zk = v[m];
            const C zm = cconj(Xp[padidx((M - k) & (M - 1))]);
            C s = cadd(zk, zm), d = csub(zk, zm);
            C e = cmulmi(cmul(d, __ldg(&tw[k])));
            store_out<T>(Q, k, uo, vo, mk<T>((T)0.5 * (s.x + e.x), (T)0.5 * (s.y + e.y)));
            if (k == 0) store_out<T>(Q, M, uo, vo, mk<T>(zk.x - zk.y, (T)0));
          }
        }
      }
    } else if (c2r) {
      if (live) {  // conj(F(conj Z))[j] = x[2j] + i x[2j+1]; real output is never exchanged: one segment
        const SegDev &sg = Q.seg[0];
        T *out = (T *)sg.base + sg.off + uo * sg.os_u + vo * sg.os_v;
        if (sg.os_d == 1 && (((uintptr_t)out) & (sizeof(C) - 1)) == 0) {
          C *oc = (C *)out;
#pragma unroll
          for (int m = 0; m < E; m++) oc[t + m * TP] = cconj(v[m]);
        } else {
#pragma unroll
          for (int m = 0; m < E; m++) {
            long long a = (long long)(2 * (t + m * TP)) * sg.os_d;
            out[a] = v[m].x;
            out[a + sg.os_d] = -v[m].y;
          }
        }
      }
    } else if (live) {
      if (Q.nseg == 1 && Q.deriv_g <= 0) {  // local stage: one base pointer, constant stride between a thread's stores
        const SegDev &sg = Q.seg[0];
        C *out = (C *)sg.base + sg.off + uo * sg.os_u + vo * sg.os_v + (long long)t * sg.os_d;
        const long long step = (long long)TP * sg.os_d;
#pragma unroll
        for (int m = 0; m < E; m++) out[m * step] = bwd ? cconj(v[m]) : v[m];
      } else {
#pragma unroll
        for (int m = 0; m < E; m++) store_out<T>(Q, t + m * TP, uo, vo, bwd ? cconj(v[m]) : v[m]);
      }
    }
  }
}

// ------------------------------------------------------------------ host side: lookup tables, one per (T, KIND, LM)
struct PipeInfo {
  void (*launch)(const StageParams &, int grid, cudaStream_t);
  const void *func;
  int threads, xs, minb;
  size_t smem;
};

template <typename T, int M, int KIND, int P, int LM> void pipe_launcher(const StageParams &Q, int grid, cudaStream_t s) {
  typedef PipeCfg<T, M, KIND, P, LM> Cfg;
  P3B_LAUNCH((pow2_pipe_kernel<T, M, KIND, P, LM>), grid, Cfg::THREADS, Cfg::smem, s, Q);
}

template <typename T, int M, int KIND, int P, int LM> const PipeInfo *pipe_info_one() {
  typedef PipeCfg<T, M, KIND, P, LM> Cfg;
  if constexpr (!Cfg::valid) {
    return nullptr;
  } else {
    static const PipeInfo info = {pipe_launcher<T, M, KIND, P, LM>, (const void *)pow2_pipe_kernel<T, M, KIND, P, LM>, Cfg::THREADS,
                                  Cfg::XS, Cfg::MINB, Cfg::smem};
    return &info;
  }
}

template <typename T, int M, int KIND, int LM> const PipeInfo *pipe_info_m(int P) {
  switch (P) {
    case 1: return pipe_info_one<T, M, KIND, 1, LM>();
    case 2: return pipe_info_one<T, M, KIND, 2, LM>();
    case 4: return pipe_info_one<T, M, KIND, 4, LM>();
    case 8: return pipe_info_one<T, M, KIND, 8, LM>();
    case 16: return pipe_info_one<T, M, KIND, 16, LM>();
  }
  return nullptr;
}

template <typename T, int KIND, int LM> const PipeInfo *pipe_info(int M, int P) {
  switch (M) {
    case 64: return pipe_info_m<T, 64, KIND, LM>(P);
    case 128: return pipe_info_m<T, 128, KIND, LM>(P);
    case 256: return pipe_info_m<T, 256, KIND, LM>(P);
    case 512: return pipe_info_m<T, 512, KIND, LM>(P);
    case 1024: return pipe_info_m<T, 1024, KIND, LM>(P);
    case 2048: return pipe_info_m<T, 2048, KIND, LM>(P);
    case 4096: return pipe_info_m<T, 4096, KIND, LM>(P);
  }
  return nullptr;
}

// defined in pow2_pipe_inst.cu, compiled once per (precision, kind) so that the instantiations build in parallel
const PipeInfo *pipe_lookup(int prec, int kind, int lm, int M, int P);

}  // namespace p3b
