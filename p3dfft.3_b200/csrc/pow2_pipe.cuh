// pow2_pipe.cuh -- TMA-fed stage kernel for power-of-two FFT cores with unit-stride input (the production fast path).
//
// Same arithmetic as pow2_stage.cuh (register-resident radix-16/8/4/2 Stockham passes, E complex values per
// thread, R2C/C2R through the half-length complex core).  Data movement:
//   * one shared-memory buffer per pencil, used IN PLACE: the pencil lands there by one bulk copy (cp.async.bulk ->
//     UBLKCP, completion on the pencil's own mbarrier: no registers, LSU wavefronts or address arithmetic are spent
//     on loads), is read into registers, and the same buffer then carries the padded Stockham exchanges (a pencil is
//     entirely in registers between passes, so nothing is lost by overwriting it);
//   * as soon as the last exchange of a tile has been read back, the pencil's leader thread issues the bulk copy of
//     the NEXT tile's pencil into the buffer, so the load is in flight during the last pass, the epilogue and all the
//     global stores of the current tile;
//   * contiguous output (TS = 0): the TP threads of a pencil form an independent group -- group-scoped named barriers
//     only, groups of one CTA drift apart, so their load, shared-memory, FP64 and store phases overlap;
//   * transposed output (TS = 1): the last exchange re-maps threads from (pencil-major) to (lanes across the tile's P
//     pencils), so the stores -- straight from registers to global memory, or to a peer's buffer over NVLink through
//     the segment table -- are runs of P elements along the OUTPUT's unit-stride dimension; only that exchange and
//     what follows it synchronise the whole CTA;
//   * twiddles of passes 2 and 3 come from compact shared-memory tables laid out [q][k] so that a warp's reads are
//     contiguous.
// Bulk copies need 16-byte aligned pencils; the host selects this kernel only then (pow2_stage.cuh otherwise, which
// also serves inputs whose unit-stride direction is not the transform dimension).
// Replaces reference FFTW execute + reorder_trans + pack_sendbuf_trans (exec.C:737-1326, 2792-2879).
#pragma once
#include "common.cuh"
#include "generic_stage.cuh"
#include "pow2_stage.cuh"

namespace p3b {

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

// ------------------------------------------------------------------ flags between kernels (SYNC = 1 variants)
#ifdef P3B_EMU
inline unsigned long long flag_ld(const unsigned long long *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void flag_st(unsigned long long *p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline unsigned long long ctr_add(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
inline void fence_sys() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void fence_proxy_async_all() {}
inline unsigned long long now_ns() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}
inline void spin_pause() {
  struct timespec ts = {0, 20000};
  nanosleep(&ts, nullptr);
}
inline void sync_trap(const char *what) {
  fprintf(stderr, "p3dfft_b200 (emulation): %s\n", what);
  abort();
}
#else
__device__ __forceinline__ unsigned long long flag_ld(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void flag_st(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ctr_add(unsigned long long *p, unsigned long long v) { return atomicAdd(p, v); }
__device__ __forceinline__ void fence_sys() { __threadfence_system(); }
// data written through the generic proxy (by other SMs or other GPUs) is about to be read by bulk copies (async proxy)
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void spin_pause() { __nanosleep(100); }
__device__ __forceinline__ void sync_trap(const char *what) {
  printf("p3dfft_b200: %s\n", what);
  __trap();
}
#endif

// have all wait sources published flag `id` for this exec?  blocking: spin until they have (trap after Y.timeout_ns if set)
__device__ __forceinline__ bool flags_ready(const SyncDev &Y, int id, bool blocking) {
  for (int j = 0; j < Y.wait_n; j++) {
    const unsigned long long *w = Y.wait_base + Y.wait_off[j] + id;
    const unsigned long long want = Y.wait_epoch[j];
    if (flag_ld(w) >= want) continue;
    if (!blocking) return false;
    const unsigned long long t0 = now_ns();
    while (flag_ld(w) < want) {
      spin_pause();
      if (Y.timeout_ns && now_ns() - t0 > Y.timeout_ns) sync_trap("timed out waiting for a tile-group flag (a peer rank died or never ran this exec)");
    }
  }
  fence_proxy_async_all();
  return true;
}

// `add` pencils of group g are complete (the caller has synchronised the threads that stored them): count them, and publish
// the group's flag to every target when it was the last contribution
template <int P> __device__ __forceinline__ void group_done(const SyncDev &Y, int g, unsigned long long add) {
  const unsigned long long total = (unsigned long long)Y.grp[g].tiles_u * Y.grp[g].tiles_v * P;
  fence_sys();
  const unsigned long long prev = ctr_add(Y.ctl + 1 + g, add);
  const int id = Y.grp[g].signal_id;
  if (prev + add == total && id >= 0) {
    // groups handed out earlier may still have a tile in flight on another CTA: wait for their counters
    for (int h = 0; h < Y.grp[g].after; h++) {
      const unsigned long long th = (unsigned long long)Y.grp[h].tiles_u * Y.grp[h].tiles_v * P;
      while (flag_ld(Y.ctl + 1 + h) < th) spin_pause();
    }
    fence_sys();
    for (int j = 0; j < Y.sig_n; j++) flag_st(Y.sig_ptr[j] + id, Y.sig_epoch[j]);
  }
}

// store of data that is not read again before it leaves the L2 (every stage output is >> the 126 MB L2).
// -DP3B_STREAM_STORES selects st.global.cs (evict-first); measured on the 1024^3 round trip it makes no difference
// (18.38 / 18.33 ms against 18.43 / 18.14 ms, alternating runs on one box), so the default stays a plain store
template <typename C> __device__ __forceinline__ void st_out(C *p, C v) {
#if defined(P3B_STREAM_STORES) && !defined(P3B_EMU)
  __stcs(p, v);
#else
  *p = v;
#endif
}

constexpr size_t kPipeSmemMax = 232448 - 1024;  // 227 KB opt-in minus the static reserve

// KIND of the kernel template: the P3DFFTCU_K_* value for C2C / R2C / C2R; kPipeR2R stands for EVERY r2r kind (DCT/DST I-IV,
// real or complex data) whose internal FFT length L = M is a power of two -- which one is read from StageParams::kind at run
// time (a CTA-uniform switch around the load and store loops; the M-point core in between is the same for all of them).
// kPipeDCT1 is the compile-time form of the one kind the BASELINE configurations use (DCT-I on complex data, the Chebyshev
// direction of config 4): every index of the even extension and every live output is known at compile time, so the kernel
// executes the C2C instruction stream (the runtime form needs three times as many instructions per pencil) and the unused
// half of the last pass is dropped
constexpr int kPipeR2R = 13;
constexpr int kPipeDCT1 = P3DFFTCU_K_DCT1;

template <typename T, int M, int KIND, int P, int TS> struct PipeCfg {
  enum { E = Pow2Cfg<M>::E, TP = M / E, THREADS = P * TP, XP0 = Pow2Smem<M>::PENCIL };
  enum { R1 = Pow2Cfg<M>::R1, R2 = Pow2Cfg<M>::R2, R3 = Pow2Cfg<M>::R3 };
  // complex-sized elements of one pencil landing in shared memory; a bulk copy moves multiples of 16 bytes, so the M+1
  // single-precision elements of a C2R pencil travel with the 8 bytes that follow them (the host checks they exist)
  enum { NIN = KIND == P3DFFTCU_K_C2R ? (sizeof(T) == 4 ? M + 2 : M + 1) : M };
  static constexpr size_t csz = 2 * sizeof(T);
  // pencil pitch (complex elements): every pencil starts 16-byte aligned (bulk-copy destination) an odd number of
  // 16-byte units after the previous one, so the lanes-across-pencils accesses of the transposed mapping spread over
  // the banks; XP0 = M + M/16 + 1 is odd
  enum { PITCH = sizeof(T) == 8 ? XP0 : ((XP0 + 1) % 4 == 2 ? XP0 + 1 : XP0 + 3) };
  enum { T2N = R1 * R2, T3N = R3 > 1 ? R3 * TP : 0 };
  static constexpr size_t bar_bytes = 384;  // P mbarriers (<= 128 bytes), then 2 x 16 words for the work items handed out dynamically
  static constexpr size_t smem = bar_bytes + ((size_t)P * PITCH + T2N + T3N) * csz;
  static constexpr bool valid = (THREADS >= 32) && (THREADS <= 1024) && (P <= 16) && (smem <= kPipeSmemMax) && (TS ? P >= 2 : true) &&
                                (TP > 32 ? P <= 15 : true);
  // register budget as in pow2_stage.cuh: 128 per thread in double, 80 in single
  // register budget as threads per SM: 128 registers per thread in double (16 complex values + temporaries), 80 in single;
  // twice as many threads when a thread holds 8 values
  enum { BUDGET = (sizeof(T) == 8 ? 512 : 768) * (E == 8 && M >= 512 ? 2 : 1), MINB = (BUDGET / THREADS) < 1 ? 1 : (BUDGET / THREADS) };
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

// second pass (Ns = R1): the twiddle of butterfly input q is T2[q][k], k = j mod R1
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
    const C *col = T2 + ((t + b * TP) & (R1 - 1));
#pragma unroll
    for (int q = 0; q < R; q++) a[q] = v[b + q * NB];
#pragma unroll
    for (int q = 1; q < R; q++) a[q] = cmul(a[q], col[q * R1]);
    Radix<T, R>::run(a);
#pragma unroll
    for (int q = 0; q < R; q++) v[b + q * NB] = a[q];
  }
}

// third pass (Ns * R = M): the twiddle w_M^{q (t + b TP)} factors into T3[q][t] and the 16th root w_16^{q b}
// (TP = M / 16), so one small table serves all NB butterflies of a thread
template <typename T, int M, int E, int R>
__device__ __forceinline__ void reg_pass3(typename cx<T>::type *v, int t, const typename cx<T>::type *T3) {
  typedef typename cx<T>::type C;
  constexpr int TP = M / E, NB = E / R;
  static_assert(E == 16 || NB == 1, "the 16th-root factor below assumes 16 values per thread (or a single butterfly)");
#ifdef P3B_SKELETON
  return;
#endif
  C w[R];
#pragma unroll
  for (int q = 1; q < R; q++) w[q] = T3[q * TP + t];
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

// wt * exp(-2 pi i m / (2E)) for a compile-time m < E (E = 16: 32nd roots; E = 8: 16th roots)
template <typename T, int E> __device__ __forceinline__ typename cx<T>::type real_twiddle(typename cx<T>::type wt, int m) {
  // cos/sin(2 pi m / 32), m = 0..15
  const double c32[16] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708, 0.70710678118654752440,
                          0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785, 0.0, -0.19509032201612826785,
                          -0.38268343236508977173, -0.55557023301960222474, -0.70710678118654752440, -0.83146961230254523708,
                          -0.92387953251128675613, -0.98078528040323044913};
  const double s32[16] = {0.0, 0.19509032201612826785, 0.38268343236508977173, 0.55557023301960222474, 0.70710678118654752440,
                          0.83146961230254523708, 0.92387953251128675613, 0.98078528040323044913, 1.0, 0.98078528040323044913,
                          0.92387953251128675613, 0.83146961230254523708, 0.70710678118654752440, 0.55557023301960222474,
                          0.38268343236508977173, 0.19509032201612826785};
  const int j = E == 16 ? m : 2 * m;
  if (j == 0) return wt;
  const T c = (T)c32[j], s = (T)s32[j];  // multiply by (c - i s)
  return mk<T>(wt.x * c + wt.y * s, wt.y * c - wt.x * s);
}

// R2C split: X[k] = ((Z[k] + conj Z[M-k]) - i w (Z[k] - conj Z[M-k])) / 2 with w = e^{-2 pi i k/N}
template <typename T, typename C> __device__ __forceinline__ C r2c_split(C zk, C zpartner, C w) {
  const C zm = cconj(zpartner);
  const C s = cadd(zk, zm), d = csub(zk, zm);
  const C e = cmulmi(cmul(d, w));
  return mk<T>((T)0.5 * (s.x + e.x), (T)0.5 * (s.y + e.y));
}


// ------------------------------------------------------------------ r2r kinds on the M-point core (KIND = kPipeR2R)
// Every r2r kind is a complex-linear map: symmetric extension z of the n inputs to L = M points (with a half-sample
// pre-twiddle for kinds III / IV), FFT_M, post-twiddle of the first n outputs (generic_stage.cuh has the same table, FFTW
// definitions init.C:1191-1607) -- so the same code serves real data and the `_COMPLEX` variants (re and im separately,
// init.C:1179-1188).  The pencil has landed in shared memory as it lies in global memory (n real or complex values);
// v[m] = z[t + m TP] is built straight from it.
template <typename T, int M, int E>
__device__ __forceinline__ void r2r_load(typename cx<T>::type *v, const typename cx<T>::type *BA, int t, const StageParams &Q) {
  typedef typename cx<T>::type C;
  constexpr int TP = M / E;
  const int n = Q.nfft;
  const bool cplx = Q.dt_in == 2;
  const T *BR = reinterpret_cast<const T *>(BA);
  const C *__restrict__ tw2 = (const C *)Q.tw2;
  const C zero = mk<T>((T)0, (T)0);
  auto X = [&](int j) -> C { return cplx ? BA[j] : mk<T>(BR[j], (T)0); };
  switch (Q.kind) {
    case P3DFFTCU_K_DCT1:  // even extension, M = 2(n-1)
#pragma unroll
      for (int m = 0; m < E; m++) {
        const int i = t + m * TP;
        v[m] = X(i < n ? i : M - i);
      }
      break;
    case P3DFFTCU_K_DST1:  // odd extension, M = 2(n+1): z_0 = z_{n+1} = 0
#pragma unroll
      for (int m = 0; m < E; m++) {
        const int i = t + m * TP;
        if (i == 0 || i == n + 1) v[m] = zero;
        else v[m] = i <= n ? X(i - 1) : cneg(X(M - 1 - i));
      }
      break;
    case P3DFFTCU_K_DCT2:  // half-sample even extension, M = 2n
#pragma unroll
      for (int m = 0; m < E; m++) {
        const int i = t + m * TP;
        v[m] = X(i < n ? i : M - 1 - i);
      }
      break;
    case P3DFFTCU_K_DST2:  // half-sample odd extension
#pragma unroll
      for (int m = 0; m < E; m++) {
        const int i = t + m * TP;
        v[m] = i < n ? X(i) : cneg(X(M - 1 - i));
      }
      break;
    case P3DFFTCU_K_DCT3:  // z_i = x_i w_i (i < n), 0 (i = n), -x_{2n-i} w_i (i > n); w = tw2
#pragma unroll
      for (int m = 0; m < E; m++) {
        const int i = t + m * TP;
        if (i == n) v[m] = zero;
        else {
          const C x = cmul(X(i < n ? i : M - i), __ldg(&tw2[i]));
          v[m] = i < n ? x : cneg(x);
        }
      }
      break;
    case P3DFFTCU_K_DST3:  // z_0 = 0, z_i = i x_{i-1} w_i (i <= n), i x_{2n-i-1} w_i (i > n)
#pragma unroll
      for (int m = 0; m < E; m++) {
        const int i = t + m * TP;
        if (i == 0) v[m] = zero;
        else v[m] = cmuli(cmul(X(i <= n ? i - 1 : M - i - 1), __ldg(&tw2[i])));
      }
      break;
    case P3DFFTCU_K_DCT4:  // z_i = x_i w_i (i < n), -x_{2n-1-i} w_i (i >= n)
#pragma unroll
      for (int m = 0; m < E; m++) {
        const int i = t + m * TP;
        const C x = cmul(X(i < n ? i : M - 1 - i), __ldg(&tw2[i]));
        v[m] = i < n ? x : cneg(x);
      }
      break;
    default:  // P3DFFTCU_K_DST4: z_i = x_i w_i (i < n), +x_{2n-1-i} w_i (i >= n)
#pragma unroll
      for (int m = 0; m < E; m++) {
        const int i = t + m * TP;
        v[m] = cmul(X(i < n ? i : M - 1 - i), __ldg(&tw2[i]));
      }
      break;
  }
}

// output index and value of core output F[idx] for the r2r kind (k outside [0, n): not an output)
template <typename T>
__device__ __forceinline__ int r2r_post(const StageParams &Q, int idx, typename cx<T>::type &y) {
  typedef typename cx<T>::type C;
  const C *__restrict__ tw2 = (const C *)Q.tw2;
  const C *__restrict__ tw3 = (const C *)Q.tw3;
  const int n = Q.n_out;
  switch (Q.kind) {
    case P3DFFTCU_K_DST1: y = cmuli(y); return idx - 1;
    case P3DFFTCU_K_DCT2: if (idx < n) y = cmul(y, __ldg(&tw2[idx])); return idx;
    case P3DFFTCU_K_DST2: if (idx >= 1 && idx <= n) y = cmuli(cmul(y, __ldg(&tw2[idx]))); return idx - 1;
    case P3DFFTCU_K_DCT4: if (idx < n) y = cmul(y, __ldg(&tw3[idx])); return idx;
    case P3DFFTCU_K_DST4: if (idx < n) y = cmuli(cmul(y, __ldg(&tw3[idx]))); return idx;
    default: return idx;  // DCT1, DCT3, DST3
  }
}

// ------------------------------------------------------------------ the kernel
// SY = 1: tile groups with wait / signal flags (SyncDev, common.cuh) -- the persistent kernels of an overlapped pair
template <typename T, int M, int KIND, int P, int TS, int SY>
__device__ __forceinline__ void pow2_pipe_body(const StageParams &Q, const SyncDev *Yp) {
  typedef typename cx<T>::type C;
  typedef PipeCfg<T, M, KIND, P, TS> Cfg;
  constexpr int E = Cfg::E, R1 = Cfg::R1, R2 = Cfg::R2, R3 = Cfg::R3;
  constexpr int TP = Cfg::TP, THREADS = Cfg::THREADS, PITCH = Cfg::PITCH;
  constexpr bool r2c = KIND == P3DFFTCU_K_R2C, c2r = KIND == P3DFFTCU_K_C2R;
  constexpr bool bwd = KIND == P3DFFTCU_K_C2C_BWD || c2r;
  constexpr int twscale = (r2c || c2r) ? 2 : 1;  // the table is exp(-2 pi i j / nfft), nfft = 2M in the real cases
  constexpr bool r2r = KIND == kPipeR2R, dct1 = KIND == kPipeDCT1;
  // one pencil (R2C: 2M reals = M complex-sized elements; r2r kinds: n real or complex values, rounded up to 16 bytes by the host)
  const unsigned bytes = (r2r || dct1) ? (unsigned)Q.pipe_bytes : (unsigned)(Cfg::NIN * Cfg::csz);
  // R2C with a third pass of radix 2, 4 or 8 (M = 512, 1024, 2048: the 1024-, 2048- and 4096-point real transforms): the
  // last pass works on NS = M/R3 columns; the thread that owns the butterflies of column j = t + TP b (b < NB/2) also takes
  // those of the mirror column NS - j, so Z[k] = Z[j + NS q] and its Hermitian partner Z[M-k] = Z[(NS-j) + NS (R3-1-q)] both
  // come out of its own registers: no exchange pass for the split (6 instead of 8 shared-memory passes, two barriers
  // fewer), and the next tile's bulk copy is issued before the last pass instead of after the split
  constexpr bool r2c_sym = r2c && E == 16 && (R3 == 2 || R3 == 4 || R3 == 8);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem_raw);  // one mbarrier per pencil: "landed"
  long long *snext = reinterpret_cast<long long *>(smem_raw + 128);               // SY: tile numbers from the dynamic counter
  C *B = reinterpret_cast<C *>(smem_raw + Cfg::bar_bytes);
  C *T2 = B + P * PITCH;   // [R2][R1]
  C *T3 = T2 + Cfg::T2N;   // [R3][TP]
  const C *__restrict__ tw = (const C *)Q.tw;
  const int tid = threadIdx.x;
  const int tile_u = Q.tile_u, tu_log2 = Q.tu_log2;
  for (int i = tid; i < Cfg::T2N; i += THREADS) T2[i] = tw[(i / R1) * (i % R1) * (M / (R1 * R2)) * twscale];
  for (int i = tid; i < Cfg::T3N; i += THREADS) T3[i] = tw[(i / TP) * (i % TP) * twscale];

  // mapping A (load side and all passes but the last): pencil-major.  mapping B (last pass and stores): the same when
  // the output is contiguous along the transform dimension, else the lanes run across the tile's pencils
  const int slotA = tid / TP, tA = tid % TP;
  const int slotB = TS ? tid % P : slotA, tB = TS ? tid / P : tA;
  int puA = slotA & (tile_u - 1), pvA = slotA >> tu_log2;  // (re-derived per work item when pencils are handed out singly)
  int puB = slotB & (tile_u - 1), pvB = slotB >> tu_log2;
  C *BA = B + slotA * PITCH, *BB = B + slotB * PITCH;
  unsigned long long *bar = bars + slotA;
  auto syncA = [&]() {  // the threads of one pencil (whole warps, or a fraction of one warp)
    if (TP <= 32) __syncwarp();
    else group_bar(1 + slotA, TP);
  };
  auto syncB = [&]() {
    if (TS) __syncthreads();
    else syncA();
  };
  // the pencil's leader thread starts the bulk copy of its pencil of tile `tl`; a pencil outside the array completes
  // the phase with zero bytes so that the waits stay uniform
  // consecutive tile numbers (= concurrently running CTAs) run along the dimension that is closer to unit stride in the
  // OUTPUT, so that the stores of neighbouring CTAs fill whole DRAM pages together
  const bool vfast = Q.vfast != 0;
  // SY: the tiles are numbered group by group (TileGroupDev); g = group of tile tl
  auto tile_origin = [&](long long tl, int g, long long &u0, long long &v0) {
    if constexpr (SY) {
      const TileGroupDev &G = Yp->grp[g];
      const long long l = tl - G.tile0;
      const long long iu = vfast ? l / G.tiles_v : l % G.tiles_u, iv = vfast ? l % G.tiles_v : l / G.tiles_u;
      u0 = G.u0 + iu * tile_u;
      v0 = G.v0 + iv * Q.tile_v;
    } else {
      const long long iu = vfast ? tl / Q.tiles_v : tl % Q.tiles_u, iv = vfast ? tl % Q.tiles_v : tl / Q.tiles_u;
      u0 = iu * tile_u;
      v0 = iv * Q.tile_v;
    }
  };
  auto group_of = [&](long long tl, int g) {
    if constexpr (SY)
      while (g + 1 < Yp->ngroups && tl >= Yp->grp[g + 1].tile0) g++;
    return g;
  };
  int gI = 0, gReady = -1;  // leaders: group of the tile being issued; groups <= gReady have had their flag seen
  long long pend = -1;      // leaders: tile whose load waits for its group's flag (issued, blocking, at the top of the loop)
  // SY, contiguous output, dynamic: the unit of work is ONE pencil (a pencil's thread group never synchronises with the others):
  // work item w = pencil w % P of tile w / P
  bool dynamic = false, dynp = false;
  if constexpr (SY) {
    dynamic = Yp->dynamic != 0;
    dynp = dynamic && !TS;
  }
  const long long nwork = dynp ? Q.ntiles * P : Q.ntiles;
  auto issue = [&](long long wl, bool blocking) {
    if (tA == 0 && wl < nwork) {
      long long tl = wl;
      if constexpr (SY && !TS) {
        if (dynp) {
          tl = wl / P;
          const int ps = (int)(wl % P);
          puA = ps & (tile_u - 1);
          pvA = ps >> tu_log2;
        }
      }
      if constexpr (SY) {
        gI = group_of(tl, gI);
        if (gI > gReady) {
          const int wid = Yp->grp[gI].wait_id;
          if (wid >= 0 && !flags_ready(*Yp, wid, blocking)) {
            pend = wl;
            return;
          }
          gReady = gI;
        }
      }
      long long u, v;
      tile_origin(tl, gI, u, v);
      u += puA;
      v += pvA;
      const bool live = SY ? (u < Yp->grp[gI].u1 && v < Yp->grp[gI].v1) : (u < Q.nu && v < Q.nv);
      fence_async_smem();
      mbar_expect_tx(bar, live ? bytes : 0u);
      if (live) {
        const long long base = u * Q.is_u + v * Q.is_v;
        const bool real_in = r2c || (r2r && Q.dt_in == 1);  // strides count elements of the input's own type
        const void *src = real_in ? (const void *)((const T *)Q.in + base) : (const void *)((const C *)Q.in + base);
        bulk_g2s(BA, src, bytes, bar);
      }
    }
  };

  // real transforms: e^{-2 pi i k/N} for k = t + m TP factors into tw[t] (per thread, loop-invariant) and the
  // compile-time 32nd root e^{-2 pi i m/32} (TP = M/16 = N/32)
  C wt = mk<T>((T)1, (T)0);
  if constexpr (r2c) wt = tw[tB];
  if constexpr (c2r) wt = tw[tA];

  if (tid < P) mbar_init(bars + tid, 1);
  // next work item from the counter; CTAs beyond keep_ctas stop once the counter has passed boost_limit (checked BEFORE
  // taking an item: whatever was taken is processed)
  auto grab = [&]() -> long long {
    if constexpr (SY) {
      if ((int)blockIdx.x >= Yp->keep_ctas && flag_ld(Yp->ctl) >= Yp->boost_limit) return (long long)1 << 60;
      return (long long)ctr_add(Yp->ctl, 1ull);
    }
    return 0;
  };
  // the slot a handed-out item travels through: one per CTA (TS) or one per pencil's thread group, double-buffered
  long long *sn = snext + (TS ? 0 : 2 * slotA);
  if constexpr (SY)
    if (dynamic && (TS ? tid == 0 : tA == 0)) sn[0] = grab();
  __syncthreads();

  long long work = blockIdx.x;  // tile number, or pencil number when pencils are handed out singly
  if constexpr (SY)
    if (dynamic) work = sn[0];
  unsigned parity = 0;
  int gS = 0, it = 0;            // SY: group of the tile being processed; iteration count
  unsigned long long cnt = 0;    // SY: pencils of group gS this thread has to account for
  issue(work, true);
  while (work < nwork) {
    long long nxt = work + gridDim.x;
    long long tile = work;
    if constexpr (SY && !TS) {
      if (dynp) {
        tile = work / P;
        const int ps = (int)(work % P);
        puA = puB = ps & (tile_u - 1);
        pvA = pvB = ps >> tu_log2;
      }
    }
    if constexpr (SY) {
      gS = group_of(tile, gS);
      // read back after the next barrier of the CTA (TS) / of the pencil's thread group
      if (dynamic && (TS ? tid == 0 : tA == 0)) sn[(it + 1) & 1] = grab();
      if (tA == 0 && pend >= 0) {  // the prefetch found the group's flag not yet set: wait for it now
        const long long t = pend;
        pend = -1;
        issue(t, true);
      }
    }
    long long uo, vo;
    tile_origin(tile, gS, uo, vo);
    uo += puB;
    vo += pvB;
    const bool live = SY ? (uo < Yp->grp[gS].u1 && vo < Yp->grp[gS].v1) : (uo < Q.nu && vo < Q.nv);
    C v[E];
    mbar_wait(bar, parity);  // this thread's pencil (mapping A) has landed
    parity ^= 1;
    // ---------------- shared memory -> registers (+ pre-processing)
    if (c2r) {
      // Z[k] = (X[k] + conj X[M-k]) + i e^{+2 pi i k/N} (X[k] - conj X[M-k]); we need conj Z for the conj-trick inverse
#pragma unroll
      for (int m = 0; m < E; m++) {
        const int k = tA + m * TP;
        C a = BA[k];
        C b = cconj(BA[M - k]);
        if (k == 0) { a.y = 0; b.y = 0; }  // FFTW's c2r ignores Im X[0] and Im X[N/2]
        C s = cadd(a, b), d = csub(a, b);
        C w = cconj(real_twiddle<T, E>(wt, m));
        C e = cmuli(cmul(d, w));
        v[m] = cconj(cadd(s, e));
      }
    } else if constexpr (r2r) {
      r2r_load<T, M, E>(v, BA, tA, Q);
    } else if constexpr (dct1) {
      // even extension of the n = M/2 + 1 complex inputs: z_i = x_i (i <= M/2), x_{M-i} (i > M/2); i = tA + m TP, TP = M/E
#pragma unroll
      for (int m = 0; m < E; m++) {
        const int i = tA + m * TP;
        if (m < E / 2) v[m] = BA[i];
        else if (m == E / 2) v[m] = BA[tA == 0 ? M / 2 : M - i];
        else v[m] = BA[M - i];
      }
    } else {
#pragma unroll
      for (int m = 0; m < E; m++) {
        C x = BA[tA + m * TP];
        v[m] = bwd ? cconj(x) : x;
      }
    }
    syncA();  // the pencil is in registers: its buffer now carries the exchanges

    // ---------------- passes
    reg_pass<T, M, E, R1, false>(v, tA, 1, tw, twscale);
    smem_scatter<T, M, E, R1>(v, BA, tA, 1);
    if constexpr (R3 > 1) {
      syncA();
      smem_gather<T, M, E>(v, BA, tA);
      reg_pass2<T, M, E, R1, R2>(v, tA, T2);
      syncA();
      smem_scatter<T, M, E, R2>(v, BA, tA, R1);
    }
    if constexpr (r2c_sym) {
      constexpr int NS = M / R3, NB = E / R3, HB = NB / 2;  // columns of the last pass, butterflies per thread, pairs per thread
      syncB();
      // gather: for pair b < HB, column A = j = t + TP b and its mirror B = NS - j (thread 0: column 0 pairs with NS/2);
      // v[(2 b + side) R3 + q] = input q of that column
      const bool col0 = tB == 0;
#pragma unroll
      for (int b = 0; b < HB; b++) {
        const int j = tB + TP * b;
        const int jp = (col0 && b == 0) ? NS / 2 : NS - j;
#pragma unroll
        for (int q = 0; q < R3; q++) {
          v[(2 * b) * R3 + q] = BB[padidx(j + NS * q)];
          v[(2 * b + 1) * R3 + q] = BB[padidx(jp + NS * q)];
        }
      }
      syncB();  // every value is in registers: the buffers are free for the next tile
      if constexpr (SY) if (dynamic) nxt = sn[(it + 1) & 1];
      issue(nxt, false);
      C w3[R3];  // w_M^{q t}
#pragma unroll
      for (int q = 1; q < R3; q++) w3[q] = T3[q * TP + tB];
      const bool seg1 = Q.nseg == 1 && Q.deriv_g <= 0;
      const SegDev &sg = Q.seg[0];
      C *out = (C *)sg.base + sg.off + uo * sg.os_u + vo * sg.os_v;
      auto put = [&](int k, C x) {
        if (!live) return;
        if (seg1) st_out(out + (long long)k * sg.os_d, x);
        else store_out<T>(Q, k, uo, vo, x);
      };
      const C one = mk<T>((T)1, (T)0);
#pragma unroll
      for (int b = 0; b < HB; b++) {
        const int j = tB + TP * b;
        const bool special = col0 && b == 0;
        C za[R3], zb[R3];
#pragma unroll
        for (int q = 0; q < R3; q++) {
          za[q] = v[(2 * b) * R3 + q];
          zb[q] = v[(2 * b + 1) * R3 + q];
        }
        // last-pass twiddles: column j: w_M^{q j} = w_M^{q t} w_16^{q b}; mirror NS - j: w_R^q conj(w_M^{q j}); column NS/2: w_{2R}^q
#pragma unroll
        for (int q = 1; q < R3; q++) {
          const C wq = mul_w16<T>(w3[q], q * b);
          za[q] = cmul(za[q], wq);
          if (special) zb[q] = mul_w16<T>(zb[q], q * (8 / R3));
          else zb[q] = mul_w16<T>(cmul(zb[q], cconj(wq)), q * (16 / R3));
        }
        Radix<T, R3>::run(za);  // Z[j + NS q]
        Radix<T, R3>::run(zb);  // Z[j' + NS q]
        // split twiddles e^{-2 pi i k/N}, k = j + NS q: wj w_{2R}^q; the partner row M - k takes -conj of it
        const C wj = real_twiddle<T, E>(wt, b);
        if (special) {
          put(0, mk<T>(za[0].x + za[0].y, (T)0));
          put(M, mk<T>(za[0].x - za[0].y, (T)0));
#pragma unroll
          for (int q = 1; q < R3; q++) put(NS * q, r2c_split<T>(za[q], za[R3 - q], mul_w16<T>(one, q * (8 / R3))));
          const C wh = real_twiddle<T, 16>(one, 8 / R3);  // e^{-2 pi i (NS/2) / N}
#pragma unroll
          for (int q = 0; q < R3; q++)
            put(NS / 2 + NS * q, r2c_split<T>(zb[q], zb[R3 - 1 - q], mul_w16<T>(wh, q * (8 / R3))));
        } else {
#pragma unroll
          for (int q = 0; q < R3; q++) {
            const C wk = mul_w16<T>(wj, q * (8 / R3));
            put(j + NS * q, r2c_split<T>(za[q], zb[R3 - 1 - q], wk));
            put(M - j - NS * q, r2c_split<T>(zb[R3 - 1 - q], za[q], mk<T>(-wk.x, wk.y)));
          }
        }
      }
    } else {
    syncB();
    smem_gather<T, M, E>(v, BB, tB);  // re-maps to the store side when TS
    if constexpr (!r2c) {
      syncB();  // every value is back in registers: the buffers are free for the next tile
      if constexpr (SY) if (dynamic) nxt = sn[(it + 1) & 1];
      issue(nxt, false);
    }
    if constexpr (R3 > 1) reg_pass3<T, M, E, R3>(v, tB, T3);
    else reg_pass2<T, M, E, R1, R2>(v, tB, T2);
    }
    // v[m] = forward core output F[tB + m*TP] of pencil slotB

    // ---------------- epilogue + stores
    if constexpr (r2c_sym) {
      // (stored above)
    } else if constexpr (r2c) {
      // X[k] = ((Z[k] + conj Z[M-k]) - i e^{-2 pi i k/N} (Z[k] - conj Z[M-k])) / 2, k = 0..M
      // (no barrier before these writes: a thread overwrites exactly the locations its own gather above has read.
      //  Tried and dropped: FFT + split in the pencil-major mapping with partners by warp shuffles and a re-mapping pass
      //  afterwards -- two CTA barriers instead of three, but 64 double shuffles and spills: 3.85 vs 4.6-4.9 TB/s)
#pragma unroll
      for (int m = 0; m < E; m++) BB[padidx(tB + m * TP)] = v[m];
      syncB();
      C xM = mk<T>((T)0, (T)0);
#pragma unroll
      for (int m = 0; m < E; m++) {
        const int k = tB + m * TP;
        const C zk = v[m];
        const C zm = cconj(BB[padidx((M - k) & (M - 1))]);
        C s = cadd(zk, zm), d = csub(zk, zm);
        C e = cmulmi(cmul(d, real_twiddle<T, E>(wt, m)));
        v[m] = mk<T>((T)0.5 * (s.x + e.x), (T)0.5 * (s.y + e.y));
        if (m == 0) xM = mk<T>(zk.x - zk.y, (T)0);  // X[M], used by the thread that owns k = 0
      }
      syncB();
      if constexpr (SY) if (dynamic) nxt = sn[(it + 1) & 1];
      issue(nxt, false);
      if (live) {
        if (Q.nseg == 1 && Q.deriv_g <= 0) {  // local stage: one base pointer, constant stride between a thread's stores
          const SegDev &sg = Q.seg[0];
          C *out = (C *)sg.base + sg.off + uo * sg.os_u + vo * sg.os_v + (long long)tB * sg.os_d;
          const long long step = (long long)TP * sg.os_d;
#pragma unroll
          for (int m = 0; m < E; m++) st_out(out + m * step, v[m]);
          if (tB == 0) st_out(out + E * step, xM);
        } else {
#pragma unroll
          for (int m = 0; m < E; m++) store_out<T>(Q, tB + m * TP, uo, vo, v[m]);
          if (tB == 0) store_out<T>(Q, M, uo, vo, xM);
        }
      }
    } else if (c2r) {
      if (live) {  // conj(F(conj Z))[j] = x[2j] + i x[2j+1]; real output is never exchanged: one segment
        const SegDev &sg = Q.seg[0];
        T *out = (T *)sg.base + sg.off + uo * sg.os_u + vo * sg.os_v;
        if (sg.os_d == 1 && (((uintptr_t)out) & (sizeof(C) - 1)) == 0) {
          C *oc = (C *)out;
#pragma unroll
          for (int m = 0; m < E; m++) st_out(oc + tB + m * TP, cconj(v[m]));
        } else {
#pragma unroll
          for (int m = 0; m < E; m++) {
            long long a = (long long)(2 * (tB + m * TP)) * sg.os_d;
            out[a] = v[m].x;
            out[a + sg.os_d] = -v[m].y;
          }
        }
      }
    } else if constexpr (dct1) {
      if (live) {  // outputs k = tB + m TP <= M/2: m < E/2, and k = M/2 from the thread with tB = 0
        if (Q.nseg == 1 && Q.deriv_g <= 0) {
          const SegDev &sg = Q.seg[0];
          C *out = (C *)sg.base + sg.off + uo * sg.os_u + vo * sg.os_v + (long long)tB * sg.os_d;
          const long long step = (long long)TP * sg.os_d;
#pragma unroll
          for (int m = 0; m < E / 2; m++) st_out(out + m * step, v[m]);
          if (tB == 0) st_out(out + (E / 2) * step, v[E / 2]);
        } else {
#pragma unroll
          for (int m = 0; m < E / 2; m++) store_out<T>(Q, tB + m * TP, uo, vo, v[m]);
          if (tB == 0) store_out<T>(Q, M / 2, uo, vo, v[E / 2]);
        }
      }
    } else if constexpr (r2r) {
      if (live) {  // only the first n of the M core outputs are results (shifted by one for the sine kinds I / II)
        const bool direct = Q.nseg == 1 && Q.deriv_g <= 0 && Q.dt_out == 2;
        const SegDev &sg = Q.seg[0];
        C *out = (C *)sg.base + sg.off + uo * sg.os_u + vo * sg.os_v;
#pragma unroll
        for (int m = 0; m < E; m++) {
          C y = v[m];
          const int k = r2r_post<T>(Q, tB + m * TP, y);
          if (k >= 0 && k < Q.n_out) {
            if (direct) st_out(out + (long long)k * sg.os_d, y);
            else store_out<T>(Q, k, uo, vo, y);
          }
        }
      }
    } else if (live) {
      if (Q.nseg == 1 && Q.deriv_g <= 0) {  // local stage: one base pointer, constant stride between a thread's stores
        const SegDev &sg = Q.seg[0];
        C *out = (C *)sg.base + sg.off + uo * sg.os_u + vo * sg.os_v + (long long)tB * sg.os_d;
        const long long step = (long long)TP * sg.os_d;
#pragma unroll
        for (int m = 0; m < E; m++) st_out(out + m * step, bwd ? cconj(v[m]) : v[m]);
      } else {
#pragma unroll
        for (int m = 0; m < E; m++) store_out<T>(Q, tB + m * TP, uo, vo, bwd ? cconj(v[m]) : v[m]);
      }
    }
    if constexpr (SY) {
      // a group that signals: count this tile's pencils; when the CTA (TS) or the pencil's thread group (!TS) leaves the
      // group, add them to the group's counter -- whoever completes it publishes the flag
      if (Yp->grp[gS].signal_id >= 0 || Yp->grp[gS].count) {
        const bool leaving = nxt >= nwork || group_of(dynp ? nxt / P : nxt, gS) != gS;
        if (TS) {
          cnt += P;
          if (leaving) {
            __syncthreads();
            if (tid == 0) group_done<P>(*Yp, gS, cnt);
            cnt = 0;
          }
        } else {
          cnt += 1;
          if (leaving) {
            syncA();
            if (tA == 0) group_done<P>(*Yp, gS, cnt);
            cnt = 0;
          }
        }
      }
      it++;
    }
    work = nxt;
  }
}

template <typename T, int M, int KIND, int P, int TS>
__global__ void __launch_bounds__(PipeCfg<T, M, KIND, P, TS>::THREADS, PipeCfg<T, M, KIND, P, TS>::MINB)
pow2_pipe_kernel(const __grid_constant__ StageParams Q) {
  pow2_pipe_body<T, M, KIND, P, TS, 0>(Q, nullptr);
}

template <typename T, int M, int KIND, int P, int TS>
__global__ void __launch_bounds__(PipeCfg<T, M, KIND, P, TS>::THREADS, PipeCfg<T, M, KIND, P, TS>::MINB)
pow2_pipe_sync_kernel(const __grid_constant__ StageParams Q, const __grid_constant__ SyncDev Y) {
  pow2_pipe_body<T, M, KIND, P, TS, 1>(Q, &Y);
}

// ------------------------------------------------------------------ host side: lookup tables, one per (T, KIND, TS)
struct PipeInfo {
  void (*launch)(const StageParams &, int grid, cudaStream_t);
  void (*launch_sync)(const StageParams &, const SyncDev &, int grid, cudaStream_t);
  const void *func, *func_sync;
  int threads, ts, minb;
  size_t smem;
};

template <typename T, int M, int KIND, int P, int TS> void pipe_launcher(const StageParams &Q, int grid, cudaStream_t s) {
  typedef PipeCfg<T, M, KIND, P, TS> Cfg;
  P3B_LAUNCH((pow2_pipe_kernel<T, M, KIND, P, TS>), grid, Cfg::THREADS, Cfg::smem, s, Q);
}

template <typename T, int M, int KIND, int P, int TS> void pipe_sync_launcher(const StageParams &Q, const SyncDev &Y, int grid, cudaStream_t s) {
  typedef PipeCfg<T, M, KIND, P, TS> Cfg;
  P3B_LAUNCH2((pow2_pipe_sync_kernel<T, M, KIND, P, TS>), grid, Cfg::THREADS, Cfg::smem, s, Q, Y);
}

template <typename T, int M, int KIND, int P, int TS> const PipeInfo *pipe_info_one() {
  typedef PipeCfg<T, M, KIND, P, TS> Cfg;
  if constexpr (!Cfg::valid) {
    return nullptr;
  } else {
    if constexpr (KIND == kPipeR2R || KIND == kPipeDCT1) {  // (no tile-group form: an r2r stage of an overlapped pair runs chunk by chunk)
      static const PipeInfo info = {pipe_launcher<T, M, KIND, P, TS>, nullptr, (const void *)pow2_pipe_kernel<T, M, KIND, P, TS>, nullptr,
                                    Cfg::THREADS, TS, Cfg::MINB, Cfg::smem};
      return &info;
    } else {
      static const PipeInfo info = {pipe_launcher<T, M, KIND, P, TS>, pipe_sync_launcher<T, M, KIND, P, TS>,
                                    (const void *)pow2_pipe_kernel<T, M, KIND, P, TS>, (const void *)pow2_pipe_sync_kernel<T, M, KIND, P, TS>,
                                    Cfg::THREADS, TS, Cfg::MINB, Cfg::smem};
      return &info;
    }
  }
}

template <typename T, int M, int KIND, int TS> const PipeInfo *pipe_info_m(int P) {
  switch (P) {
    case 1: return pipe_info_one<T, M, KIND, 1, TS>();
    case 2: return pipe_info_one<T, M, KIND, 2, TS>();
    case 4: return pipe_info_one<T, M, KIND, 4, TS>();
    case 8: return pipe_info_one<T, M, KIND, 8, TS>();
    case 16: return pipe_info_one<T, M, KIND, 16, TS>();
  }
  return nullptr;
}

template <typename T, int KIND, int TS> const PipeInfo *pipe_info(int M, int P) {
  switch (M) {
    case 64: return pipe_info_m<T, 64, KIND, TS>(P);
    case 128: return pipe_info_m<T, 128, KIND, TS>(P);
    case 256: return pipe_info_m<T, 256, KIND, TS>(P);
    case 512: return pipe_info_m<T, 512, KIND, TS>(P);
    case 1024: return pipe_info_m<T, 1024, KIND, TS>(P);
    case 2048: return pipe_info_m<T, 2048, KIND, TS>(P);
    case 4096: return pipe_info_m<T, 4096, KIND, TS>(P);
  }
  return nullptr;
}

// defined in pow2_pipe_inst.cu, compiled once per (precision, kind) so that the instantiations build in parallel
const PipeInfo *pipe_lookup(int prec, int kind, int ts, int M, int P);

}  // namespace p3b
