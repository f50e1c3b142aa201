// pow2_pipe.cuh -- software-pipelined stage kernel for power-of-two FFT cores (the production fast path).
//
// Same arithmetic as pow2_stage.cuh (register-resident radix-16/8/4/2 Stockham passes, E complex values per
// thread, padded shared-memory exchanges, R2C/C2R through the half-length complex core) but restructured so
// that HBM traffic and arithmetic overlap inside ONE resident CTA per SM:
//   * a persistent CTA walks its tiles of P pencils; while tile i is being transformed, tile i+1 is already
//     streaming from global into the staging buffer S with cp.async (LDGSTS: no registers held, addresses in
//     the INPUT's coalescing order, 16-byte granules where the layout allows);
//   * at the top of an iteration the tile is read out of S into registers (one LDS per value, applying the
//     kind's pre-processing), S is released and the prefetch of the next tile is issued immediately, so a full
//     tile of loads is in flight during all the passes, exchanges and stores of the current tile;
//   * passes exchange through a second buffer X; when S + X do not fit in 227 KB (1024-point double with 8
//     pencils, 2048-point single) X holds half a tile and the exchange runs in two half-steps;
//   * stores go straight from registers to global (or to a peer's buffer over NVLink through the segment
//     table) in the OUTPUT's coalescing order: the thread -> (pencil, slot) mapping is chosen for the store side.
// The transform kind, the load granule and the tile size are template parameters: no per-element branching.
// Replaces reference FFTW execute + reorder_trans + pack_sendbuf_trans (exec.C:737-1326, 2792-2879).
#pragma once
#include "common.cuh"
#include "generic_stage.cuh"
#include "pow2_stage.cuh"

namespace p3b {

// ------------------------------------------------------------------ cp.async wrappers
#ifdef P3B_EMU
template <int N> inline void cp_async(void *dst, const void *src) { memcpy(dst, src, N); }
inline void cp_async_commit() {}
inline void cp_async_wait_all() {}
#else
template <int N> __device__ __forceinline__ void cp_async(void *smem_dst, const void *gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  if (N == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(s), "l"(gsrc), "n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#endif

// load granule: LD_ELEM = one complex element of the core (for R2C: two adjacent reals), LD_REAL = one real (R2C whose
// transform dimension is not the unit-stride one, or whose rows are not 2-real aligned)
enum { LD_ELEM = 0, LD_REAL = 1 };

constexpr size_t kPipeSmemMax = 232448 - 1024;  // 227 KB opt-in minus the static reserve

template <typename T, int M, int P> struct PipeCfg {
  enum { E = Pow2Cfg<M>::E, TP = M / E, THREADS = P * TP, PITCH = Pow2Smem<M>::PENCIL };
  static constexpr size_t pencil_bytes = (size_t)PITCH * 2 * sizeof(T);
  static constexpr bool fits1 = 2 * P * pencil_bytes <= kPipeSmemMax;
  static constexpr bool fits2 = (P >= 2) && (P + P / 2) * pencil_bytes <= kPipeSmemMax;
  static constexpr bool valid = (THREADS >= 32) && (THREADS <= 1024) && (fits1 || fits2);
  enum { XS = fits1 ? 1 : 2, PX = P / XS };
  static constexpr size_t smem = (size_t)(P + PX) * pencil_bytes;
  // register budget as in pow2_stage.cuh: 128 per thread in double, 80 in single
  enum { BUDGET = sizeof(T) == 8 ? 512 : 768, MINB = (BUDGET / THREADS) < 1 ? 1 : (BUDGET / THREADS) };
};

// ------------------------------------------------------------------ stores
template <typename T> __device__ __forceinline__ typename cx<T>::type apply_deriv(typename cx<T>::type val, int k, int g) {
  T kap = (T)deriv_kappa(k, g);
  return mk<T>(-kap * val.y, kap * val.x);
}

// ------------------------------------------------------------------ the kernel
template <typename T, int M, int KIND, int P, int LD>
__global__ void __launch_bounds__(PipeCfg<T, M, P>::THREADS, PipeCfg<T, M, P>::MINB)
pow2_pipe_kernel(const __grid_constant__ StageParams Q) {
  typedef typename cx<T>::type C;
  typedef PipeCfg<T, M, P> Cfg;
  typedef Pow2Cfg<M> R;
  constexpr int E = R::E, R1 = R::R1, R2 = R::R2, R3 = R::R3;
  constexpr int TP = Cfg::TP, THREADS = Cfg::THREADS, PITCH = Cfg::PITCH, XS = Cfg::XS, PX = Cfg::PX;
  constexpr bool r2c = KIND == P3DFFTCU_K_R2C, c2r = KIND == P3DFFTCU_K_C2R;
  constexpr bool bwd = KIND == P3DFFTCU_K_C2C_BWD || c2r;
  constexpr int twscale = (r2c || c2r) ? 2 : 1;  // the table is exp(-2 pi i j / nfft), nfft = 2M in the real cases
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C *S = reinterpret_cast<C *>(smem_raw);
  C *X = S + P * PITCH;
  const C *__restrict__ tw = (const C *)Q.tw;
  const int tid = threadIdx.x;
  const int tile_u = Q.tile_u, tu_log2 = Q.tu_log2, tile_v = Q.tile_v;

  // compute/store mapping: thread -> (pencil slot in tile, FFT slot t)
  int slot, t;
  if (Q.store_ord == ORD_D) { slot = tid / TP; t = tid % TP; }
  else if (Q.store_ord == ORD_U) { slot = tid % P; t = tid / P; }  // slot = pv * tile_u + pu: u fastest
  else { int pl = tid % P; t = tid / P; slot = (pl % tile_v) * tile_u + pl / tile_v; }  // v fastest across lanes
  const int pu = slot & (tile_u - 1), pv = slot >> tu_log2;
  const C *Sp = S + slot * PITCH;
  C *Xp = X + (slot % PX) * PITCH;
  const int xh = slot / PX;  // half-step of the exchange this pencil takes part in

  // prefetch of one tile into S, in the input's coalescing order
  auto prefetch = [&](long long tile) {
    const long long u0 = (tile % Q.tiles_u) * tile_u, v0 = (tile / Q.tiles_u) * tile_v;
    constexpr int GPP = (LD == LD_REAL) ? 2 * M : (c2r ? M + 1 : M);  // granules per pencil
    constexpr int TOTAL = P * GPP;
#pragma unroll 4
    for (int idx = tid; idx < TOTAL; idx += THREADS) {
      int g, pl;
      if (Q.load_ord == ORD_D) { pl = idx / GPP; g = idx - pl * GPP; }
      else { g = idx / P; pl = idx % P; }
      int lu, lv;
      if (Q.load_ord == ORD_V) { lv = pl % tile_v; lu = pl / tile_v; }
      else { lu = pl & (tile_u - 1); lv = pl >> tu_log2; }
      if (u0 + lu >= Q.nu || v0 + lv >= Q.nv) continue;
      const long long base = (u0 + lu) * Q.is_u + (v0 + lv) * Q.is_v;
      C *dst = S + (lv * tile_u + lu) * PITCH;
      if (LD == LD_REAL) {
        cp_async<sizeof(T)>((T *)(dst + padidx(g >> 1)) + (g & 1), (const T *)Q.in + base + (long long)g * Q.is_d);
      } else if (r2c) {
        cp_async<sizeof(C)>(dst + padidx(g), (const T *)Q.in + base + 2 * g);
      } else {
        cp_async<sizeof(C)>(dst + padidx(g), (const C *)Q.in + base + (long long)g * Q.is_d);
      }
    }
    cp_async_commit();
  };

  // one exchange through X: scatter in Stockham order, gather in slot order (XS half-steps when X holds half a tile)
  auto exchange = [&](C *v, auto scatter, bool lead_sync) {
#pragma unroll
    for (int h = 0; h < XS; h++) {
      if (h > 0 || lead_sync) __syncthreads();  // X is free again
      if (XS == 1 || xh == h) scatter(v);
      __syncthreads();
      if (XS == 1 || xh == h) smem_gather<T, M, E>(v, Xp, t);
    }
  };

  long long tile = blockIdx.x;
  if (tile < Q.ntiles) prefetch(tile);
  for (; tile < Q.ntiles; tile += gridDim.x) {
    const long long u0 = (tile % Q.tiles_u) * tile_u, v0 = (tile / Q.tiles_u) * tile_v;
    const long long uo = u0 + pu, vo = v0 + pv;
    const bool live = uo < Q.nu && vo < Q.nv;
    C v[E];
    cp_async_wait_all();
    __syncthreads();  // the whole tile has landed in S
    // ---------------- S -> registers (+ pre-processing)
    if (c2r) {
      // Z[k] = (X[k] + conj X[M-k]) + i e^{+2 pi i k/N} (X[k] - conj X[M-k]); we need conj Z for the conj-trick inverse
#pragma unroll
      for (int m = 0; m < E; m++) {
        const int k = t + m * TP;
        C a = Sp[padidx(k)];
        C b = cconj(Sp[padidx(M - k)]);
        if (k == 0) { a.y = 0; b.y = 0; }  // FFTW's c2r ignores Im X[0] and Im X[N/2]
        C s = cadd(a, b), d = csub(a, b);
        C w = cconj(__ldg(&tw[k]));
        C e = cmuli(cmul(d, w));
        v[m] = cconj(cadd(s, e));
      }
    } else {
#pragma unroll
      for (int m = 0; m < E; m++) {
        C x = Sp[padidx(t + m * TP)];
        v[m] = bwd ? cconj(x) : x;
      }
    }
    __syncthreads();  // everyone has read S: release it to the next tile
    if (tile + gridDim.x < Q.ntiles) prefetch(tile + gridDim.x);

    // ---------------- passes
    reg_pass<T, M, E, R1, false>(v, t, 1, tw, twscale);
    exchange(v, [&](C *w) { smem_scatter<T, M, E, R1>(w, Xp, t, 1); }, false);
    reg_pass<T, M, E, R2, true>(v, t, R1, tw, twscale);
    if (R3 > 1) {
      exchange(v, [&](C *w) { smem_scatter<T, M, E, R2>(w, Xp, t, R1); }, true);
      reg_pass<T, M, E, (R3 > 1 ? R3 : 2), true>(v, t, R1 * R2, tw, twscale);
    }
    // v[m] = forward core output F[t + m*TP] of this thread's pencil

    // ---------------- epilogue + stores
    if (r2c) {
      // X[k] = ((Z[k] + conj Z[M-k]) - i e^{-2 pi i k/N} (Z[k] - conj Z[M-k])) / 2, k = 0..M
#pragma unroll
      for (int h = 0; h < XS; h++) {
        __syncthreads();  // X is free
        if (XS == 1 || xh == h) {
#pragma unroll
          for (int m = 0; m < E; m++) Xp[padidx(t + m * TP)] = v[m];
        }
        __syncthreads();
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

// ------------------------------------------------------------------ host side: lookup tables, one per (T, KIND, LD)
struct PipeInfo {
  void (*launch)(const StageParams &, int grid, cudaStream_t);
  const void *func;
  int threads, xs, minb;
  size_t smem;
};

template <typename T, int M, int KIND, int P, int LD> void pipe_launcher(const StageParams &Q, int grid, cudaStream_t s) {
  typedef PipeCfg<T, M, P> Cfg;
  P3B_LAUNCH((pow2_pipe_kernel<T, M, KIND, P, LD>), grid, Cfg::THREADS, Cfg::smem, s, Q);
}

template <typename T, int M, int KIND, int P, int LD> const PipeInfo *pipe_info_one() {
  typedef PipeCfg<T, M, P> Cfg;
  if constexpr (!Cfg::valid) {
    return nullptr;
  } else {
    static const PipeInfo info = {pipe_launcher<T, M, KIND, P, LD>, (const void *)pow2_pipe_kernel<T, M, KIND, P, LD>, Cfg::THREADS,
                                  Cfg::XS, Cfg::MINB, Cfg::smem};
    return &info;
  }
}

template <typename T, int M, int KIND, int LD> const PipeInfo *pipe_info_m(int P) {
  switch (P) {
    case 2: return pipe_info_one<T, M, KIND, 2, LD>();
    case 4: return pipe_info_one<T, M, KIND, 4, LD>();
    case 8: return pipe_info_one<T, M, KIND, 8, LD>();
    case 16: return pipe_info_one<T, M, KIND, 16, LD>();
  }
  return nullptr;
}

template <typename T, int KIND, int LD> const PipeInfo *pipe_info(int M, int P) {
  switch (M) {
    case 64: return pipe_info_m<T, 64, KIND, LD>(P);
    case 128: return pipe_info_m<T, 128, KIND, LD>(P);
    case 256: return pipe_info_m<T, 256, KIND, LD>(P);
    case 512: return pipe_info_m<T, 512, KIND, LD>(P);
    case 1024: return pipe_info_m<T, 1024, KIND, LD>(P);
    case 2048: return pipe_info_m<T, 2048, KIND, LD>(P);
    case 4096: return pipe_info_m<T, 4096, KIND, LD>(P);
  }
  return nullptr;
}

// defined in pow2_pipe_inst.cu, compiled once per (precision, kind) so that the ~150 instantiations build in parallel
const PipeInfo *pipe_lookup(int prec, int kind, int ld, int M, int P);

}  // namespace p3b
