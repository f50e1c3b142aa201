// mixed_pipe.cuh -- TMA-fed stage kernel for smooth lengths M = Q * MC: one odd factor Q in {3, 5, 7, 9, 15} times a power of two
// MC in {64 ... 1024} (192 = 3 * 64, 768 = 3 * 256, 640 = 5 * 128, 1536 = 3 * 512, 896 = 7 * 128, 1152 = 9 * 128, 1920 = 15 * 128, ...; real
// transforms of twice those), wherever Q * MC / 16 threads per pencil make whole warps within one CTA.
//
// Cooley-Tukey split n = MC n1 + n2, k = k1 + Q k2:
//     X[k1 + Q k2] = sum_{n2} w_MC^{n2 k2} { w_M^{n2 k1} sum_{n1} x[MC n1 + n2] w_Q^{n1 k1} }
// A pencil is handled by Q groups of TPC = MC/16 threads.  The radix-Q step (the braces) runs first, in one of two forms:
// register butterflies -- a thread loads the Q values x[MC n1 + n2] of one n2, evaluates the odd-length DFT by its real
// symmetry and writes the Q outputs, twiddled, back in place, so that group k1 finds its row contiguous (dft_odd below) --
// or, the kernel's first form, kept where it measured faster (single precision, Q = 3): group k1 builds ITS row straight
// from the pencil as it landed (Q loads and Q-1 complex multiplies per value, by definition).  Either way group k1 then
// holds its row times the twiddle w_M^{n2 k1} and runs the SAME register-resident MC-point core as the power-of-two kernel (pow2_pipe.cuh:
// radix-16 passes, padded in-place exchanges, smem twiddle tables) on its own sub-buffer.  Data movement is the power-of-two
// kernel's: one bulk copy (cp.async.bulk + mbarrier) per pencil, issued for the NEXT tile as soon as the last exchange has
// been read back; transposed stores re-map threads to lanes-across-pencils in that exchange.  Barriers are CTA-wide (a
// pencil's 3 * 16 threads are not whole warps).  Replaces Bluestein (two FFTs of the next power of two >= 2M) for these
// lengths: the reference's FFTW plans of init.C:1146-1607 / templ.C:1283-1366 for sizes such as the 768^3 runs of
// extra/makejob.py:131-134.
#pragma once
#include "pow2_pipe.cuh"

namespace p3b {

// ------------------------------------------------------------------ radix-Q butterfly (odd Q) in registers
// cos / sin of 2 pi m / Q, m < Q, as compile-time constants (indices are constant after unrolling)
template <int Q> struct QRoots;
template <> struct QRoots<3> {
  static __host__ __device__ constexpr double c(int m) { constexpr double t[3] = {1.0, -0.4999999999999998, -0.5000000000000004}; return t[m]; }
  static __host__ __device__ constexpr double s(int m) { constexpr double t[3] = {0.0, 0.8660254037844387, -0.8660254037844384}; return t[m]; }
};
template <> struct QRoots<5> {
  static __host__ __device__ constexpr double c(int m) { constexpr double t[5] = {1.0, 0.30901699437494745, -0.8090169943749473, -0.8090169943749476, 0.30901699437494723}; return t[m]; }
  static __host__ __device__ constexpr double s(int m) { constexpr double t[5] = {0.0, 0.9510565162951535, 0.5877852522924732, -0.587785252292473, -0.9510565162951536}; return t[m]; }
};
template <> struct QRoots<7> {
  static __host__ __device__ constexpr double c(int m) { constexpr double t[7] = {1.0, 0.6234898018587336, -0.22252093395631434, -0.900968867902419, -0.9009688679024191, -0.2225209339563146, 0.6234898018587334}; return t[m]; }
  static __host__ __device__ constexpr double s(int m) { constexpr double t[7] = {0.0, 0.7818314824680298, 0.9749279121818236, 0.43388373911755823, -0.433883739117558, -0.9749279121818236, -0.7818314824680299}; return t[m]; }
};
template <> struct QRoots<9> {
  static __host__ __device__ constexpr double c(int m) { constexpr double t[9] = {1.0, 0.766044443118978, 0.17364817766693041, -0.4999999999999998, -0.9396926207859083, -0.9396926207859084, -0.5000000000000004, 0.17364817766692997, 0.7660444431189778}; return t[m]; }
  static __host__ __device__ constexpr double s(int m) { constexpr double t[9] = {0.0, 0.6427876096865393, 0.984807753012208, 0.8660254037844387, 0.3420201433256689, -0.34202014332566866, -0.8660254037844384, -0.9848077530122081, -0.6427876096865396}; return t[m]; }
};
template <> struct QRoots<15> {
  static __host__ __device__ constexpr double c(int m) { constexpr double t[15] = {1.0, 0.9135454576426009, 0.6691306063588582, 0.30901699437494745, -0.10452846326765333, -0.4999999999999998, -0.8090169943749473, -0.9781476007338057, -0.9781476007338057, -0.8090169943749476, -0.5000000000000004, -0.10452846326765423, 0.30901699437494723, 0.6691306063588585, 0.913545457642601}; return t[m]; }
  static __host__ __device__ constexpr double s(int m) { constexpr double t[15] = {0.0, 0.40673664307580015, 0.7431448254773941, 0.9510565162951535, 0.9945218953682734, 0.8660254037844387, 0.5877852522924732, 0.20791169081775931, -0.20791169081775907, -0.587785252292473, -0.8660254037844384, -0.9945218953682733, -0.9510565162951536, -0.743144825477394, -0.40673664307580015}; return t[m]; }
};

// The radix-Q step runs as register butterflies (one thread transforms the Q values x[MC n1 + n2], n1 < Q, of one n2 and
// writes them back in place) for Q >= BflyMinQ<T>::value; below that every thread group evaluates its own output row by definition
// from the whole pencil (Q - 1 complex multiplies per value: FP64-issue-bound from Q = 5 on).  Measured on B200, 1-GPU cubes
// in double, by definition -> butterflies (profiles/r02v_mixab_*.txt): 768^3 R2C 5.12 -> 4.80 ms, 640^3 C2C 6.20 -> 5.48,
// 896^3 19.1 -> 14.8, 1152^3 R2C 28.8 -> 21.3, 960^3 35.4 -> 23.0; single precision, Q = 3: 3.32 -> 3.34 (the 3 x 256
// stages lose 4-10 %, the 3 x 128 stage gains 7 %), hence by definition there.  -DP3B_MIX_BFLY_MINQ_F64/_F32 override.
#ifndef P3B_MIX_BFLY_MINQ_F64
#define P3B_MIX_BFLY_MINQ_F64 3
#endif
#ifndef P3B_MIX_BFLY_MINQ_F32
#define P3B_MIX_BFLY_MINQ_F32 5
#endif
template <typename T> struct BflyMinQ { enum { value = sizeof(T) == 8 ? P3B_MIX_BFLY_MINQ_F64 : P3B_MIX_BFLY_MINQ_F32 }; };

// forward DFT of odd length Q by its real symmetry: a_j = x_j + x_{Q-j}, b_j = x_j - x_{Q-j} (j <= H = (Q-1)/2),
//   y_k = x_0 + sum_j a_j cos(2 pi j k / Q) - i sum_j b_j sin(2 pi j k / Q),  y_{Q-k} = the same with + i:
// (Q-1)^2 real multiply-adds instead of (Q-1)^2 complex multiplies.  emit(k, y_k) receives the outputs.
template <typename T, int Q, typename Emit> __device__ __forceinline__ void dft_odd(const typename cx<T>::type *x, Emit emit) {
  typedef typename cx<T>::type C;
  constexpr int H = (Q - 1) / 2;
  C a[H], b[H];
  C y0 = x[0];
#pragma unroll
  for (int j = 1; j <= H; j++) {
    a[j - 1] = cadd(x[j], x[Q - j]);
    b[j - 1] = csub(x[j], x[Q - j]);
    y0 = cadd(y0, a[j - 1]);
  }
  emit(0, y0);
#pragma unroll
  for (int k = 1; k <= H; k++) {
    T cr = x[0].x, ci = x[0].y, dr = (T)0, di = (T)0;
#pragma unroll
    for (int j = 1; j <= H; j++) {
      const T c = (T)QRoots<Q>::c((j * k) % Q), sn = (T)QRoots<Q>::s((j * k) % Q);
      cr += c * a[j - 1].x;
      ci += c * a[j - 1].y;
      dr += sn * b[j - 1].x;
      di += sn * b[j - 1].y;
    }
    emit(k, mk<T>(cr + di, ci - dr));      // c_k - i d_k
    emit(Q - k, mk<T>(cr - di, ci + dr));  // c_k + i d_k
  }
}

template <typename T, int MC, int Q, int KIND, int P, int TS> struct MixCfg {
  enum { E = Pow2Cfg<MC>::E, TPC = MC / E, TP = Q * TPC, THREADS = P * TP, M = Q * MC };
  enum { R1 = Pow2Cfg<MC>::R1, R2 = Pow2Cfg<MC>::R2, R3 = Pow2Cfg<MC>::R3 };
  enum { PITCHC = Pow2Smem<MC>::PENCIL };  // padded sub-buffer of one group (odd)
  // complex-sized elements of one landed pencil (C2R: M + 1, in single precision with the 8 bytes that follow)
  enum { NIN = KIND == P3DFFTCU_K_C2R ? (sizeof(T) == 4 ? M + 2 : M + 1) : M };
  enum { PITCH0 = Q * PITCHC };  // = M + M/16 + Q: holds the landed pencil and the natural-order padded R2C split buffer
  enum { PITCH = sizeof(T) == 8 ? PITCH0 : (PITCH0 % 2 ? PITCH0 + 1 : PITCH0) };  // pencils start 16-byte aligned
  enum { T2N = R1 * R2, T3N = R3 > 1 ? R3 * TPC : 0 };
  static constexpr size_t csz = 2 * sizeof(T);
  static constexpr size_t bar_bytes = 128;
  static constexpr size_t smem = bar_bytes + ((size_t)P * PITCH + T2N + T3N) * csz;
  static constexpr bool valid = (E == 16 || MC == 64) && (THREADS % 32 == 0) && (THREADS >= 64) && (THREADS <= 768) && (P <= 16) &&
                                (smem <= kPipeSmemMax) && (TS ? P >= 2 : true) && (sizeof(T) == 8 ? THREADS <= 512 : true);
  // small CTAs (3 x 128 cores: 24 threads per pencil) share an SM in pairs: two independent CTAs overlap their phases
  enum { MINB = THREADS <= 256 ? 2 : 1 };
};

template <typename T, int MC, int Q, int KIND, int P, int TS>
__global__ void __launch_bounds__(MixCfg<T, MC, Q, KIND, P, TS>::THREADS, MixCfg<T, MC, Q, KIND, P, TS>::MINB)
mixed_pipe_kernel(const __grid_constant__ StageParams S) {
  typedef typename cx<T>::type C;
  typedef MixCfg<T, MC, Q, KIND, P, TS> Cfg;
  constexpr int E = Cfg::E, R1 = Cfg::R1, R2 = Cfg::R2, R3 = Cfg::R3;
  constexpr int TPC = Cfg::TPC, TP = Cfg::TP, THREADS = Cfg::THREADS, PITCH = Cfg::PITCH, PITCHC = Cfg::PITCHC, M = Cfg::M;
  constexpr bool r2c = KIND == P3DFFTCU_K_R2C, c2r = KIND == P3DFFTCU_K_C2R;
  constexpr bool bwd = KIND == P3DFFTCU_K_C2C_BWD || c2r;
  constexpr int twscale = (r2c || c2r) ? 2 : 1;  // the table is exp(-2 pi i j / nfft), nfft = 2M in the real cases
  constexpr unsigned bytes = (unsigned)(Cfg::NIN * Cfg::csz);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem_raw);
  C *B = reinterpret_cast<C *>(smem_raw + Cfg::bar_bytes);
  C *T2 = B + P * PITCH;  // [R2][R1]: w_MC^{q k}
  C *T3 = T2 + Cfg::T2N;  // [R3][TPC]
  const C *__restrict__ tw = (const C *)S.tw;
  const int tid = threadIdx.x;
  const int tile_u = S.tile_u, tu_log2 = S.tu_log2;
  for (int i = tid; i < Cfg::T2N; i += THREADS) T2[i] = tw[(i / R1) * (i % R1) * (MC / (R1 * R2)) * Q * twscale];
  for (int i = tid; i < Cfg::T3N; i += THREADS) T3[i] = tw[(i / TPC) * (i % TPC) * Q * twscale];

  // mapping A (load side, radix-Q step, all core passes but the last): pencil-major; mapping B (last pass and stores): the
  // same for contiguous output, lanes across the tile's pencils for transposed output
  const int slotA = tid / TP, tA = tid % TP;
  const int slotB = TS ? tid % P : slotA, tB = TS ? tid / P : tA;
  // (store side with the group index fastest -- consecutive threads holding consecutive outputs k = qB + Q (tcB + TPC m) --
  //  was measured slower on both store forms: the gather from the Q sub-buffers then conflicts on the banks, 768-point C2C
  //  1.58 vs 1.38 ms transposed, 1.68 vs 1.62 ms contiguous; profiles/r02b_/r02c_ncu_full_768.summary.csv)
  const int qA = tA / TPC, tcA = tA % TPC, qB = tB / TPC, tcB = tB % TPC;
  const int puA = slotA & (tile_u - 1), pvA = slotA >> tu_log2;
  const int puB = slotB & (tile_u - 1), pvB = slotB >> tu_log2;
  C *BA = B + slotA * PITCH, *BB = B + slotB * PITCH;
  unsigned long long *bar = bars + slotA;
  const bool vfast = S.vfast != 0;
  auto tile_origin = [&](long long tl, long long &u0, long long &v0) {
    const long long iu = vfast ? tl / S.tiles_v : tl % S.tiles_u, iv = vfast ? tl % S.tiles_v : tl / S.tiles_u;
    u0 = iu * tile_u;
    v0 = iv * S.tile_v;
  };
  auto issue = [&](long long tl) {
    if (tA == 0 && tl < S.ntiles) {
      long long u, v;
      tile_origin(tl, u, v);
      u += puA;
      v += pvA;
      const bool live = u < S.nu && v < S.nv;
      fence_async_smem();
      mbar_expect_tx(bar, live ? bytes : 0u);
      if (live) {
        const long long base = u * S.is_u + v * S.is_v;
        const void *src = r2c ? (const void *)((const T *)S.in + base) : (const void *)((const C *)S.in + base);
        bulk_g2s(BA, src, bytes, bar);
      }
    }
  };
  if (tid < P) mbar_init(bars + tid, 1);
  __syncthreads();

  unsigned parity = 0;
  issue(blockIdx.x);
  for (long long tile = blockIdx.x; tile < S.ntiles; tile += gridDim.x) {
    const long long nxt = tile + gridDim.x;
    long long uo, vo;
    tile_origin(tile, uo, vo);
    uo += puB;
    vo += pvB;
    const bool live = uo < S.nu && vo < S.nv;
    if constexpr (Q >= BflyMinQ<T>::value) {
      // the register butterflies below read every pencil of the tile: wait for all of them (one phase per pencil and tile)
#pragma unroll 1
      for (int p = 0; p < P; p++) mbar_wait(bars + p, parity);
    } else {
      mbar_wait(bar, parity);
    }
    parity ^= 1;

    if constexpr (c2r) {
      // Hermitian pre-processing once per pencil, in place: Z[j] = (X[j] + conj X[M-j]) + i e^{+2 pi i j/N} (X[j] - conj X[M-j]),
      // stored as conj Z for the conj-trick inverse.  Pairs (j, M-j), j = tA + TP m' < M/2; e^{+2 pi i (M-j)/N} = -conj(e^{...j})
#pragma unroll
      for (int mp = 0; mp < E / 2; mp++) {
        const int j = tA + TP * mp;
        C xa = BA[j], xb = BA[M - j];
        if (j == 0) { xa.y = 0; xb.y = 0; }  // FFTW's c2r ignores Im X[0] and Im X[N/2]
        const C w = cconj(__ldg(&tw[j]));   // e^{+2 pi i j/N}
        {
          const C b = cconj(xb), s = cadd(xa, b), d = csub(xa, b);
          BA[j] = cconj(cadd(s, cmuli(cmul(d, w))));
        }
        if (j > 0) {
          const C b = cconj(xa), s = cadd(xb, b), d = csub(xb, b);
          BA[M - j] = cconj(cadd(s, cmuli(cmul(d, cneg(cconj(w))))));
        }
      }
      if (tA == 0) {  // j = M/2 pairs with itself: e^{+2 pi i (M/2)/N} = i
        const C x = BA[M / 2], b = cconj(x);
        const C s = cadd(x, b), d = csub(x, b);
        BA[M / 2] = cconj(cadd(s, cmuli(cmuli(d))));
      }
      __syncthreads();
    }
    // ---------------- radix-Q step: v[m] = w_M^{n2 qA} sum_{n1} x[MC n1 + n2] w_Q^{n1 qA}, n2 = tcA + TPC m, for group qA
    auto X = [&](int j) -> C {
      const C x = BA[j];
      return (bwd && !c2r) ? cconj(x) : x;
    };
    C v[E];
    if constexpr (Q >= BflyMinQ<T>::value) {
      // register butterflies: the P * MC butterflies of the tile are dealt to the CTA's threads (b = pencil * MC + n2); butterfly
      // n2 reads x[MC n1 + n2], n1 < Q, and writes y[k1] w_M^{n2 k1} back to x[MC k1 + n2] -- locations no other butterfly
      // touches; group k1 then finds its row contiguous at [MC k1, MC k1 + MC)
#pragma unroll 1
      for (int bf = tid; bf < P * MC; bf += THREADS) {
        const int n2 = bf & (MC - 1);
        C *Bp = B + (bf / MC) * PITCH;
        C x[Q];
#pragma unroll
        for (int n1 = 0; n1 < Q; n1++) {
          const C z = Bp[n1 * MC + n2];
          x[n1] = (bwd && !c2r) ? cconj(z) : z;
        }
        dft_odd<T, Q>(x, [&](int k1, const C &y) { Bp[k1 * MC + n2] = k1 ? cmul(y, __ldg(&tw[n2 * k1 * twscale])) : y; });
      }
      __syncthreads();
#pragma unroll
      for (int m = 0; m < E; m++) v[m] = BA[qA * MC + tcA + TPC * m];
    } else {
#pragma unroll
      for (int m = 0; m < E; m++) v[m] = X(tcA + TPC * m);
#pragma unroll
      for (int n1 = 1; n1 < Q; n1++) {
        const C wq = __ldg(&tw[((n1 * qA) % Q) * MC * twscale]);  // w_Q^{n1 qA}
#pragma unroll
        for (int m = 0; m < E; m++) {
          const C x = cmul(X(n1 * MC + tcA + TPC * m), wq);
          v[m] = cadd(v[m], x);
        }
      }
      if (qA > 0) {
#pragma unroll
        for (int m = 0; m < E; m++) v[m] = cmul(v[m], __ldg(&tw[(tcA + TPC * m) * qA * twscale]));
      }
    }
    __syncthreads();  // the pencil is in registers: its buffer now carries the exchanges of the Q cores

    // ---------------- MC-point core of group qA on its sub-buffer
    C *Bq = BA + qA * PITCHC;
    reg_pass<T, MC, E, R1, false>(v, tcA, 1, tw, 1);
    smem_scatter<T, MC, E, R1>(v, Bq, tcA, 1);
    if constexpr (R3 > 1) {
      __syncthreads();
      smem_gather<T, MC, E>(v, Bq, tcA);
      reg_pass2<T, MC, E, R1, R2>(v, tcA, T2);
      __syncthreads();
      smem_scatter<T, MC, E, R2>(v, Bq, tcA, R1);
    }
    __syncthreads();
    smem_gather<T, MC, E>(v, BB + qB * PITCHC, tcB);  // re-maps to the store side when TS
    __syncthreads();  // every value is back in registers
    if constexpr (!r2c) issue(nxt);
    if constexpr (R3 > 1) reg_pass3<T, MC, E, R3>(v, tcB, T3);
    else reg_pass2<T, MC, E, R1, R2>(v, tcB, T2);
    // v[m] = forward transform output F[k], k = qB + Q (tcB + TPC m), of pencil slotB

    // ---------------- epilogue + stores
    if constexpr (r2c) {
      // X[k] = ((Z[k] + conj Z[M-k]) - i e^{-2 pi i k/N} (Z[k] - conj Z[M-k])) / 2, k = 0..M: partners through the buffer in
      // natural (padded) order
#pragma unroll
      for (int m = 0; m < E; m++) BB[padidx(qB + Q * (tcB + TPC * m))] = v[m];
      __syncthreads();
      C xM = mk<T>((T)0, (T)0);
#pragma unroll
      for (int m = 0; m < E; m++) {
        const int k = qB + Q * (tcB + TPC * m);
        const C zk = v[m];
        const C zp = BB[padidx(k == 0 ? 0 : M - k)];
        v[m] = r2c_split<T>(zk, zp, __ldg(&tw[k]));
        if (k == 0) xM = mk<T>(zk.x - zk.y, (T)0);
      }
      __syncthreads();
      issue(nxt);
      if (live) {
#pragma unroll
        for (int m = 0; m < E; m++) store_out<T>(S, qB + Q * (tcB + TPC * m), uo, vo, v[m]);
        if (qB == 0 && tcB == 0) store_out<T>(S, M, uo, vo, xM);
      }
    } else if constexpr (c2r) {
      if (live) {  // conj(F(conj Z))[j] = x[2j] + i x[2j+1]; real output is never exchanged: one segment
        const SegDev &sg = S.seg[0];
        T *out = (T *)sg.base + sg.off + uo * sg.os_u + vo * sg.os_v;
        if (sg.os_d == 1 && (((uintptr_t)out) & (sizeof(C) - 1)) == 0) {
          C *oc = (C *)out;
#pragma unroll
          for (int m = 0; m < E; m++) st_out(oc + qB + Q * (tcB + TPC * m), cconj(v[m]));
        } else {
#pragma unroll
          for (int m = 0; m < E; m++) {
            const long long a = (long long)(2 * (qB + Q * (tcB + TPC * m))) * sg.os_d;
            out[a] = v[m].x;
            out[a + sg.os_d] = -v[m].y;
          }
        }
      }
    } else if (live) {
      if (S.nseg == 1 && S.deriv_g <= 0) {  // local stage: one base pointer
        const SegDev &sg = S.seg[0];
        C *out = (C *)sg.base + sg.off + uo * sg.os_u + vo * sg.os_v;
#pragma unroll
        for (int m = 0; m < E; m++) st_out(out + (long long)(qB + Q * (tcB + TPC * m)) * sg.os_d, bwd ? cconj(v[m]) : v[m]);
      } else {
#pragma unroll
        for (int m = 0; m < E; m++) store_out<T>(S, qB + Q * (tcB + TPC * m), uo, vo, bwd ? cconj(v[m]) : v[m]);
      }
    }
  }
}

// ------------------------------------------------------------------ host side: lookup tables, one TU per (precision, kind)
template <typename T, int MC, int Q, int KIND, int P, int TS> void mixed_launcher(const StageParams &S, int grid, cudaStream_t s) {
  typedef MixCfg<T, MC, Q, KIND, P, TS> Cfg;
  P3B_LAUNCH((mixed_pipe_kernel<T, MC, Q, KIND, P, TS>), grid, Cfg::THREADS, Cfg::smem, s, S);
}

template <typename T, int MC, int Q, int KIND, int P, int TS> const PipeInfo *mixed_info_one() {
  typedef MixCfg<T, MC, Q, KIND, P, TS> Cfg;
  if constexpr (!Cfg::valid) {
    return nullptr;
  } else {
    static const PipeInfo info = {mixed_launcher<T, MC, Q, KIND, P, TS>, nullptr, (const void *)mixed_pipe_kernel<T, MC, Q, KIND, P, TS>,
                                  nullptr, Cfg::THREADS, TS, 1, Cfg::smem};
    return &info;
  }
}

template <typename T, int MC, int Q, int KIND, int TS> const PipeInfo *mixed_info_p(int P) {
  switch (P) {
    case 2: return mixed_info_one<T, MC, Q, KIND, 2, TS>();
    case 4: return mixed_info_one<T, MC, Q, KIND, 4, TS>();
    case 8: return mixed_info_one<T, MC, Q, KIND, 8, TS>();
    case 16: return mixed_info_one<T, MC, Q, KIND, 16, TS>();
  }
  return nullptr;
}

template <typename T, int Q, int KIND, int TS> const PipeInfo *mixed_info_mc(int MC, int P) {
  switch (MC) {
    case 64: return mixed_info_p<T, 64, Q, KIND, TS>(P);  // 8 values per thread (192 = 3 * 64, 320, 448, 576, 960)
    case 128: return mixed_info_p<T, 128, Q, KIND, TS>(P);
    case 256: return mixed_info_p<T, 256, Q, KIND, TS>(P);
    case 512: return mixed_info_p<T, 512, Q, KIND, TS>(P);
    case 1024: return mixed_info_p<T, 1024, Q, KIND, TS>(P);
  }
  return nullptr;
}

template <typename T, int KIND, int TS> const PipeInfo *mixed_info(int Q, int MC, int P) {
  switch (Q) {
    case 3: return mixed_info_mc<T, 3, KIND, TS>(MC, P);
    case 5: return mixed_info_mc<T, 5, KIND, TS>(MC, P);
    case 7: return mixed_info_mc<T, 7, KIND, TS>(MC, P);
    case 9: return mixed_info_mc<T, 9, KIND, TS>(MC, P);    // 1152 = 9 * 128, 2304, 4608: the same radix-Q step, Q need not be prime
    case 15: return mixed_info_mc<T, 15, KIND, TS>(MC, P);  // 1920 = 15 * 128, 3840
  }
  return nullptr;
}

// values per thread of the MC-point core: 16, or 8 for the 64-point core
inline int mixed_values_per_thread(int MC) { return MC == 64 ? 8 : 16; }

// defined in mixed_pipe_inst.cu, compiled once per (precision, kind)
const PipeInfo *mixed_lookup(int prec, int kind, int ts, int Q, int MC, int P);

}  // namespace p3b
