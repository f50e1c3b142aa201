// pow2_tload.cuh -- power-of-two stage kernel for inputs whose unit-stride dimension is NOT the transform dimension
// (a user array in a non-default memory order as the first stage's input, the 1D API along a strided dimension):
// the tile is fetched by TMA *tensor* copies (cp.async.bulk.tensor, one 2-D box per 256 rows) instead of plain loads.
//
// A tile = P pencils that are neighbours along the input's unit-stride dimension.  One box row = the P values those
// pencils hold at one index of the transform dimension (P * 16 B = 128 B for complex double), so the tile lands in shared
// memory as [row][P] -- already transposed with respect to the pencils.  Threads read it with the lanes running across the
// pencils (conflict-free, 512 contiguous bytes per warp), do the first radix pass in that mapping and scatter into the
// per-pencil padded buffers: the FIRST exchange re-maps to the pencil-major mapping of the power-of-two pipe kernel
// (pow2_pipe.cuh), whose remaining passes, transposed or contiguous stores, segment table, fused derivative and next-tile
// prefetch follow unchanged.  No registers, LSU wavefronts or address arithmetic are spent on the strided loads, and DRAM
// sees 128-byte rows instead of per-thread 16-byte accesses.  Kinds: C2C forward / backward, R2C (the first stages of the
// reference's transforms: exec.C:737-1326 cases with TRANS_IN reordering).  CTA-wide barriers.
#pragma once
#include "pow2_pipe.cuh"

namespace p3b {

// kernel-side handle of a TMA tensor map (CUtensorMap is 128 opaque bytes, 64-byte aligned); the emulation keeps the plain
// geometry instead
struct alignas(64) TMapArg {
#ifdef P3B_EMU
  const unsigned char *base;
  long long stride1, stride2;      // bytes between consecutive indices of dims 1 and 2
  unsigned dim0, dim1, dim2;       // extents in elements
  unsigned box0, box1, box2;       // box extents in elements
  unsigned elem_bytes;
  unsigned char pad_[64];
#else
  unsigned char bytes[128];
#endif
};

#ifdef P3B_EMU
inline void tensor_g2s(void *dst, const TMapArg *tm, int c0, int c1, int c2, unsigned long long *) {
  unsigned char *d = (unsigned char *)dst;
  for (unsigned s = 0; s < tm->box2; s++)
    for (unsigned r = 0; r < tm->box1; r++)
      for (unsigned e = 0; e < tm->box0; e++) {
        const long long i0 = (long long)c0 + e, i1 = (long long)c1 + r, i2 = (long long)c2 + s;
        unsigned char *o = d + (((size_t)s * tm->box1 + r) * tm->box0 + e) * tm->elem_bytes;
        if (i0 < tm->dim0 && i1 < tm->dim1 && i2 < tm->dim2)
          memcpy(o, tm->base + i0 * tm->elem_bytes + i1 * tm->stride1 + i2 * tm->stride2, tm->elem_bytes);
        else memset(o, 0, tm->elem_bytes);  // out-of-bounds elements are zero-filled
      }
}
#else
__device__ __forceinline__ void tensor_g2s(void *smem_dst, const TMapArg *tm, int c0, int c1, int c2, unsigned long long *bar) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst), b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(d),
               "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(b)
               : "memory");
}
#endif

template <typename T, int M, int KIND, int P, int TS> struct TLoadCfg {
  enum { E = Pow2Cfg<M>::E, TP = M / E, THREADS = P * TP, XP0 = Pow2Smem<M>::PENCIL };
  enum { R1 = Pow2Cfg<M>::R1, R2 = Pow2Cfg<M>::R2, R3 = Pow2Cfg<M>::R3 };
  enum { PITCH = sizeof(T) == 8 ? XP0 : ((XP0 + 1) % 4 == 2 ? XP0 + 1 : XP0 + 3) };  // as PipeCfg
  enum { T2N = R1 * R2, T3N = R3 > 1 ? R3 * TP : 0 };
  // rows of the landed tile: M complex rows, or 2M real rows for R2C; one box holds at most 256 rows
  enum { ROWS = KIND == P3DFFTCU_K_R2C ? 2 * M : M, BOXROWS = ROWS < 256 ? ROWS : 256, NBOX = ROWS / BOXROWS };
  static constexpr size_t csz = 2 * sizeof(T);
  static constexpr size_t row_bytes = (KIND == P3DFFTCU_K_R2C ? sizeof(T) : csz) * P;
  static constexpr size_t bar_bytes = 128;
  static constexpr size_t smem = bar_bytes + ((size_t)P * PITCH + T2N + T3N) * csz;  // (P * PITCH >= P * M: holds the landed tile)
  static constexpr bool valid = (THREADS >= 64) && (THREADS <= 1024) && (THREADS % 32 == 0) && (P >= 2) && (P <= 16) &&
                                (smem <= kPipeSmemMax) && (row_bytes >= 16) && (row_bytes % 16 == 0) && (E == 16 || E == 8);
  enum { BUDGET = sizeof(T) == 8 ? 512 : 768, MINB = (BUDGET / THREADS) < 1 ? 1 : (BUDGET / THREADS) };
};

template <typename T, int M, int KIND, int P, int TS>
__global__ void __launch_bounds__(TLoadCfg<T, M, KIND, P, TS>::THREADS, TLoadCfg<T, M, KIND, P, TS>::MINB)
pow2_tload_kernel(const __grid_constant__ StageParams Q, const __grid_constant__ TMapArg tmap) {
  typedef typename cx<T>::type C;
  typedef TLoadCfg<T, M, KIND, P, TS> Cfg;
  constexpr int E = Cfg::E, R1 = Cfg::R1, R2 = Cfg::R2, R3 = Cfg::R3;
  constexpr int TP = Cfg::TP, THREADS = Cfg::THREADS, PITCH = Cfg::PITCH;
  constexpr bool r2c = KIND == P3DFFTCU_K_R2C;
  constexpr bool bwd = KIND == P3DFFTCU_K_C2C_BWD;
  constexpr int twscale = r2c ? 2 : 1;
  constexpr unsigned tile_bytes = (unsigned)(Cfg::ROWS * Cfg::row_bytes);
#ifdef P3B_EMU
  unsigned char *tl_smem = smem_raw;
#else
  extern __shared__ __align__(128) unsigned char tl_smem[];  // own name (other kernels of this unit declare 16 bytes): TMA tensor copies land 128-byte aligned
#endif
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(tl_smem);
  C *B = reinterpret_cast<C *>(tl_smem + Cfg::bar_bytes);
  C *T2 = B + P * PITCH;
  C *T3 = T2 + Cfg::T2N;
  const C *__restrict__ tw = (const C *)Q.tw;
  const int tid = threadIdx.x;
  for (int i = tid; i < Cfg::T2N; i += THREADS) T2[i] = tw[(i / R1) * (i % R1) * (M / (R1 * R2)) * twscale];
  for (int i = tid; i < Cfg::T3N; i += THREADS) T3[i] = tw[(i / TP) * (i % TP) * twscale];

  // the tile's P pencils are neighbours along the input's unit-stride dimension: u (Q.load_ord == ORD_U) or v
  const bool along_u = Q.load_ord == ORD_U;
  // mapping L (landed tile, first pass): lanes across pencils; mapping A (middle passes): pencil-major; mapping S (last
  // pass, stores): lanes across pencils again for transposed stores (same pencils: the output's unit-stride dimension is
  // the input's), else pencil-major
  const int slotL = tid % P, tL = tid / P;
  const int slotA = tid / TP, tA = tid % TP;
  const int slotS = TS ? slotL : slotA, tS = TS ? tL : tA;
  C *BL = B + slotL * PITCH, *BA = B + slotA * PITCH, *BS = B + slotS * PITCH;
  const long long tiles_unit = along_u ? Q.tiles_u : Q.tiles_v;  // tiles along the unit-stride dimension
  auto tile_origin = [&](long long tl, long long &u0, long long &v0) {
    const long long a = tl % tiles_unit, b = tl / tiles_unit;  // consecutive tiles run along the unit-stride dimension
    u0 = along_u ? a * P : b;
    v0 = along_u ? b : a * P;
  };
  auto issue = [&](long long tl) {
    if (tid == 0 && tl < Q.ntiles) {
      long long u0, v0;
      tile_origin(tl, u0, v0);
      const long long cu = along_u ? u0 : v0, co = along_u ? v0 : u0;  // coordinate along the unit dimension / the other one
      fence_async_smem();
      mbar_expect_tx(bar, tile_bytes);
      const int c0 = (int)(cu * (r2c ? 1 : 2));  // inner coordinate in scalars of type T
#pragma unroll 1
      for (int bx = 0; bx < Cfg::NBOX; bx++) {
        // dims 1 and 2 of the map are ordered by stride: the transform dimension is dim 2 when Q.tl_swap is set
        const int cd = bx * Cfg::BOXROWS;
        tensor_g2s(reinterpret_cast<unsigned char *>(B) + (size_t)bx * Cfg::BOXROWS * Cfg::row_bytes, &tmap, c0, Q.tl_swap ? (int)co : cd,
                   Q.tl_swap ? cd : (int)co, bar);
      }
    }
  };
  C wt = mk<T>((T)1, (T)0);
  if constexpr (r2c) wt = tw[tS];
  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();

  unsigned parity = 0;
  issue(blockIdx.x);
  for (long long tile = blockIdx.x; tile < Q.ntiles; tile += gridDim.x) {
    const long long nxt = tile + gridDim.x;
    long long u0, v0;
    tile_origin(tile, u0, v0);
    const long long uo = u0 + (along_u ? slotS : 0), vo = v0 + (along_u ? 0 : slotS);
    const bool live = uo < Q.nu && vo < Q.nv;
    C v[E];
    mbar_wait(bar, parity);
    parity ^= 1;
    // ---------------- landed tile [row][P] -> registers, lanes across pencils
    if constexpr (r2c) {
      const T *LR = reinterpret_cast<const T *>(B);
#pragma unroll
      for (int m = 0; m < E; m++) {
        const int j = tL + m * TP;
        v[m] = mk<T>(LR[(2 * j) * P + slotL], LR[(2 * j + 1) * P + slotL]);
      }
    } else {
#pragma unroll
      for (int m = 0; m < E; m++) {
        const C x = B[(tL + m * TP) * P + slotL];
        v[m] = bwd ? cconj(x) : x;
      }
    }
    __syncthreads();  // the tile is in registers: the same memory now carries the per-pencil padded exchange buffers
    // ---------------- first pass in the landed mapping; its exchange re-maps to pencil-major
    reg_pass<T, M, E, R1, false>(v, tL, 1, tw, twscale);
    smem_scatter<T, M, E, R1>(v, BL, tL, 1);
    __syncthreads();
    if constexpr (R3 > 1) {
      smem_gather<T, M, E>(v, BA, tA);
      reg_pass2<T, M, E, R1, R2>(v, tA, T2);
      __syncthreads();
      smem_scatter<T, M, E, R2>(v, BA, tA, R1);
      __syncthreads();
    }
    smem_gather<T, M, E>(v, BS, tS);
    if constexpr (!r2c) {
      __syncthreads();  // every value is back in registers: the memory is free for the next tile
      issue(nxt);
    }
    if constexpr (R3 > 1) reg_pass3<T, M, E, R3>(v, tS, T3);
    else reg_pass2<T, M, E, R1, R2>(v, tS, T2);
    // v[m] = forward core output F[tS + m*TP] of pencil slotS

    if constexpr (r2c) {
      // X[k] = ((Z[k] + conj Z[M-k]) - i e^{-2 pi i k/N} (Z[k] - conj Z[M-k])) / 2, k = 0..M (partners through the buffer;
      // a thread overwrites exactly the locations its own gather has read)
#pragma unroll
      for (int m = 0; m < E; m++) BS[padidx(tS + m * TP)] = v[m];
      __syncthreads();
      C xM = mk<T>((T)0, (T)0);
#pragma unroll
      for (int m = 0; m < E; m++) {
        const int k = tS + m * TP;
        const C zk = v[m];
        const C zm = cconj(BS[padidx((M - k) & (M - 1))]);
        const C s = cadd(zk, zm), d = csub(zk, zm);
        const C e = cmulmi(cmul(d, real_twiddle<T, E>(wt, m)));
        v[m] = mk<T>((T)0.5 * (s.x + e.x), (T)0.5 * (s.y + e.y));
        if (m == 0) xM = mk<T>(zk.x - zk.y, (T)0);
      }
      __syncthreads();
      issue(nxt);
      if (live) {
#pragma unroll
        for (int m = 0; m < E; m++) store_out<T>(Q, tS + m * TP, uo, vo, v[m]);
        if (tS == 0) store_out<T>(Q, M, uo, vo, xM);
      }
    } else if (live) {
      if (Q.nseg == 1 && Q.deriv_g <= 0) {
        const SegDev &sg = Q.seg[0];
        C *out = (C *)sg.base + sg.off + uo * sg.os_u + vo * sg.os_v + (long long)tS * sg.os_d;
        const long long step = (long long)TP * sg.os_d;
#pragma unroll
        for (int m = 0; m < E; m++) st_out(out + m * step, bwd ? cconj(v[m]) : v[m]);
      } else {
#pragma unroll
        for (int m = 0; m < E; m++) store_out<T>(Q, tS + m * TP, uo, vo, bwd ? cconj(v[m]) : v[m]);
      }
    }
  }
}

// ------------------------------------------------------------------ host side
struct TLoadInfo {
  void (*launch)(const StageParams &, const TMapArg &, int grid, cudaStream_t);
  const void *func;
  int threads, ts, minb, rows, boxrows;
  size_t smem;
};

template <typename T, int M, int KIND, int P, int TS> void tload_launcher(const StageParams &Q, const TMapArg &tm, int grid, cudaStream_t s) {
  typedef TLoadCfg<T, M, KIND, P, TS> Cfg;
  P3B_LAUNCH2((pow2_tload_kernel<T, M, KIND, P, TS>), grid, Cfg::THREADS, Cfg::smem, s, Q, tm);
}

template <typename T, int M, int KIND, int P, int TS> const TLoadInfo *tload_info_one() {
  typedef TLoadCfg<T, M, KIND, P, TS> Cfg;
  if constexpr (!Cfg::valid) {
    return nullptr;
  } else {
    static const TLoadInfo info = {tload_launcher<T, M, KIND, P, TS>, (const void *)pow2_tload_kernel<T, M, KIND, P, TS>, Cfg::THREADS, TS,
                                   Cfg::MINB, Cfg::ROWS, Cfg::BOXROWS, Cfg::smem};
    return &info;
  }
}

template <typename T, int M, int KIND, int TS> const TLoadInfo *tload_info_m(int P) {
  switch (P) {
    case 4: return tload_info_one<T, M, KIND, 4, TS>();
    case 8: return tload_info_one<T, M, KIND, 8, TS>();
    case 16: return tload_info_one<T, M, KIND, 16, TS>();
  }
  return nullptr;
}

template <typename T, int KIND, int TS> const TLoadInfo *tload_info(int M, int P) {
  switch (M) {
    case 64: return tload_info_m<T, 64, KIND, TS>(P);
    case 128: return tload_info_m<T, 128, KIND, TS>(P);
    case 256: return tload_info_m<T, 256, KIND, TS>(P);
    case 512: return tload_info_m<T, 512, KIND, TS>(P);
    case 1024: return tload_info_m<T, 1024, KIND, TS>(P);
    case 2048: return tload_info_m<T, 2048, KIND, TS>(P);
    case 4096: return tload_info_m<T, 4096, KIND, TS>(P);
  }
  return nullptr;
}

// defined in pow2_tload_inst.cu, compiled once per (precision, kind in 1..3)
const TLoadInfo *tload_lookup(int prec, int kind, int ts, int M, int P);

}  // namespace p3b
