// gpu_layer.cu -- implementation of the p3dfftcu_* C ABI (include/p3dfft_b200.h): device selection,
// memory, twiddle tables, stage planning (kernel variant + tile shape) and launch, stand-alone
// spectral derivative, CUDA-IPC peer mapping and the stream-ordered peer barrier.
#include <cuda_runtime.h>
#ifndef P3B_EMU
#include <cuda.h>  // CUtensorMap types only; cuTensorMapEncodeTiled is reached through cudaGetDriverEntryPoint
#endif
#include <time.h>

#include <atomic>
#include <cmath>
#include <condition_variable>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "generic_stage.cuh"
#include "pow2_stage.cuh"
#include "pow2_pipe.cuh"
#include "mixed_pipe.cuh"
#include "pow2_tload.cuh"
#include "fastcore_stage.cuh"

using namespace p3b;

namespace p3b {
#define P3B_DECL_PIPE(p, k) const PipeInfo *pipe_lookup_p##p##_##k(int ts, int M, int P);
P3B_DECL_PIPE(4, 1) P3B_DECL_PIPE(4, 2) P3B_DECL_PIPE(4, 3) P3B_DECL_PIPE(4, 4) P3B_DECL_PIPE(4, 5) P3B_DECL_PIPE(4, 13)
P3B_DECL_PIPE(8, 1) P3B_DECL_PIPE(8, 2) P3B_DECL_PIPE(8, 3) P3B_DECL_PIPE(8, 4) P3B_DECL_PIPE(8, 5) P3B_DECL_PIPE(8, 13)
#undef P3B_DECL_PIPE
const PipeInfo *pipe_lookup(int prec, int kind, int ts, int M, int P) {
  // r2r kinds: one kernel family with the kind read at run time (the caller asks for kPipeDCT1 to get the compile-time DCT-I)
  if (kind >= P3DFFTCU_K_DCT1 && kind != kPipeDCT1) kind = kPipeR2R;
  if (prec == 4) switch (kind) {
      case 1: return pipe_lookup_p4_1(ts, M, P);
      case 2: return pipe_lookup_p4_2(ts, M, P);
      case 3: return pipe_lookup_p4_3(ts, M, P);
      case 4: return pipe_lookup_p4_4(ts, M, P);
      case 5: return pipe_lookup_p4_5(ts, M, P);
      case 13: return pipe_lookup_p4_13(ts, M, P);
    }
  if (prec == 8) switch (kind) {
      case 1: return pipe_lookup_p8_1(ts, M, P);
      case 2: return pipe_lookup_p8_2(ts, M, P);
      case 3: return pipe_lookup_p8_3(ts, M, P);
      case 4: return pipe_lookup_p8_4(ts, M, P);
      case 5: return pipe_lookup_p8_5(ts, M, P);
      case 13: return pipe_lookup_p8_13(ts, M, P);
    }
  return nullptr;
}
#define P3B_DECL_TL(p, k) const TLoadInfo *tload_lookup_p##p##_##k(int ts, int M, int P);
P3B_DECL_TL(4, 1) P3B_DECL_TL(4, 2) P3B_DECL_TL(4, 3) P3B_DECL_TL(8, 1) P3B_DECL_TL(8, 2) P3B_DECL_TL(8, 3)
#undef P3B_DECL_TL
const TLoadInfo *tload_lookup(int prec, int kind, int ts, int M, int P) {
  if (prec == 4) switch (kind) {
      case 1: return tload_lookup_p4_1(ts, M, P);
      case 2: return tload_lookup_p4_2(ts, M, P);
      case 3: return tload_lookup_p4_3(ts, M, P);
    }
  if (prec == 8) switch (kind) {
      case 1: return tload_lookup_p8_1(ts, M, P);
      case 2: return tload_lookup_p8_2(ts, M, P);
      case 3: return tload_lookup_p8_3(ts, M, P);
    }
  return nullptr;
}
#define P3B_DECL_MIX(p, k) const PipeInfo *mixed_lookup_p##p##_##k(int ts, int Q, int MC, int P);
P3B_DECL_MIX(4, 1) P3B_DECL_MIX(4, 2) P3B_DECL_MIX(4, 3) P3B_DECL_MIX(4, 4)
P3B_DECL_MIX(8, 1) P3B_DECL_MIX(8, 2) P3B_DECL_MIX(8, 3) P3B_DECL_MIX(8, 4)
#undef P3B_DECL_MIX
const PipeInfo *mixed_lookup(int prec, int kind, int ts, int Q, int MC, int P) {
  if (prec == 4) switch (kind) {
      case 1: return mixed_lookup_p4_1(ts, Q, MC, P);
      case 2: return mixed_lookup_p4_2(ts, Q, MC, P);
      case 3: return mixed_lookup_p4_3(ts, Q, MC, P);
      case 4: return mixed_lookup_p4_4(ts, Q, MC, P);
    }
  if (prec == 8) switch (kind) {
      case 1: return mixed_lookup_p8_1(ts, Q, MC, P);
      case 2: return mixed_lookup_p8_2(ts, Q, MC, P);
      case 3: return mixed_lookup_p8_3(ts, Q, MC, P);
      case 4: return mixed_lookup_p8_4(ts, Q, MC, P);
    }
  return nullptr;
}
}  // namespace p3b

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};
int g_device = -1;
int g_num_sms = 148;
size_t g_smem_optin = 227 * 1024;

int fail(const char *what, cudaError_t e) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return 1;
}
int failmsg(const std::string &m) {
  g_err = m;
  return 1;
}
#define CK(call)                                   \
  do {                                             \
    cudaError_t e__ = (call);                      \
    if (e__ != cudaSuccess) return fail(#call, e__); \
  } while (0)

// ------------------------------------------------------------------ twiddle tables
// which: 0 -> exp(-2 pi i j/n), j<n;  1 -> exp(-i pi j/(2n)), j<2n;  2 -> exp(-i pi (2j+1)/(4n)), j<n
std::mutex g_tw_mu;
std::map<std::tuple<int, int, int>, void *> g_tw;

template <typename T> int build_table(int which, int n, void **out) {
  int len = which == 1 ? 2 * n : n;
  std::vector<T> h(2 * (size_t)len);
  const long double pi = 3.14159265358979323846264338327950288L;
  for (int j = 0; j < len; j++) {
    long double a;
    if (which == 0) a = -2.0L * pi * (long double)j / (long double)n;
    else if (which == 1) a = -pi * (long double)j / (2.0L * (long double)n);
    else a = -pi * (long double)(2 * j + 1) / (4.0L * (long double)n);
    h[2 * j] = (T)cosl(a);
    h[2 * j + 1] = (T)sinl(a);
  }
  void *d = nullptr;
  CK(cudaMalloc(&d, h.size() * sizeof(T)));
  CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = d;
  return 0;
}

// Bluestein tables for length L on a power-of-two core of M points (M >= 2L-1):
//   chirp[j] = exp(-i pi j^2 / L), j < L;  bhat = FFT_M(h) / M with h[i] = h[M-i] = conj(chirp[i]) for i < L, 0 elsewhere
std::map<std::tuple<int, int, int>, std::pair<void *, void *>> g_blue;

void fft_ld(std::vector<long double> &re, std::vector<long double> &im) {  // in-place radix-2, forward, length 2^k
  const size_t n = re.size();
  for (size_t i = 1, j = 0; i < n; i++) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
  }
  const long double pi = 3.14159265358979323846264338327950288L;
  for (size_t len = 2; len <= n; len <<= 1) {
    for (size_t i = 0; i < n; i += len)
      for (size_t k = 0; k < len / 2; k++) {
        const long double a = -2.0L * pi * (long double)k / (long double)len, wr = cosl(a), wi = sinl(a);
        const size_t p = i + k, q = i + k + len / 2;
        const long double xr = re[q] * wr - im[q] * wi, xi = re[q] * wi + im[q] * wr;
        re[q] = re[p] - xr; im[q] = im[p] - xi;
        re[p] += xr; im[p] += xi;
      }
  }
}

template <typename T> int build_blue(int L, int M, void **chirp_out, void **bhat_out) {
  const long double pi = 3.14159265358979323846264338327950288L;
  std::vector<long double> cr(L), ci(L), hr(M, 0.0L), hi(M, 0.0L);
  for (int j = 0; j < L; j++) {
    const long long q = ((long long)j * j) % (2LL * L);  // j^2 mod 2L keeps the argument small
    const long double a = -pi * (long double)q / (long double)L;
    cr[j] = cosl(a);
    ci[j] = sinl(a);
    hr[j] = cr[j];
    hi[j] = -ci[j];
    if (j > 0) { hr[M - j] = cr[j]; hi[M - j] = -ci[j]; }
  }
  fft_ld(hr, hi);
  std::vector<T> hc(2 * (size_t)L), hb(2 * (size_t)M);
  for (int j = 0; j < L; j++) { hc[2 * j] = (T)cr[j]; hc[2 * j + 1] = (T)ci[j]; }
  for (int j = 0; j < M; j++) { hb[2 * j] = (T)(hr[j] / M); hb[2 * j + 1] = (T)(hi[j] / M); }
  void *dc = nullptr, *db = nullptr;
  CK(cudaMalloc(&dc, hc.size() * sizeof(T)));
  CK(cudaMemcpy(dc, hc.data(), hc.size() * sizeof(T), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&db, hb.size() * sizeof(T)));
  CK(cudaMemcpy(db, hb.data(), hb.size() * sizeof(T), cudaMemcpyHostToDevice));
  *chirp_out = dc;
  *bhat_out = db;
  return 0;
}

int get_blue(int L, int M, int prec, const void **chirp, const void **bhat) {
  std::lock_guard<std::mutex> lk(g_tw_mu);
  auto key = std::make_tuple(L, M, prec);
  auto it = g_blue.find(key);
  if (it == g_blue.end()) {
    void *c = nullptr, *b = nullptr;
    int rc = prec == 8 ? build_blue<double>(L, M, &c, &b) : build_blue<float>(L, M, &c, &b);
    if (rc) return rc;
    it = g_blue.emplace(key, std::make_pair(c, b)).first;
  }
  *chirp = it->second.first;
  *bhat = it->second.second;
  return 0;
}

int get_table(int which, int n, int prec, const void **out) {
  std::lock_guard<std::mutex> lk(g_tw_mu);
  auto key = std::make_tuple(which, n, prec);
  auto it = g_tw.find(key);
  if (it != g_tw.end()) {
    *out = it->second;
    return 0;
  }
  void *d = nullptr;
  int rc = prec == 8 ? build_table<double>(which, n, &d) : build_table<float>(which, n, &d);
  if (rc) return rc;
  g_tw[key] = d;
  *out = d;
  return 0;
}

// ------------------------------------------------------------------ deriv + barrier kernels
struct DerivParams {
  const void *in;
  void *out;
  long long total;
  int sd0, sd1, ldir, g, gstart;
};

template <typename T> __global__ void deriv_kernel(const __grid_constant__ DerivParams p) {
  typedef typename cx<T>::type C;
  const C *in = (const C *)p.in;
  C *out = (C *)p.out;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += (long long)gridDim.x * blockDim.x) {
    int c;
    if (p.ldir == 0) c = (int)(i % p.sd0);
    else if (p.ldir == 1) c = (int)((i / p.sd0) % p.sd1);
    else c = (int)(i / ((long long)p.sd0 * p.sd1));
    T kap = (T)deriv_kappa(c + p.gstart, p.g);
    C x = in[i];
    out[i] = mk<T>(-kap * x.y, kap * x.x);
  }
}

struct BarrierParams {
  unsigned long long *peer[64];
  int slot[64];
  unsigned long long epoch[64];
  unsigned long long *mine;
  int n, my_slot;
  unsigned long long timeout_ns;  // 0: wait for ever
};

struct PublishParams {
  unsigned long long *ptr[P3B_MAXSRC];
  unsigned long long epoch[P3B_MAXSRC];
  int ids[P3B_MAXGRP];
  int n, nids;
};

#ifndef P3B_EMU
__global__ void peer_barrier_kernel(const __grid_constant__ BarrierParams p) {
  int j = threadIdx.x;
  if (j >= p.n) return;
  __threadfence_system();
  unsigned long long *remote = p.peer[j] + p.my_slot;
  const unsigned long long epoch = p.epoch[j];
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote), "l"(epoch) : "memory");
  const unsigned long long *mine = p.mine + p.slot[j];
  unsigned long long t0, t1, v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
    if (v >= epoch) break;
    // like MPI_Alltoallv the barrier waits for ever by default (a peer may legitimately be late: a checkpoint, a debugger,
    // a lazy module load); P3DFFT_B200_PEER_TIMEOUT_S=<seconds> makes a dead peer fail loudly instead of hanging the GPU
    if (p.timeout_ns) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > p.timeout_ns) {
        printf("p3dfft_b200: peer barrier timed out (slot %d waiting for slot %d, epoch %llu, seen %llu)\n", p.my_slot, p.slot[j],
               epoch, v);
        __trap();
      }
    }
    __nanosleep(200);
  }
}

__global__ void flags_publish_kernel(const __grid_constant__ PublishParams p) {
  const int j = threadIdx.x / P3B_MAXGRP, i = threadIdx.x % P3B_MAXGRP;
  if (j >= p.n || i >= p.nids) return;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.ptr[j] + p.ids[i]), "l"(p.epoch[j]) : "memory");
}
#endif

unsigned long long g_peer_timeout_ns = [] {
  const char *e = getenv("P3DFFT_B200_PEER_TIMEOUT_S");
  const double s = e ? atof(e) : 0.0;
  return s > 0 ? (unsigned long long)(s * 1e9) : 0ull;
}();

// ------------------------------------------------------------------ stage object
enum Variant { V_GENERIC = 0, V_POW2 = 1, V_PIPE = 2, V_FAST = 3, V_TLOAD = 4 };

struct PipePlan {
  const p3b::PipeInfo *info = nullptr;
  int M = 0, P = 0, ld = 0, grid = 0;
  int tile_u = 1, tile_v = 1, tu_log2 = 0, load_ord = 0, store_ord = 0;
  long long tiles_u = 0, tiles_v = 0, ntiles = 0;
  int vfast = 0;
  int bytes = 0;  // r2r kinds: bytes of one pencil's bulk copy
};

// tensor-load kernel (pow2_tload.cuh): tile geometry + the TMA tensor map of the input, encoded for the last `in` pointer
struct TLoadPlan {
  const p3b::TLoadInfo *info = nullptr;
  int M = 0, P = 0, grid = 0, along_u = 0, store_ord = 0, swap = 0;
  long long tiles_u = 0, tiles_v = 0, ntiles = 0;
  // geometry of the map: scalars of `prec` bytes; dim0 = the input's unit-stride dimension, dim1 = transform dimension, dim2 = the other
  unsigned long long dim[3] = {0, 0, 0}, stride_b[2] = {0, 0};
  unsigned box0 = 0, box1 = 0;
  const void *encoded_for = nullptr;
  p3b::TMapArg map;
};

struct FastPlan {
  p3b::FastInfo info;
  int M = 0, blue = 0, threads = 0, grid = 0;
  size_t smem = 0;
};

}  // namespace

struct p3dfftcu_stage_s {
  p3dfftcu_stage_desc d;
  StageParams P;
  Variant variant;
  Variant fallback = V_GENERIC;  // V_PIPE only: the variant that takes inputs a bulk copy cannot (pointer not 16-byte aligned)
  int threads;
  int grid;
  size_t smem;
  Pow2Plan pw;
  PipePlan pp;
  FastPlan fp;
  TLoadPlan tl;
  bool have_pw = false;  // the non-pipelined pow2 kernel is kept as the fallback for unaligned user pointers
  bool empty = false;    // no local pencils on this rank: exec is a no-op
  std::string name;
};

namespace {

void factorize(int L, int *fac, int *nfac) {
  int n = 0;
  while (L % 4 == 0) { fac[n++] = 4; L /= 4; }
  while (L % 2 == 0) { fac[n++] = 2; L /= 2; }
  for (int p = 3; (long long)p * p <= L; p += 2)
    while (L % p == 0) { fac[n++] = p; L /= p; }
  if (L > 1) fac[n++] = L;
  *nfac = n;
}

int internal_length(int kind, int n) {
  switch (kind) {
    case P3DFFTCU_K_DCT1: return 2 * (n - 1);
    case P3DFFTCU_K_DST1: return 2 * (n + 1);
    case P3DFFTCU_K_DCT2: case P3DFFTCU_K_DST2: case P3DFFTCU_K_DCT3: case P3DFFTCU_K_DST3:
    case P3DFFTCU_K_DCT4: case P3DFFTCU_K_DST4: return 2 * n;
    default: return n;
  }
}

// which of (d,u,v) is the unit-stride direction: 0 d, 1 u, 2 v
int fastest(long long sd, long long su, long long sv, long long nd, long long nu, long long nv) {
  long long best = -1;
  int which = 0;
  long long s[3] = {sd, su, sv}, n[3] = {nd, nu, nv};
  for (int i = 0; i < 3; i++) {
    if (n[i] <= 1) continue;
    if (best < 0 || s[i] < best) { best = s[i]; which = i; }
  }
  return which;
}

template <typename T> int setup_generic(p3dfftcu_stage_s *st) {
  const p3dfftcu_stage_desc &d = st->d;
  StageParams &P = st->P;
  size_t csz = 2 * sizeof(T);
  P.lstride = P.L + 1;
  size_t per_pencil = 2 * (size_t)P.lstride * csz;
  size_t budget = g_smem_optin - 1024;
  int pmax = (int)(budget / per_pencil);
  if (pmax < 1) return failmsg("transform length too large for the shared-memory stage kernel");
  int fin = fastest(d.is_d, d.is_u, d.is_v, d.n_in, d.nu, d.nv);
  int fout = fastest(d.seg[0].os_d, d.seg[0].os_u, d.seg[0].os_v, d.seg[0].k1 - d.seg[0].k0, d.nu, d.nv);
  int run = (int)(128 / csz);  // pencils per tile for 128-byte runs across pencils
  if (run < 4) run = 4;
  int tu = 1, tv = 1;
  auto clampll = [](long long a, long long b) { return (int)(a < b ? a : b); };
  bool needU = (fin == 1 || fout == 1), needV = (fin == 2 || fout == 2);
  if (needU && needV) {
    int side = 1;
    while ((side + 1) * (side + 1) <= pmax && side + 1 <= 8) side++;
    tu = clampll(side, d.nu);
    tv = clampll(pmax / tu > 8 ? 8 : pmax / tu, d.nv);
  } else if (needU) {
    tu = clampll(pmax < run ? pmax : run, d.nu);
  } else if (needV) {
    tv = clampll(pmax < run ? pmax : run, d.nv);
  } else {
    int want = 4096 / (P.L > 0 ? P.L : 1);
    if (want < 1) want = 1;
    if (want > pmax) want = pmax;
    tu = clampll(want, d.nu);
    tv = clampll(want / tu > 0 ? want / tu : 1, d.nv);
  }
  P.tile_u = tu;
  P.tile_v = tv;
  P.load_ord = fin == 0 ? ORD_D : (fin == 1 ? ORD_U : ORD_V);
  P.store_ord = fout == 0 ? ORD_D : (fout == 1 ? ORD_U : ORD_V);
  P.tiles_u = (d.nu + tu - 1) / tu;
  P.ntiles = P.tiles_u * ((d.nv + tv - 1) / tv);
  st->threads = 256;
  st->smem = (size_t)tu * tv * per_pencil;
  auto kern = generic_stage_kernel<T>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_smem_optin));
  int occ = 1;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, st->threads, st->smem));
  if (occ < 1) occ = 1;
  long long g = (long long)g_num_sms * occ;
  st->grid = (int)(P.ntiles < g ? P.ntiles : g);
  if (st->grid < 1) st->grid = 1;
  char nm[160];
  snprintf(nm, sizeof nm, "generic<%s> L=%d tile=%dx%d load=%d store=%d smem=%zu grid=%d", sizeof(T) == 8 ? "f64" : "f32",
           P.L, tu, tv, P.load_ord, P.store_ord, st->smem, st->grid);
  st->name = nm;
  return 0;
}



int ilog2(int x) {
  int l = 0;
  while ((1 << l) < x) l++;
  return l;
}

// fastcore kernel (fastcore_stage.cuh): any kind whose internal FFT length L is a power of two in 64..4096, or any L with
// 2L-1 <= 4096 through Bluestein; fills st->P's tile shape like the generic kernel.  0 ok, <0 not applicable, >0 error
int fast_setup(p3dfftcu_stage_s *st) {
  const p3dfftcu_stage_desc &d = st->d;
  StageParams &P = st->P;
  if (d.kind == P3DFFTCU_K_EMPTY) return -1;
  const int L = P.L;
  if (L < 24) return -1;  // tiny transforms: the generic kernel's direct butterflies are as good
  int M = 64, blue = 0;
  if ((L & (L - 1)) == 0 && L >= 64 && L <= 4096) M = L;
  else {
    blue = 1;
    while (M < 2 * L - 1) M *= 2;
    if (M > 4096) return -1;
  }
  FastPlan &fp = st->fp;
  if (!fast_lookup(d.prec, M, blue, &fp.info)) return -1;
  const size_t csz = (size_t)d.prec * 2;
  const int TP = fp.info.tp;
  int fin = fastest(d.is_d, d.is_u, d.is_v, d.n_in, d.nu, d.nv);
  int fout = fastest(d.seg[0].os_d, d.seg[0].os_u, d.seg[0].os_v, d.seg[0].k1 - d.seg[0].k0, d.nu, d.nv);
  const bool needU = fin == 1 || fout == 1, needV = fin == 2 || fout == 2;
  // pencils per tile: 128-byte runs across pencils when a side is transposed (as many as 512 threads and the shared
  // memory hold, at most 16); 256-thread CTAs otherwise, so that two or three of them overlap their phases on an SM
  int np = (needU || needV) ? 512 / TP : 256 / TP;
  if (np > 16) np = 16;
  if (np < 1) np = 1;
  while (np > 1 && ((size_t)np * fp.info.pitch + fp.info.table_elems) * csz > g_smem_optin - 1024) np /= 2;
  if (((size_t)np * fp.info.pitch + fp.info.table_elems) * csz > g_smem_optin - 1024) return -1;
  if (np * TP < 32) return -1;
  int tu = np, tv = 1;
  if (needU && needV) {
    tu = 1;
    while (tu * tu < np) tu *= 2;
    tv = np / tu;
  } else if (needV) {
    tv = np;
    tu = 1;
  }
  P.tile_u = tu;
  P.tile_v = tv;
  P.tu_log2 = ilog2(tu);
  P.load_ord = fin == 0 ? ORD_D : (fin == 1 ? ORD_U : ORD_V);
  P.store_ord = fout == 0 ? ORD_D : (fout == 1 ? ORD_U : ORD_V);
  P.tiles_u = (d.nu + tu - 1) / tu;
  P.ntiles = P.tiles_u * ((d.nv + tv - 1) / tv);
  if (get_table(0, M, d.prec, &P.tw_core)) return 1;
  if (blue && get_blue(L, M, d.prec, &P.chirp, &P.bhat)) return 1;
  fp.M = M;
  fp.blue = blue;
  fp.threads = np * TP;
  fp.smem = ((size_t)np * fp.info.pitch + fp.info.table_elems) * csz;
  // one kernel function serves stages with different pencils per tile: allow the largest request once and for all
  if (cudaFuncSetAttribute(fp.info.func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_smem_optin) != cudaSuccess) return 1;
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fp.info.func, fp.threads, fp.smem) != cudaSuccess) return 1;
  if (occ < 1) occ = 1;
  long long g = (long long)g_num_sms * occ;
  fp.grid = (int)(P.ntiles < g ? (P.ntiles > 0 ? P.ntiles : 1) : g);
  char nm[200];
  snprintf(nm, sizeof nm, "fastcore<%s,L=%d,M=%d%s> threads=%d tile=%dx%d load=%d store=%d smem=%zu grid=%d occ=%d",
           d.prec == 8 ? "f64" : "f32", L, M, blue ? ",bluestein" : "", fp.threads, tu, tv, P.load_ord, P.store_ord, fp.smem, fp.grid, occ);
  st->name = nm;
  return 0;
}

// picks the pipelined kernel (pow2_pipe.cuh) for a stage; returns 0 ok, <0 not applicable, >0 CUDA error
int pipe_setup(const p3dfftcu_stage_desc &d, PipePlan *pp, std::string *name) {
  const bool r2r = d.kind >= P3DFFTCU_K_DCT1;
  const bool real = d.kind == P3DFFTCU_K_R2C || d.kind == P3DFFTCU_K_C2R;
  int M, mixQ = 0, mixMC = 0;
  if (r2r) {  // the kind's symmetric extension has a power-of-two length: DCT-I of 2^k+1 points, DST-I of 2^k-1, II-IV of 2^k
    M = internal_length(d.kind, d.nfft);
    if (M < 64 || M > 4096 || (M & (M - 1))) return -1;
  } else if (pow2_supported(d)) {
    M = real ? d.nfft / 2 : d.nfft;
  } else {
    // smooth lengths: one odd factor 3, 5, 7, 9 or 15 times a power of two 64...1024 (mixed_pipe.cuh)
    if (d.kind < P3DFFTCU_K_C2C_FWD || d.kind > P3DFFTCU_K_C2R) return -1;
    if (real && (d.nfft % 2 || (d.kind == P3DFFTCU_K_C2R && d.nseg != 1))) return -1;
    M = real ? d.nfft / 2 : d.nfft;
    const char *nomix = getenv("P3DFFT_B200_NO_MIXED");
    if (nomix && atoi(nomix)) return -1;
    for (int q : {3, 5, 7, 9, 15})
      if (M % q == 0) {
        const int mc = M / q;
        if (mc >= 64 && mc <= 1024 && (mc & (mc - 1)) == 0) { mixQ = q; mixMC = mc; }
      }
    if (!mixQ) return -1;
  }
  int fin = fastest(d.is_d, d.is_u, d.is_v, d.n_in, d.nu, d.nv);
  int fout = fastest(d.seg[0].os_d, d.seg[0].os_u, d.seg[0].os_v, d.seg[0].k1 - d.seg[0].k0, d.nu, d.nv);
  if (fin != 0 || d.is_d != 1) return -1;  // bulk copies move whole pencils: the transform dimension must be unit-stride
  // 16-byte aligned pencils of a multiple of 16 bytes.  esz = bytes of one input element
  const long long esz = (long long)d.prec * d.dt_in;
  if ((d.is_u * esz) % 16 || (d.is_v * esz) % 16) return -1;
  const long long pencil_bytes = (long long)d.n_in * esz, copy_bytes = (pencil_bytes + 15) / 16 * 16;
  if (pencil_bytes % 16) {
    // single-precision C2R (M+1 elements of 8 bytes), r2r pencils of 2^k+-1 values: the copy takes up to 15 bytes more, which
    // must belong to the array: true when the rows are padded (library-owned intermediates, planner.cpp) so that every
    // pencil is followed by a gap
    long long pitch = 0;  // smallest stride between two pencils
    if (d.nu > 1) pitch = d.is_u;
    if (d.nv > 1 && (pitch == 0 || d.is_v < pitch)) pitch = d.is_v;
    const bool roomy = pitch * esz >= copy_bytes;
    if (!(((d.kind == P3DFFTCU_K_C2R && d.prec == 4) || r2r) && roomy)) return -1;
  }
  pp->bytes = (int)copy_bytes;
  const int ts = fout != 0;
  const size_t csz = (size_t)d.prec * (r2r ? d.dt_out : 2);  // element size of the output runs
  const int E = mixQ ? mixed_values_per_thread(mixMC) : pow2_values_per_thread(M), TP = M / E;
  // transposed stores: runs of 128 bytes across the tile's pencils; contiguous stores: 256 threads per CTA
  // (single precision, local stages: 8 pencils = 64-byte runs with twice the CTAs per SM measured 4-14 % faster than 16
  //  pencils; exchange stages keep 128-byte runs where the tile fits: peer stores of 64-byte runs reach 580 instead of
  //  717 GB/s, tools/microbench/peer_store_bench.cu)
  int want = ts ? ((128 / csz > 8 && d.nseg == 1) ? 8 : (int)(128 / csz)) : (256 / TP > 0 ? 256 / TP : 1);
  if (d.whole_sm_ctas) {  // 512 threads at 128 registers (double) fill an SM
    const int w512 = 512 / TP > 0 ? 512 / TP : 1;
    if (w512 > want) want = w512;
  }
  if (const char *e = getenv(ts ? "P3DFFT_B200_POW2_PENCILS" : "P3DFFT_B200_POW2_PENCILS_PM")) {
    int w = atoi(e);
    if (w > 0) want = w;
  }
  const long long ext = fout == 2 ? d.nv : d.nu;
  while (want > 1 && want / 2 >= ext) want /= 2;
  if (want > 16) want = 16;
  const PipeInfo *info = nullptr;
  // DCT-I on complex data has a compile-time form; every other r2r kind (and DCT-I on real data) takes the run-time one
  const int lkind = r2r ? ((d.kind == P3DFFTCU_K_DCT1 && d.dt_in == 2 && d.dt_out == 2) ? kPipeDCT1 : kPipeR2R) : d.kind;
  auto lookup = [&](int p) { return mixQ ? mixed_lookup(d.prec, d.kind, ts, mixQ, mixMC, p) : pipe_lookup(d.prec, lkind, ts, M, p); };
  if (mixQ && !ts) want = 384 / TP >= 2 ? 384 / TP : 2;  // (CTA-wide barriers: one CTA of up to 384 threads per SM)
  if (mixQ) {  // round down to a power of two
    int w2 = 1;
    while (w2 * 2 <= want) w2 *= 2;
    want = w2;
  }
  int P = want;
  for (; P >= 1; P /= 2) {
    info = lookup(P);
    if (info && info->smem <= g_smem_optin) break;
    info = nullptr;
  }
  if (!info) {  // small cores need several pencils to fill a warp
    for (P = want * 2; P <= 16 && !info; P *= 2) {
      info = lookup(P);
      if (info && info->smem > g_smem_optin) info = nullptr;
      if (info) break;
    }
  }
  if (!info) return -1;
  int tu = P, tv = 1;
  if (fout == 2) {
    tv = P;
    tu = 1;
  }
  pp->info = info;
  pp->M = M;
  pp->P = P;
  pp->ld = ts;
  pp->tile_u = tu;
  pp->tile_v = tv;
  pp->tu_log2 = ilog2(tu);
  pp->load_ord = ORD_D;
  pp->store_ord = fout == 0 ? ORD_D : (fout == 1 ? ORD_U : ORD_V);
  pp->tiles_u = (d.nu + tu - 1) / tu;
  pp->tiles_v = (d.nv + tv - 1) / tv;
  pp->ntiles = pp->tiles_u * pp->tiles_v;
  pp->vfast = d.nv > 1 && (d.nu <= 1 || d.seg[0].os_v < d.seg[0].os_u);
  if (cudaFuncSetAttribute(info->func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)info->smem) != cudaSuccess) return 1;
  if (info->func_sync && cudaFuncSetAttribute(info->func_sync, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)info->smem) != cudaSuccess) return 1;
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, info->func, info->threads, info->smem) != cudaSuccess) return 1;
  if (occ < 1) occ = 1;
  long long g = (long long)g_num_sms * occ;
  pp->grid = (int)(pp->ntiles < g ? (pp->ntiles > 0 ? pp->ntiles : 1) : g);
  char nm[220];
  char mix[32] = "";
  if (mixQ) snprintf(mix, sizeof mix, ",%dx%d", mixQ, mixMC);
  snprintf(nm, sizeof nm, "pipe<%s,M=%d,P=%d,%s%s%s> threads=%d tile=%dx%d%s store=%d smem=%zu grid=%d occ=%d",
           d.prec == 8 ? "f64" : "f32", M, P, ts ? "transposed" : "contiguous", r2r ? (lkind == kPipeDCT1 ? ",r2r:dct1" : ",r2r") : "", mix,
           info->threads, tu, tv,
           pp->vfast ? " v-fast" : "", pp->store_ord, info->smem, pp->grid, occ);
  *name = nm;
  return 0;
}

// picks the tensor-load kernel (pow2_tload.cuh) for a power-of-two stage whose input is unit-stride along u or v instead of
// the transform dimension; returns 0 ok, <0 not applicable, >0 CUDA error
int tload_setup(const p3dfftcu_stage_desc &d, TLoadPlan *tl, std::string *name) {
  if (!pow2_supported(d) || d.kind == P3DFFTCU_K_C2R) return -1;
  const char *off = getenv("P3DFFT_B200_NO_TLOAD");
  if (off && atoi(off)) return -1;
  const bool r2c = d.kind == P3DFFTCU_K_R2C;
  const int M = r2c ? d.nfft / 2 : d.nfft;
  const int fin = fastest(d.is_d, d.is_u, d.is_v, d.n_in, d.nu, d.nv);
  const int fout = fastest(d.seg[0].os_d, d.seg[0].os_u, d.seg[0].os_v, d.seg[0].k1 - d.seg[0].k0, d.nu, d.nv);
  if (fin == 0) return -1;
  const bool along_u = fin == 1;
  if ((along_u ? d.is_u : d.is_v) != 1) return -1;
  const long long esz = (long long)d.prec * d.dt_in;
  const long long s_other = along_u ? d.is_v : d.is_u, n_unit = along_u ? d.nu : d.nv, n_other = along_u ? d.nv : d.nu;
  if ((d.is_d * esz) % 16 || (n_other > 1 && (s_other * esz) % 16)) return -1;
  const int ts = fout != 0;
  if (ts && fout != fin) return -1;  // transposed stores run across the SAME pencils: the output's unit-stride dimension must be the input's
  int want = (int)(128 / esz);        // 128-byte box rows
  if (want > 16) want = 16;
  while (want > 4 && want / 2 >= n_unit) want /= 2;
  const TLoadInfo *info = nullptr;
  int P = want;
  for (; P >= 4; P /= 2) {
    info = tload_lookup(d.prec, d.kind, ts, M, P);
    if (info && info->smem <= g_smem_optin && (P * esz) % 16 == 0) break;
    info = nullptr;
  }
  if (!info) return -1;
  tl->info = info;
  tl->M = M;
  tl->P = P;
  tl->along_u = along_u ? 1 : 0;
  tl->store_ord = fout == 0 ? ORD_D : (fout == 1 ? ORD_U : ORD_V);
  tl->tiles_u = along_u ? (d.nu + P - 1) / P : d.nu;
  tl->tiles_v = along_u ? d.nv : (d.nv + P - 1) / P;
  tl->ntiles = tl->tiles_u * tl->tiles_v;
  // tensor map geometry in scalars of `prec` bytes: dim0 = the unit-stride dimension; dims 1 and 2 ordered by stride
  const unsigned long long sc = r2c ? 1 : 2;
  const unsigned long long sd = (unsigned long long)(d.is_d * esz), so = (unsigned long long)((n_other > 1 ? s_other : d.is_d * d.n_in) * esz);
  tl->dim[0] = (unsigned long long)n_unit * sc;
  tl->box0 = (unsigned)(P * sc);
  tl->box1 = (unsigned)info->boxrows;
  const bool swap = so < sd;  // the other dimension has the smaller stride: it becomes dim 1
  tl->swap = swap ? 1 : 0;
  tl->dim[1] = swap ? (unsigned long long)n_other : (unsigned long long)d.n_in;
  tl->dim[2] = swap ? (unsigned long long)d.n_in : (unsigned long long)n_other;
  tl->stride_b[0] = swap ? so : sd;
  tl->stride_b[1] = swap ? sd : so;
  tl->encoded_for = nullptr;
  if (cudaFuncSetAttribute(info->func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)info->smem) != cudaSuccess) return 1;
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, info->func, info->threads, info->smem) != cudaSuccess) return 1;
  if (occ < 1) occ = 1;
  const long long g = (long long)g_num_sms * occ;
  tl->grid = (int)(tl->ntiles < g ? (tl->ntiles > 0 ? tl->ntiles : 1) : g);
  char nm[220];
  snprintf(nm, sizeof nm, "tload<%s,M=%d,P=%d,%s> threads=%d tile=%s%d rows=%d box=%ux%u%s smem=%zu grid=%d occ=%d", d.prec == 8 ? "f64" : "f32",
           M, P, ts ? "transposed" : "contiguous", info->threads, along_u ? "u" : "v", P, info->rows, tl->box0, tl->box1, swap ? " d=dim2" : "",
           info->smem, tl->grid, occ);
  *name = nm;
  return 0;
}

// encodes the TMA tensor map of the stage's input for `in` (cached: exec loops call with the same pointer)
int tload_encode(TLoadPlan *tl, const p3dfftcu_stage_desc &d, const void *in) {
  if (tl->encoded_for == in) return 0;
  const bool swap = tl->swap != 0;
#ifdef P3B_EMU
  TMapArg &m = tl->map;
  memset(&m, 0, sizeof m);
  m.base = (const unsigned char *)in;
  m.elem_bytes = (unsigned)d.prec;
  m.dim0 = (unsigned)tl->dim[0]; m.dim1 = (unsigned)tl->dim[1]; m.dim2 = (unsigned)tl->dim[2];
  m.stride1 = (long long)tl->stride_b[0]; m.stride2 = (long long)tl->stride_b[1];
  m.box0 = tl->box0;
  m.box1 = swap ? 1u : tl->box1;
  m.box2 = swap ? tl->box1 : 1u;
#else
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                               const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = [] {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess) {
      cudaGetLastError();
      fn = nullptr;
    }
    return (EncodeFn)fn;
  }();
  if (!encode) return failmsg("cuTensorMapEncodeTiled is not available");
  static_assert(sizeof(CUtensorMap) == sizeof(TMapArg), "tensor map size");
  cuuint64_t gdim[3] = {tl->dim[0], tl->dim[1], tl->dim[2]};
  cuuint64_t gstr[2] = {tl->stride_b[0], tl->stride_b[1]};
  cuuint32_t box[3] = {tl->box0, swap ? 1u : tl->box1, swap ? tl->box1 : 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(reinterpret_cast<CUtensorMap *>(&tl->map), d.prec == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                      const_cast<void *>(in), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char m[160];
    snprintf(m, sizeof m, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return failmsg(m);
  }
#endif
  tl->encoded_for = in;
  return 0;
}

}  // namespace

extern "C" {

const char *p3dfftcu_last_error(void) { return g_err.c_str(); }

int p3dfftcu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int p3dfftcu_init(int device) {
  int n = p3dfftcu_device_count();
  if (n <= 0) return failmsg("no CUDA device available");
  if (device < 0) {
    const char *lr = getenv("LOCAL_RANK");
    if (!lr) lr = getenv("P3DFFT_RANK");
    device = lr ? atoi(lr) % n : 0;
    int cur = 0;
    if (!lr && cudaGetDevice(&cur) == cudaSuccess) device = cur;
  }
  CK(cudaSetDevice(device));
  g_device = device;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  g_num_sms = prop.multiProcessorCount;
  g_smem_optin = prop.sharedMemPerBlockOptin;
  if (prop.major < 10) {
    char m[200];
    snprintf(m, sizeof m, "device %d (%s, sm_%d%d) is not a Blackwell sm_100 part; this library ships sm_100a code only",
             device, prop.name, prop.major, prop.minor);
    return failmsg(m);
  }
  return 0;
}

int p3dfftcu_malloc(void **ptr, size_t bytes) {
  cudaError_t e = cudaMalloc(ptr, bytes ? bytes : 16);
  if (e != cudaSuccess) {
    size_t fr = 0, tot = 0;
    cudaGetLastError();
    cudaMemGetInfo(&fr, &tot);
    char m[200];
    snprintf(m, sizeof m, "cudaMalloc of %zu bytes: %s (device memory: %zu free of %zu)", bytes, cudaGetErrorString(e), fr, tot);
    return failmsg(m);
  }
  return 0;
}
int p3dfftcu_free(void *ptr) {
  if (ptr) CK(cudaFree(ptr));
  return 0;
}
int p3dfftcu_memset(void *ptr, int value, size_t bytes, void *stream) {
  CK(cudaMemsetAsync(ptr, value, bytes, (cudaStream_t)stream));
  return 0;
}
int p3dfftcu_memcpy(void *dst, const void *src, size_t bytes, int kind, void *stream) {
  cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : (kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
  CK(cudaMemcpyAsync(dst, src, bytes, k, (cudaStream_t)stream));
  return 0;
}
int p3dfftcu_stream_sync(void *stream) {
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}
int p3dfftcu_pointer_is_device(const void *ptr) {
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? 1 : 0;
}

// ---- page-locked user arrays (LRU cache of cudaHostRegister'ed ranges) and the staging ring for arrays that cannot be locked
namespace {
struct PinRange {
  uintptr_t base;
  size_t bytes;
  unsigned long long used;
};
std::vector<PinRange> g_pins;
unsigned long long g_pin_tick = 0;
size_t pin_total() {
  size_t t = 0;
  for (const PinRange &r : g_pins) t += r.bytes;
  return t;
}
void pin_drop(size_t i) {
#ifndef P3B_EMU
  cudaHostUnregister((void *)g_pins[i].base);
  cudaGetLastError();
#endif
  g_pins.erase(g_pins.begin() + (long)i);
}
const size_t kRingChunk = (size_t)32 << 20;
const int kRingSlots = 4;
void *g_ring[kRingSlots] = {nullptr, nullptr, nullptr, nullptr};
cudaEvent_t g_ring_ev[kRingSlots];

// CPU side of the staging ring: one memcpy split over a few threads (a single core copies ~10 GB/s, PCIe 5 x16 moves ~50)
class CopyPool {
 public:
  explicit CopyPool(int n) {
    for (int i = 1; i < n; i++) workers_.emplace_back([this, i] { loop(i); });
    parts_.resize((size_t)n);
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      gen_++;
    }
    cv_.notify_all();
    for (std::thread &t : workers_) t.join();
  }
  void copy(void *dst, const void *src, size_t n) {
    const size_t T = parts_.size();
    if (T <= 1 || n < ((size_t)1 << 20)) {
      memcpy(dst, src, n);
      return;
    }
    const size_t per = ((n + T - 1) / T + 4095) / 4096 * 4096;
    {
      std::lock_guard<std::mutex> lk(mu_);
      for (size_t i = 0; i < T; i++) {
        const size_t off = i * per < n ? i * per : n, len = off + per < n ? per : n - off;
        parts_[i] = Part{(char *)dst + off, (const char *)src + off, len};
      }
      pending_ = (int)T - 1;
      gen_++;
    }
    cv_.notify_all();
    if (parts_[0].n) memcpy(parts_[0].dst, parts_[0].src, parts_[0].n);
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
  }

 private:
  struct Part {
    char *dst;
    const char *src;
    size_t n;
  };
  void loop(int i) {
    unsigned long long seen = 0;
    for (;;) {
      Part p;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        p = parts_[(size_t)i];
      }
      if (p.n) memcpy(p.dst, p.src, p.n);
      {
        std::lock_guard<std::mutex> lk(mu_);
        pending_--;
      }
      done_.notify_one();
    }
  }
  std::vector<std::thread> workers_;
  std::vector<Part> parts_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  unsigned long long gen_ = 0;
  int pending_ = 0;
  bool stop_ = false;
};
CopyPool *g_pool = nullptr;
int g_ranks_on_host = 1;
CopyPool &copy_pool() {
  if (!g_pool) {
    // default: the host's cores shared among the ranks of this job, at most 16 threads (measured on a 16-core box, 1024^3
    // double round trip, pinned 660 ms: 2 / 8 / 16 threads -> 2099 / 1034 / 830 ms; profiles/r02b_bench_c3_ring*.json)
    const char *e = getenv("P3DFFT_B200_HOST_THREADS");
    const int hw = (int)std::thread::hardware_concurrency();
    int n = e ? atoi(e) : (hw > 0 ? hw / (g_ranks_on_host > 0 ? g_ranks_on_host : 1) : 4);
    if (!e && n > 16) n = 16;
    if (hw > 0 && n > hw) n = hw;
    if (n < 1) n = 1;
    g_pool = new CopyPool(n);
  }
  return *g_pool;
}
}  // namespace

int p3dfftcu_host_is_pinned(const void *ptr) {
#ifdef P3B_EMU
  (void)ptr;
  return 1;
#else
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return at.type == cudaMemoryTypeHost ? 1 : 0;
#endif
}

int p3dfftcu_host_pin(const void *ptr, size_t bytes) {
#ifdef P3B_EMU
  (void)ptr; (void)bytes;
  return 0;
#else
  if (!ptr || !bytes) return 0;
  const uintptr_t a = (uintptr_t)ptr, e = a + bytes;
  for (size_t i = 0; i < g_pins.size(); i++) {
    PinRange &r = g_pins[i];
    if (a >= r.base && e <= r.base + r.bytes) {  // known range
      r.used = ++g_pin_tick;
      return 0;
    }
  }
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) cudaGetLastError();
  else if (at.type == cudaMemoryTypeHost) {
    // page-locked by the application (cudaMallocHost / its own cudaHostRegister) -- unless it is one of OUR ranges that this
    // request outgrows: then the range is re-registered at the new size below
    bool ours = false;
    for (size_t i = g_pins.size(); i-- > 0;)
      if (a < g_pins[i].base + g_pins[i].bytes && e > g_pins[i].base) {
        ours = true;
        pin_drop(i);
      }
    if (!ours) return 0;
  }
  static const size_t cap = [] {
    const char *v = getenv("P3DFFT_B200_HOST_PIN_MAX_GB");
    const double gb = v ? atof(v) : 64.0;
    return (size_t)(gb * 1073741824.0);
  }();
  if (bytes > cap) return 1;
  for (size_t i = g_pins.size(); i-- > 0;)  // a stale overlapping range (the application re-allocated): drop it
    if (a < g_pins[i].base + g_pins[i].bytes && e > g_pins[i].base) pin_drop(i);
  while (!g_pins.empty() && (g_pins.size() >= 32 || pin_total() + bytes > cap)) {
    size_t lru = 0;
    for (size_t i = 1; i < g_pins.size(); i++)
      if (g_pins[i].used < g_pins[lru].used) lru = i;
    pin_drop(lru);
  }
  cudaError_t rc = cudaHostRegister((void *)ptr, bytes, cudaHostRegisterDefault);
  if (rc != cudaSuccess) {
    cudaGetLastError();
    return 1;
  }
  g_pins.push_back(PinRange{a, bytes, ++g_pin_tick});
  return 0;
#endif
}

int p3dfftcu_host_unpin(const void *ptr) {
  const uintptr_t a = (uintptr_t)ptr;
  for (size_t i = g_pins.size(); i-- > 0;)
    if (a >= g_pins[i].base && a < g_pins[i].base + g_pins[i].bytes) pin_drop(i);
  return 0;
}

void p3dfftcu_host_ranks_hint(int ranks_on_this_host) { g_ranks_on_host = ranks_on_this_host > 0 ? ranks_on_this_host : 1; }

int p3dfftcu_host_unpin_all(void) {
  while (!g_pins.empty()) pin_drop(g_pins.size() - 1);
#ifndef P3B_EMU
  for (int i = 0; i < kRingSlots; i++)
    if (g_ring[i]) {
      cudaFreeHost(g_ring[i]);
      cudaEventDestroy(g_ring_ev[i]);
      g_ring[i] = nullptr;
    }
  delete g_pool;
  g_pool = nullptr;
#endif
  return 0;
}

int p3dfftcu_memcpy_staged(void *dst, const void *src, size_t bytes, int kind, void *stream) {
#ifdef P3B_EMU
  return p3dfftcu_memcpy(dst, src, bytes, kind, stream) || p3dfftcu_stream_sync(stream);
#else
  cudaStream_t cs = (cudaStream_t)stream;
  for (int i = 0; i < kRingSlots; i++)
    if (!g_ring[i]) {
      CK(cudaMallocHost(&g_ring[i], kRingChunk));
      CK(cudaEventCreateWithFlags(&g_ring_ev[i], cudaEventDisableTiming));
    }
  CopyPool &pool = copy_pool();
  const size_t nchunk = (bytes + kRingChunk - 1) / kRingChunk;
  auto span = [&](size_t c, size_t *off, size_t *len) {
    *off = c * kRingChunk;
    *len = bytes - *off < kRingChunk ? bytes - *off : kRingChunk;
  };
  if (kind == 0) {  // host -> device: CPU copy into slot c % slots (once its previous DMA has drained), then DMA
    for (size_t c = 0; c < nchunk; c++) {
      size_t off, len;
      span(c, &off, &len);
      const int b = (int)(c % kRingSlots);
      if (c >= (size_t)kRingSlots) CK(cudaEventSynchronize(g_ring_ev[b]));
      pool.copy(g_ring[b], (const char *)src + off, len);
      CK(cudaMemcpyAsync((char *)dst + off, g_ring[b], len, cudaMemcpyHostToDevice, cs));
      CK(cudaEventRecord(g_ring_ev[b], cs));
    }
    CK(cudaStreamSynchronize(cs));
  } else {  // device -> host: the DMAs of the next slots run while the CPU copies chunk c out of the ring
    const size_t ahead = (size_t)kRingSlots - 1;
    for (size_t c = 0; c < nchunk + ahead; c++) {
      if (c < nchunk) {
        size_t off, len;
        span(c, &off, &len);
        const int b = (int)(c % kRingSlots);
        CK(cudaMemcpyAsync(g_ring[b], (const char *)src + off, len, cudaMemcpyDeviceToHost, cs));
        CK(cudaEventRecord(g_ring_ev[b], cs));
      }
      if (c >= ahead) {
        size_t off, len;
        span(c - ahead, &off, &len);
        const int b = (int)((c - ahead) % kRingSlots);
        CK(cudaEventSynchronize(g_ring_ev[b]));
        pool.copy((char *)dst + off, g_ring[b], len);
      }
    }
  }
  return 0;
#endif
}

int p3dfftcu_stage_create(const p3dfftcu_stage_desc *desc, p3dfftcu_stage *out) {
  const p3dfftcu_stage_desc &d = *desc;
  if (d.prec != 4 && d.prec != 8) return failmsg("stage: prec must be 4 or 8");
  if (d.nseg < 1 || d.nseg > P3DFFTCU_MAXSEG) return failmsg("stage: bad segment count");
  if (d.kind < 0 || d.kind > P3DFFTCU_K_DST4) return failmsg("stage: unknown kind");
  if (d.nfft < 1) return failmsg("stage: transform length must be positive");
  if (d.kind == P3DFFTCU_K_DCT1 && d.nfft < 2) return failmsg("stage: DCT-I needs n >= 2");
  p3dfftcu_stage_s *st = new p3dfftcu_stage_s();
  st->d = d;
  StageParams &P = st->P;
  memset(&P, 0, sizeof P);
  P.nu = d.nu; P.nv = d.nv; P.is_d = d.is_d; P.is_u = d.is_u; P.is_v = d.is_v;
  P.kind = d.kind; P.dt_in = d.dt_in; P.dt_out = d.dt_out;
  P.nfft = d.nfft; P.n_in = d.n_in; P.n_out = d.n_out;
  P.L = internal_length(d.kind, d.nfft);
  P.nseg = d.nseg;
  for (int s = 0; s < d.nseg; s++) {
    P.seg[s].k0 = d.seg[s].k0; P.seg[s].k1 = d.seg[s].k1;
    P.seg[s].off = d.seg[s].off; P.seg[s].os_d = d.seg[s].os_d; P.seg[s].os_u = d.seg[s].os_u; P.seg[s].os_v = d.seg[s].os_v;
    P.seg[s].base = nullptr;
  }
  if (d.nu <= 0 || d.nv <= 0 || d.n_in <= 0) {
    // this rank holds no pencil of the stage (more ranks than planes along a distributed dimension): nothing to launch
    st->variant = V_GENERIC;
    st->threads = 0;
    st->grid = 0;
    st->smem = 0;
    P.ntiles = 0;
    st->name = "empty (no local pencils)";
    st->empty = true;
    *out = st;
    return 0;
  }
  int rc = 0;
  if (d.kind != P3DFFTCU_K_EMPTY) {
    factorize(P.L, P.fac, &P.nfac);
    rc = get_table(0, P.L, d.prec, &P.tw);
    if (!rc && d.kind >= P3DFFTCU_K_DCT2) rc = get_table(1, d.nfft, d.prec, &P.tw2);
    if (!rc && d.kind >= P3DFFTCU_K_DCT4) rc = get_table(2, d.nfft, d.prec, &P.tw3);
  }
  st->variant = V_GENERIC;
  if (!rc) {
    const char *force = getenv("P3DFFT_B200_FORCE_GENERIC");
    bool allow_fast = !(force && atoi(force));
    if (allow_fast && pow2_supported(d)) {
      rc = pow2_setup(d, g_num_sms, g_smem_optin, &st->pw, &st->name);
      if (!rc) {
        st->variant = V_POW2;
        st->have_pw = true;
      } else if (rc < 0) rc = 0;  // negative: not applicable, fall through to the generic kernel
      const char *nopipe = getenv("P3DFFT_B200_NO_PIPE");
      if (!rc && st->have_pw && !(nopipe && atoi(nopipe))) {
        std::string pname;
        int prc = pipe_setup(d, &st->pp, &pname);
        if (prc > 0) rc = failmsg("pipelined stage kernel setup failed");
        else if (prc == 0) {
          st->variant = V_PIPE;
          st->name = pname;
        } else {
          // unit stride along u or v instead of the transform dimension: TMA tensor loads of [row][P] tiles
          int trc = tload_setup(d, &st->tl, &pname);
          if (trc > 0) rc = failmsg("tensor-load stage kernel setup failed");
          else if (trc == 0) {
            st->variant = V_TLOAD;
            st->name = pname;
          }
        }
      }
    }
    if (!rc && st->variant == V_GENERIC && allow_fast) {  // r2r kinds, other lengths: the register core (+ Bluestein)
      const char *nofast = getenv("P3DFFT_B200_NO_FASTCORE");
      if (!(nofast && atoi(nofast))) {
        int frc = fast_setup(st);
        if (frc > 0) rc = failmsg("fastcore stage kernel setup failed");
        else if (frc == 0) st->variant = V_FAST;
      }
    }
    if (!rc && st->variant == V_GENERIC) rc = d.prec == 8 ? setup_generic<double>(st) : setup_generic<float>(st);
    // r2r kinds whose symmetric extension has a power-of-two length, unit-stride aligned pencils: the TMA-fed kernel; the
    // variant chosen above stays as the fallback for input pointers a bulk copy cannot take
    const char *nopipe = getenv("P3DFFT_B200_NO_PIPE"), *nor2r = getenv("P3DFFT_B200_NO_PIPE_R2R");
    const bool r2r_pipe = d.kind >= P3DFFTCU_K_DCT1 && !(nor2r && atoi(nor2r));
    const bool mixed_pipe = d.kind >= P3DFFTCU_K_C2C_FWD && d.kind <= P3DFFTCU_K_C2R && !pow2_supported(d);  // 3|5|7 x 2^k
    if (!rc && allow_fast && !(nopipe && atoi(nopipe)) && (r2r_pipe || mixed_pipe)) {
      std::string pname;
      int prc = pipe_setup(d, &st->pp, &pname);
      if (prc > 0) rc = failmsg("pipelined stage kernel setup failed");
      else if (prc == 0) {
        st->fallback = st->variant;
        st->variant = V_PIPE;
        st->name = pname;
      }
    }
  }
  if (rc) {
    delete st;
    return rc;
  }
  *out = st;
  return 0;
}

int p3dfftcu_stage_destroy(p3dfftcu_stage st) {
  delete st;
  return 0;
}

const char *p3dfftcu_stage_variant(p3dfftcu_stage st) { return st->name.c_str(); }

int p3dfftcu_stage_exec(p3dfftcu_stage st, const void *in, void *const *dst, int ndst, int deriv_g, void *stream) {
  return p3dfftcu_stage_exec_capped(st, in, dst, ndst, deriv_g, stream, 0);
}

int p3dfftcu_stage_exec_capped(p3dfftcu_stage st, const void *in, void *const *dst, int ndst, int deriv_g, void *stream,
                               int max_ctas) {
  if (st->empty) return 0;
  if (deriv_g > 0 && st->d.dt_out != 2) return failmsg("stage: spectral derivative needs complex output");
  StageParams P = st->P;
  P.in = in;
  P.deriv_g = deriv_g;
  for (int s = 0; s < P.nseg; s++) {
    int slot = st->d.seg[s].slot;
    if (slot < 0 || slot >= ndst || !dst[slot]) return failmsg("stage: missing destination buffer for a segment");
    P.seg[s].base = dst[slot];
  }
  if (P.ntiles == 0 && (st->variant == V_GENERIC || st->variant == V_FAST)) return 0;  // (the generic plan's tile count; other variants carry their own)
  cudaStream_t cs = (cudaStream_t)stream;
  Variant variant = st->variant;
  if (variant == V_PIPE) {
    // bulk copies need 16-byte aligned sources; an odd user pointer takes the non-pipelined kernel instead
    if (((uintptr_t)in) % 16) variant = st->have_pw ? V_POW2 : st->fallback;
  }
  if (variant == V_TLOAD) {
    // tensor maps need a 16-byte aligned base; a failed encode (driver too old, geometry refused) retires the variant
    if (((uintptr_t)in) % 16 || !st->tl.info || tload_encode(&st->tl, st->d, in)) {
      if (st->tl.info && !(((uintptr_t)in) % 16)) st->tl.info = nullptr;
      variant = V_POW2;
    }
  }
  if (variant == V_TLOAD) {
    const TLoadPlan &tl = st->tl;
    if (tl.ntiles > 0) {
      P.tile_u = tl.along_u ? tl.P : 1; P.tile_v = tl.along_u ? 1 : tl.P; P.tu_log2 = ilog2(P.tile_u);
      P.load_ord = tl.along_u ? ORD_U : ORD_V; P.store_ord = tl.store_ord;
      P.tiles_u = tl.tiles_u; P.tiles_v = tl.tiles_v; P.ntiles = tl.ntiles; P.tl_swap = tl.swap;
      int grid = tl.grid;
      if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
      tl.info->launch(P, tl.map, grid, cs);
    }
  } else if (variant == V_PIPE) {
    const PipePlan &pp = st->pp;
    if (pp.ntiles > 0) {
      P.tile_u = pp.tile_u; P.tile_v = pp.tile_v; P.tu_log2 = pp.tu_log2;
      P.load_ord = pp.load_ord; P.store_ord = pp.store_ord;
      P.tiles_u = pp.tiles_u; P.tiles_v = pp.tiles_v; P.vfast = pp.vfast; P.ntiles = pp.ntiles;
      P.pipe_bytes = pp.bytes;
      int grid = pp.grid;
      if (st->d.nseg > 1) {  // tuning: cap the CTAs of a fused exchange stage (NVLink-bound: may not need every SM)
        static const int xcap = getenv("P3DFFT_B200_XGRID") ? atoi(getenv("P3DFFT_B200_XGRID")) : 0;
        if (xcap > 0 && grid > xcap) grid = xcap;
      }
      if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
      pp.info->launch(P, grid, cs);
    }
  } else if (variant == V_FAST) {
    const FastPlan &fp = st->fp;
    if (P.ntiles > 0) {
      int grid = fp.grid;
      if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
      fp.info.launch(P, grid, fp.threads, fp.smem, cs);
    }
  } else if (variant == V_POW2) {
    Pow2Plan pw = st->pw;
    if (max_ctas > 0 && pw.grid > max_ctas) pw.grid = max_ctas;
    int rc = pow2_launch(pw, P, cs);
    if (rc) return failmsg(std::string("pow2 stage launch failed: ") + cudaGetErrorString(cudaGetLastError()));
  } else {
    const int grid = (max_ctas > 0 && st->grid > max_ctas) ? max_ctas : st->grid;
    if (st->d.prec == 8) P3B_LAUNCH(generic_stage_kernel<double>, grid, st->threads, st->smem, cs, P);
    else P3B_LAUNCH(generic_stage_kernel<float>, grid, st->threads, st->smem, cs, P);
  }
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}

// 0: no tile-group form; 1: yes, tiles by CTA striding (all CTAs must be resident); 2: yes, tiles from a counter (launches that
// share one control block share the work)
int p3dfftcu_stage_sync_capable(p3dfftcu_stage st) {
  if (st->empty || st->variant != V_PIPE || st->pp.ntiles <= 0 || !st->pp.info->launch_sync) return 0;
  if (st->pp.ld) return 2;  // transposed output: one CTA barrier per tile anyway, whole tiles are handed out
#ifdef P3B_EMU
  return 1;  // (the emulation's group barriers are CTA barriers: thread groups cannot take different numbers of pencils)
#else
  // contiguous output: a pencil's thread group takes single pencils, provided it is made of whole warps
  return st->pp.info->threads / st->pp.P >= 32 ? 2 : 1;
#endif
}

int p3dfftcu_stage_exec_sync(p3dfftcu_stage st, const void *in, void *const *dst, int ndst, int deriv_g, void *stream,
                             int max_ctas, const p3dfftcu_sync *sy) {
  if (!p3dfftcu_stage_sync_capable(st)) return failmsg("stage: this kernel variant has no tile-group form");
  if (((uintptr_t)in) % 16) return failmsg("stage: the tile-group form needs a 16-byte aligned input");
  if (deriv_g > 0 && st->d.dt_out != 2) return failmsg("stage: spectral derivative needs complex output");
  if (sy->ngroups < 1 || sy->ngroups > P3B_MAXGRP || sy->wait_n > P3B_MAXSRC || sy->sig_n > P3B_MAXSRC)
    return failmsg("stage: bad tile-group table");
  StageParams P = st->P;
  P.in = in;
  P.deriv_g = deriv_g;
  for (int s = 0; s < P.nseg; s++) {
    int slot = st->d.seg[s].slot;
    if (slot < 0 || slot >= ndst || !dst[slot]) return failmsg("stage: missing destination buffer for a segment");
    P.seg[s].base = dst[slot];
  }
  const PipePlan &pp = st->pp;
  P.tile_u = pp.tile_u; P.tile_v = pp.tile_v; P.tu_log2 = pp.tu_log2;
  P.load_ord = pp.load_ord; P.store_ord = pp.store_ord;
  P.tiles_u = pp.tiles_u; P.tiles_v = pp.tiles_v; P.vfast = pp.vfast;
  SyncDev Y;
  memset(&Y, 0, sizeof Y);
  Y.ngroups = sy->ngroups;
  Y.dynamic = p3dfftcu_stage_sync_capable(st) == 2 ? 1 : 0;
  Y.keep_ctas = 1 << 30;
  Y.ctl = (unsigned long long *)sy->ctl;
  Y.timeout_ns = g_peer_timeout_ns;
  Y.wait_base = (const unsigned long long *)sy->wait_base;
  Y.wait_n = sy->wait_n;
  Y.sig_n = sy->sig_n;
  for (int j = 0; j < sy->wait_n; j++) { Y.wait_off[j] = sy->wait_off[j]; Y.wait_epoch[j] = sy->wait_epoch[j]; }
  for (int j = 0; j < sy->sig_n; j++) { Y.sig_ptr[j] = (unsigned long long *)sy->sig_ptr[j]; Y.sig_epoch[j] = sy->sig_epoch[j]; }
  long long t0 = 0;
  for (int g = 0; g < sy->ngroups; g++) {
    const p3dfftcu_group &G = sy->grp[g];
    if (G.u0 < 0 || G.v0 < 0 || G.u1 > st->d.nu || G.v1 > st->d.nv) return failmsg("stage: tile group outside the stage");
    TileGroupDev &D = Y.grp[g];
    D.u0 = G.u0; D.u1 = G.u1; D.v0 = G.v0; D.v1 = G.v1;
    D.tiles_u = G.u1 > G.u0 ? (G.u1 - G.u0 + pp.tile_u - 1) / pp.tile_u : 0;
    D.tiles_v = G.v1 > G.v0 ? (G.v1 - G.v0 + pp.tile_v - 1) / pp.tile_v : 0;
    if (D.tiles_u == 0 || D.tiles_v == 0) D.tiles_u = D.tiles_v = 0;  // (an empty group: the host publishes its flag)
    D.tile0 = t0;
    D.wait_id = G.wait_id;
    D.signal_id = G.signal_id;
    D.count = G.count;
    D.after = G.after;
    t0 += (long long)D.tiles_u * D.tiles_v;
  }
  P.ntiles = t0;
  if (t0 == 0) return 0;
  const long long per_tile = (Y.dynamic && !pp.ld) ? pp.P : 1;  // contiguous-output kernels hand out single pencils
  int grid;
  {
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pp.info->func_sync, pp.info->threads, pp.info->smem) != cudaSuccess || occ < 1) occ = 1;
    grid = g_num_sms * occ;
  }
  if (t0 * per_tile < grid) grid = (int)(t0 * per_tile);
  if (max_ctas > 0 && grid > max_ctas) {
    if (Y.dynamic && sy->boost_ctas > max_ctas && sy->boost_groups > 0 && sy->boost_groups < sy->ngroups) {
      Y.keep_ctas = max_ctas;
      Y.boost_limit = (unsigned long long)(Y.grp[sy->boost_groups].tile0 * per_tile);
      if (grid > sy->boost_ctas) grid = sy->boost_ctas;
    } else grid = max_ctas;
  }
  pp.info->launch_sync(P, Y, grid, (cudaStream_t)stream);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}

int p3dfftcu_flags_publish(void *const *ptrs, const unsigned long long *epochs, int n, const int *ids, int nids, void *stream) {
  if (n > P3B_MAXSRC || nids > P3B_MAXGRP) return failmsg("flags publish: too many targets");
  if (n <= 0 || nids <= 0) return 0;
#ifndef P3B_EMU
  PublishParams p;
  memset(&p, 0, sizeof p);
  for (int j = 0; j < n; j++) { p.ptr[j] = (unsigned long long *)ptrs[j]; p.epoch[j] = epochs[j]; }
  for (int i = 0; i < nids; i++) p.ids[i] = ids[i];
  p.n = n; p.nids = nids;
  flags_publish_kernel<<<1, P3B_MAXSRC * P3B_MAXGRP, 0, (cudaStream_t)stream>>>(p);
  g_launches++;
  CK(cudaGetLastError());
#else
  (void)stream;
  for (int j = 0; j < n; j++)
    for (int i = 0; i < nids; i++) __atomic_store_n((unsigned long long *)ptrs[j] + ids[i], epochs[j], __ATOMIC_RELEASE);
#endif
  return 0;
}

void p3dfftcu_set_peer_timeout(double seconds) { g_peer_timeout_ns = seconds > 0 ? (unsigned long long)(seconds * 1e9) : 0ull; }

int p3dfftcu_deriv(const void *in, void *out, int prec, const int sd[3], int ldir, int g, int gstart, void *stream) {
  long long total = (long long)sd[0] * sd[1] * sd[2];
  if (total == 0) return 0;
  int threads = 256;
  long long want = (total + threads - 1) / threads;
  int grid = (int)(want < (long long)g_num_sms * 16 ? want : (long long)g_num_sms * 16);
  DerivParams p = {in, out, total, sd[0], sd[1], ldir, g, gstart};
  cudaStream_t cs = (cudaStream_t)stream;
  if (prec == 8) P3B_LAUNCH(deriv_kernel<double>, grid, threads, 0, cs, p);
  else P3B_LAUNCH(deriv_kernel<float>, grid, threads, 0, cs, p);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
}

int p3dfftcu_stream_create(void **stream, int high_priority) {
  int lo = 0, hi = 0;
  CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // numerically lower = greater priority
  cudaStream_t s;
  CK(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high_priority ? hi : lo));
  *stream = s;
  return 0;
}
int p3dfftcu_stream_destroy(void *stream) {
  if (stream) CK(cudaStreamDestroy((cudaStream_t)stream));
  return 0;
}
int p3dfftcu_stream_wait_event(void *stream, void *ev) {
  CK(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)ev, 0));
  return 0;
}
int p3dfftcu_num_sms(void) { return g_num_sms; }

int p3dfftcu_event_create(void **ev) {
  cudaEvent_t e;
  CK(cudaEventCreate(&e));
  *ev = e;
  return 0;
}
int p3dfftcu_event_destroy(void *ev) {
  CK(cudaEventDestroy((cudaEvent_t)ev));
  return 0;
}
int p3dfftcu_event_record(void *ev, void *stream) {
  CK(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)stream));
  return 0;
}
int p3dfftcu_event_elapsed(void *ev0, void *ev1, float *ms) {
  CK(cudaEventSynchronize((cudaEvent_t)ev1));
  CK(cudaEventElapsedTime(ms, (cudaEvent_t)ev0, (cudaEvent_t)ev1));
  return 0;
}

int p3dfftcu_ipc_export(void *ptr, char handle[P3DFFTCU_IPC_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) <= P3DFFTCU_IPC_BYTES, "ipc handle size");
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, ptr));
  memset(handle, 0, P3DFFTCU_IPC_BYTES);
  memcpy(handle, &h, sizeof h);
  return 0;
}
int p3dfftcu_ipc_open(const char handle[P3DFFTCU_IPC_BYTES], void **ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  CK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
int p3dfftcu_ipc_close(void *ptr) {
  CK(cudaIpcCloseMemHandle(ptr));
  return 0;
}

int p3dfftcu_peer_barrier(void *const *peer_flags, const int *peer_slots, int n, void *my_flags, int my_slot,
                          const unsigned long long *epochs, void *stream) {
  if (n > 64) return failmsg("peer barrier: too many peers");
  BarrierParams p;
  memset(&p, 0, sizeof p);
  for (int j = 0; j < n; j++) {
    p.peer[j] = (unsigned long long *)peer_flags[j];
    p.slot[j] = peer_slots[j];
    p.epoch[j] = epochs[j];
  }
  p.mine = (unsigned long long *)my_flags;
  p.n = n;
  p.my_slot = my_slot;
  p.timeout_ns = g_peer_timeout_ns;
#ifndef P3B_EMU
  peer_barrier_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(p);
  g_launches++;
  CK(cudaGetLastError());
  return 0;
#else
  // CPU emulation: same protocol with host atomics on the shared-memory "device" buffers
  (void)stream;
  for (int j = 0; j < n; j++) __atomic_store_n(p.peer[j] + p.my_slot, p.epoch[j], __ATOMIC_RELEASE);
  struct timespec t0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int j = 0; j < n; j++) {
    long spins = 0;
    while (__atomic_load_n(p.mine + p.slot[j], __ATOMIC_ACQUIRE) < p.epoch[j]) {
      if (++spins > 2000) {
        struct timespec ts = {0, 100000}, t1;
        nanosleep(&ts, nullptr);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        const double el = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
        if (p.timeout_ns && el * 1e9 > (double)p.timeout_ns) return failmsg("peer barrier timed out (emulation)");
      }
    }
  }
  return 0;
#endif
}

long long p3dfftcu_launch_count(void) { return g_launches.load(); }

}  // extern "C"
