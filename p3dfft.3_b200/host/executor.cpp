// executor.cpp -- runs a planned transform: buffer chain, per-stage launches, fused exchanges,
// host<->device staging, stand-alone spectral derivative.
//
// Replaces the reference executor (build/exec.C:101-223 stage walk and buffer ping-pong, :297-515
// dispatch, :2299-2341 / :2668-2769 pack + MPI_Alltoallv + unpack).  Differences by design:
//  * every stage is ONE kernel launch that reads its input array once and writes its output once;
//  * an exchange is not a separate collective: the stage kernel stores each peer's block straight into
//    that peer's work buffer over NVLink (buffers mapped with CUDA IPC), bracketed by two stream-ordered
//    peer barriers, so there is no pack buffer, no unpack pass and no host synchronisation;
//  * intermediate arrays live in two library-owned device buffers, never in the user's `out`.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>

#include "plan.h"

namespace p3dfft {
namespace b200 {

bool timers_on();

static void fatal(const Plan *pl, const char *what) {
  fprintf(stderr, "p3dfft_b200 fatal: %s: %s\n", what, p3dfftcu_last_error());
  fflush(stderr);
  MPI_Abort(pl ? pl->comm : MPI_COMM_WORLD, 1);
}
#define GPU(call, pl, what) \
  do {                      \
    if (call) b200::fatal(pl, what); \
  } while (0)

static Workspace g_ws;
Workspace &workspace() { return g_ws; }

static bool ws_init(Workspace &ws, std::string *err) {
  if (ws.world_size) return true;
  int flag = 0;
  MPI_Initialized(&flag);
  ws.world_rank = 0;
  ws.world_size = 1;
  if (flag) {
    MPI_Comm_rank(MPI_COMM_WORLD, &ws.world_rank);
    MPI_Comm_size(MPI_COMM_WORLD, &ws.world_size);
  }
  if (ws.world_size > WS_MAX_RANKS) {
    *err = "more than 64 ranks: the peer flag arrays of this build hold 64 slots";
    ws.world_size = 0;
    return false;
  }
  p3dfftcu_host_ranks_hint(ws.world_size);  // (one host: every rank of the world shares its cores)
  ws.peers.assign(ws.world_size, PeerMap());
  ws.epoch_with.assign(ws.world_size, 0);
  return true;
}

static void close_peer(Workspace &ws, int wr, bool flags_too) {
  PeerMap &pm = ws.peers[wr];
  if (wr != ws.world_rank) {
    for (int w = 0; w < 2; w++)
      if (pm.buf[w]) p3dfftcu_ipc_close(pm.buf[w]);
    if (flags_too && pm.flags) p3dfftcu_ipc_close(pm.flags);
  }
  pm.buf[0] = pm.buf[1] = nullptr;
  pm.gen = 0;
  if (flags_too) pm.flags = nullptr;
}

// all ranks of comm agree: true only if `ok` holds everywhere (keeps the failure paths collective)
static bool all_ok(bool ok, MPI_Comm comm) {
  int l = ok ? 1 : 0, g = l;
  MPI_Allreduce(&l, &g, 1, MPI_INT, MPI_MIN, comm);
  return g != 0;
}

bool workspace_reserve(long long bytes, MPI_Comm comm, int nranks, int rank, std::string *err, std::vector<int> *world_of) {
  Workspace &ws = g_ws;
  bool ok = ws_init(ws, err);
  if (!all_ok(ok, comm)) {
    if (ok) *err = "workspace initialisation failed on another rank";
    return false;
  }
  // who is who: rank in comm -> rank in MPI_COMM_WORLD
  std::vector<int> wof(nranks, 0);
  MPI_Allgather(&ws.world_rank, 1, MPI_INT, wof.data(), 1, MPI_INT, comm);
  if (world_of) *world_of = wof;
  // agree on the capacity: every rank of comm ends up with at least the largest request
  long long want = bytes > ws.bytes ? bytes : ws.bytes, wmax = want;
  MPI_Allreduce(&want, &wmax, 1, MPI_LONG_LONG, MPI_MAX, comm);
  const bool grow = ws.bytes < wmax || !ws.buf[0];
  if (grow) {
    ok = p3dfftcu_stream_sync(current_stream()) == 0;
    if (!ok) *err = p3dfftcu_last_error();
  }
  // the members that are about to re-allocate: everybody drops the mappings of their old buffers first
  std::vector<int> grows(nranks, 0);
  int g = grow ? 1 : 0;
  MPI_Allgather(&g, 1, MPI_INT, grows.data(), 1, MPI_INT, comm);
  for (int r = 0; r < nranks; r++)
    if (grows[r] && wof[r] != ws.world_rank && ws.peers[wof[r]].gen) close_peer(ws, wof[r], false);
  MPI_Barrier(comm);
  if (grow && ok) {
    for (int w = 0; w < 2; w++) {
      if (ws.buf[w]) {
        // a rank outside comm may still have the old buffer mapped: keep it until cleanup unless the whole world is here
        if (nranks == ws.world_size) p3dfftcu_free(ws.buf[w]);
        else ws.retired.push_back(ws.buf[w]);
      }
      ws.buf[w] = nullptr;
      if (ok && p3dfftcu_malloc(&ws.buf[w], (size_t)wmax)) {
        *err = std::string("workspace allocation failed: ") + p3dfftcu_last_error();
        ok = false;
      }
    }
    if (ok) {
      ws.bytes = wmax;
      ws.gen++;
      ws.peers[ws.world_rank].buf[0] = ws.buf[0];
      ws.peers[ws.world_rank].buf[1] = ws.buf[1];
      ws.peers[ws.world_rank].gen = ws.gen;
    } else ws.bytes = 0;
  }
  if (ok && !ws.flags) {
    if (p3dfftcu_malloc(&ws.flags, 8 * WS_FLAG_WORDS) || p3dfftcu_memset(ws.flags, 0, 8 * WS_FLAG_WORDS, current_stream()) ||
        p3dfftcu_stream_sync(current_stream())) {
      *err = p3dfftcu_last_error();
      ok = false;
    } else ws.peers[ws.world_rank].flags = ws.flags;
  }
  if (nranks > 1) {
    // exchange generation + IPC handles [buf0, buf1, flags]; a failed rank still takes part, then everybody fails together
    const size_t rec = 16 + 3 * P3DFFTCU_IPC_BYTES;
    std::vector<char> mine(rec, 0), all((size_t)nranks * rec);
    unsigned long long gen = ok ? ws.gen : 0;
    memcpy(&mine[0], &gen, 8);
    if (ok && (p3dfftcu_ipc_export(ws.buf[0], &mine[16]) || p3dfftcu_ipc_export(ws.buf[1], &mine[16 + P3DFFTCU_IPC_BYTES]) ||
               p3dfftcu_ipc_export(ws.flags, &mine[16 + 2 * P3DFFTCU_IPC_BYTES]))) {
      *err = std::string("CUDA IPC export failed: ") + p3dfftcu_last_error();
      ok = false;
      gen = 0;
      memcpy(&mine[0], &gen, 8);
    }
    MPI_Allgather(mine.data(), (int)rec, MPI_BYTE, all.data(), (int)rec, MPI_BYTE, comm);
    for (int r = 0; r < nranks; r++) {
      const char *h = &all[(size_t)r * rec];
      unsigned long long pgen;
      memcpy(&pgen, h, 8);
      if (pgen == 0) ok = false;  // that rank failed
      const int wr = wof[r];
      if (!ok || wr == ws.world_rank) continue;
      PeerMap &pm = ws.peers[wr];
      if (pm.gen == pgen) continue;
      if (pm.gen) close_peer(ws, wr, false);
      if (p3dfftcu_ipc_open(h + 16, &pm.buf[0]) || p3dfftcu_ipc_open(h + 16 + P3DFFTCU_IPC_BYTES, &pm.buf[1]) ||
          (!pm.flags && p3dfftcu_ipc_open(h + 16 + 2 * P3DFFTCU_IPC_BYTES, &pm.flags))) {
        *err = std::string("CUDA IPC open failed (peer access between the GPUs is required): ") + p3dfftcu_last_error();
        ok = false;
      } else pm.gen = pgen;
    }
  }
  if (!all_ok(ok, comm)) {
    if (ok) *err = "workspace setup failed on another rank";
    return false;
  }
  return true;
}

static void *g_bounce = nullptr;
static long long g_bounce_bytes = 0;
void *workspace_bounce(long long bytes) {
  if (g_bounce_bytes < bytes) {
    p3dfftcu_stream_sync(current_stream());
    if (g_bounce) p3dfftcu_free(g_bounce);
    g_bounce = nullptr;
    g_bounce_bytes = 0;
    if (p3dfftcu_malloc(&g_bounce, (size_t)bytes)) return nullptr;
    g_bounce_bytes = bytes;
  }
  return g_bounce;
}

// device staging of host arrays (and of device arrays a bulk copy cannot read), shared by all plans: [0] input, [1] output
static void *g_stage[2] = {nullptr, nullptr};
static long long g_stage_bytes[2] = {0, 0};
void *workspace_stage(int which, long long bytes) {
  if (bytes < 16) bytes = 16;  // (a rank without local elements still gets a valid pointer)
  if (g_stage_bytes[which] < bytes) {
    p3dfftcu_stream_sync(current_stream());
    if (g_stage[which]) p3dfftcu_free(g_stage[which]);
    g_stage[which] = nullptr;
    g_stage_bytes[which] = 0;
    if (p3dfftcu_malloc(&g_stage[which], (size_t)bytes)) return nullptr;
    g_stage_bytes[which] = bytes;
  }
  return g_stage[which];
}

void workspace_release() {
  Workspace &ws = g_ws;
  for (int i = 0; i < 2; i++) {
    if (g_stage[i]) p3dfftcu_free(g_stage[i]);
    g_stage[i] = nullptr;
    g_stage_bytes[i] = 0;
  }
  if (gpu_ready()) {
    p3dfftcu_stream_sync(current_stream());
    p3dfftcu_host_unpin_all();
  }
  if (g_bounce) p3dfftcu_free(g_bounce);
  g_bounce = nullptr;
  g_bounce_bytes = 0;
  if (!ws.buf[0] && !ws.flags) return;
  p3dfftcu_stream_sync(current_stream());
  int flag = 0;
  MPI_Initialized(&flag);
  for (int r = 0; r < (int)ws.peers.size(); r++) close_peer(ws, r, true);
  if (ws.world_size > 1 && flag) MPI_Barrier(MPI_COMM_WORLD);
  for (int w = 0; w < 2; w++) {
    if (ws.buf[w]) p3dfftcu_free(ws.buf[w]);
    ws.buf[w] = nullptr;
  }
  for (size_t i = 0; i < ws.retired.size(); i++) p3dfftcu_free(ws.retired[i]);
  ws.retired.clear();
  if (ws.flags) p3dfftcu_free(ws.flags);
  ws.flags = nullptr;
  ws.bytes = 0;
  ws.gen = 0;
  ws.world_size = 0;
  ws.peers.clear();
  ws.epoch_with.clear();
}

// world rank of the q-th peer of an exchange stage
static inline int peer_wr(const Plan *pl, const StagePlan &st, size_t q) { return pl->world_of[st.peers[q].peer_world]; }

static void peer_barrier(Plan *pl, const StagePlan &st, void *stream) {
  Workspace &ws = g_ws;
  int n = (int)st.peers.size();
  if (n > WS_MAX_RANKS) fatal(pl, "peer barrier among more than 64 ranks");
  void *pf[WS_MAX_RANKS];
  int slots[WS_MAX_RANKS];
  unsigned long long ep[WS_MAX_RANKS];
  for (int q = 0; q < n; q++) {
    const int wr = peer_wr(pl, st, q);
    pf[q] = (char *)ws.peers[wr].flags + 8 * WS_FLAG_BARRIER0;
    slots[q] = wr;
    ep[q] = ++ws.epoch_with[wr];
  }
  GPU(p3dfftcu_peer_barrier(pf, slots, n, ws.flags, ws.world_rank, ep, stream), pl, "peer barrier");
}

// ---- host arrays.  The reference's users pass ordinary heap arrays (sample/C++/test3D_r2c.C:197-205), i.e. pageable memory,
// which a bare cudaMemcpy moves at a fraction of the PCIe rate.  Modes (P3DFFT_B200_HOST_STAGING or p3dfft_b200_set_host_staging):
//   ring      (default) copy through a ring of pinned chunks, the CPU side split over a few threads and overlapped with the DMA;
//             touches nothing of the user's address space
//   register  page-lock the user's array on first use and remember the range (p3dfftcu_host_pin): full PCIe rate from the second
//             call on, but the application must call p3dfft_b200_host_release() before it frees such an array (a stale
//             registration of a re-used address range would make the DMA engine read the old pages)
//   plain     a bare cudaMemcpyAsync
// Arrays the application page-locked itself (cudaHostAlloc / cudaHostRegister) always take the direct asynchronous copy.
enum { HOST_REGISTER = 0, HOST_RING = 1, HOST_PLAIN = 2 };
static int g_host_mode = -1;
static int parse_host_mode(const char *e) {
  if (e && !strcmp(e, "register")) return HOST_REGISTER;
  if (e && !strcmp(e, "plain")) return HOST_PLAIN;
  return HOST_RING;
}
void set_host_staging(const char *mode) { g_host_mode = parse_host_mode(mode); }
static int host_mode() {
  if (g_host_mode < 0) g_host_mode = parse_host_mode(getenv("P3DFFT_B200_HOST_STAGING"));
  return g_host_mode;
}
// true: copy with the ring (synchronous); false: [host, host+bytes) can be copied by an asynchronous cudaMemcpy
static bool host_needs_ring(const void *host, size_t bytes) {
  const int m = host_mode();
  if (m == HOST_PLAIN || bytes < ((size_t)1 << 20)) return false;
  if (p3dfftcu_host_is_pinned(host) == 1) return false;  // page-locked by the application, or registered earlier
  if (m == HOST_REGISTER) return p3dfftcu_host_pin(host, bytes) != 0;
  return true;
}
static void copy_in(const Plan *pl, void *dev, const void *host, size_t bytes, void *stream) {
  if (host_needs_ring(host, bytes)) GPU(p3dfftcu_memcpy_staged(dev, host, bytes, 0, stream), pl, "host->device copy (staging ring)");
  else GPU(p3dfftcu_memcpy(dev, host, bytes, 0, stream), pl, "host->device copy");
}
// returns when `host` holds the data
static void copy_out(const Plan *pl, void *host, const void *dev, size_t bytes, void *stream) {
  if (host_needs_ring(host, bytes)) GPU(p3dfftcu_memcpy_staged(host, dev, bytes, 1, stream), pl, "device->host copy (staging ring)");
  else {
    GPU(p3dfftcu_memcpy(host, dev, bytes, 1, stream), pl, "device->host copy");
    GPU(p3dfftcu_stream_sync(stream), pl, "stream synchronise");
  }
}

// CTA budgets of an overlapped pair: the exchange stage is NVLink-bound (SM-issued peer stores top out at 717 GB/s per GPU
// whatever the run length >= 128 B and from 74 CTAs on: tools/microbench/peer_store_bench.cu) and keeps its speed on about
// half of the SMs (2 GPUs: 148 -> 74 CTAs costs 3 %); the local stage gets the rest
static void pair_caps(int npeers, int *xcap, int *lcap) {
  static int sms = p3dfftcu_num_sms();
  static int xs = [] {
    const char *e = getenv("P3DFFT_B200_OVERLAP_XSMS");
    int v = e ? atoi(e) : 0;
    return v > 0 ? v : 0;
  }();
  // measured (1024^3 double, persistent pair kernels, profiles/r02_pairs_*): half of the SMs is best on 2, 4 and 8 GPUs
  // (8 GPUs: 56 / 74 / 92 SMs -> 4.06 / 3.87 / 3.91 ms per round trip; all-to-all stores alone reach 636 GB/s from 74 CTAs
  // and 669 GB/s from 148, tools/microbench/a2a_bench.cu)
  (void)npeers;
  int x = xs > 0 ? xs : sms / 2;
  if (x >= sms) x = sms - 1;
  if (x < 1) x = 1;
  *xcap = x;        // the chunk stages of a pair are planned with CTAs that fill an SM (whole_sm_ctas): one CTA per SM,
  *lcap = sms - x;  // so the two kernels can never crowd each other out whatever the dispatch order
}

// runs stages s (first) and s+1 of an overlapped pair; `ssrc` = input of stage s, `ldst` = output of the local stage
static void run_pair(Plan *pl, size_t s, const void *ssrc, void *ldst, const int deriv_g[2], void *stream) {
  Workspace &ws = g_ws;
  StagePlan &first = pl->stages[s], &second = pl->stages[s + 1];
  const bool l_first = first.pair == StagePlan::PAIR_L_THEN_X;
  StagePlan &L = l_first ? first : second, &X = l_first ? second : first;
  const size_t xs = l_first ? s + 1 : s;  // index of the exchange stage: it writes the peers' buffers w
  const int w = (int)(xs & 1);
  const int gL = deriv_g[l_first ? 0 : 1], gX = deriv_g[l_first ? 1 : 0];
  void *xdsts[P3DFFTCU_MAXSEG];
  for (size_t q = 0; q < X.peers.size(); q++) xdsts[q] = ws.peers[peer_wr(pl, X, q)].buf[w];
  void *xstream = pl->xstream;
  std::vector<void *> &ev = pl->sync_events;  // [0] fork, [1] join, [2 + c] chunk c of the first stage is complete
  // P3DFFT_B200_OVERLAP_TRACE=1: start/end events of every chunk on its stream, printed (rank 0) relative to the fork
  static const bool trace = getenv("P3DFFT_B200_OVERLAP_TRACE") && atoi(getenv("P3DFFT_B200_OVERLAP_TRACE"));
  static std::vector<void *> tev;
  auto mark = [&](size_t i, void *st) {
    if (!trace) return;
    while (tev.size() <= i) {
      void *e = nullptr;
      GPU(p3dfftcu_event_create(&e), pl, "event");
      tev.push_back(e);
    }
    GPU(p3dfftcu_event_record(tev[i], st), pl, "event");
  };
  const size_t C = X.chunks.size();
  int xcap, lcap;
  pair_caps((int)X.peers.size(), &xcap, &lcap);
  GPU(p3dfftcu_event_record(ev[0], stream), pl, "event");
  GPU(p3dfftcu_stream_wait_event(xstream, ev[0]), pl, "stream wait");
  peer_barrier(pl, X, xstream);  // every peer has finished reading its buffer w
  if (l_first) {
    // main stream: L chunk by chunk into my work buffer; side stream: X on each chunk as soon as it is complete
    void *lbuf = ws.buf[s & 1];
    for (size_t c = 0; c < C; c++) {
      mark(4 * c + 0, stream);
      if (L.chunks[c].handle) {
        void *d1[1] = {lbuf};
        GPU(p3dfftcu_stage_exec_capped(L.chunks[c].handle, (const char *)ssrc + L.chunks[c].in_off_bytes, d1, 1, gL, stream,
                                       c == 0 ? 0 : lcap), pl, "stage launch");
      }
      mark(4 * c + 1, stream);
      GPU(p3dfftcu_event_record(ev[2 + c], stream), pl, "event");
      GPU(p3dfftcu_stream_wait_event(xstream, ev[2 + c]), pl, "stream wait");
      mark(4 * c + 2, xstream);
      if (X.chunks[c].handle)
        GPU(p3dfftcu_stage_exec_capped(X.chunks[c].handle, (const char *)lbuf + X.chunks[c].in_off_bytes, xdsts, (int)X.peers.size(),
                                       gX, xstream, c + 1 == C ? 0 : xcap), pl, "stage launch");
      mark(4 * c + 3, xstream);
    }
    peer_barrier(pl, X, xstream);  // every peer's blocks have landed in my buffer w
  } else {
    // side stream: X chunk by chunk, each followed by a barrier (chunk c of every peer has landed); main stream: L on
    // each chunk of the received array
    for (size_t c = 0; c < C; c++) {
      mark(4 * c + 2, xstream);
      if (X.chunks[c].handle)
        GPU(p3dfftcu_stage_exec_capped(X.chunks[c].handle, (const char *)ssrc + X.chunks[c].in_off_bytes, xdsts, (int)X.peers.size(),
                                       gX, xstream, c == 0 ? 0 : xcap), pl, "stage launch");
      mark(4 * c + 3, xstream);
      peer_barrier(pl, X, xstream);
      GPU(p3dfftcu_event_record(ev[2 + c], xstream), pl, "event");
      GPU(p3dfftcu_stream_wait_event(stream, ev[2 + c]), pl, "stream wait");
      mark(4 * c + 0, stream);
      if (L.chunks[c].handle) {
        void *d1[1] = {ldst};
        GPU(p3dfftcu_stage_exec_capped(L.chunks[c].handle, (const char *)ws.buf[w] + L.chunks[c].in_off_bytes, d1, 1, gL, stream,
                                       c + 1 == C ? 0 : lcap), pl, "stage launch");
      }
      mark(4 * c + 1, stream);
    }
  }
  GPU(p3dfftcu_event_record(ev[1], xstream), pl, "event");
  GPU(p3dfftcu_stream_wait_event(stream, ev[1]), pl, "stream wait");
  if (trace) {
    mark(4 * C, stream);
    GPU(p3dfftcu_stream_sync(stream), pl, "stream synchronise");
    if (pl->rank == 0) {
      fprintf(stderr, "pair trace (%s first, %zu chunks, xcap %d lcap %d), ms since fork:", l_first ? "L" : "X", C, xcap, lcap);
      for (size_t c = 0; c < C; c++) {
        float t[4];
        for (int i = 0; i < 4; i++) GPU(p3dfftcu_event_elapsed(ev[0], tev[4 * c + i], &t[i]), pl, "event");
        fprintf(stderr, "  [%zu] L %.2f-%.2f X %.2f-%.2f", c, t[0], t[1], t[2], t[3]);
      }
      float te;
      GPU(p3dfftcu_event_elapsed(ev[0], tev[4 * C], &te), pl, "event");
      fprintf(stderr, "  end %.2f\n", te);
    }
  }
}

// The same pair as ONE persistent launch per stage: the chunks are tile groups of the two kernels, which run side by side on
// disjoint sets of SMs; "chunk c is complete" travels as a flag word written from inside the producing kernel (to this
// GPU's flag array when the local stage feeds the exchange stage, to every peer's over NVLink when the exchange stage
// feeds the peers' local stages) and is awaited by the thread that issues the consuming kernel's bulk loads.  No launch and
// no barrier kernel per chunk; the CTAs freed by the exchange kernel join the local stage's tile pool.
// (zdst != nullptr: the L -> X -> Z triple of plan.h, Z writing zdst)
static void run_pair_sync(Plan *pl, size_t s, const void *ssrc, void *ldst, const int deriv_g[3], void *stream, void *zdst) {
  Workspace &ws = g_ws;
  StagePlan &first = pl->stages[s], &second = pl->stages[s + 1];
  const bool l_first = first.pair == StagePlan::PAIR_L_THEN_X;
  StagePlan &L = l_first ? first : second, &X = l_first ? second : first;
  const size_t xs = l_first ? s + 1 : s;
  const int w = (int)(xs & 1);
  const int gL = deriv_g[l_first ? 0 : 1], gX = deriv_g[l_first ? 1 : 0];
  const int np = (int)X.peers.size();
  void *xdsts[P3DFFTCU_MAXSEG];
  for (int q = 0; q < np; q++) xdsts[q] = ws.peers[peer_wr(pl, X, q)].buf[w];
  void *xstream = pl->xstream;
  std::vector<void *> &ev = pl->sync_events;
  const size_t C = X.chunk_range.size();
  int xcap, lcap;
  pair_caps(np, &xcap, &lcap);
  if (((uintptr_t)ssrc) % 16) {
    // bulk copies need 16-byte aligned pencils and every rank must follow the same protocol: an odd user pointer is staged
    void *st_in = workspace_stage(0, first.in_bytes);
    if (!st_in) fatal(pl, "staging allocation");
    GPU(p3dfftcu_memcpy(st_in, ssrc, (size_t)first.in_bytes, 2, stream), pl, "device copy");
    ssrc = st_in;
  }
  // group tables: chunk c of a stage = the range [c0, c1) of chunk_dim, everything along its other pencil dimension
  p3dfftcu_sync sl, sx;
  memset(&sl, 0, sizeof sl);
  memset(&sx, 0, sizeof sx);
  auto fill = [&](const StagePlan &st, p3dfftcu_sync &sy) {
    sy.ngroups = (int)C;
    const bool along_u = st.chunk_dim == st.u;
    for (size_t c = 0; c < C; c++) {
      p3dfftcu_group &g = sy.grp[c];
      const int c0 = st.chunk_range[c].first, c1 = st.chunk_range[c].second;
      g.u0 = along_u ? c0 : 0;
      g.u1 = along_u ? c1 : (int)st.desc.nu;
      g.v0 = along_u ? 0 : c0;
      g.v1 = along_u ? (int)st.desc.nv : c1;
      g.wait_id = g.signal_id = -1;
    }
  };
  fill(L, sl);
  fill(X, sx);
  sl.ctl = pl->ctl;
  sx.ctl = (char *)pl->ctl + 8 * 32;
  GPU(p3dfftcu_memset(pl->ctl, 0, 8 * 96, stream), pl, "memset");
  static const int sms = p3dfftcu_num_sms();
  // P3DFFT_B200_OVERLAP_TRACE=1: when each kernel of the pair ended, relative to the fork (rank 0)
  static const bool trace = getenv("P3DFFT_B200_OVERLAP_TRACE") && atoi(getenv("P3DFFT_B200_OVERLAP_TRACE"));
  static std::vector<void *> tev;
  const char *tname[8];
  int ntr = 0;
  auto mark = [&](const char *name, void *st) {
    if (!trace || ntr >= 8) return;
    while ((int)tev.size() <= ntr) {
      void *e = nullptr;
      GPU(p3dfftcu_event_create(&e), pl, "event");
      tev.push_back(e);
    }
    tname[ntr] = name;
    GPU(p3dfftcu_event_record(tev[ntr++], st), pl, "event");
  };
  GPU(p3dfftcu_event_record(ev[0], stream), pl, "event");
  GPU(p3dfftcu_stream_wait_event(xstream, ev[0]), pl, "stream wait");
  peer_barrier(pl, X, xstream);  // every peer has finished reading its buffer w
  const bool x_live = X.pair_handle && p3dfftcu_stage_sync_capable(X.pair_handle);
  const bool l_live = L.pair_handle && p3dfftcu_stage_sync_capable(L.pair_handle);
  if (l_first && X.triple && zdst) {
    // three kernels: L (z-chunks) -> X (first the leading z-chunks behind L, then the rest of z piece by piece along L's
    // transform dimension a, each piece published to the peers) -> Z (piece k once every peer has published it)
    StagePlan &Z = pl->stages[s + 2];
    const int c1 = X.tri_c1, K = (int)X.tri_a_range.size(), a = L.dim;
    const unsigned long long le = ++ws.local_epoch;
    // L: chunks [0, c1) as they are, the remaining chunks merged into one group
    const bool l_u = L.chunk_dim == L.u;
    sl.ngroups = c1 + 1;
    {
      p3dfftcu_group &g = sl.grp[c1];
      const int z0 = L.chunk_range[c1].first, z1 = L.chunk_range[C - 1].second;
      g.u0 = l_u ? z0 : 0;
      g.u1 = l_u ? z1 : (int)L.desc.nu;
      g.v0 = l_u ? 0 : z0;
      g.v1 = l_u ? (int)L.desc.nv : z1;
    }
    for (int c = 0; c <= c1; c++) sl.grp[c].signal_id = c;
    sl.sig_n = 1;
    sl.sig_ptr[0] = (char *)ws.flags + 8 * WS_FLAG_LOCAL0;
    sl.sig_epoch[0] = le;
    sl.boost_ctas = sms;
    sl.boost_groups = 1;
    // X: chunks [0, c1) wait for L's flags and are counted; then K pieces (a-range x remaining z) wait for L's last group
    const bool x_u = X.chunk_dim == X.u;  // (the other pencil dimension of X is a)
    sx.ngroups = c1 + K;
    for (int c = 0; c < c1; c++) {
      sx.grp[c].wait_id = c;
      sx.grp[c].count = 1;
    }
    int empty_ids[P3DFFTCU_MAXGRP], nempty = 0;
    const int z0 = X.chunk_range[c1].first, z1 = X.chunk_range[C - 1].second;
    for (int k = 0; k < K; k++) {
      p3dfftcu_group &g = sx.grp[c1 + k];
      const int a0 = X.tri_a_range[k].first, a1 = X.tri_a_range[k].second;
      g.u0 = x_u ? z0 : a0;
      g.u1 = x_u ? z1 : a1;
      g.v0 = x_u ? a0 : z0;
      g.v1 = x_u ? a1 : z1;
      g.wait_id = c1;
      g.signal_id = k;
      g.count = 0;
      g.after = c1;
      if (!x_live || g.u1 <= g.u0 || g.v1 <= g.v0) empty_ids[nempty++] = k;
    }
    sx.wait_n = 1;
    sx.wait_base = ws.flags;
    sx.wait_off[0] = WS_FLAG_LOCAL0;
    sx.wait_epoch[0] = le;
    // Z: piece k = a-range k, everything along its other pencil dimension
    p3dfftcu_sync sz;
    memset(&sz, 0, sizeof sz);
    sz.ngroups = K;
    sz.ctl = (char *)pl->ctl + 8 * 64;
    const bool z_u = a == Z.u;
    for (int k = 0; k < K; k++) {
      p3dfftcu_group &g = sz.grp[k];
      const int a0 = X.tri_a_range[k].first, a1 = X.tri_a_range[k].second;
      g.u0 = z_u ? a0 : 0;
      g.u1 = z_u ? a1 : (int)Z.desc.nu;
      g.v0 = z_u ? 0 : a0;
      g.v1 = z_u ? (int)Z.desc.nv : a1;
      g.wait_id = k;
      g.signal_id = -1;
    }
    sx.sig_n = sz.wait_n = np;
    sz.wait_base = ws.flags;
    for (int q = 0; q < np; q++) {
      const int wr = peer_wr(pl, X, q);
      const unsigned long long e = ++ws.epoch_with[wr];
      sx.sig_ptr[q] = (char *)ws.peers[wr].flags + 8 * (WS_FLAG_GROUP0 + ws.world_rank * WS_FLAGS_PER_SRC);
      sx.sig_epoch[q] = e;
      sz.wait_off[q] = WS_FLAG_GROUP0 + wr * WS_FLAGS_PER_SRC;
      sz.wait_epoch[q] = e;
    }
    const bool z_live = Z.pair_handle && p3dfftcu_stage_sync_capable(Z.pair_handle);
    void *d1[1] = {ws.buf[s & 1]}, *dz[1] = {zdst};
    if (l_live) GPU(p3dfftcu_stage_exec_sync(L.pair_handle, ssrc, d1, 1, gL, stream, lcap, &sl), pl, "stage launch");
    mark("L", stream);
    mark("barrier", xstream);
    if (nempty) GPU(p3dfftcu_flags_publish(sx.sig_ptr, sx.sig_epoch, np, empty_ids, nempty, xstream), pl, "flag publish");
    if (x_live) GPU(p3dfftcu_stage_exec_sync(X.pair_handle, ws.buf[s & 1], xdsts, np, gX, xstream, xcap, &sx), pl, "stage launch");
    mark("X", xstream);
    if (z_live) {
      GPU(p3dfftcu_stage_exec_sync(Z.pair_handle, ws.buf[w], dz, 1, deriv_g[2], stream, lcap, &sz), pl, "stage launch");
      mark("Za", stream);
      if (p3dfftcu_stage_sync_capable(Z.pair_handle) == 2)  // the SMs the exchange kernel leaves join Z's work pool
        GPU(p3dfftcu_stage_exec_sync(Z.pair_handle, ws.buf[w], dz, 1, deriv_g[2], xstream, xcap, &sz), pl, "stage launch");
      mark("Zb", xstream);
    }
  } else if (l_first) {
    // L publishes chunk c in this GPU's flag array, X waits for it
    const unsigned long long le = ++ws.local_epoch;
    for (size_t c = 0; c < C; c++) {
      sl.grp[c].signal_id = (int)c;
      sx.grp[c].wait_id = (int)c;
    }
    sl.boost_ctas = sms;  // the first chunk on every SM: nothing else can run before it is complete
    sl.boost_groups = 1;
    sl.sig_n = 1;
    sl.sig_ptr[0] = (char *)ws.flags + 8 * WS_FLAG_LOCAL0;
    sl.sig_epoch[0] = le;
    sx.wait_n = 1;
    sx.wait_base = ws.flags;
    sx.wait_off[0] = WS_FLAG_LOCAL0;
    sx.wait_epoch[0] = le;
    void *d1[1] = {ws.buf[s & 1]};
    if (l_live) GPU(p3dfftcu_stage_exec_sync(L.pair_handle, ssrc, d1, 1, gL, stream, lcap, &sl), pl, "stage launch");
    mark("L", stream);
    mark("barrier", xstream);
    if (x_live) GPU(p3dfftcu_stage_exec_sync(X.pair_handle, ws.buf[s & 1], xdsts, np, gX, xstream, xcap, &sx), pl, "stage launch");
    mark("X", xstream);
    peer_barrier(pl, X, xstream);  // every peer's blocks have landed in my buffer w
    mark("end barrier", xstream);
  } else {
    // X publishes chunk c in every peer's flag array (row = my world rank), L waits for chunk c of every peer
    int empty_ids[P3DFFTCU_MAXGRP], nempty = 0;
    for (size_t c = 0; c < C; c++) {
      sx.grp[c].signal_id = (int)c;
      sl.grp[c].wait_id = (int)c;
      const p3dfftcu_group &g = sx.grp[c];
      if (!x_live || g.u1 <= g.u0 || g.v1 <= g.v0) empty_ids[nempty++] = (int)c;  // nothing to send: the host publishes it
    }
    sx.sig_n = sl.wait_n = np;
    sl.wait_base = ws.flags;
    for (int q = 0; q < np; q++) {
      const int wr = peer_wr(pl, X, q);
      const unsigned long long e = ++ws.epoch_with[wr];
      sx.sig_ptr[q] = (char *)ws.peers[wr].flags + 8 * (WS_FLAG_GROUP0 + ws.world_rank * WS_FLAGS_PER_SRC);
      sx.sig_epoch[q] = e;
      sl.wait_off[q] = WS_FLAG_GROUP0 + wr * WS_FLAGS_PER_SRC;
      sl.wait_epoch[q] = e;
    }
    mark("barrier", xstream);
    if (nempty) GPU(p3dfftcu_flags_publish(sx.sig_ptr, sx.sig_epoch, np, empty_ids, nempty, xstream), pl, "flag publish");
    if (x_live) GPU(p3dfftcu_stage_exec_sync(X.pair_handle, ssrc, xdsts, np, gX, xstream, xcap, &sx), pl, "stage launch");
    mark("X", xstream);
    void *d1[1] = {ldst};
    if (l_live) {
      GPU(p3dfftcu_stage_exec_sync(L.pair_handle, ws.buf[w], d1, 1, gL, stream, lcap, &sl), pl, "stage launch");
      mark("La", stream);
      // the SMs the exchange kernel leaves join the local stage (same tile counter) when its kernel hands tiles out dynamically
      if (p3dfftcu_stage_sync_capable(L.pair_handle) == 2)
        GPU(p3dfftcu_stage_exec_sync(L.pair_handle, ws.buf[w], d1, 1, gL, xstream, xcap, &sl), pl, "stage launch");
      mark("Lb", xstream);
    }
  }
  GPU(p3dfftcu_event_record(ev[1], xstream), pl, "event");
  GPU(p3dfftcu_stream_wait_event(stream, ev[1]), pl, "stream wait");
  if (trace) {
    GPU(p3dfftcu_stream_sync(stream), pl, "stream synchronise");
    if (pl->rank == 0) {
      fprintf(stderr, "pair-sync trace (%s%s, xcap %d lcap %d), kernel end times in ms since the fork:", l_first ? "L first" : "X first",
              (l_first && X.triple && zdst) ? ", triple" : "", xcap, lcap);
      for (int i = 0; i < ntr; i++) {
        float t = 0;
        GPU(p3dfftcu_event_elapsed(ev[0], tev[i], &t), pl, "event");
        fprintf(stderr, "  %s %.3f", tname[i], t);
      }
      fprintf(stderr, "\n");
    }
  }
}

static void add_timer(const StagePlan &st, bool deriv, double sec) {
  if (st.kind == P3DFFTCU_K_EMPTY) {
    if (st.exchange) timers.alltoall += sec;
    else timers.reorder_out += sec;
  } else if (st.exchange) {
    if (deriv) timers.packsend_deriv += sec;
    else timers.packsend_trans += sec;
  } else {
    bool same = st.in.mo[0] == st.out.mo[0] && st.in.mo[1] == st.out.mo[1] && st.in.mo[2] == st.out.mo[2];
    if (deriv) (same ? timers.trans_deriv : timers.reorder_deriv) += sec;
    else (same ? timers.trans_exec : timers.reorder_trans) += sec;
  }
}

void plan_exec(Plan *pl, const void *in, void *out, int idir, bool OW) {
  if (!pl || !pl->ok) {
    printf("Error in exec: transform plan is not set\n");
    return;
  }
  if (!gpu_ready()) {
    fprintf(stderr,
            "p3dfft_b200 fatal: no usable CUDA device (or the CUDA layer failed to initialise); this library has no CPU "
            "execution path\n");
    MPI_Abort(pl->comm, 1);
  }
  if (in == (const void *)out && !OW)
    printf("Warning in transform3D:exec_deriv: input and output are the same, but overwrite priviledge is not set\n");
  void *stream = current_stream();
  Workspace &ws = g_ws;
  const bool in_dev = p3dfftcu_pointer_is_device(in) == 1;
  const bool out_dev = p3dfftcu_pointer_is_device(out) == 1;
  const void *src = in;
  void *dst = out;
  if (!in_dev) {  // host arrays are staged in device buffers shared by all plans of the process
    void *st_in = workspace_stage(0, pl->in_bytes);
    if (!st_in) fatal(pl, "staging allocation");
    copy_in(pl, st_in, in, (size_t)pl->in_bytes, stream);
    src = st_in;
  }
  if (!out_dev) {
    dst = workspace_stage(1, pl->out_bytes);
    if (!dst) fatal(pl, "staging allocation");
  }
  const size_t S = pl->stages.size();
  const bool timing = timers_on();
  std::vector<void *> &ev = pl->events;
  size_t ev0 = 0;  // first event of this exec: every exec since the last read keeps its own S+1 events (up to 64 execs)
  if (timing) {
    if (pl->timed_execs >= 64) plan_collect_times(pl, false);  // the event ring is full: fold its 64 execs into the sums first
    ev0 = (size_t)pl->timed_execs * (S + 1);
    while (ev.size() < ev0 + S + 1) {
      void *e = nullptr;
      GPU(p3dfftcu_event_create(&e), pl, "event");
      ev.push_back(e);
    }
    GPU(p3dfftcu_event_record(ev[ev0], stream), pl, "event");
  }
  int deriv_stage = -1;
  auto deriv_len = [&](size_t s) {
    const StagePlan &st = pl->stages[s];
    if (idir >= 0 && st.dim == idir && st.kind != P3DFFTCU_K_EMPTY) {
      if (st.dt_out != 2) printf("Error in exec_deriv: expected complex output type\n");
      else {
        deriv_stage = (int)s;
        return st.kind == P3DFFTCU_K_R2C ? (st.n_out - 1) * 2 : st.n_out;  // exec.C:240-246
      }
    }
    return 0;
  };
  for (size_t s = 0; s < S; s++) {
    StagePlan &st = pl->stages[s];
    const void *ssrc = s == 0 ? src : ws.buf[(s - 1) & 1];
    if (st.pair != StagePlan::PAIR_NONE && s + 1 < S) {
      // overlapped pair: the local stage's output is my work buffer, or the user's array when it is the last stage
      const bool l_first = st.pair == StagePlan::PAIR_L_THEN_X;
      const size_t ls = l_first ? s : s + 1;
      void *ldst = ls + 1 == S ? dst : ws.buf[ls & 1];
      if (l_first || (const void *)ldst != ssrc) {  // (in-place call whose output would overwrite X's input: run in sequence)
        const bool triple = l_first && st.pair_sync && pl->stages[s + 1].triple && s + 3 == S;
        const int dg[3] = {deriv_len(s), deriv_len(s + 1), triple ? deriv_len(s + 2) : 0};
        if (st.pair_sync) run_pair_sync(pl, s, ssrc, ldst, dg, stream, triple ? dst : nullptr);
        else run_pair(pl, s, ssrc, ldst, dg, stream);
        if (triple) {  // the whole triple is booked on its first stage
          if (timing)
            for (size_t e = s + 1; e <= s + 3; e++) GPU(p3dfftcu_event_record(ev[ev0 + e], stream), pl, "event");
          s += 2;
          continue;
        }
        if (l_first && s + 2 == S)  // the exchange stage is the plan's last: its result sits in my work buffer
          GPU(p3dfftcu_memcpy(dst, ws.buf[(s + 1) & 1], (size_t)pl->stages[s + 1].out_bytes, 2, stream), pl, "device copy");
        if (timing) {  // the pair is timed as a whole: its duration is booked on the first stage, zero on the second
          GPU(p3dfftcu_event_record(ev[ev0 + s + 1], stream), pl, "event");
          GPU(p3dfftcu_event_record(ev[ev0 + s + 2], stream), pl, "event");
        }
        s++;
        continue;
      }
    }
    const bool last = s + 1 == S;
    const int deriv_g = deriv_len(s);
    if (st.exchange) {
      const int w = (int)(s & 1);
      void *dsts[P3DFFTCU_MAXSEG];
      for (size_t q = 0; q < st.peers.size(); q++) dsts[q] = ws.peers[peer_wr(pl, st, q)].buf[w];
      peer_barrier(pl, st, stream);  // every peer has finished reading its buffer w
      GPU(p3dfftcu_stage_exec(st.handle, ssrc, dsts, (int)st.peers.size(), deriv_g, stream), pl, "stage launch");
      peer_barrier(pl, st, stream);  // every peer's block has landed in my buffer w
      if (last) GPU(p3dfftcu_memcpy(dst, ws.buf[w], (size_t)st.out_bytes, 2, stream), pl, "device copy");
    } else {
      void *sdst = last ? dst : ws.buf[s & 1];
      const bool bounce = last && ssrc == (const void *)sdst;  // single stage, in == out
      if (bounce) {
        sdst = workspace_bounce(st.out_bytes);
        if (!sdst) fatal(pl, "scratch allocation for an in-place single-stage transform");
      }
      void *dsts[1] = {sdst};
      GPU(p3dfftcu_stage_exec(st.handle, ssrc, dsts, 1, deriv_g, stream), pl, "stage launch");
      if (bounce) GPU(p3dfftcu_memcpy(dst, sdst, (size_t)st.out_bytes, 2, stream), pl, "device copy");
    }
    if (timing) GPU(p3dfftcu_event_record(ev[ev0 + s + 1], stream), pl, "event");
  }
  if (!out_dev) {
    copy_out(pl, out, dst, (size_t)pl->out_bytes, stream);
  } else if (!in_dev) {
    GPU(p3dfftcu_stream_sync(stream), pl, "stream synchronise");  // the host input may be reused by the caller
  }
  if (timing) pl->timed_execs++;
  pl->events_valid = timing;
  pl->last_deriv_stage = deriv_stage;
}

// per-stage milliseconds averaged over the execs since the last read (waits for them to finish); also feeds p3dfft::timers.
// final = false: only folds the recorded execs into the running sums (the event ring holds 64 execs)
void plan_collect_times(Plan *pl, bool final) {
  const size_t S = pl->stages.size();
  if (pl->acc_ms.size() != S) pl->acc_ms.assign(S, 0.0);
  if (pl->events_valid && pl->timed_execs >= 1) {
    for (size_t s = 0; s < S; s++) {
      double sum = 0;
      for (int x = 0; x < pl->timed_execs; x++) {
        float ms = 0;
        const size_t e = (size_t)x * (S + 1) + s;
        GPU(p3dfftcu_event_elapsed(pl->events[e], pl->events[e + 1], &ms), pl, "event");
        sum += ms;
      }
      pl->acc_ms[s] += sum;
      add_timer(pl->stages[s], (int)s == pl->last_deriv_stage, sum * 1e-3);
    }
    pl->acc_execs += pl->timed_execs;
  }
  pl->events_valid = false;
  pl->timed_execs = 0;
  if (final && pl->acc_execs > 0) {
    for (size_t s = 0; s < S; s++) pl->stage_ms[s] = (float)(pl->acc_ms[s] / pl->acc_execs);
    pl->acc_ms.assign(S, 0.0);
    pl->acc_execs = 0;
  }
}

}  // namespace b200

// Stand-alone derivative (reference build/deriv.C:85-185).  The reference selects the storage dimension
// with the INVERSE permutation (deriv.C:90-94: ldir = i such that MemOrder[i] == idir); for the cyclic
// orders {1,2,0} / {2,0,1} that differs from MemOrder[idir].  Kept as is so results match the reference;
// set P3DFFT_B200_DERIV_LDIR=memorder to use MemOrder[idir] instead.
template <class Type> void compute_deriv(Type *in, Type *out, DataGrid *gr, int idir) {
  if (!b200::gpu_ready()) {
    fprintf(stderr, "p3dfft_b200 fatal: no usable CUDA device; this library has no CPU execution path\n");
    MPI_Abort(gr->Pgrid->mpi_comm_glob, 1);
  }
  int sd[3], ldir = 0;
  for (int i = 0; i < 3; i++) {
    sd[gr->MemOrder[i]] = gr->Ldims[i];
    if (gr->MemOrder[i] == idir) ldir = i;
  }
  const char *mode = getenv("P3DFFT_B200_DERIV_LDIR");
  if (mode && !strcmp(mode, "memorder")) ldir = gr->MemOrder[idir];
  int g = gr->dim_conj_sym == idir ? (gr->Gdims[idir] - 1) * 2 : gr->Gdims[idir];
  const int prec = b200::tinfo<Type>::prec;
  const size_t bytes = (size_t)sd[0] * sd[1] * sd[2] * 2 * prec;
  void *stream = b200::current_stream();
  const bool in_dev = p3dfftcu_pointer_is_device(in) == 1, out_dev = p3dfftcu_pointer_is_device(out) == 1;
  void *din = (void *)in, *dout = (void *)out;
  if (!in_dev || !out_dev) {
    // host arrays: one device scratch kept across calls (the kernel works in place on it)
    void *scratch = b200::workspace_bounce((long long)bytes);
    if (!scratch) b200::fatal(nullptr, "staging allocation");
    if (!in_dev) {
      b200::copy_in(nullptr, scratch, in, bytes, stream);
      din = scratch;
    }
    if (!out_dev) dout = scratch;
  }
  GPU(p3dfftcu_deriv(din, dout, prec, sd, ldir, g, gr->GlobStart[idir], stream), nullptr, "derivative kernel");
  if (!out_dev) b200::copy_out(nullptr, out, dout, bytes, stream);
  else if (!in_dev) GPU(p3dfftcu_stream_sync(stream), nullptr, "stream synchronise");
}

template void compute_deriv<mycomplex>(mycomplex *, mycomplex *, DataGrid *, int);
template void compute_deriv<complex_double>(complex_double *, complex_double *, DataGrid *, int);

}  // namespace p3dfft
