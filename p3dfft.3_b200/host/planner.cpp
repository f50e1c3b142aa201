// planner.cpp -- turns (grid1, grid2, three 1D types) into a list of fused stages.
//
// Contract kept from the reference planner (build/templ.C:91-802): types[i] acts along logical
// dimension i; a real-to-complex type runs first and halves that dimension to N/2+1, a complex-to-real
// type runs last (templ.C:639-655, 246-253); a 1D transform only runs on a dimension that is entirely
// local; distributions change by swapping the Dmap entries of a local and a distributed dimension
// inside one sub-communicator (templ.C:282-346).  Input and output layouts are exactly grid1 / grid2.
// Everything in between is this build's own choice: a small exhaustive search picks the operation
// order with the fewest exchanges and the intermediate storage orders that keep every stage's global
// loads and stores coalesced (the reference's swap0 heuristic and its five reorder cases, templ.C:816-858
// and exec.C:737-2032, have no counterpart here).  Every exchange is fused into the stage that
// transforms the dimension being given away; an exchange with no transform becomes a copy stage.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <sstream>

#include "plan.h"

namespace p3dfft {
namespace b200 {

namespace {

struct Op {
  int is_x;  // 0: transform dim d; 1: exchange (a becomes distributed, b becomes local)
  int d, a, b;
};

void block(int n, int p, int idx, int *st, int *sz) {
  int base = n / p, nlow = p - n % p;
  *st = idx < nlow ? idx * base : nlow * base + (idx - nlow) * (base + 1);
  *sz = idx < nlow ? base : base + 1;
}

struct Search {
  const ProcGrid *pg;
  const gen_trans_type *ty[3];
  int r2c_dim, c2r_dim;
  int target_dmap[3];
  std::vector<std::vector<Op>> found;
  int maxx;

  bool local(const int dmap[3], int i) const { return pg->ProcDims[dmap[i]] == 1; }
  bool goal(const int dmap[3], int done) const {
    if (done != 7) return false;
    for (int i = 0; i < 3; i++) {
      bool l1 = local(dmap, i), l2 = pg->ProcDims[target_dmap[i]] == 1;
      if (l1 != l2) return false;
      if (!l1 && dmap[i] != target_dmap[i]) return false;
    }
    return true;
  }
  void dfs(int dmap[3], int done, int nx, std::vector<Op> &path, int depth) {
    if (goal(dmap, done)) {
      found.push_back(path);
      return;
    }
    if (depth > 8) return;
    for (int d = 0; d < 3; d++) {
      if (done & (1 << d)) continue;
      if (!local(dmap, d)) continue;
      if (r2c_dim >= 0 && d != r2c_dim && !(done & (1 << r2c_dim))) continue;
      if (d == c2r_dim && (done | (1 << d)) != 7) continue;
      path.push_back(Op{0, d, -1, -1});
      dfs(dmap, done | (1 << d), nx, path, depth + 1);
      path.pop_back();
    }
    if (nx >= maxx) return;
    for (int a = 0; a < 3; a++) {
      if (!local(dmap, a)) continue;
      for (int b = 0; b < 3; b++) {
        if (b == a || local(dmap, b)) continue;
        if (!path.empty() && path.back().is_x && path.back().a == b && path.back().b == a) continue;
        std::swap(dmap[a], dmap[b]);
        path.push_back(Op{1, -1, a, b});
        dfs(dmap, done, nx + 1, path, depth + 1);
        path.pop_back();
        std::swap(dmap[a], dmap[b]);
      }
    }
  }
};

const int kPerm[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};

int lead_dim(const int mo[3], const int ld[3]) {
  // logical dimension with the smallest storage rank among those of extent > 1
  int best = -1;
  for (int i = 0; i < 3; i++)
    if (ld[i] > 1 && (best < 0 || mo[i] < mo[best])) best = i;
  return best < 0 ? 0 : best;
}

// P3DFFT_B200_COST_ROWROW / P3DFFT_B200_COST_TLOAD: A/B overrides of two weights of the layout search (development)
double env_cost(const char *name, double dflt) {
  const char *e = getenv(name);
  return (e && atof(e) > 0) ? atof(e) : dflt;
}

double stage_cost(int d, const int mo_in[3], const int ld_in[3], const int mo_out[3], const int ld_out[3]) {
  const double c_rowrow = env_cost("P3DFFT_B200_COST_ROWROW", 1.30), c_tload = env_cost("P3DFFT_B200_COST_TLOAD", 1.12);
  int fi = lead_dim(mo_in, ld_in), fo = lead_dim(mo_out, ld_out);
  double c;
  if (fi == d && fo == d) c = 1.00;
  else if (fi == fo) c = c_rowrow;  // row-granular loads AND stores (measured slowest: 128-byte bulk copies per row)
  else if (fi == d) {
    c = 1.08;  // contiguous (TMA-prefetched) loads + transposed stores: stores do not stall the pipeline
    // the stores of a tile are one short run per output index k: keep consecutive k close in memory (d second in the
    // output order) so a tile touches a few pages instead of one 2 MB page per run (measured 5.9 vs 5.1 TB/s)
    int before = 0;
    for (int e = 0; e < 3; e++)
      if (e != d && ld_out[e] > 1 && mo_out[e] < mo_out[d]) before++;
    if (before >= 2) c += 0.05;
  }
  else if (fo == d) {
    c = c_tload;  // transposed loads (TMA tensor boxes: one short row per input index) + contiguous stores
    // the mirror image of the rule above: consecutive input indices close in memory (d second in the input order)
    int before = 0;
    for (int e = 0; e < 3; e++)
      if (e != d && ld_in[e] > 1 && mo_in[e] < mo_in[d]) before++;
    if (before >= 2) c += 0.05;
  }
  else c = 1.60;
  if (mo_in[0] == mo_out[0] && mo_in[1] == mo_out[1] && mo_in[2] == mo_out[2]) c -= 0.02;
  return c;
}

struct ProtoStage {
  int kind, dim, dt_in, dt_out, nfft, n_in, n_out;
  bool exchange;
  int xb;        // dim gathered by the exchange
  int comm_dim;  // processor-grid dim of the exchange
  int gd_in[3], gd_out[3];
  int dmap_in[3], dmap_out[3];
  int ld_in[3], ld_out[3];
  int rep_in[3], rep_out[3];  // largest block of each dimension over all ranks: rank-independent input of the layout choice
};

void local_dims(const ProcGrid *pg, const int gd[3], const int dmap[3], int ld[3]) {
  for (int i = 0; i < 3; i++) {
    int st;
    block(gd[i], pg->ProcDims[dmap[i]], pg->grid_id_cart[dmap[i]], &st, &ld[i]);
  }
}

// every rank must pick the same intermediate storage orders (senders address their peers' arrays), so the
// layout search sees the largest block of each dimension instead of this rank's own block
void rep_dims(const ProcGrid *pg, const int gd[3], const int dmap[3], int rd[3]) {
  for (int i = 0; i < 3; i++) {
    int p = pg->ProcDims[dmap[i]];
    rd[i] = gd[i] / p + (gd[i] % p ? 1 : 0);
  }
}

bool build_protos(const std::vector<Op> &ops, const DataGrid &g1, const DataGrid &g2, const gen_trans_type *ty[3], int dt_in,
                  std::vector<ProtoStage> *out, std::string *err) {
  const ProcGrid *pg = g1.Pgrid;
  int gd[3], dmap[3], dt = dt_in;
  for (int i = 0; i < 3; i++) {
    gd[i] = g1.Gdims[i];
    dmap[i] = g1.Dmap[i];
  }
  out->clear();
  for (size_t i = 0; i < ops.size(); i++) {
    ProtoStage s;
    memset(&s, 0, sizeof s);
    s.exchange = false;
    s.xb = s.comm_dim = -1;
    memcpy(s.gd_in, gd, sizeof gd);
    memcpy(s.dmap_in, dmap, sizeof dmap);
    const Op &op = ops[i];
    if (!op.is_x) {
      const gen_trans_type *t = ty[op.d];
      s.kind = t->kind;
      s.dim = op.d;
      s.dt_in = dt;
      if (t->dt1 != dt) {
        *err = "datatypes of consecutive 1D transforms do not match";
        return false;
      }
      s.dt_out = t->dt2;
      s.n_in = gd[op.d];
      if (t->dt1 < t->dt2) {  // R2C: N real -> N/2+1 complex
        s.nfft = gd[op.d];
        s.n_out = s.nfft / 2 + 1;
      } else if (t->dt1 > t->dt2) {  // C2R: real length comes from the output grid when consistent
        int n2 = g2.Gdims[op.d];
        s.nfft = (n2 / 2 + 1 == s.n_in) ? n2 : (s.n_in - 1) * 2;
        s.n_out = s.nfft;
      } else {
        s.nfft = s.n_out = gd[op.d];
      }
      gd[op.d] = s.n_out;
      dt = s.dt_out;
      if (i + 1 < ops.size() && ops[i + 1].is_x && ops[i + 1].a == op.d) {  // fuse the exchange that gives d away
        const Op &x = ops[++i];
        s.exchange = true;
        s.xb = x.b;
        s.comm_dim = dmap[x.b];
        std::swap(dmap[x.a], dmap[x.b]);
      }
    } else {
      s.kind = P3DFFTCU_K_EMPTY;
      s.dim = op.a;
      s.dt_in = s.dt_out = dt;
      s.nfft = s.n_in = s.n_out = gd[op.a];
      s.exchange = true;
      s.xb = op.b;
      s.comm_dim = dmap[op.b];
      std::swap(dmap[op.a], dmap[op.b]);
    }
    memcpy(s.gd_out, gd, sizeof gd);
    memcpy(s.dmap_out, dmap, sizeof dmap);
    local_dims(pg, s.gd_in, s.dmap_in, s.ld_in);
    local_dims(pg, s.gd_out, s.dmap_out, s.ld_out);
    rep_dims(pg, s.gd_in, s.dmap_in, s.rep_in);
    rep_dims(pg, s.gd_out, s.dmap_out, s.rep_out);
    out->push_back(s);
  }
  for (int i = 0; i < 3; i++)
    if (gd[i] != g2.Gdims[i]) {
      char m[200];
      snprintf(m, sizeof m, "output grid dimension %d is %d but the transform produces %d", i, g2.Gdims[i], gd[i]);
      *err = m;
      return false;
    }
  if (out->empty()) {  // nothing to do but (maybe) reorder: one copy stage along the output's leading dimension
    ProtoStage s;
    memset(&s, 0, sizeof s);
    s.kind = P3DFFTCU_K_EMPTY;
    {
      int rd[3];
      rep_dims(pg, gd, dmap, rd);
      s.dim = lead_dim(g2.MemOrder, rd);
    }
    s.dt_in = s.dt_out = dt;
    s.nfft = s.n_in = s.n_out = gd[s.dim];
    s.xb = s.comm_dim = -1;
    memcpy(s.gd_in, gd, sizeof gd);
    memcpy(s.gd_out, gd, sizeof gd);
    memcpy(s.dmap_in, dmap, sizeof dmap);
    memcpy(s.dmap_out, dmap, sizeof dmap);
    local_dims(pg, gd, dmap, s.ld_in);
    local_dims(pg, gd, dmap, s.ld_out);
    rep_dims(pg, gd, dmap, s.rep_in);
    rep_dims(pg, gd, dmap, s.rep_out);
    out->push_back(s);
  }
  return true;
}

// choose the storage order of every intermediate array; returns total cost
double choose_layouts(const std::vector<ProtoStage> &ps, const int mo1[3], const int mo2[3], std::vector<int> *perm_idx) {
  size_t S = ps.size();
  std::vector<int> cur(S + 1, 0), best;
  double bestc = 1e30;
  size_t ninter = S - 1;
  long long combos = 1;
  for (size_t i = 0; i < ninter; i++) combos *= 6;
  for (long long c = 0; c < combos; c++) {
    long long x = c;
    double tot = 0;
    const int *prev = mo1;
    for (size_t s = 0; s < S; s++) {
      const int *next = (s + 1 == S) ? mo2 : kPerm[x % 6];
      if (s + 1 < S) {
        cur[s + 1] = (int)(x % 6);
        x /= 6;
      }
      tot += stage_cost(ps[s].dim, prev, ps[s].rep_in, next, ps[s].rep_out);
      prev = next;
    }
    if (tot < bestc - 1e-9) {
      bestc = tot;
      best = cur;
    }
  }
  *perm_idx = best;
  return bestc;
}

int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return e ? atoi(e) : dflt;
}

void fill_stage(StagePlan &st, const ProtoStage &p, const int mo_in[3], const int mo_out[3], const ProcGrid *pg, int prec,
                bool inter_in, bool inter_out) {
  // intermediate (library-owned) arrays get 128-byte aligned rows; P3DFFT_B200_NO_PAD=1 keeps them dense
  const bool padding = !env_int("P3DFFT_B200_NO_PAD", 0);
  const int pad_in = (inter_in && padding) ? 128 / (p.dt_in * prec) : 0, pad_out = (inter_out && padding) ? 128 / (p.dt_out * prec) : 0;
  st.kind = p.kind;
  st.dim = p.dim;
  st.u = (p.dim + 1) % 3;
  st.v = (p.dim + 2) % 3;
  if (st.u > st.v) std::swap(st.u, st.v);
  st.dt_in = p.dt_in;
  st.dt_out = p.dt_out;
  st.nfft = p.nfft;
  st.n_in = p.n_in;
  st.n_out = p.n_out;
  st.exchange = p.exchange;
  st.xdim_gather = p.xb;
  st.comm_dim = p.comm_dim;
  st.ref_kind = p.exchange ? (p.kind == P3DFFTCU_K_EMPTY ? MPI_ONLY : TRANSMPI) : TRANS_ONLY;
  st.in.set(p.ld_in, mo_in, pad_in);
  st.out.set(p.ld_out, mo_out, pad_out);
  st.in_bytes = st.in.span * p.dt_in * prec;
  st.out_bytes = st.out.span * p.dt_out * prec;
  p3dfftcu_stage_desc &d = st.desc;
  memset(&d, 0, sizeof d);
  d.kind = p.kind;
  d.prec = prec;
  d.dt_in = p.dt_in;
  d.dt_out = p.dt_out;
  d.nfft = p.nfft;
  d.n_in = p.n_in;
  d.n_out = p.n_out;
  d.nu = p.ld_in[st.u];
  d.nv = p.ld_in[st.v];
  d.is_d = st.in.stride[st.dim];
  d.is_u = st.in.stride[st.u];
  d.is_v = st.in.stride[st.v];
  st.peers.clear();
  if (!p.exchange) {
    d.nseg = 1;
    d.seg[0].k0 = 0;
    d.seg[0].k1 = p.n_out;
    d.seg[0].slot = 0;
    d.seg[0].off = 0;
    d.seg[0].os_d = st.out.stride[st.dim];
    d.seg[0].os_u = st.out.stride[st.u];
    d.seg[0].os_v = st.out.stride[st.v];
    return;
  }
  const int c = p.comm_dim, np = pg->ProcDims[c], me = pg->grid_id_cart[c];
  int b_st, b_sz;
  block(p.gd_in[p.xb], np, me, &b_st, &b_sz);  // my block of the gathered dimension before the exchange
  d.nseg = np;
  for (int q = 0; q < np; q++) {
    PeerSeg ps;
    int coords[3] = {pg->grid_id_cart[0], pg->grid_id_cart[1], pg->grid_id_cart[2]};
    coords[c] = q;
    ps.peer_world = pg->rank_of(coords);
    ps.peer_sub = q;
    int kst, ksz;
    block(p.gd_out[p.dim], np, q, &kst, &ksz);
    ps.k0 = kst;
    ps.k1 = kst + ksz;
    int ldq[3] = {p.ld_out[0], p.ld_out[1], p.ld_out[2]};
    ldq[p.dim] = ksz;
    ps.lay.set(ldq, mo_out, pad_out);
    ps.b_off = b_st;
    st.peers.push_back(ps);
    d.seg[q].k0 = ps.k0;
    d.seg[q].k1 = ps.k1;
    d.seg[q].slot = q;
    d.seg[q].off = (long long)b_st * ps.lay.stride[p.xb];
    d.seg[q].os_d = ps.lay.stride[st.dim];
    d.seg[q].os_u = ps.lay.stride[st.u];
    d.seg[q].os_v = ps.lay.stride[st.v];
  }
}

// ---- overlap pairs: cut an exchange stage X and a neighbouring local stage L into the same chunks

// chunk [c0, c1) of stage st along logical dimension cdim (one of st.u, st.v)
bool make_chunk(const StagePlan &st, int prec, int cdim, int c0, int c1, StagePlan::Chunk *out, std::string *err) {
  out->handle = nullptr;
  out->in_off_bytes = 0;
  if (c1 <= c0) return true;
  p3dfftcu_stage_desc d = st.desc;
  d.whole_sm_ctas = 1;  // the two stages of a pair split the SMs by CTA count
  const bool along_u = cdim == st.u;
  long long in_off;
  if (along_u) {
    d.nu = c1 - c0;
    in_off = (long long)c0 * d.is_u;
    for (int q = 0; q < d.nseg; q++) d.seg[q].off += (long long)c0 * d.seg[q].os_u;
  } else {
    d.nv = c1 - c0;
    in_off = (long long)c0 * d.is_v;
    for (int q = 0; q < d.nseg; q++) d.seg[q].off += (long long)c0 * d.seg[q].os_v;
  }
  out->in_off_bytes = in_off * st.dt_in * prec;
  if (p3dfftcu_stage_create(&d, &out->handle)) {
    *err = std::string("chunk stage setup failed: ") + p3dfftcu_last_error();
    return false;
  }
  return true;
}

bool plan_overlap(Plan *pl, const std::vector<ProtoStage> &protos) {
  const size_t S = pl->stages.size();
  if (pl->nranks < 2 || S < 2 || !env_int("P3DFFT_B200_OVERLAP", 1)) return true;
  int nchunks = env_int("P3DFFT_B200_OVERLAP_CHUNKS", 8);
  if (nchunks < 2) return true;
  if (nchunks > 16) nchunks = 16;
  std::vector<bool> used(S, false);
  for (size_t x = 0; x < S; x++) {
    StagePlan &X = pl->stages[x];
    if (!X.exchange || X.kind == P3DFFTCU_K_EMPTY || used[x]) continue;
    // candidate partners: the local stage before (its input must not be the buffer the peers write: it has to be the
    // user's array, i.e. stage 0) or after (its output must not be the buffer X reads: last stage, or X is stage 0)
    int l = -1, mode = StagePlan::PAIR_NONE;
    if (x >= 1 && x - 1 == 0 && !used[x - 1] && !pl->stages[x - 1].exchange && pl->stages[x - 1].kind != P3DFFTCU_K_EMPTY) {
      l = (int)x - 1;
      mode = StagePlan::PAIR_L_THEN_X;
    } else if (x + 1 < S && !used[x + 1] && !pl->stages[x + 1].exchange && pl->stages[x + 1].kind != P3DFFTCU_K_EMPTY &&
               (x + 2 == S || x == 0)) {
      l = (int)x + 1;
      mode = StagePlan::PAIR_X_THEN_L;
    }
    if (l < 0) continue;
    StagePlan &L = pl->stages[l];
    if (L.dim == X.dim) continue;
    const int cdim = 3 - L.dim - X.dim;
    // X before L: the chunks are ranges of the RECEIVED array, so cdim must keep its distribution across the exchange
    if (mode == StagePlan::PAIR_X_THEN_L && cdim == X.xdim_gather) continue;
    // rank-independent chunk size from the largest block of cdim (every rank must run the same number of barriers);
    // multiples of 16 keep the tiles of both stages whole
    const ProtoStage &px = protos[x];
    const int rep = mode == StagePlan::PAIR_L_THEN_X ? px.rep_in[cdim] : px.rep_out[cdim];
    int cs = (rep + nchunks - 1) / nchunks;
    int align = env_int("P3DFFT_B200_OVERLAP_ALIGN", 16);  // (tests use 1 to cut small and uneven grids)
    if (align < 1) align = 1;
    cs = (cs + align - 1) / align * align;
    if (cs <= 0 || rep < 2 * cs) continue;  // too small to be worth cutting
    const int mine = mode == StagePlan::PAIR_L_THEN_X ? X.in.ldims[cdim] : X.out.ldims[cdim];
    StagePlan &first = mode == StagePlan::PAIR_L_THEN_X ? L : X;
    first.pair = mode;
    L.chunk_dim = X.chunk_dim = cdim;
    // one persistent launch per stage (tile groups + flags) when the kernels of BOTH stages have that form on EVERY rank
    // (a rank without local pencils counts as capable); otherwise one launch + one peer barrier per chunk
    bool sync_ok = env_int("P3DFFT_B200_PAIR_SYNC", 1) != 0;
    StagePlan *both[2] = {&L, &X};
    for (int i = 0; i < 2 && sync_ok; i++) {
      p3dfftcu_stage_desc d = both[i]->desc;
      d.whole_sm_ctas = 1;  // the two kernels split the SMs by CTA count
      if (p3dfftcu_stage_create(&d, &both[i]->pair_handle)) {
        pl->error = std::string("pair stage setup failed: ") + p3dfftcu_last_error();
        return false;
      }
      const bool empty = d.nu <= 0 || d.nv <= 0 || d.n_in <= 0;
      if (!empty && !p3dfftcu_stage_sync_capable(both[i]->pair_handle)) sync_ok = false;
    }
    int ok_l = sync_ok ? 1 : 0, ok_g = ok_l;
    MPI_Allreduce(&ok_l, &ok_g, 1, MPI_INT, MPI_MIN, pl->comm);
    L.pair_sync = X.pair_sync = ok_g != 0;
    // chunk boundaries (rank-independent, clipped to this rank's block below).  With flags a chunk costs nothing, so the
    // first and the last full chunk are halved: the pipeline fills and drains on half a chunk
    std::vector<int> bnd;
    const int C0 = (rep + cs - 1) / cs;
    const bool halves = L.pair_sync && env_int("P3DFFT_B200_PAIR_HALVES", 1) && cs % 2 == 0 && (cs / 2) % std::max(1, align / 2) == 0 &&
                        C0 >= 2 && C0 + 2 <= P3DFFTCU_MAXGRP && C0 + 2 <= WS_FLAGS_PER_SRC;
    for (int c = 0; c <= C0; c++) {
      bnd.push_back(c * cs);
      if (halves && (c == 0 || c == C0 - 1)) bnd.push_back(c * cs + cs / 2);
    }
    const int C = (int)bnd.size() - 1;
    L.chunks.resize(C);
    X.chunks.resize(C);
    L.chunk_range.resize(C);
    X.chunk_range.resize(C);
    for (int c = 0; c < C; c++) L.chunk_range[c] = X.chunk_range[c] = std::make_pair(std::min(bnd[c], mine), std::min(bnd[c + 1], mine));
    if (!L.pair_sync)
      for (int c = 0; c < C; c++) {
        const int c0 = L.chunk_range[c].first, c1 = L.chunk_range[c].second;
        if (!make_chunk(L, pl->prec, cdim, c0, c1, &L.chunks[c], &pl->error)) return false;
        if (!make_chunk(X, pl->prec, cdim, c0, c1, &X.chunks[c], &pl->error)) return false;
      }
    used[x] = used[l] = true;
    // L -> X -> Z triple (see plan.h): the stage after X transforms the dimension X gathers, as the plan's last stage
    if (L.pair_sync && mode == StagePlan::PAIR_L_THEN_X && x + 2 == S && env_int("P3DFFT_B200_TRIPLE", 1)) {
      StagePlan &Z = pl->stages[x + 1];
      const int a = L.dim;
      bool tri_ok = !Z.exchange && Z.kind != P3DFFTCU_K_EMPTY && Z.dim == X.xdim_gather && Z.dim == cdim && C >= 4;
      if (tri_ok) {
        p3dfftcu_stage_desc d = Z.desc;
        d.whole_sm_ctas = 1;
        if (p3dfftcu_stage_create(&d, &Z.pair_handle)) {
          pl->error = std::string("pair stage setup failed: ") + p3dfftcu_last_error();
          return false;
        }
        const bool empty = d.nu <= 0 || d.nv <= 0 || d.n_in <= 0;
        if (!empty && !p3dfftcu_stage_sync_capable(Z.pair_handle)) tri_ok = false;
      }
      int t_l = tri_ok ? 1 : 0, t_g = t_l;
      MPI_Allreduce(&t_l, &t_g, 1, MPI_INT, MPI_MIN, pl->comm);
      if (t_g) {
        // second phase = the last chunks, about `frac` of the chunk dimension; cut along a into K pieces (multiples of 16)
        // the second phase cannot start before L is complete: the first phase has to last as long as L does on its half
        // of the SMs (~37 GB/s per SM) while X sends at ~630 GB/s (measured, profiles/r02_pairs_4gpu_trace.txt)
        int pct = 50;
        {
          double lb = 1, xb = 0;
          for (int i = 0; i < 3; i++) lb *= protos[l].rep_in[i];
          lb = lb * protos[l].dt_in * pl->prec;
          double ob = 1;
          for (int i = 0; i < 3; i++) ob *= protos[l].rep_out[i];
          lb += ob * protos[l].dt_out * pl->prec;
          const int np = pl->pgrid->ProcDims[px.comm_dim];
          xb = ob * protos[l].dt_out * pl->prec * (np - 1) / np;
          const double t_l = lb / (74 * 36.8e9), t_x = xb / 630e9;
          const double f1 = std::min(0.8, std::max(0.5, t_x > 0 ? t_l / t_x : 0.5));
          pct = (int)(100 * (1 - f1) + 0.5);
        }
        const int frac_pct = std::min(90, std::max(10, env_int("P3DFFT_B200_TRIPLE_PCT", pct)));
        int c1 = C;
        while (c1 > 1 && (bnd[C] - bnd[c1 - 1]) * 100 <= (long long)bnd[C] * frac_pct) c1--;
        if (c1 == C) c1 = C - 1;  // (at least the last chunk)
        const int K = std::min(std::max(2, env_int("P3DFFT_B200_TRIPLE_PIECES", 4)), 8);
        const int rep_a = px.rep_in[a], mine_a = X.in.ldims[a];
        int ka = (rep_a + K - 1) / K;
        ka = (ka + 15) / 16 * 16;
        const int Ka = (rep_a + ka - 1) / ka;
        if (c1 >= 1 && c1 < C && Ka >= 2 && c1 + 1 + Ka <= P3DFFTCU_MAXGRP) {
          X.triple = true;
          X.tri_c1 = c1;
          for (int k = 0; k < Ka; k++) X.tri_a_range.push_back(std::make_pair(std::min(k * ka, mine_a), std::min((k + 1) * ka, mine_a)));
          used[x + 1] = true;
        }
      }
    }
  }
  bool any = false;
  for (size_t s = 0; s < S; s++) any = any || pl->stages[s].pair != StagePlan::PAIR_NONE;
  if (any) {
    if (p3dfftcu_malloc(&pl->ctl, 8 * 96)) {
      pl->error = std::string("pair control block: ") + p3dfftcu_last_error();
      return false;
    }
    if (p3dfftcu_stream_create(&pl->xstream, 1)) {
      pl->error = std::string("side stream: ") + p3dfftcu_last_error();
      return false;
    }
    for (int i = 0; i < 2 + P3DFFTCU_MAXGRP; i++) {
      void *e = nullptr;
      if (p3dfftcu_event_create(&e)) {
        pl->error = std::string("event: ") + p3dfftcu_last_error();
        return false;
      }
      pl->sync_events.push_back(e);
    }
  }
  return true;
}

bool finish_plan(Plan *pl, const std::vector<ProtoStage> &protos, const std::vector<int> &perm) {
  const int *mo1 = pl->g1->MemOrder, *mo2 = pl->g2->MemOrder;
  size_t S = protos.size();
  pl->stages.resize(S);
  for (size_t s = 0; s < S; s++) {
    const int *mi = s == 0 ? mo1 : kPerm[perm[s]];
    const int *mo = s + 1 == S ? mo2 : kPerm[perm[s + 1]];
    fill_stage(pl->stages[s], protos[s], mi, mo, pl->pgrid, pl->prec, s > 0, s + 1 < S);
    if (pl->stages[s].desc.nseg > P3DFFTCU_MAXSEG) {
      pl->error = "too many ranks in one exchange sub-communicator for this build (max 32)";
      return false;
    }
  }
  pl->in_bytes = pl->stages.front().in_bytes;
  pl->out_bytes = pl->stages.back().out_bytes;
  long long wb = 0;
  for (size_t s = 0; s < S; s++) wb = std::max(wb, pl->stages[s].out_bytes);
  wb = std::max(wb, pl->in_bytes);
  long long wmax = wb;
  MPI_Allreduce(&wb, &wmax, 1, MPI_LONG_LONG, MPI_MAX, pl->comm);
  pl->work_bytes = (wmax + 511) & ~511LL;
  pl->stage_ms.assign(S, 0.f);
  if (gpu_ready()) {
    bool ok = true;
    for (size_t s = 0; s < S && ok; s++) {
      if (p3dfftcu_stage_create(&pl->stages[s].desc, &pl->stages[s].handle)) {
        pl->error = std::string("stage setup failed: ") + p3dfftcu_last_error();
        ok = false;
      }
    }
    if (ok && !plan_overlap(pl, protos)) ok = false;
    // planning is collective: a failure on one rank must fail the plan everywhere instead of leaving the others waiting in
    // the workspace exchange below
    int ok_local = ok ? 1 : 0, ok_all = ok_local;
    MPI_Allreduce(&ok_local, &ok_all, 1, MPI_INT, MPI_MIN, pl->comm);
    if (!ok_all) {
      if (ok) pl->error = "stage setup failed on another rank";
      return false;
    }
    // a single local stage writes straight into the user's array: no work buffers (an in == out call takes a lazily
    // allocated scratch, executor.cpp)
    bool need_ws = S > 1 || pl->nranks > 1;
    for (size_t s = 0; s < S; s++) need_ws = need_ws || pl->stages[s].exchange;
    std::string err;
    if (need_ws) {
      if (!workspace_reserve(pl->work_bytes, pl->comm, pl->nranks, pl->rank, &err, &pl->world_of)) {
        pl->error = err;
        return false;
      }
    } else pl->world_of.assign(1, 0);
  }
  return true;
}

}  // namespace

Plan::Plan()
    : ok(false), prec(0), dt_in(0), dt_out(0), nranks(1), rank(0), comm(MPI_COMM_NULL), g1(nullptr), g2(nullptr), pgrid(nullptr),
      in_bytes(0), out_bytes(0), work_bytes(0), timed_execs(0), events_valid(false), acc_execs(0), last_deriv_stage(-1), xstream(nullptr), ctl(nullptr) {}

Plan::~Plan() {
  for (size_t s = 0; s < stages.size(); s++) {
    if (stages[s].handle) p3dfftcu_stage_destroy(stages[s].handle);
    if (stages[s].pair_handle) p3dfftcu_stage_destroy(stages[s].pair_handle);
    for (size_t c = 0; c < stages[s].chunks.size(); c++)
      if (stages[s].chunks[c].handle) p3dfftcu_stage_destroy(stages[s].chunks[c].handle);
  }
  for (size_t i = 0; i < events.size(); i++) p3dfftcu_event_destroy(events[i]);
  for (size_t i = 0; i < sync_events.size(); i++) p3dfftcu_event_destroy(sync_events[i]);
  if (xstream) p3dfftcu_stream_destroy(xstream);
  if (ctl) p3dfftcu_free(ctl);
  delete g1;
  delete g2;
  delete pgrid;
}

static Plan *new_plan(const DataGrid &g1, const DataGrid &g2, int dt_in, int dt_out, int prec) {
  Plan *pl = new Plan();
  pl->prec = prec;
  pl->dt_in = dt_in;
  pl->dt_out = dt_out;
  pl->pgrid = new ProcGrid(*g1.Pgrid);
  pl->g1 = new DataGrid(g1);
  pl->g2 = new DataGrid(g2);
  pl->g1->Pgrid = pl->pgrid;
  pl->g2->Pgrid = pl->pgrid;
  pl->comm = pl->pgrid->mpi_comm_glob;
  pl->nranks = pl->pgrid->numtasks;
  pl->rank = pl->pgrid->taskid;
  return pl;
}

Plan *plan3d_create(const DataGrid &g1, const DataGrid &g2, const trans_type3D *type, int dt_in, int dt_out, int prec) {
  if (!(*g1.Pgrid == *g2.Pgrid)) {
    printf("Error in transform3D: processor grids dont match\n");
    MPI_Abort(g1.Pgrid->mpi_comm_glob, 0);
  }
  Plan *pl = new_plan(g1, g2, dt_in, dt_out, prec);
  if (!type->is_set) {
    pl->error = "3D transform type is not set";
    printf("Error in transform3D: %s\n", pl->error.c_str());
    return pl;
  }
  if (g1.nd == 3) {
    cout << "Three-dimensional decomposition is presently not supported" << endl;
    pl->error = "3D decomposition not supported";
    return pl;
  }
  Search S;
  S.pg = pl->pgrid;
  S.r2c_dim = S.c2r_dim = -1;
  int done0 = 0;
  for (int i = 0; i < 3; i++) {
    S.ty[i] = types1D[type->types[i]];
    S.target_dmap[i] = g2.Dmap[i];
    if (S.ty[i]->prec != prec) pl->error = "precision of the 1D types differs from the array type";
    if (S.ty[i]->dt1 < S.ty[i]->dt2) {
      if (S.r2c_dim >= 0) printf("ERror in transform3D: more than one real-to-complex 1D transform\n");
      S.r2c_dim = i;
    } else if (S.ty[i]->dt1 > S.ty[i]->dt2) {
      if (S.c2r_dim >= 0) printf("ERror in transform3D: more than one complex-to-real 1D transforms\n");
      S.c2r_dim = i;
    }
    if (S.ty[i]->is_empty) done0 |= 1 << i;
  }
  if (S.r2c_dim >= 0 && S.c2r_dim >= 0) {
    printf("Error in transform3D: can't have both R2C and C2R transforms\n");
    pl->error = "both R2C and C2R in one 3D transform";
  }
  if (!pl->error.empty()) {
    printf("Error in transform3D: %s\n", pl->error.c_str());
    return pl;
  }
  int first = S.r2c_dim >= 0 ? S.r2c_dim : -1;
  if (first >= 0 && S.ty[first]->dt1 != dt_in) printf("Error in transform3D: input datatypes don't match\n");
  int dmap[3] = {g1.Dmap[0], g1.Dmap[1], g1.Dmap[2]};
  for (S.maxx = 0; S.maxx <= 4 && S.found.empty(); S.maxx++) {
    std::vector<Op> path;
    S.dfs(dmap, done0, 0, path, 0);
  }
  if (S.found.empty()) {
    pl->error = "no sequence of transforms and exchanges reaches the requested output distribution";
    printf("Error in transform3D: %s\n", pl->error.c_str());
    return pl;
  }
  double bestc = 1e30;
  std::vector<ProtoStage> bestp;
  std::vector<int> bestperm;
  std::string err;
  for (size_t i = 0; i < S.found.size(); i++) {
    std::vector<ProtoStage> ps;
    if (!build_protos(S.found[i], g1, g2, S.ty, dt_in, &ps, &err)) continue;
    if (ps.back().dt_out != dt_out) {
      err = "output datatype does not match the 3D transform type";
      continue;
    }
    std::vector<int> perm;
    double c = choose_layouts(ps, g1.MemOrder, g2.MemOrder, &perm);
    int unfused = 0;
    for (size_t s = 0; s < ps.size(); s++)
      if (ps[s].exchange && ps[s].kind == P3DFFTCU_K_EMPTY) unfused++;
    c += 1.0 * unfused;
    if (c < bestc - 1e-9) {
      bestc = c;
      bestp = ps;
      bestperm = perm;
    }
  }
  if (bestp.empty()) {
    pl->error = err.empty() ? "planning failed" : err;
    printf("Error in transform3D: %s\n", pl->error.c_str());
    return pl;
  }
  pl->ok = finish_plan(pl, bestp, bestperm);
  if (!pl->ok) printf("Error in transform3D: %s\n", pl->error.c_str());
  return pl;
}

Plan *plan1d_create(const DataGrid &g1, const DataGrid &g2, const gen_trans_type *type, int dim, int dt_in, int dt_out, int prec) {
  Plan *pl = new_plan(g1, g2, dt_in, dt_out, prec);
  if (dim < 0 || dim > 2 || g1.Ldims[dim] != g1.Gdims[dim] || g2.Ldims[dim] != g2.Gdims[dim]) {
    printf("Error in transplan: dimensions dont match %d, %d, %d\n", dim >= 0 && dim < 3 ? g1.Ldims[dim] : -1,
           dim >= 0 && dim < 3 ? g2.Ldims[dim] : -1, dim);
    pl->error = "transform dimension is not local";
    return pl;
  }
  ProtoStage s;
  memset(&s, 0, sizeof s);
  s.kind = type->kind;
  s.dim = dim;
  s.dt_in = type->dt1;
  s.dt_out = type->dt2;
  s.n_in = g1.Gdims[dim];
  if (type->dt1 < type->dt2) {
    s.nfft = g1.Gdims[dim];
    s.n_out = s.nfft / 2 + 1;
  } else if (type->dt1 > type->dt2) {
    s.nfft = g2.Gdims[dim];
    s.n_out = s.nfft;
    if (s.nfft / 2 + 1 != s.n_in) {
      printf("Error in transplan: dimension too small %d, N=%d\n", g1.Gdims[dim], s.nfft);
      pl->error = "complex-to-real sizes inconsistent";
      return pl;
    }
  } else {
    s.nfft = s.n_out = g1.Gdims[dim];
  }
  if (g2.Gdims[dim] != s.n_out) {
    printf("Error in transplan: dimension too small %d, N=%d\n", g2.Gdims[dim], s.nfft);
    pl->error = "output grid size along the transform dimension is inconsistent";
    return pl;
  }
  s.exchange = false;
  s.xb = s.comm_dim = -1;
  for (int i = 0; i < 3; i++) {
    s.gd_in[i] = g1.Gdims[i];
    s.gd_out[i] = g2.Gdims[i];
    s.dmap_in[i] = g1.Dmap[i];
    s.dmap_out[i] = g2.Dmap[i];
    s.ld_in[i] = g1.Ldims[i];
    s.ld_out[i] = g2.Ldims[i];
    if (i != dim && g1.Ldims[i] != g2.Ldims[i]) pl->error = "1D transform cannot change the distribution";
  }
  if (!pl->error.empty()) {
    printf("Error in transplan: %s\n", pl->error.c_str());
    return pl;
  }
  std::vector<ProtoStage> ps(1, s);
  std::vector<int> perm(2, 0);
  pl->ok = finish_plan(pl, ps, perm);
  if (!pl->ok) printf("Error in transplan: %s\n", pl->error.c_str());
  return pl;
}

void plan_destroy(Plan *p) { delete p; }
bool plan_ok(const Plan *p) { return p && p->ok; }
void plan_dims(const Plan *p, int dims1[3], int dims2[3]) {
  for (int i = 0; i < 3; i++) {
    dims1[i] = p->g1->Ldims[i];
    dims2[i] = p->g2->Ldims[i];
  }
}

static void put3(std::ostringstream &o, const char *key, const int a[3]) { o << "\"" << key << "\":[" << a[0] << "," << a[1] << "," << a[2] << "]"; }
static void put3l(std::ostringstream &o, const char *key, const long long a[3]) {
  o << "\"" << key << "\":[" << a[0] << "," << a[1] << "," << a[2] << "]";
}

std::string describe(const Plan &p) {
  std::ostringstream o;
  o << "{\"ok\":" << (p.ok ? "true" : "false") << ",\"prec\":" << p.prec << ",\"dt_in\":" << p.dt_in << ",\"dt_out\":" << p.dt_out
    << ",\"rank\":" << p.rank << ",\"nranks\":" << p.nranks << ",\"in_bytes\":" << p.in_bytes << ",\"out_bytes\":" << p.out_bytes
    << ",\"work_bytes\":" << p.work_bytes << ",\"error\":\"" << p.error << "\",\"stages\":[";
  for (size_t s = 0; s < p.stages.size(); s++) {
    const StagePlan &st = p.stages[s];
    if (s) o << ",";
    o << "{\"kind\":" << st.kind << ",\"ref_kind\":" << st.ref_kind << ",\"dim\":" << st.dim << ",\"u\":" << st.u << ",\"v\":" << st.v
      << ",\"dt_in\":" << st.dt_in << ",\"dt_out\":" << st.dt_out << ",\"nfft\":" << st.nfft << ",\"n_in\":" << st.n_in
      << ",\"n_out\":" << st.n_out << ",\"exchange\":" << (st.exchange ? "true" : "false") << ",\"gather_dim\":" << st.xdim_gather
      << ",\"comm_dim\":" << st.comm_dim << ",";
    put3(o, "in_ldims", st.in.ldims);
    o << ",";
    put3(o, "in_mo", st.in.mo);
    o << ",";
    put3l(o, "in_stride", st.in.stride);
    o << ",";
    put3(o, "out_ldims", st.out.ldims);
    o << ",";
    put3(o, "out_mo", st.out.mo);
    o << ",\"pair\":" << st.pair << ",\"chunk_dim\":" << st.chunk_dim << ",\"chunks\":" << st.chunks.size()
      << ",\"pair_sync\":" << (st.pair_sync ? "true" : "false");
    if (st.pair_handle) o << ",\"pair_variant\":\"" << p3dfftcu_stage_variant(st.pair_handle) << "\"";
    if (st.triple) o << ",\"triple\":true,\"tri_c1\":" << st.tri_c1 << ",\"tri_pieces\":" << st.tri_a_range.size();
    o << ",\"variant\":\"" << (st.handle ? p3dfftcu_stage_variant(st.handle) : "") << "\",\"segs\":[";
    for (int q = 0; q < st.desc.nseg; q++) {
      const p3dfftcu_seg &g = st.desc.seg[q];
      if (q) o << ",";
      o << "{\"k0\":" << g.k0 << ",\"k1\":" << g.k1 << ",\"slot\":" << g.slot << ",\"off\":" << g.off << ",\"os_d\":" << g.os_d
        << ",\"os_u\":" << g.os_u << ",\"os_v\":" << g.os_v << ",\"peer_world\":" << (st.exchange ? st.peers[q].peer_world : p.rank);
      if (st.exchange) {
        o << ",";
        put3(o, "peer_ldims", st.peers[q].lay.ldims);
      }
      o << "}";
    }
    o << "]}";
  }
  o << "]}";
  return o.str();
}

// kept for API compatibility (reference templ.C:627-802): reports the order in which this build's planner
// would run the three 1D transforms for a transform between gr1 and gr2
bool find_order(int L[3], const trans_type3D *tp, const DataGrid *gr1, const DataGrid *gr2, bool *return_steps) {
  *return_steps = false;
  int n = 0;
  bool used[3] = {false, false, false};
  for (int i = 0; i < 3; i++) {
    const gen_trans_type *t = types1D[tp->types[i]];
    if (t->dt1 < t->dt2) {
      L[0] = i;
      used[i] = true;
    }
  }
  int last = -1;
  for (int i = 0; i < 3; i++) {
    const gen_trans_type *t = types1D[tp->types[i]];
    if (t->dt1 > t->dt2) last = i;
  }
  if (used[0] || used[1] || used[2]) n = 1;
  // local dimensions of the input first, the dimension local in the output last
  for (int pass = 0; pass < 2; pass++)
    for (int i = 0; i < 3; i++) {
      if (used[i] || i == last) continue;
      bool loc = gr1->Pdims[i] == 1;
      if ((pass == 0) == loc) {
        L[n++] = i;
        used[i] = true;
      }
    }
  if (last >= 0) L[n++] = last;
  (void)gr2;
  return gr1->Pdims[L[0]] != 1;
}

}  // namespace b200

bool find_order(int L[3], const trans_type3D *tp, const DataGrid *gr1, const DataGrid *gr2, bool *return_steps) {
  return b200::find_order(L, tp, gr1, gr2, return_steps);
}

}  // namespace p3dfft
