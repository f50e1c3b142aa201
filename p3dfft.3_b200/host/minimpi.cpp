// minimpi.cpp -- implementation of include/compat/mpi.h: a one-host MPI subset over POSIX
// shared memory.  It provides exactly what the P3DFFT++ API and samples need
// (reference build/init.C:1631-1672 Cartesian topology + sub-communicators,
// build/exec.C:2317 MPI_Alltoallv, samples' Bcast/Reduce/Barrier/Dims_create) so that the
// host code and the unmodified reference samples link without an MPI installation.
// One process per rank; collectives copy through per-rank shared data files.
#include "mpi.h"

#include <atomic>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <algorithm>

#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

namespace {

constexpr int kMaxRanks = 64;
constexpr int kMaxCtx = 4096;

struct CtxShared {
  std::atomic<int> count;
  std::atomic<int> sense;
};
struct SlotShared {
  std::atomic<uint64_t> data_bytes;  // current size of this rank's data file
  std::atomic<int> scratch[4];
};
struct Ctl {
  std::atomic<int> ready;
  int nranks;
  std::atomic<int> next_ctx;
  std::atomic<int> abort_flag;
  CtxShared ctx[kMaxCtx];
  SlotShared slot[kMaxRanks];
};

struct Comm {
  bool valid = false;
  int ctx = -1;
  int rank = 0;
  std::vector<int> members;  // world ranks, in communicator order
  int ncart = 0;
  int cdims[3] = {1, 1, 1};
  int local_sense = 0;
};

struct PeerMap {
  int fd = -1;
  char *ptr = nullptr;
  uint64_t bytes = 0;
};

bool g_inited = false, g_finalized = false;
int g_rank = 0, g_size = 1;
std::string g_session;
Ctl *g_ctl = nullptr;
std::vector<Comm> g_comms;
PeerMap g_peer[kMaxRanks];

[[noreturn]] void die(const char *msg) {
  fprintf(stderr, "[minimpi rank %d] fatal: %s\n", g_rank, msg);
  if (g_ctl) g_ctl->abort_flag.store(1);
  _exit(3);
}

std::string ctl_name() { return "/p3dfft_mpi_" + g_session; }
std::string data_name(int r) { return "/p3dfft_mpi_" + g_session + "_d" + std::to_string(r); }

void relax(int &spins) {
  if (++spins < 2000) {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  } else if (spins < 20000) {
    sched_yield();
  } else {
    struct timespec ts = {0, 50000};
    nanosleep(&ts, nullptr);
    if (g_ctl && g_ctl->abort_flag.load()) {
      fprintf(stderr, "[minimpi rank %d] another rank aborted; exiting\n", g_rank);
      _exit(3);
    }
  }
}

void cleanup_files() {
  if (g_size <= 1 || g_session.empty()) return;
  shm_unlink(data_name(g_rank).c_str());
  if (g_rank == 0) shm_unlink(ctl_name().c_str());
}

// grow (never shrink) this rank's data file to >= bytes and return its mapping
char *own_area(uint64_t bytes) {
  PeerMap &m = g_peer[g_rank];
  if (m.ptr && m.bytes >= bytes) return m.ptr;
  uint64_t nb = std::max<uint64_t>(bytes, 1 << 20);
  nb = std::max(nb, m.bytes * 2);
  nb = (nb + 4095) & ~uint64_t(4095);
  if (m.fd < 0) {
    m.fd = shm_open(data_name(g_rank).c_str(), O_CREAT | O_RDWR, 0600);
    if (m.fd < 0) die("shm_open(data) failed");
  }
  if (m.ptr) munmap(m.ptr, m.bytes);
  if (ftruncate(m.fd, (off_t)nb) != 0) die("ftruncate(data) failed (is /dev/shm full?)");
  m.ptr = (char *)mmap(nullptr, nb, PROT_READ | PROT_WRITE, MAP_SHARED, m.fd, 0);
  if (m.ptr == MAP_FAILED) die("mmap(data) failed");
  m.bytes = nb;
  g_ctl->slot[g_rank].data_bytes.store(nb);
  return m.ptr;
}

// mapping of a peer's data file, valid after a barrier that follows the peer's publish
const char *peer_area(int r) {
  if (r == g_rank) return g_peer[r].ptr;
  PeerMap &m = g_peer[r];
  uint64_t nb = g_ctl->slot[r].data_bytes.load();
  if (m.ptr && m.bytes == nb) return m.ptr;
  if (m.ptr) munmap(m.ptr, m.bytes);
  if (m.fd < 0) {
    m.fd = shm_open(data_name(r).c_str(), O_RDONLY, 0600);
    if (m.fd < 0) die("shm_open(peer data) failed");
  }
  m.ptr = (char *)mmap(nullptr, nb, PROT_READ, MAP_SHARED, m.fd, 0);
  if (m.ptr == MAP_FAILED) die("mmap(peer data) failed");
  m.bytes = nb;
  return m.ptr;
}

Comm &get(MPI_Comm c) {
  if (!g_inited) {
    int a = 0;
    MPI_Init(&a, nullptr);
  }
  if (c < 0 || c >= (int)g_comms.size() || !g_comms[c].valid) die("invalid communicator");
  return g_comms[c];
}

void barrier(Comm &c) {
  int n = (int)c.members.size();
  if (n <= 1) return;
  CtxShared &s = g_ctl->ctx[c.ctx];
  int sense = !c.local_sense;
  c.local_sense = sense;
  if (s.count.fetch_add(1) + 1 == n) {
    s.count.store(0);
    s.sense.store(sense);
  } else {
    int spins = 0;
    while (s.sense.load() != sense) relax(spins);
  }
}

size_t dt_size(MPI_Datatype dt) {
  switch (dt) {
    case MPI_CHAR: case MPI_BYTE: return 1;
    case MPI_INT: case MPI_UNSIGNED: case MPI_FLOAT: return 4;
    case MPI_LONG: case MPI_LONG_LONG: case MPI_UNSIGNED_LONG: case MPI_DOUBLE: case MPI_COMPLEX: return 8;
    case MPI_DOUBLE_COMPLEX: return 16;
  }
  die("unknown datatype");
}

template <class T> void red(T *acc, const T *x, int n, MPI_Op op) {
  for (int i = 0; i < n; i++) {
    switch (op) {
      case MPI_SUM: acc[i] += x[i]; break;
      case MPI_PROD: acc[i] *= x[i]; break;
      case MPI_MAX: if (x[i] > acc[i]) acc[i] = x[i]; break;
      case MPI_MIN: if (x[i] < acc[i]) acc[i] = x[i]; break;
      default: die("unknown reduction op");
    }
  }
}
void reduce_into(void *acc, const void *x, int n, MPI_Datatype dt, MPI_Op op) {
  switch (dt) {
    case MPI_INT: red((int *)acc, (const int *)x, n, op); break;
    case MPI_UNSIGNED: red((unsigned *)acc, (const unsigned *)x, n, op); break;
    case MPI_LONG: case MPI_LONG_LONG: red((long long *)acc, (const long long *)x, n, op); break;
    case MPI_UNSIGNED_LONG: red((unsigned long long *)acc, (const unsigned long long *)x, n, op); break;
    case MPI_FLOAT: red((float *)acc, (const float *)x, n, op); break;
    case MPI_DOUBLE: red((double *)acc, (const double *)x, n, op); break;
    case MPI_CHAR: case MPI_BYTE: red((char *)acc, (const char *)x, n, op); break;
    default: die("reduction on unsupported datatype");
  }
}

int new_comm(const std::vector<int> &members, int myrank, int ctx) {
  Comm c;
  c.valid = true;
  c.ctx = ctx;
  c.rank = myrank;
  c.members = members;
  g_comms.push_back(c);
  return (int)g_comms.size() - 1;
}

// collective over parent: members with equal color form a new communicator ordered by (key, parent rank)
int split(Comm &p, int color, int key) {
  int n = (int)p.members.size();
  if (n == 1) return new_comm(p.members, 0, 0);
  int *mine = (int *)own_area(2 * sizeof(int));
  mine[0] = color;
  mine[1] = key;
  barrier(p);
  std::vector<std::pair<std::pair<int, int>, int>> grp;  // ((key, parent rank), world rank)
  for (int i = 0; i < n; i++) {
    const int *q = (const int *)peer_area(p.members[i]);
    if (q[0] == color) grp.push_back({{q[1], i}, p.members[i]});
  }
  barrier(p);
  std::sort(grp.begin(), grp.end());
  std::vector<int> members;
  int myrank = -1;
  for (size_t i = 0; i < grp.size(); i++) {
    if (grp[i].second == g_rank) myrank = (int)i;
    members.push_back(grp[i].second);
  }
  int leader = members[0];
  if (leader == g_rank) {
    int id = g_ctl->next_ctx.fetch_add(1);
    if (id >= kMaxCtx) die("out of communicator contexts");
    g_ctl->ctx[id].count.store(0);
    g_ctl->ctx[id].sense.store(0);
    g_ctl->slot[g_rank].scratch[0].store(id);
  }
  barrier(p);
  int id = g_ctl->slot[leader].scratch[0].load();
  barrier(p);
  return new_comm(members, myrank, id);
}

}  // namespace

extern "C" {

int MPI_Init(int *, char ***) {
  if (g_inited) return MPI_SUCCESS;
  const char *r = getenv("P3DFFT_RANK"), *n = getenv("P3DFFT_NRANKS"), *s = getenv("P3DFFT_SESSION");
  if (!r || !n) {
    r = getenv("RANK");
    n = getenv("WORLD_SIZE");
  }
  g_rank = r ? atoi(r) : 0;
  g_size = n ? atoi(n) : 1;
  if (g_size < 1 || g_rank < 0 || g_rank >= g_size || g_size > kMaxRanks) {
    fprintf(stderr, "[minimpi] bad rank/size %d/%d\n", g_rank, g_size);
    _exit(3);
  }
  if (s) g_session = s;
  else {
    const char *port = getenv("MASTER_PORT");
    g_session = std::string(port ? port : "0") + "_" + std::to_string((long)getppid());
  }
  g_inited = true;
  Comm world;
  world.valid = true;
  world.ctx = 0;
  world.rank = g_rank;
  for (int i = 0; i < g_size; i++) world.members.push_back(i);
  g_comms.push_back(world);
  Comm self;
  self.valid = true;
  self.ctx = 1;
  self.rank = 0;
  self.members.push_back(g_rank);
  g_comms.push_back(self);
  if (g_size == 1) return MPI_SUCCESS;

  std::string name = ctl_name();
  int fd = -1;
  if (g_rank == 0) {
    shm_unlink(name.c_str());
    fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) die("shm_open(ctl) failed");
    if (ftruncate(fd, sizeof(Ctl)) != 0) die("ftruncate(ctl) failed");
  } else {
    int spins = 0;
    double t0 = MPI_Wtime();
    for (;;) {
      fd = shm_open(name.c_str(), O_RDWR, 0600);
      if (fd >= 0) {
        struct stat st;
        if (fstat(fd, &st) == 0 && (size_t)st.st_size >= sizeof(Ctl)) break;
        close(fd);
      }
      relax(spins);
      if (MPI_Wtime() - t0 > 600.0) die("timed out waiting for rank 0 (600 s)");
    }
  }
  g_ctl = (Ctl *)mmap(nullptr, sizeof(Ctl), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  if (g_ctl == MAP_FAILED) die("mmap(ctl) failed");
  close(fd);
  if (g_rank == 0) {
    g_ctl->nranks = g_size;
    g_ctl->next_ctx.store(2);
    g_ctl->abort_flag.store(0);
    g_ctl->ready.store(0x600DF00D);
  } else {
    int spins = 0;
    while (g_ctl->ready.load() != 0x600DF00D) relax(spins);
  }
  own_area(1 << 20);
  atexit(cleanup_files);
  barrier(g_comms[0]);
  return MPI_SUCCESS;
}

int MPI_Initialized(int *flag) {
  *flag = g_inited ? 1 : 0;
  return MPI_SUCCESS;
}

int MPI_Finalize(void) {
  if (!g_inited || g_finalized) return MPI_SUCCESS;
  g_finalized = true;
  if (g_size > 1) {
    barrier(g_comms[0]);
    cleanup_files();
  }
  return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm, int code) {
  fprintf(stderr, "[minimpi rank %d] MPI_Abort(%d)\n", g_rank, code);
  if (g_ctl) g_ctl->abort_flag.store(1);
  cleanup_files();
  _exit(code ? code : 1);
}

double MPI_Wtime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int MPI_Comm_rank(MPI_Comm comm, int *rank) {
  *rank = get(comm).rank;
  return MPI_SUCCESS;
}
int MPI_Comm_size(MPI_Comm comm, int *size) {
  *size = (int)get(comm).members.size();
  return MPI_SUCCESS;
}
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *out) {
  Comm p = get(comm);
  int id = split(get(comm), 0, p.rank);
  g_comms[id].ncart = p.ncart;
  memcpy(g_comms[id].cdims, p.cdims, sizeof(p.cdims));
  *out = id;
  return MPI_SUCCESS;
}
int MPI_Comm_free(MPI_Comm *comm) {
  if (*comm >= 2 && *comm < (int)g_comms.size()) g_comms[*comm].valid = false;
  *comm = MPI_COMM_NULL;
  return MPI_SUCCESS;
}
int MPI_Comm_compare(MPI_Comm a, MPI_Comm b, int *result) {
  Comm &x = get(a), &y = get(b);
  if (a == b) *result = MPI_IDENT;
  else if (x.members == y.members) *result = MPI_CONGRUENT;
  else {
    std::vector<int> s = x.members, t = y.members;
    std::sort(s.begin(), s.end());
    std::sort(t.begin(), t.end());
    *result = (s == t) ? MPI_SIMILAR : MPI_UNEQUAL;
  }
  return MPI_SUCCESS;
}
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm *out) {
  *out = split(get(comm), color, key);
  return MPI_SUCCESS;
}
MPI_Comm MPI_Comm_f2c(MPI_Fint f) { return (MPI_Comm)f; }
MPI_Fint MPI_Comm_c2f(MPI_Comm c) { return (MPI_Fint)c; }

// balanced factorisation, non-increasing order, honouring preset (non-zero) entries
int MPI_Dims_create(int nnodes, int ndims, int *dims) {
  int rem = nnodes, nfree = 0;
  for (int i = 0; i < ndims; i++) {
    if (dims[i] > 0) {
      if (rem % dims[i]) die("MPI_Dims_create: preset dims do not divide nnodes");
      rem /= dims[i];
    } else nfree++;
  }
  if (nfree == 0) return MPI_SUCCESS;
  std::vector<int> f(nfree, 1), primes;
  for (int p = 2, m = rem; m > 1;) {
    if (m % p == 0) { primes.push_back(p); m /= p; } else p++;
  }
  for (int i = (int)primes.size() - 1; i >= 0; i--) {  // largest prime first onto the smallest factor
    int k = (int)(std::min_element(f.begin(), f.end()) - f.begin());
    f[k] *= primes[i];
  }
  std::sort(f.begin(), f.end(), [](int a, int b) { return a > b; });
  for (int i = 0, k = 0; i < ndims; i++)
    if (dims[i] <= 0) dims[i] = f[k++];
  return MPI_SUCCESS;
}

int MPI_Cart_create(MPI_Comm comm, int ndims, const int *dims, const int *, int, MPI_Comm *out) {
  if (ndims > 3) die("MPI_Cart_create: at most 3 dimensions supported");
  Comm p = get(comm);
  int prod = 1;
  for (int i = 0; i < ndims; i++) prod *= dims[i];
  if (prod != (int)p.members.size()) die("MPI_Cart_create: grid size differs from communicator size");
  int id = split(get(comm), 0, p.rank);
  g_comms[id].ncart = ndims;
  for (int i = 0; i < 3; i++) g_comms[id].cdims[i] = i < ndims ? dims[i] : 1;
  *out = id;
  return MPI_SUCCESS;
}
int MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int *coords) {
  Comm &c = get(comm);
  int r = rank;
  for (int i = c.ncart - 1; i >= 0; i--) {  // row-major: last dimension varies fastest
    if (i < maxdims) coords[i] = r % c.cdims[i];
    r /= c.cdims[i];
  }
  return MPI_SUCCESS;
}
int MPI_Cart_rank(MPI_Comm comm, const int *coords, int *rank) {
  Comm &c = get(comm);
  int r = 0;
  for (int i = 0; i < c.ncart; i++) r = r * c.cdims[i] + ((coords[i] % c.cdims[i]) + c.cdims[i]) % c.cdims[i];
  *rank = r;
  return MPI_SUCCESS;
}
int MPI_Cart_sub(MPI_Comm comm, const int *remain, MPI_Comm *out) {
  Comm p = get(comm);
  int co[3] = {0, 0, 0};
  MPI_Cart_coords(comm, p.rank, 3, co);
  int color = 0, key = 0, nd = 0, nd_dims[3] = {1, 1, 1};
  for (int i = 0; i < p.ncart; i++) {
    if (remain[i]) {
      key = key * p.cdims[i] + co[i];
      nd_dims[nd++] = p.cdims[i];
    } else color = color * p.cdims[i] + co[i];
  }
  int id = split(get(comm), color, key);
  g_comms[id].ncart = nd;
  memcpy(g_comms[id].cdims, nd_dims, sizeof(nd_dims));
  *out = id;
  return MPI_SUCCESS;
}

int MPI_Barrier(MPI_Comm comm) {
  barrier(get(comm));
  return MPI_SUCCESS;
}

int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm comm) {
  Comm &c = get(comm);
  if (c.members.size() <= 1) return MPI_SUCCESS;
  size_t nb = (size_t)count * dt_size(dt);
  if (c.rank == root) memcpy(own_area(nb), buf, nb);
  barrier(c);
  if (c.rank != root) memcpy(buf, peer_area(c.members[root]), nb);
  barrier(c);
  return MPI_SUCCESS;
}

static int reduce_impl(const void *sbuf, void *rbuf, int count, MPI_Datatype dt, MPI_Op op, int root, bool all, Comm &c) {
  size_t nb = (size_t)count * dt_size(dt);
  const void *src = (sbuf == MPI_IN_PLACE) ? rbuf : sbuf;
  if (c.members.size() <= 1) {
    if (src != rbuf) memcpy(rbuf, src, nb);
    return MPI_SUCCESS;
  }
  memcpy(own_area(nb), src, nb);
  barrier(c);
  if (all || c.rank == root) {
    std::vector<char> acc(nb);
    memcpy(acc.data(), peer_area(c.members[0]), nb);
    for (size_t i = 1; i < c.members.size(); i++) reduce_into(acc.data(), peer_area(c.members[i]), count, dt, op);
    memcpy(rbuf, acc.data(), nb);
  }
  barrier(c);
  return MPI_SUCCESS;
}
int MPI_Reduce(const void *sbuf, void *rbuf, int count, MPI_Datatype dt, MPI_Op op, int root, MPI_Comm comm) {
  return reduce_impl(sbuf, rbuf, count, dt, op, root, false, get(comm));
}
int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm) {
  return reduce_impl(sbuf, rbuf, count, dt, op, 0, true, get(comm));
}

int MPI_Allgather(const void *sbuf, int scount, MPI_Datatype sdt, void *rbuf, int rcount, MPI_Datatype rdt, MPI_Comm comm) {
  Comm &c = get(comm);
  size_t nb = (size_t)scount * dt_size(sdt), rb = (size_t)rcount * dt_size(rdt);
  if (c.members.size() <= 1) {
    memcpy(rbuf, sbuf, nb);
    return MPI_SUCCESS;
  }
  memcpy(own_area(nb), sbuf, nb);
  barrier(c);
  for (size_t i = 0; i < c.members.size(); i++) memcpy((char *)rbuf + i * rb, peer_area(c.members[i]), rb);
  barrier(c);
  return MPI_SUCCESS;
}
int MPI_Gather(const void *sbuf, int scount, MPI_Datatype sdt, void *rbuf, int rcount, MPI_Datatype rdt, int root, MPI_Comm comm) {
  Comm &c = get(comm);
  size_t nb = (size_t)scount * dt_size(sdt), rb = (size_t)rcount * dt_size(rdt);
  if (c.members.size() <= 1) {
    memcpy(rbuf, sbuf, nb);
    return MPI_SUCCESS;
  }
  memcpy(own_area(nb), sbuf, nb);
  barrier(c);
  if (c.rank == root)
    for (size_t i = 0; i < c.members.size(); i++) memcpy((char *)rbuf + i * rb, peer_area(c.members[i]), rb);
  barrier(c);
  return MPI_SUCCESS;
}

int MPI_Alltoallv(const void *sbuf, const int *scounts, const int *sdispls, MPI_Datatype sdt,
                  void *rbuf, const int *rcounts, const int *rdispls, MPI_Datatype rdt, MPI_Comm comm) {
  Comm &c = get(comm);
  int n = (int)c.members.size();
  size_t ss = dt_size(sdt), rs = dt_size(rdt);
  if (n == 1) {
    memcpy((char *)rbuf + (size_t)rdispls[0] * rs, (const char *)sbuf + (size_t)sdispls[0] * ss, (size_t)scounts[0] * ss);
    return MPI_SUCCESS;
  }
  // published record: int64 cnt[n] (bytes), int64 off[n] (bytes into payload), payload
  size_t hdr = 2 * (size_t)n * sizeof(int64_t), total = 0;
  for (int j = 0; j < n; j++) total += (size_t)scounts[j] * ss;
  char *area = own_area(hdr + total);
  int64_t *cnt = (int64_t *)area, *off = cnt + n;
  size_t pos = 0;
  for (int j = 0; j < n; j++) {
    cnt[j] = (int64_t)scounts[j] * (int64_t)ss;
    off[j] = (int64_t)pos;
    memcpy(area + hdr + pos, (const char *)sbuf + (size_t)sdispls[j] * ss, (size_t)cnt[j]);
    pos += (size_t)cnt[j];
  }
  barrier(c);
  for (int i = 0; i < n; i++) {
    const char *pa = peer_area(c.members[i]);
    const int64_t *pc = (const int64_t *)pa, *po = pc + n;
    size_t want = (size_t)rcounts[i] * rs;
    if ((size_t)pc[c.rank] != want) die("MPI_Alltoallv: send/recv count mismatch");
    memcpy((char *)rbuf + (size_t)rdispls[i] * rs, pa + hdr + po[c.rank], want);
  }
  barrier(c);
  return MPI_SUCCESS;
}
int MPI_Alltoall(const void *sbuf, int scount, MPI_Datatype sdt, void *rbuf, int rcount, MPI_Datatype rdt, MPI_Comm comm) {
  int n = (int)get(comm).members.size();
  std::vector<int> sc(n, scount), sd(n), rc(n, rcount), rd(n);
  for (int i = 0; i < n; i++) {
    sd[i] = i * scount;
    rd[i] = i * rcount;
  }
  return MPI_Alltoallv(sbuf, sc.data(), sd.data(), sdt, rbuf, rc.data(), rd.data(), rdt, comm);
}

int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm) { die("MPI_Send is not implemented in the mini-MPI"); }
int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) { die("MPI_Recv is not implemented in the mini-MPI"); }
int MPI_Irecv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { die("MPI_Irecv is not implemented in the mini-MPI"); }
int MPI_Isend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { die("MPI_Isend is not implemented in the mini-MPI"); }
int MPI_Waitall(int, MPI_Request *, MPI_Status *) { die("MPI_Waitall is not implemented in the mini-MPI"); }

}  // extern "C"
