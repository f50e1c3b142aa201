// plan.h -- internal host-side representation of a planned 3D (or 1D) transform.
// A Plan is a list of stages; a stage is one launch of a fused stage kernel (1D transform along one
// locally-held dimension + storage reorder + optional exchange).  Replaces the reference's
// stage / transplan / MPIplan / trans_MPIplan class family (include/p3dfft.h:529-650).
#pragma once
#include <string>
#include <vector>

#include "p3dfft.h"

namespace p3dfft {
namespace b200 {

// storage layout of a local 3D block: extent and element stride of each LOGICAL dimension
struct Layout {
  int ldims[3];
  int mo[3];  // MemOrder: storage rank of logical dim i
  long long stride[3];
  long long count() const { return (long long)ldims[0] * ldims[1] * ldims[2]; }
  long long span;  // elements from the first to one past the last, padding included
  // pad_elems > 0 (library-owned intermediate arrays only): the rows along the unit-stride dimension start every
  // multiple of pad_elems elements (128 bytes), so that 513-element rows of a half-complex array stay sector- and
  // bulk-copy-aligned; user-visible arrays are always dense
  void set(const int ld[3], const int mo_[3], int pad_elems = 0);
};

struct PeerSeg {
  int peer_world;  // rank in the global communicator that owns the destination
  int peer_sub;    // its index in the exchange sub-communicator
  int k0, k1;      // range of the transform-dimension output index it receives
  Layout lay;      // layout of the receiver's array
  int b_off;       // where my block starts along the gathered dimension in the receiver's array
};

struct StagePlan {
  int kind;             // P3DFFTCU_K_*
  int ref_kind;         // TRANS_ONLY / MPI_ONLY / TRANSMPI (reporting)
  int dim;              // logical transform (pencil) dimension d
  int u, v;             // the two other logical dimensions
  int dt_in, dt_out;
  int nfft, n_in, n_out;
  Layout in;            // layout of the input block on this rank
  Layout out;           // layout of the output block on this rank (exchange: what this rank RECEIVES)
  bool exchange;
  int xdim_gather;      // logical dim that becomes local (b); valid if exchange
  int comm_dim;         // processor-grid dimension of the sub-communicator
  std::vector<PeerSeg> peers;
  p3dfftcu_stage_desc desc;
  p3dfftcu_stage handle;
  long long out_bytes;  // bytes of this rank's output block
  long long in_bytes;
  // overlap of a fused exchange stage X with a neighbouring local stage L (north_star: "overlapped with the FFT of the
  // next pencil batch"): both stages are cut into the same chunks along a dimension neither of them transforms and run on
  // two streams, chunk c of the later stage waiting for chunk c of the earlier one
  enum { PAIR_NONE = 0, PAIR_L_THEN_X = 1, PAIR_X_THEN_L = 2 };
  int pair;             // set on the FIRST stage of a pair (the second one is the next stage); PAIR_NONE otherwise
  int chunk_dim;        // logical dimension the pair is cut along
  struct Chunk {
    p3dfftcu_stage handle;   // nullptr: empty on this rank (the barrier of the chunk still runs)
    long long in_off_bytes;  // byte offset of the chunk in the stage's input array
  };
  std::vector<Chunk> chunks;  // on BOTH stages of a pair, same count on every rank
  // the same pair as ONE persistent launch per stage (tile groups = the chunks; wait / signal flags inside the kernels
  // instead of one launch + one peer barrier per chunk): used when every rank's kernels have the tile-group form
  std::vector<std::pair<int, int> > chunk_range;  // [c0, c1) of every chunk along chunk_dim on this rank
  p3dfftcu_stage pair_handle;                     // the whole stage planned with CTAs that fill an SM
  bool pair_sync;                                 // (on both stages) agreed by all ranks at plan time
  // local stage L -> exchange stage X -> local stage Z as THREE persistent kernels (set on X; forward slab plans): the
  // first tri_c1 chunks of X follow L chunk by chunk; the rest of the chunk dimension is then cut along L's transform
  // dimension instead (tri_a_range), X publishes each of those pieces to the peers and Z -- which transforms the gathered
  // dimension and needs every sender's share of a piece, nothing else -- runs on it while X is still sending the next
  bool triple;
  int tri_c1;
  std::vector<std::pair<int, int> > tri_a_range;  // [a0, a1) along L.dim on this rank
  StagePlan()
      : handle(nullptr), pair(PAIR_NONE), chunk_dim(-1), pair_handle(nullptr), pair_sync(false), triple(false), tri_c1(0) {}
};

struct Plan {
  bool ok;
  int prec;
  int dt_in, dt_out;
  int nranks, rank;
  MPI_Comm comm;
  std::vector<int> world_of;  // rank in comm -> rank in MPI_COMM_WORLD (the workspace's peer tables are keyed by world rank)
  DataGrid *g1, *g2;
  ProcGrid *pgrid;
  std::vector<StagePlan> stages;
  std::string error;
  long long in_bytes, out_bytes;  // user-visible array sizes on this rank
  long long work_bytes;           // per work buffer, max over stages and ranks
  std::vector<float> stage_ms;
  std::vector<void *> events;  // (S+1) CUDA events per exec recorded around the stages while timers are on
  int timed_execs;             // execs recorded since the stage times were last read (events [0, timed_execs*(S+1)))
  bool events_valid;
  std::vector<double> acc_ms;  // sums of the execs already folded out of the event ring (it holds 64 execs)
  int acc_execs;
  int last_deriv_stage;
  // overlapped pairs: high-priority side stream for the exchange stage, fork/join and per-chunk events
  void *xstream;
  std::vector<void *> sync_events;
  void *ctl;  // device words for the persistent pair kernels: two blocks of 32 (tile counter + group completion counts)
  Plan();
  ~Plan();
};

std::string describe(const Plan &p);
void plan_collect_times(Plan *p, bool final = true);

// process-global device workspace shared by all plans: two ping-pong buffers and one flag array, mapped into the peers
// with CUDA IPC.  Everything that names a peer is keyed by its rank in MPI_COMM_WORLD, so plans on different
// (sub-)communicators neither alias each other's slots nor disturb each other's epochs.
enum {
  WS_MAX_RANKS = 64,
  WS_FLAG_BARRIER0 = 0,                                         // [world rank]: peer barrier words
  WS_FLAGS_PER_SRC = 32,
  WS_FLAG_GROUP0 = WS_MAX_RANKS,                                // [source world rank][id]: tile-group flags set by peers
  WS_FLAG_LOCAL0 = WS_FLAG_GROUP0 + WS_MAX_RANKS * WS_FLAGS_PER_SRC,  // [id]: tile-group flags between two kernels of this GPU
  WS_FLAG_WORDS = WS_FLAG_LOCAL0 + 64
};
struct PeerMap {
  void *buf[2];            // the peer's work buffers mapped here (own rank: the local pointers)
  void *flags;             // its flag array
  unsigned long long gen;  // generation of its buffers these mappings belong to (0 = not mapped)
  PeerMap() : flags(nullptr), gen(0) { buf[0] = buf[1] = nullptr; }
};
struct Workspace {
  void *buf[2];
  long long bytes;
  unsigned long long gen;        // grows every time this rank re-allocates its buffers
  void *flags;                   // WS_FLAG_WORDS 8-byte words, allocated once
  int world_rank, world_size;
  std::vector<PeerMap> peers;    // [world rank]
  // [world rank] number of barriers / flag-synchronised execs this rank has run TOGETHER with that rank: both sides count
  // the same collective calls, so the value is the epoch both expect
  std::vector<unsigned long long> epoch_with;
  unsigned long long local_epoch;  // flags between two kernels of this GPU
  std::vector<void *> retired;     // old buffers a rank outside the reserving communicator may still have mapped
  Workspace() : bytes(0), gen(0), flags(nullptr), world_rank(0), world_size(0), local_epoch(0) { buf[0] = buf[1] = nullptr; }
};
Workspace &workspace();
// collective over comm: make the workspace at least `bytes` per buffer on every rank of comm and (re-)map the members'
// buffers; fails on all ranks together
bool workspace_reserve(long long bytes, MPI_Comm comm, int nranks, int rank, std::string *err, std::vector<int> *world_of);
// lazily grown local scratch of this rank alone (single-stage in == out calls)
void *workspace_bounce(long long bytes);
// device staging buffers for host arrays: which = 0 input, 1 output; grown on demand, shared by all plans
void *workspace_stage(int which, long long bytes);
void workspace_release();

void set_host_staging(const char *mode);
bool gpu_ready();
void *current_stream();

}  // namespace b200
}  // namespace p3dfft
