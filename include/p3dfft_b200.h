/* p3dfft_b200.h -- the two C ABIs that are new in the B200 build.
 *
 * (1) p3dfft_b200_*  : public extensions next to the reference's C API (Cwrap.h): stream control,
 *     synchronisation, plan introspection, launch counters.  Nothing in the reference corresponds to
 *     them; they exist because in/out may now be device pointers.
 * (2) p3dfftcu_*     : the thin GPU layer the C++ host code calls.  It replaces, one for one, the places
 *     where the reference calls into FFTW and MPI on its hot path:
 *       - fftw[f]_plan_many_dft[_r2c|_c2r] / fftw[f]_plan_many_r2r   (build/templ.C:1283-1366,
 *         build/init.C:1191-1607)                                  -> p3dfftcu_stage_create
 *       - fftw[f]_execute_dft[_r2c|_c2r] / fftw[f]_execute_r2r       (build/init.C:1146-1188) fused with
 *         transplan::reorder_trans / reorder_deriv / reorder_out     (build/exec.C:737-2032) and
 *         pack_sendbuf_trans + MPI_Alltoallv + unpack_recvbuf        (build/exec.C:2358-2957, :2698)
 *                                                                    -> p3dfftcu_stage_exec
 *       - compute_deriv<T>                                           (build/deriv.C:85-185)
 *                                                                    -> p3dfftcu_deriv
 *     Plain pointers and sizes only; host code never includes CUDA headers.
 * All p3dfftcu_* functions return 0 on success, non-zero on error (message via p3dfftcu_last_error()).
 */
#ifndef P3DFFT_B200_ABI_H
#define P3DFFT_B200_ABI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- public extensions */
const char *p3dfft_b200_version(void);
/* CUDA stream (cudaStream_t) every later exec call is enqueued on; NULL = legacy default stream */
void p3dfft_b200_set_stream(void *cuda_stream);
/* wait until all work enqueued by exec calls with device pointers has finished */
void p3dfft_b200_sync(void);
/* number of kernels this library has launched since load (stage kernels, copies excluded) */
long long p3dfft_b200_kernel_launches(void);
/* JSON description of a 3D plan (stages, layouts, exchange segments); returns bytes needed */
size_t p3dfft_b200_describe_plan3d(int plan, char *buf, size_t buflen);
size_t p3dfft_b200_describe_plan1d(int plan, char *buf, size_t buflen);
/* per-stage CUDA-event timing: 1 = on.  Read back with p3dfft_b200_stage_times (ms per stage,
   averaged over the execs of that plan since the previous read, at most 64) */
void p3dfft_b200_enable_timers(int on);
int p3dfft_b200_stage_times(int plan, float *ms, int max_stages);
/* 1 if a usable CUDA device was found at p3dfft_setup() (plans can be built and inspected without one) */
int p3dfft_b200_have_device(void);
/* How host arrays passed to exec calls cross PCIe (also the environment variable P3DFFT_B200_HOST_STAGING):
 *   "ring" (default)  through a ring of pinned chunks, CPU copy on P3DFFT_B200_HOST_THREADS threads (default: cores / ranks, at
 *                     most 16) overlapped with the DMA
 *   "register"        page-lock the array on first use (cudaHostRegister) and remember the range: full link rate from the second
 *                     call on; the application must call p3dfft_b200_host_release(ptr) before it frees such an array
 *   "plain"           a bare cudaMemcpyAsync
 * Arrays the application page-locked itself always take the direct asynchronous copy. */
void p3dfft_b200_set_host_staging(const char *mode);
void p3dfft_b200_host_release(const void *ptr);

/* ---------------------------------------------------------------- thin GPU layer */
enum {
  P3DFFTCU_K_EMPTY = 0,   /* copy / reorder / pack only */
  P3DFFTCU_K_C2C_FWD = 1, /* Y_k = sum x_j exp(-2 pi i jk/N)      (FFTW_FORWARD)  */
  P3DFFTCU_K_C2C_BWD = 2, /* Y_k = sum x_j exp(+2 pi i jk/N)      (FFTW_BACKWARD) */
  P3DFFTCU_K_R2C = 3,     /* real N -> complex N/2+1 */
  P3DFFTCU_K_C2R = 4,     /* complex N/2+1 -> real N, unnormalised */
  P3DFFTCU_K_DCT1 = 5,    /* FFTW_REDFT00 */
  P3DFFTCU_K_DST1 = 6,    /* FFTW_RODFT00 */
  P3DFFTCU_K_DCT2 = 7,    /* FFTW_REDFT10 */
  P3DFFTCU_K_DST2 = 8,    /* FFTW_RODFT10 */
  P3DFFTCU_K_DCT3 = 9,    /* FFTW_REDFT01 */
  P3DFFTCU_K_DST3 = 10,   /* FFTW_RODFT01 */
  P3DFFTCU_K_DCT4 = 11,   /* FFTW_REDFT11 */
  P3DFFTCU_K_DST4 = 12    /* FFTW_RODFT11 */
};

#define P3DFFTCU_MAXSEG 32

/* Where the outputs k in [k0,k1) of every pencil go: element (k,u,v) is written to
 * dst[slot] + (off + (k-k0)*os_d + u*os_u + v*os_v) elements.  One segment per exchange peer
 * (slot selects that peer's buffer); a purely local stage has a single segment. */
typedef struct p3dfftcu_seg {
  int k0, k1;
  int slot;
  int pad_;
  long long off, os_d, os_u, os_v;
} p3dfftcu_seg;

/* One stage = batched 1D transform of nu*nv pencils along dimension d, read with strides is_*,
 * written through the segment table.  Strides are in elements of the respective data type. */
typedef struct p3dfftcu_stage_desc {
  int kind;          /* P3DFFTCU_K_* */
  int prec;          /* 4 or 8 */
  int dt_in, dt_out; /* 1 real, 2 complex (interleaved) */
  int nfft;          /* logical transform length (FFTW's n) */
  int n_in, n_out;   /* elements per pencil read / written */
  int nseg;
  long long nu, nv;
  long long is_d, is_u, is_v;
  p3dfftcu_seg seg[P3DFFTCU_MAXSEG];
  int whole_sm_ctas; /* != 0: prefer a kernel shape whose CTA fills an SM (one CTA per SM), so that the CTA cap of
                        p3dfftcu_stage_exec_capped partitions the SMs cleanly between two overlapped stages */
  int pad2_;
} p3dfftcu_stage_desc;

typedef struct p3dfftcu_stage_s *p3dfftcu_stage;

const char *p3dfftcu_last_error(void);
int p3dfftcu_device_count(void);
/* bind this process to a device; device < 0 picks LOCAL_RANK (or P3DFFT_RANK) modulo device count */
int p3dfftcu_init(int device);
int p3dfftcu_malloc(void **ptr, size_t bytes);
int p3dfftcu_free(void *ptr);
int p3dfftcu_memset(void *ptr, int value, size_t bytes, void *stream);
/* kind: 0 host->device, 1 device->host, 2 device->device; asynchronous on `stream` */
int p3dfftcu_memcpy(void *dst, const void *src, size_t bytes, int kind, void *stream);
int p3dfftcu_stream_sync(void *stream);
/* 1 device memory, 0 host memory (pageable or pinned), <0 error */
int p3dfftcu_pointer_is_device(const void *ptr);
/* ---- host arrays of exec calls (reference semantics: `in` / `out` are ordinary heap arrays, sample/C++/test3D_r2c.C:197-205).
 * 1 if ptr is page-locked host memory (cudaHostAlloc / cudaHostRegister, by anybody), 0 if pageable */
int p3dfftcu_host_is_pinned(const void *ptr);
/* "register" mode: page-lock the user's array (cudaHostRegister) and remember the range, so later execs on the same array
 * copy at the pinned rate.  Returns 0 when [ptr, ptr+bytes) is page-locked now (by this call, an earlier one, or the
 * application), non-zero when it could not be locked (the caller then uses p3dfftcu_memcpy_staged).  Ranges are released by
 * p3dfftcu_host_unpin / _unpin_all, or evicted least-recently-used beyond 32 ranges / P3DFFT_B200_HOST_PIN_MAX_GB (64). */
int p3dfftcu_host_pin(const void *ptr, size_t bytes);
/* release the registration of the range that contains ptr (p3dfft_b200_host_release) */
int p3dfftcu_host_unpin(const void *ptr);
int p3dfftcu_host_unpin_all(void);
/* host <-> device copy of a PAGEABLE array through a double-buffered ring of pinned chunks (the CPU copy of chunk c+1
 * overlaps the DMA of chunk c); kind 0 host->device, 1 device->host; returns when the copy is complete */
int p3dfftcu_memcpy_staged(void *dst, const void *src, size_t bytes, int kind, void *stream);
/* how many ranks of the job share this host: the ring's copy threads default to cores / ranks (at most 16) */
void p3dfftcu_host_ranks_hint(int ranks_on_this_host);

int p3dfftcu_stage_create(const p3dfftcu_stage_desc *desc, p3dfftcu_stage *out);
int p3dfftcu_stage_destroy(p3dfftcu_stage st);
/* deriv_g > 0: multiply output k by i*k (k<g/2), 0 (k==g/2), i*(k-g) (k>g/2)  (exec.C:228-287) */
int p3dfftcu_stage_exec(p3dfftcu_stage st, const void *in, void *const *dst, int ndst, int deriv_g, void *stream);
/* same, with the number of CTAs capped at max_ctas (<= 0: no cap): lets two stages share the SMs when they overlap */
int p3dfftcu_stage_exec_capped(p3dfftcu_stage st, const void *in, void *const *dst, int ndst, int deriv_g, void *stream,
                               int max_ctas);
/* ---- one persistent launch per stage of an overlapped pair (replaces "one launch + one peer barrier per chunk"):
 * the stage's pencils are cut into groups -- rectangles [u0,u1) x [v0,v1) of the (u, v) pencil plane, processed in table
 * order.  A group may wait for a flag before its pencils are loaded (its input comes from another kernel that is still
 * running: the neighbouring stage on this GPU, or the exchange stages of the peer GPUs) and may publish a flag once all
 * its outputs are stored (to this GPU's flag words or, over NVLink, to the peers').  Flag words hold epochs that only grow.
 * ctl: >= 1 + ngroups zeroed 8-byte device words ([0] tile counter, [1 + g] completion count of group g); launches that
 * share one ctl share the tiles (late-joining CTAs on another stream). */
#define P3DFFTCU_MAXGRP 24
typedef struct p3dfftcu_group {
  int u0, u1, v0, v1;
  int wait_id;   /* >= 0: flag id that every wait source must have published; < 0: no wait */
  int signal_id; /* >= 0: flag id published to every signal target on completion; < 0: none */
  int count;     /* != 0: count the group's completed pencils although it publishes nothing (see `after`) */
  int after;     /* publish only once the groups [0, after) are complete as well (they must count or signal) */
} p3dfftcu_group;
typedef struct p3dfftcu_sync {
  int ngroups;
  p3dfftcu_group grp[P3DFFTCU_MAXGRP];
  void *ctl;
  /* kernels that take their work from the counter only (p3dfftcu_stage_sync_capable() == 2): launch with up to boost_ctas
   * CTAs (<= 0: no boost); those beyond max_ctas retire once the first boost_groups groups have been handed out, so the
   * stage starts on the whole GPU and then leaves room for its partner kernel */
  int boost_ctas, boost_groups;
  int wait_n;                             /* wait sources: flag word (j, id) = ((uint64*)wait_base)[wait_off[j] + id] */
  const void *wait_base;
  int wait_off[P3DFFTCU_MAXSEG];
  unsigned long long wait_epoch[P3DFFTCU_MAXSEG];
  int sig_n;                              /* signal targets: flag word (j, id) = ((uint64*)sig_ptr[j])[id] */
  void *sig_ptr[P3DFFTCU_MAXSEG];
  unsigned long long sig_epoch[P3DFFTCU_MAXSEG];
} p3dfftcu_sync;
/* 1 if the kernel variant picked for this stage can run with tile groups (the TMA-fed power-of-two kernels) */
int p3dfftcu_stage_sync_capable(p3dfftcu_stage st);
int p3dfftcu_stage_exec_sync(p3dfftcu_stage st, const void *in, void *const *dst, int ndst, int deriv_g, void *stream,
                             int max_ctas, const p3dfftcu_sync *sy);
/* writes epoch[j] into word ids[i] of every flag array ptrs[j] (groups that are empty on this rank still have to be
 * published), stream-ordered */
int p3dfftcu_flags_publish(void *const *ptrs, const unsigned long long *epochs, int n, const int *ids, int nids, void *stream);
/* seconds a kernel waits for a peer (barrier or flag) before it traps; 0 = for ever (default; P3DFFT_B200_PEER_TIMEOUT_S) */
void p3dfftcu_set_peer_timeout(double seconds);

/* human-readable name of the kernel variant picked for this stage */
const char *p3dfftcu_stage_variant(p3dfftcu_stage st);

/* out = i*kappa*in along storage dimension ldir of a complex array with storage extents sd[3];
 * kappa from global index gstart+local index and full length g (deriv.C:85-185) */
int p3dfftcu_deriv(const void *in, void *out, int prec, const int sd[3], int ldir, int g, int gstart, void *stream);

/* side stream for overlapped stages: high_priority != 0 asks for the device's greatest priority */
int p3dfftcu_stream_create(void **stream, int high_priority);
int p3dfftcu_stream_destroy(void *stream);
int p3dfftcu_stream_wait_event(void *stream, void *ev);
int p3dfftcu_num_sms(void);

/* events (per-stage timers, cross-stream dependencies) */
int p3dfftcu_event_create(void **ev);
int p3dfftcu_event_destroy(void *ev);
int p3dfftcu_event_record(void *ev, void *stream);
int p3dfftcu_event_elapsed(void *ev0, void *ev1, float *ms);

/* peer memory for the fused exchange: export a cudaMalloc'd buffer, open a peer's export */
#define P3DFFTCU_IPC_BYTES 64
int p3dfftcu_ipc_export(void *ptr, char handle[P3DFFTCU_IPC_BYTES]);
int p3dfftcu_ipc_open(const char handle[P3DFFTCU_IPC_BYTES], void **ptr);
int p3dfftcu_ipc_close(void *ptr);
/* stream-ordered barrier among n peers (self included or not): writes epochs[j] into word `my_slot` of
 * peer_flags[j], then waits until word peer_slots[j] of my_flags reaches epochs[j].  Flag arrays are 8-byte
 * words, zero-initialised, one word per world rank; epochs[j] = number of barriers this rank and peer j have run
 * together (it grows from call to call and both sides count the same calls). */
int p3dfftcu_peer_barrier(void *const *peer_flags, const int *peer_slots, int n, void *my_flags, int my_slot,
                          const unsigned long long *epochs, void *stream);

long long p3dfftcu_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
