/* p3dfft.h -- public C++/C interface of the B200-native P3DFFT++ transform path.
 *
 * Drop-in for the reference's include/p3dfft.h (namespace p3dfft, lines 140-995) restricted to the
 * 3D-transform execution path: setup/cleanup, ProcGrid, DataGrid, trans_type3D,
 * transform3D<T1,T2>::exec / exec_deriv, transplan<T1,T2>::exec, compute_deriv, inv_mo, the type-ID
 * variables, and (via Cwrap.h / Fwrap.h) the C and Fortran entry points.  Everything behind these
 * declarations is new: the templates below are thin shells over a type-erased plan object that runs
 * hand-written sm_100a kernels through the C ABI in p3dfft_b200.h.  No FFTW, no CPU path.
 *
 * in/out pointers may be HOST pointers (reference semantics; data is staged through the GPU and the
 * call returns when `out` is complete) or DEVICE pointers (no staging; the call is stream-ordered and
 * returns immediately - see p3dfft_b200_sync()).
 */
#ifndef P3DFFT_B200_P3DFFT_H
#define P3DFFT_B200_P3DFFT_H

#include "mpi.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define Type3D int
#define Type1D int
#define Plan3D int
#define Plan1D int

/* C-visible transform type IDs (reference build/init.C:84-89); valid after p3dfft_setup() */
#ifdef __cplusplus
extern "C" {
#endif
extern int P3DFFT_EMPTY_TYPE_SINGLE, P3DFFT_EMPTY_TYPE_DOUBLE, P3DFFT_EMPTY_TYPE_SINGLE_COMPLEX, P3DFFT_EMPTY_TYPE_DOUBLE_COMPLEX;
extern int P3DFFT_R2CFFT_S, P3DFFT_R2CFFT_D, P3DFFT_C2RFFT_S, P3DFFT_C2RFFT_D;
extern int P3DFFT_CFFT_FORWARD_S, P3DFFT_CFFT_FORWARD_D, P3DFFT_CFFT_BACKWARD_S, P3DFFT_CFFT_BACKWARD_D;
#define P3DFFT_B200_DECL_R2R(K)                                                                  \
  extern int P3DFFT_##K##_REAL_S, P3DFFT_##K##_REAL_D, P3DFFT_##K##_COMPLEX_S, P3DFFT_##K##_COMPLEX_D;
P3DFFT_B200_DECL_R2R(DCT1) P3DFFT_B200_DECL_R2R(DST1) P3DFFT_B200_DECL_R2R(DCT2) P3DFFT_B200_DECL_R2R(DST2)
P3DFFT_B200_DECL_R2R(DCT3) P3DFFT_B200_DECL_R2R(DST3) P3DFFT_B200_DECL_R2R(DCT4) P3DFFT_B200_DECL_R2R(DST4)
#ifdef __cplusplus
}
#endif

#ifdef __cplusplus

#include <iostream>
#include <vector>
#include <typeinfo>
#include <complex>

namespace p3dfft {

using namespace std;  // the reference header does this (p3dfft.h:142) and its samples rely on it

static const int REAL = 1;
static const int COMPLEX = 2;
static const int TRANS_ONLY = 1;
static const int MPI_ONLY = 2;
static const int TRANSMPI = 3;

typedef complex<float> mycomplex;
typedef complex<double> complex_double;

extern int EMPTY_TYPE_SINGLE, EMPTY_TYPE_DOUBLE, EMPTY_TYPE_SINGLE_COMPLEX, EMPTY_TYPE_DOUBLE_COMPLEX;
extern int R2CFFT_S, R2CFFT_D, C2RFFT_S, C2RFFT_D, CFFT_FORWARD_S, CFFT_FORWARD_D, CFFT_BACKWARD_S, CFFT_BACKWARD_D;
#define P3DFFT_B200_DECL_R2R_NS(K) extern int K##_REAL_S, K##_REAL_D, K##_COMPLEX_S, K##_COMPLEX_D;
P3DFFT_B200_DECL_R2R_NS(DCT1) P3DFFT_B200_DECL_R2R_NS(DST1) P3DFFT_B200_DECL_R2R_NS(DCT2) P3DFFT_B200_DECL_R2R_NS(DST2)
P3DFFT_B200_DECL_R2R_NS(DCT3) P3DFFT_B200_DECL_R2R_NS(DST3) P3DFFT_B200_DECL_R2R_NS(DCT4) P3DFFT_B200_DECL_R2R_NS(DST4)

void setup();
void cleanup();
void inv_mo(int mo[3], int imo[3]);
void rel_change(int *imo1, int *imo2, int *mc);
int arcmp(int *A, int *B, int N);

/* One registered 1D transform (reference gen_trans_type / trans_type1D, p3dfft.h:235-270,482-512).
 * `kind` is this build's dispatch key (values in p3dfft_b200.h, P3DFFTCU_K_*). */
class gen_trans_type {
 public:
  int isign;
  bool is_set, is_empty;
  int dt1, dt2;  // 1 = real, 2 = complex, before and after
  int prec;      // 4 or 8
  int kind;
  const char *name;
  gen_trans_type(const char *name_, int kind_, int dt1_, int dt2_, int prec_, int isign_ = 0, bool empty_ = false)
      : isign(isign_), is_set(true), is_empty(empty_), dt1(dt1_), dt2(dt2_), prec(prec_), kind(kind_), name(name_) {}
  bool operator==(const gen_trans_type &r) const {
    return r.isign == isign && r.is_set == is_set && r.dt1 == dt1 && r.dt2 == dt2 && r.kind == kind && r.prec == prec;
  }
};

/* Processor grid (reference p3dfft.h:653-699, init.C:1631-1696): 3D Cartesian rank layout, row-major,
 * plus one sub-communicator per grid dimension. */
class ProcGrid {
 public:
  int taskid, numtasks;
  int nd;
  int ProcDims[3];
  int grid_id_cart[3];
  MPI_Comm mpi_comm_glob;
  MPI_Comm mpi_comm_cart;
  MPI_Comm mpicomm[3];
  ProcGrid(int procdims[3], MPI_Comm mpi_comm_init);
  ProcGrid(const ProcGrid &rhs);
  ~ProcGrid();
  bool operator==(const ProcGrid &P) const;
  /* world rank (within mpi_comm_glob) of the process at Cartesian coordinates c */
  int rank_of(const int c[3]) const { return (c[0] * ProcDims[1] + c[1]) * ProcDims[2] + c[2]; }

 private:
  ProcGrid &operator=(const ProcGrid &);
};

/* Distributed 3D array descriptor (reference p3dfft.h:702-755, init.C:1699-1863). */
class DataGrid {
 public:
  int nd;
  int Gdims[3];
  int dim_conj_sym;
  int MemOrder[3];
  int Ldims[3];
  ProcGrid *Pgrid;
  int Pdims[3];
  int Dmap[3];
  int L[3];
  int D[3];
  int GlobStart[3];
  int grid_id[3];
  bool is_set;
  bool IsLocal(int dim) const { return dim >= 0 && dim <= 2 && Pdims[dim] == 1; }
  DataGrid(int *gdims_, int dim_conj_sym_, ProcGrid *pgrid, int *dmap, int *mem_order);
  DataGrid(const DataGrid &rhs);
  DataGrid() : is_set(false) {}
  ~DataGrid() {}
  void set_gdims(int gdims[3]) {
    for (int i = 0; i < 3; i++) Gdims[i] = gdims[i];
    InitPencil();
  }
  void get_gdims(int gdims[3]) const {
    for (int i = 0; i < 3; i++) gdims[i] = Gdims[i];
  }
  void set_mo(int mo[3]) {
    for (int i = 0; i < 3; i++) MemOrder[i] = mo[i];
  }
  /* block distribution tables: first index of / number of points owned by position p along dim i */
  int block_start(int i, int p) const { return st_[i][p]; }
  int block_size(int i, int p) const { return sz_[i][p]; }
  long long local_count() const { return (long long)Ldims[0] * Ldims[1] * Ldims[2]; }

 private:
  void InitPencil();
  vector<int> st_[3], sz_[3];
};

class trans_type3D {
 public:
  char *name;
  int prec;
  bool is_set;
  int types[3];
  trans_type3D(gen_trans_type *types_[3]);
  trans_type3D(int types_IDs[3]);
  trans_type3D(const trans_type3D &rhs);
  ~trans_type3D();

 private:
  void init(const int ids[3]);
};

bool find_order(int L[3], const trans_type3D *tp, const DataGrid *gr1, const DataGrid *gr2, bool *return_steps);

namespace b200 {
struct Plan;  // type-erased stage list + device resources (p3dfft.3_b200/host/plan.h)
Plan *plan3d_create(const DataGrid &g1, const DataGrid &g2, const trans_type3D *type, int dt_in, int dt_out, int prec);
Plan *plan1d_create(const DataGrid &g1, const DataGrid &g2, const gen_trans_type *type, int dim, int dt_in, int dt_out, int prec);
void plan_destroy(Plan *);
bool plan_ok(const Plan *);
void plan_exec(Plan *, const void *in, void *out, int idir, bool OW);
void plan_dims(const Plan *, int dims1[3], int dims2[3]);
template <class T> struct tinfo;
template <> struct tinfo<float> { enum { dt = 1, prec = 4 }; };
template <> struct tinfo<double> { enum { dt = 1, prec = 8 }; };
template <> struct tinfo<mycomplex> { enum { dt = 2, prec = 4 }; };
template <> struct tinfo<complex_double> { enum { dt = 2, prec = 8 }; };
}  // namespace b200

/* Base of everything kept in the stored_trans1D registry (reference p3dfft.h:529-546). */
class stage {
 public:
  int stage_prec;
  int dt1, dt2;
  int dims1[3], dims2[3];
  stage *next;
  int kind;
  b200::Plan *impl;
  stage() : next(NULL), impl(NULL) {}
  virtual ~stage() {
    if (impl) b200::plan_destroy(impl);
  }
  void run(const void *in, void *out, int deriv_dim, bool OW) { b200::plan_exec(impl, in, out, deriv_dim, OW); }
};

extern vector<gen_trans_type *> types1D;

/* 1D transform along one local dimension, with optional change of storage order
 * (reference transplan, p3dfft.h:581-622, exec.C:522-702). */
template <class Type1, class Type2> class transplan : public stage {
 public:
  bool is_empty;
  int trans_dim;
  int mo1[3], mo2[3];
  bool is_set;
  transplan(const DataGrid &gr1, const DataGrid &gr2, const gen_trans_type *type, int d) { init_tr(gr1, gr2, type, d); }
  transplan(const DataGrid &gr1, const DataGrid &gr2, int type_ID, int d) {
    if (type_ID < 0 || type_ID >= (int)types1D.size() || !types1D[type_ID]->is_set) {
      cout << "Error in trans_plan: 1D transform type no set" << endl;
      is_set = false;
      return;
    }
    init_tr(gr1, gr2, types1D[type_ID], d);
  }
  void init_tr(const DataGrid &gr1, const DataGrid &gr2, const gen_trans_type *type, int d) {
    is_set = false;
    is_empty = type->is_empty;
    trans_dim = d;
    kind = TRANS_ONLY;
    stage_prec = type->prec;
    dt1 = type->dt1;
    dt2 = type->dt2;
    for (int i = 0; i < 3; i++) {
      dims1[i] = gr1.Ldims[i];
      dims2[i] = gr2.Ldims[i];
      mo1[i] = gr1.MemOrder[i];
      mo2[i] = gr2.MemOrder[i];
    }
    if ((int)b200::tinfo<Type1>::dt != type->dt1 || (int)b200::tinfo<Type2>::dt != type->dt2 ||
        (int)b200::tinfo<Type1>::prec != type->prec || (int)b200::tinfo<Type2>::prec != type->prec) {
      cout << "Error in transplan: template types do not match the 1D transform type" << endl;
      return;
    }
    impl = b200::plan1d_create(gr1, gr2, type, d, type->dt1, type->dt2, type->prec);
    is_set = impl && b200::plan_ok(impl);
  }
  void exec(char *in, char *out, bool OW = false) { run(in, out, -1, OW); }
  void exec_deriv(char *in, char *out, bool OW = false) { run(in, out, trans_dim, OW); }
};

class gen_transform3D {
 public:
  int prec;
  int dt1, dt2;
  bool OW;
  b200::Plan *impl;
  gen_transform3D() : impl(NULL) {}
  virtual ~gen_transform3D() {
    if (impl) b200::plan_destroy(impl);
  }
};

/* 3D transform between two distributed layouts (reference transform3D, p3dfft.h:877-895,
 * templ.C:91-420 planner, exec.C:101-295 executor). */
template <class Type1, class Type2> class transform3D : public gen_transform3D {
  bool is_set;

 public:
  transform3D(const DataGrid &grid1_, const DataGrid &grid2_, const trans_type3D *type) {
    prec = b200::tinfo<Type1>::prec;
    dt1 = b200::tinfo<Type1>::dt;
    dt2 = b200::tinfo<Type2>::dt;
    OW = false;
    if ((int)b200::tinfo<Type2>::prec != prec) cout << "Error in transform3D: precisions don't match!" << endl;
    impl = b200::plan3d_create(grid1_, grid2_, type, dt1, dt2, prec);
    is_set = impl && b200::plan_ok(impl);
  }
  void exec(Type1 *in, Type2 *out, bool OW_ = false) { exec_deriv(in, out, -1, OW_); }
  void exec_deriv(Type1 *in, Type2 *out, int idir, bool OW_ = false) {
    if (!is_set) {
      cout << "Error in transform3D::exec: plan is not set" << endl;
      return;
    }
    b200::plan_exec(impl, in, out, idir, OW_);
  }
};

/* stand-alone spectral derivative of an already transformed distributed array (reference deriv.C:85-185) */
template <class Type> void compute_deriv(Type *in, Type *out, DataGrid *gr, int idir);

/* stage-time accumulators with the reference's categories (p3dfft.h:958-989); filled from CUDA events
 * when p3dfft_b200_enable_timers(1) was called */
class timer {
 public:
  double reorder_deriv, reorder_trans, reorder_out, reorder_in, trans_exec, trans_deriv, packsend, packsend_trans,
      packsend_deriv, unpackrecv, alltoall;
  void init();
  void print(MPI_Comm);
};
extern timer timers;

extern vector<gen_transform3D *> stored_trans3D;
extern vector<stage *> stored_trans1D;
extern vector<trans_type3D> types3D;
extern vector<ProcGrid *> stored_proc_grids;
extern vector<DataGrid *> stored_data_grids;

}  // namespace p3dfft

extern "C" {
#endif

#include "Cwrap.h"
#include "Fwrap.h"
#include "p3dfft_b200.h"

#ifdef __cplusplus
}
#endif
#endif
