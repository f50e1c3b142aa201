// cwrap.cpp -- the extern "C" surface: C wrappers (Grid* based), Fortran twins (integer handles, by
// reference, 1-based idir) and the p3dfft_b200_* extensions.  Same symbols and semantics as the
// reference's build/wrap.C:88-794; handles are indices into the process-global registries.
#include <cstdio>
#include <cstring>

#include "plan.h"

using namespace p3dfft;

namespace p3dfft {
namespace b200 {
void set_stream(void *s);
void set_timers(bool on);
}  // namespace b200
}  // namespace p3dfft

namespace {

// a 3D plan created through the C API: type-erased, remembers what the typed C++ class would know
struct CPlan3D : public gen_transform3D {};
struct CPlan1D : public stage {};

DataGrid *grid_from_c(const Grid *g) {
  ProcGrid *pg = stored_proc_grids[g->pgrid];
  return new DataGrid((int *)g->Gdims, g->dim_conj_sym, pg, (int *)g->Dmap, (int *)g->MemOrder);
}

int make_plan3d(const DataGrid &g1, const DataGrid &g2, int tp) {
  const trans_type3D *t3 = &types3D[tp];
  int dt1, dt2;
  // with an R2C the input is real and the output complex; with a C2R the reverse; otherwise both equal
  bool has_r2c = false, has_c2r = false;
  int same = types1D[t3->types[0]]->dt1;
  for (int i = 0; i < 3; i++) {
    const gen_trans_type *t = types1D[t3->types[i]];
    if (t->dt1 < t->dt2) has_r2c = true;
    else if (t->dt1 > t->dt2) has_c2r = true;
    else same = t->dt1;
  }
  dt1 = has_r2c ? 1 : (has_c2r ? 2 : same);
  dt2 = has_r2c ? 2 : (has_c2r ? 1 : same);
  CPlan3D *p = new CPlan3D();
  p->prec = t3->prec;
  p->dt1 = dt1;
  p->dt2 = dt2;
  p->OW = false;
  p->impl = b200::plan3d_create(g1, g2, t3, dt1, dt2, t3->prec);
  stored_trans3D.push_back(p);
  return (int)stored_trans3D.size() - 1;
}

int make_plan1d(const DataGrid &g1, const DataGrid &g2, int type_ID, int d) {
  const gen_trans_type *t = types1D[type_ID];
  CPlan1D *p = new CPlan1D();
  p->stage_prec = t->prec;
  p->dt1 = t->dt1;
  p->dt2 = t->dt2;
  p->kind = TRANS_ONLY;
  for (int i = 0; i < 3; i++) {
    p->dims1[i] = g1.Ldims[i];
    p->dims2[i] = g2.Ldims[i];
  }
  p->impl = b200::plan1d_create(g1, g2, t, d, t->dt1, t->dt2, t->prec);
  stored_trans1D.push_back(p);
  return (int)stored_trans1D.size() - 1;
}

void exec3d(Plan3D plan, const void *in, void *out, int idir, int OW, int prec, const char *who) {
  if (plan < 0 || plan >= (int)stored_trans3D.size()) {
    printf("Error in %s: invalid plan handle %d\n", who, plan);
    return;
  }
  gen_transform3D *t = stored_trans3D[plan];
  if (t->prec != prec) {
    printf("ERror in %s: expecting %s precision data\n", who, prec == 4 ? "single" : "double");
    MPI_Abort(MPI_COMM_WORLD, 0);
  }
  b200::plan_exec(t->impl, in, out, idir, OW != 0);
}

void exec1d(int plan, const void *in, void *out, int OW, const char *who) {
  if (plan < 0 || plan >= (int)stored_trans1D.size()) {
    printf("Error in %s: invalid plan handle %d\n", who, plan);
    return;
  }
  stored_trans1D[plan]->run(in, out, -1, OW != 0);
}

size_t copy_out(const std::string &s, char *buf, size_t buflen) {
  if (buf && buflen) {
    size_t n = s.size() < buflen - 1 ? s.size() : buflen - 1;
    memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return s.size() + 1;
}

}  // namespace

extern "C" {

void p3dfft_setup() {
  p3dfft::setup();
  P3DFFT_EMPTY_TYPE_SINGLE = EMPTY_TYPE_SINGLE;
  P3DFFT_EMPTY_TYPE_DOUBLE = EMPTY_TYPE_DOUBLE;
  P3DFFT_EMPTY_TYPE_SINGLE_COMPLEX = EMPTY_TYPE_SINGLE_COMPLEX;
  P3DFFT_EMPTY_TYPE_DOUBLE_COMPLEX = EMPTY_TYPE_DOUBLE_COMPLEX;
  P3DFFT_R2CFFT_S = R2CFFT_S;
  P3DFFT_R2CFFT_D = R2CFFT_D;
  P3DFFT_C2RFFT_S = C2RFFT_S;
  P3DFFT_C2RFFT_D = C2RFFT_D;
  P3DFFT_CFFT_FORWARD_S = CFFT_FORWARD_S;
  P3DFFT_CFFT_FORWARD_D = CFFT_FORWARD_D;
  P3DFFT_CFFT_BACKWARD_S = CFFT_BACKWARD_S;
  P3DFFT_CFFT_BACKWARD_D = CFFT_BACKWARD_D;
#define COPY_R2R(K)                       \
  P3DFFT_##K##_REAL_S = K##_REAL_S;       \
  P3DFFT_##K##_REAL_D = K##_REAL_D;       \
  P3DFFT_##K##_COMPLEX_S = K##_COMPLEX_S; \
  P3DFFT_##K##_COMPLEX_D = K##_COMPLEX_D;
  COPY_R2R(DCT1) COPY_R2R(DST1) COPY_R2R(DCT2) COPY_R2R(DST2) COPY_R2R(DCT3) COPY_R2R(DST3) COPY_R2R(DCT4) COPY_R2R(DST4)
#undef COPY_R2R
}

void p3dfft_cleanup() { p3dfft::cleanup(); }

Type3D p3dfft_init_3Dtype(int types[3]) {
  types3D.push_back(trans_type3D(types));
  return (int)types3D.size() - 1;
}

Plan3D p3dfft_plan_3Dtrans(Grid *Cgr1, Grid *Cgr2, Type3D tp) {
  DataGrid *g1 = grid_from_c(Cgr1), *g2 = grid_from_c(Cgr2);
  int id = make_plan3d(*g1, *g2, tp);
  delete g1;
  delete g2;
  return id;
}

int p3dfft_plan_1Dtrans(Grid *Cgr1, Grid *Cgr2, int type_ID, int d) {
  DataGrid *g1 = grid_from_c(Cgr1), *g2 = grid_from_c(Cgr2);
  int id = make_plan1d(*g1, *g2, type_ID, d);
  delete g1;
  delete g2;
  return id;
}

int find_grid(int gdims[3], int pgrid, int *dmap, int mem_order[3]) {
  for (size_t i = 0; i < stored_data_grids.size(); i++) {
    DataGrid *gr = stored_data_grids[i];
    if (!arcmp(gr->Gdims, gdims, 3) && *(gr->Pgrid) == *stored_proc_grids[pgrid] && !arcmp(gr->Dmap, dmap, 3) &&
        !arcmp(gr->MemOrder, mem_order, 3))
      return (int)i;
  }
  return -1;
}

int p3dfft_init_proc_grid(int pdims[3], MPI_Comm comm) {
  stored_proc_grids.push_back(new ProcGrid(pdims, comm));
  return (int)stored_proc_grids.size() - 1;
}

Grid *p3dfft_init_data_grid(int gdims[3], int dim_conj_sym, int pgrid_id, int dmap[3], int mem_order[3]) {
  ProcGrid *pg = stored_proc_grids[pgrid_id];
  DataGrid gr(gdims, dim_conj_sym, pg, dmap, mem_order);
  Grid *c = new Grid;
  c->nd = gr.nd;
  c->dim_conj_sym = dim_conj_sym;
  c->pgrid = pgrid_id;
  c->taskid = pg->taskid;
  c->numtasks = pg->numtasks;
  c->mpi_comm_glob = pg->mpi_comm_glob;
  for (int i = 0; i < 3; i++) {
    c->Gdims[i] = gdims[i];
    c->MemOrder[i] = mem_order[i];
    c->Dmap[i] = dmap[i];
    c->Ldims[i] = gr.Ldims[i];
    c->grid_id[i] = gr.grid_id[i];
    c->GlobStart[i] = gr.GlobStart[i];
    c->ProcDims[i] = pg->ProcDims[i];
  }
  return c;
}

void p3dfft_free_data_grid(Grid *gr) { delete gr; }

// Deviation from the reference (wrap.C:339-344), on purpose: the reference erases the vector slot, which
// silently renumbers every later processor-grid handle.  Here the slot is kept (set to NULL) so other
// handles stay valid.
void p3dfft_free_proc_grid(int pgrid_id) {
  if (pgrid_id < 0 || pgrid_id >= (int)stored_proc_grids.size()) return;
  delete stored_proc_grids[pgrid_id];
  stored_proc_grids[pgrid_id] = NULL;
}

void p3dfft_inv_mo(int mo[3], int imo[3]) { p3dfft::inv_mo(mo, imo); }

// declared by the reference (Cwrap.h:113) but never defined there; provided for completeness:
// dumps the entries of a double array larger than 1e-7, "(i j k) value", to the file `label`
void p3dfft_write_buf(double *buf, char *label, int sz[3], int mo[3]) {
  FILE *fp = fopen(label, "w");
  if (!fp) return;
  double *p = buf;
  for (int k = 0; k < sz[mo[2]]; k++)
    for (int j = 0; j < sz[mo[1]]; j++)
      for (int i = 0; i < sz[mo[0]]; i++) {
        if (*p > 1.e-7 || *p < -1.e-7) fprintf(fp, "(%d %d %d) %lg\n", i, j, k, *p);
        p++;
      }
  fclose(fp);
}

void p3dfft_exec_3Dtrans_single(Plan3D plan, float *in, float *out, int OW) { exec3d(plan, in, out, -1, OW, 4, "p3dfft_exec_3Dtrans_single"); }
void p3dfft_exec_3Dtrans_double(Plan3D plan, double *in, double *out, int OW) { exec3d(plan, in, out, -1, OW, 8, "p3dfft_exec_3Dtrans_double"); }
void p3dfft_exec_3Dderiv_single(Plan3D plan, float *in, float *out, int idir, int OW) { exec3d(plan, in, out, idir, OW, 4, "p3dfft_exec_3Dderiv_single"); }
void p3dfft_exec_3Dderiv_double(Plan3D plan, double *in, double *out, int idir, int OW) { exec3d(plan, in, out, idir, OW, 8, "p3dfft_exec_3Dderiv_double"); }
void p3dfft_exec_1Dtrans_double(int plan, double *in, double *out, int OW) { exec1d(plan, in, out, OW, "p3dfft_exec_1Dtrans_double"); }
void p3dfft_exec_1Dtrans_single(int plan, float *in, float *out, int OW) { exec1d(plan, in, out, OW, "p3dfft_exec_1Dtrans_single"); }

void p3dfft_compute_deriv_single(float *in, float *out, Grid *Cgrid, int idir) {
  DataGrid *g = grid_from_c(Cgrid);
  compute_deriv<mycomplex>((mycomplex *)in, (mycomplex *)out, g, idir);
  delete g;
}
void p3dfft_compute_deriv_double(double *in, double *out, Grid *Cgrid, int idir) {
  DataGrid *g = grid_from_c(Cgrid);
  compute_deriv<complex_double>((complex_double *)in, (complex_double *)out, g, idir);
  delete g;
}

// ------------------------------------------------------------------ Fortran twins (wrap.C:575-791)
void p3dfft_init_3Dtype_f(int *type, int types[3]) { *type = p3dfft_init_3Dtype(types); }

void p3dfft_plan_1Dtrans_f(int *plan, int *Fgr1, int *Fgr2, int *type_ID, int *d) {
  *plan = make_plan1d(*stored_data_grids[*Fgr1], *stored_data_grids[*Fgr2], *type_ID, *d);
}
void p3dfft_plan_3Dtrans_f(Plan3D *plan, int *Fgr1, int *Fgr2, Type3D *tp) {
  *plan = make_plan3d(*stored_data_grids[*Fgr1], *stored_data_grids[*Fgr2], *tp);
}
int p3dfft_init_proc_grid_f(int *pdims, int *mpicomm) { return p3dfft_init_proc_grid(pdims, MPI_Comm_f2c(*mpicomm)); }

void p3dfft_init_data_grid_f(int *mygrid, int *ldims, int *glob_start, int *gdims, int *dim_conj_sym, int *pgrid_id, int *dmap,
                             int *mem_order) {
  int num = find_grid(gdims, *pgrid_id, dmap, mem_order);
  if (num < 0) {
    if (*pgrid_id < 0 || *pgrid_id >= (int)stored_proc_grids.size() || !stored_proc_grids[*pgrid_id]) {
      printf("Error in p3dfft_init_data_grid_f: invalid processor grid %d\n", *pgrid_id);
      *mygrid = -1;
      return;
    }
    stored_data_grids.push_back(new DataGrid(gdims, *dim_conj_sym, stored_proc_grids[*pgrid_id], dmap, mem_order));
    num = (int)stored_data_grids.size() - 1;
  }
  DataGrid *g = stored_data_grids[num];
  memcpy(ldims, g->Ldims, 3 * sizeof(int));
  memcpy(glob_start, g->GlobStart, 3 * sizeof(int));
  *mygrid = num;
}

void p3dfft_exec_3Dtrans_double_f(Plan3D *plan, double *in, double *out, int *OW) { p3dfft_exec_3Dtrans_double(*plan, in, out, *OW); }
void p3dfft_exec_3Dtrans_single_f(Plan3D *plan, float *in, float *out, int *OW) { p3dfft_exec_3Dtrans_single(*plan, in, out, *OW); }
void p3dfft_exec_3Dderiv_double_f(Plan3D *plan, double *in, double *out, int *idir, int *OW) {
  p3dfft_exec_3Dderiv_double(*plan, in, out, *idir - 1, *OW);
}
void p3dfft_exec_3Dderiv_single_f(Plan3D *plan, float *in, float *out, int *idir, int *OW) {
  p3dfft_exec_3Dderiv_single(*plan, in, out, *idir - 1, *OW);
}
void p3dfft_exec_1Dtrans_double_f(int *plan, double *in, double *out, int *OW) { p3dfft_exec_1Dtrans_double(*plan, in, out, *OW); }
void p3dfft_exec_1Dtrans_single_f(int *plan, float *in, float *out, int *OW) { p3dfft_exec_1Dtrans_single(*plan, in, out, *OW); }
void p3dfft_compute_deriv_single_f(float *in, float *out, int *igrid, int *idir) {
  compute_deriv<mycomplex>((mycomplex *)in, (mycomplex *)out, stored_data_grids[*igrid], *idir - 1);
}
void p3dfft_compute_deriv_double_f(double *in, double *out, int *igrid, int *idir) {
  compute_deriv<complex_double>((complex_double *)in, (complex_double *)out, stored_data_grids[*igrid], *idir - 1);
}

// ------------------------------------------------------------------ extensions
const char *p3dfft_b200_version(void) { return "p3dfft.3_b200 0.1 (sm_100a)"; }
void p3dfft_b200_set_stream(void *s) { b200::set_stream(s); }
void p3dfft_b200_sync(void) {
  if (b200::gpu_ready() && p3dfftcu_stream_sync(b200::current_stream())) {
    fprintf(stderr, "p3dfft_b200 fatal: %s\n", p3dfftcu_last_error());
    MPI_Abort(MPI_COMM_WORLD, 1);
  }
}
long long p3dfft_b200_kernel_launches(void) { return p3dfftcu_launch_count(); }
size_t p3dfft_b200_describe_plan3d(int plan, char *buf, size_t buflen) {
  if (plan < 0 || plan >= (int)stored_trans3D.size() || !stored_trans3D[plan]->impl) return copy_out("{\"ok\":false}", buf, buflen);
  return copy_out(b200::describe(*stored_trans3D[plan]->impl), buf, buflen);
}
size_t p3dfft_b200_describe_plan1d(int plan, char *buf, size_t buflen) {
  if (plan < 0 || plan >= (int)stored_trans1D.size() || !stored_trans1D[plan]->impl) return copy_out("{\"ok\":false}", buf, buflen);
  return copy_out(b200::describe(*stored_trans1D[plan]->impl), buf, buflen);
}
void p3dfft_b200_enable_timers(int on) { b200::set_timers(on != 0); }
int p3dfft_b200_stage_times(int plan, float *ms, int max_stages) {
  if (plan < 0 || plan >= (int)stored_trans3D.size() || !stored_trans3D[plan]->impl) return 0;
  b200::plan_collect_times(stored_trans3D[plan]->impl);
  const std::vector<float> &t = stored_trans3D[plan]->impl->stage_ms;
  int n = (int)t.size() < max_stages ? (int)t.size() : max_stages;
  for (int i = 0; i < n; i++) ms[i] = t[i];
  return (int)t.size();
}
int p3dfft_b200_have_device(void) { return b200::gpu_ready() ? 1 : 0; }
void p3dfft_b200_set_host_staging(const char *mode) { b200::set_host_staging(mode); }
void p3dfft_b200_host_release(const void *ptr) {
  if (!b200::gpu_ready()) return;
  p3dfftcu_stream_sync(b200::current_stream());
  p3dfftcu_host_unpin(ptr);
}

}  // extern "C"
