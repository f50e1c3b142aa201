// registry.cpp -- process-global registries and the 44-entry table of 1D transform types.
// Mirrors the reference's setup()/cleanup() (build/init.C:121-753, 827-1101) and the handle vectors of
// build/init.C:99-105: IDs are positions in types1D, in this fixed order:
//   0-3 EMPTY_{S,D,SC,DC}; 4-7 R2C_S R2C_D C2R_S C2R_D; 8-11 CFFT_FWD_S/D CFFT_BWD_S/D;
//   then for k = 1..4: DCTk {REAL_S, REAL_D, COMPLEX_S, COMPLEX_D}, DSTk {same}.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "plan.h"

int P3DFFT_EMPTY_TYPE_SINGLE, P3DFFT_EMPTY_TYPE_DOUBLE, P3DFFT_EMPTY_TYPE_SINGLE_COMPLEX, P3DFFT_EMPTY_TYPE_DOUBLE_COMPLEX;
int P3DFFT_R2CFFT_S, P3DFFT_R2CFFT_D, P3DFFT_C2RFFT_S, P3DFFT_C2RFFT_D;
int P3DFFT_CFFT_FORWARD_S, P3DFFT_CFFT_FORWARD_D, P3DFFT_CFFT_BACKWARD_S, P3DFFT_CFFT_BACKWARD_D;
#define DEF_R2R(K) int P3DFFT_##K##_REAL_S, P3DFFT_##K##_REAL_D, P3DFFT_##K##_COMPLEX_S, P3DFFT_##K##_COMPLEX_D;
DEF_R2R(DCT1) DEF_R2R(DST1) DEF_R2R(DCT2) DEF_R2R(DST2) DEF_R2R(DCT3) DEF_R2R(DST3) DEF_R2R(DCT4) DEF_R2R(DST4)
#undef DEF_R2R

namespace p3dfft {

int EMPTY_TYPE_SINGLE, EMPTY_TYPE_DOUBLE, EMPTY_TYPE_SINGLE_COMPLEX, EMPTY_TYPE_DOUBLE_COMPLEX;
int R2CFFT_S, R2CFFT_D, C2RFFT_S, C2RFFT_D, CFFT_FORWARD_S, CFFT_FORWARD_D, CFFT_BACKWARD_S, CFFT_BACKWARD_D;
#define DEF_R2R(K) int K##_REAL_S, K##_REAL_D, K##_COMPLEX_S, K##_COMPLEX_D;
DEF_R2R(DCT1) DEF_R2R(DST1) DEF_R2R(DCT2) DEF_R2R(DST2) DEF_R2R(DCT3) DEF_R2R(DST3) DEF_R2R(DCT4) DEF_R2R(DST4)
#undef DEF_R2R

vector<gen_trans_type *> types1D;
vector<gen_transform3D *> stored_trans3D;
vector<stage *> stored_trans1D;
vector<trans_type3D> types3D;
vector<ProcGrid *> stored_proc_grids;
vector<DataGrid *> stored_data_grids;
timer timers;

namespace b200 {
static bool g_gpu_ready = false;
static void *g_stream = nullptr;
static bool g_timers_on = false;
bool gpu_ready() { return g_gpu_ready; }
void *current_stream() { return g_stream; }
void set_stream(void *s) { g_stream = s; }
bool timers_on() { return g_timers_on; }
void set_timers(bool on) { g_timers_on = on; }
}  // namespace b200

static int reg(const char *name, int kind, int dt1, int dt2, int prec, int isign = 0, bool empty = false) {
  types1D.push_back(new gen_trans_type(name, kind, dt1, dt2, prec, isign, empty));
  return (int)types1D.size() - 1;
}

void setup() {
  if (!types1D.empty()) return;  // idempotent
  int flag = 0;
  MPI_Initialized(&flag);
  if (!flag) {
    int argc = 0;
    MPI_Init(&argc, NULL);
  }
  EMPTY_TYPE_SINGLE = reg("Empty Type Single", P3DFFTCU_K_EMPTY, REAL, REAL, 4, 0, true);
  EMPTY_TYPE_DOUBLE = reg("Empty Type Double", P3DFFTCU_K_EMPTY, REAL, REAL, 8, 0, true);
  EMPTY_TYPE_SINGLE_COMPLEX = reg("Empty Type Single Complex", P3DFFTCU_K_EMPTY, COMPLEX, COMPLEX, 4, 0, true);
  EMPTY_TYPE_DOUBLE_COMPLEX = reg("Empty Type Double Complex", P3DFFTCU_K_EMPTY, COMPLEX, COMPLEX, 8, 0, true);
  R2CFFT_S = reg("Real-to-complex Fourier Transform, single precision", P3DFFTCU_K_R2C, REAL, COMPLEX, 4);
  R2CFFT_D = reg("Real-to-complex Fourier Transform, double precision", P3DFFTCU_K_R2C, REAL, COMPLEX, 8);
  C2RFFT_S = reg("Complex-to-real Fourier Transform, single precision", P3DFFTCU_K_C2R, COMPLEX, REAL, 4);
  C2RFFT_D = reg("Complex-to-real Fourier Transform, double precision", P3DFFTCU_K_C2R, COMPLEX, REAL, 8);
  CFFT_FORWARD_S = reg("Complex forward Fourier Transform, single precision", P3DFFTCU_K_C2C_FWD, COMPLEX, COMPLEX, 4, -1);
  CFFT_FORWARD_D = reg("Complex forward Fourier Transform, double precision", P3DFFTCU_K_C2C_FWD, COMPLEX, COMPLEX, 8, -1);
  CFFT_BACKWARD_S = reg("Complex backward Fourier Transform, single precision", P3DFFTCU_K_C2C_BWD, COMPLEX, COMPLEX, 4, 1);
  CFFT_BACKWARD_D = reg("Complex backward Fourier Transform, double precision", P3DFFTCU_K_C2C_BWD, COMPLEX, COMPLEX, 8, 1);
#define REG_R2R(K, KIND, LABEL)                                                           \
  K##_REAL_S = reg(LABEL ", real, single precision", KIND, REAL, REAL, 4);                \
  K##_REAL_D = reg(LABEL ", real, double precision", KIND, REAL, REAL, 8);                \
  K##_COMPLEX_S = reg(LABEL ", complex, single precision", KIND, COMPLEX, COMPLEX, 4);    \
  K##_COMPLEX_D = reg(LABEL ", complex, double precision", KIND, COMPLEX, COMPLEX, 8);
  REG_R2R(DCT1, P3DFFTCU_K_DCT1, "Cosine transform DCT-I")
  REG_R2R(DST1, P3DFFTCU_K_DST1, "Sine transform DST-I")
  REG_R2R(DCT2, P3DFFTCU_K_DCT2, "Cosine transform DCT-II")
  REG_R2R(DST2, P3DFFTCU_K_DST2, "Sine transform DST-II")
  REG_R2R(DCT3, P3DFFTCU_K_DCT3, "Cosine transform DCT-III")
  REG_R2R(DST3, P3DFFTCU_K_DST3, "Sine transform DST-III")
  // The reference registers its four DCT4 IDs with the DCT-I planner (build/init.C:640,652,664,676), so a
  // program asking for "DCT4" there gets FFTW_REDFT00.  Reproduced by default so results stay identical;
  // P3DFFT_B200_TRUE_DCT4=1 selects the real DCT-IV (FFTW_REDFT11) kernel instead.
  {
    const char *e = getenv("P3DFFT_B200_TRUE_DCT4");
    const int k4 = (e && atoi(e)) ? P3DFFTCU_K_DCT4 : P3DFFTCU_K_DCT1;
    REG_R2R(DCT4, k4, "Cosine transform DCT-IV")
  }
  REG_R2R(DST4, P3DFFTCU_K_DST4, "Sine transform DST-IV")
#undef REG_R2R

  // bind the device.  A host without a GPU can still build and inspect plans; exec will abort.
  b200::g_gpu_ready = false;
  const char *plan_only = getenv("P3DFFT_B200_PLAN_ONLY");  // build and describe plans without touching a device
  if (!(plan_only && atoi(plan_only)) && p3dfftcu_device_count() > 0) {
    if (p3dfftcu_init(-1) == 0) b200::g_gpu_ready = true;
    else fprintf(stderr, "p3dfft_b200: %s\n", p3dfftcu_last_error());
  }
}

void cleanup() {
  for (size_t i = 0; i < stored_trans3D.size(); i++) delete stored_trans3D[i];
  stored_trans3D.clear();
  for (size_t i = 0; i < stored_trans1D.size(); i++) delete stored_trans1D[i];
  stored_trans1D.clear();
  types3D.clear();
  for (size_t i = 0; i < stored_data_grids.size(); i++) delete stored_data_grids[i];
  stored_data_grids.clear();
  for (size_t i = 0; i < stored_proc_grids.size(); i++) delete stored_proc_grids[i];
  stored_proc_grids.clear();
  for (size_t i = 0; i < types1D.size(); i++) delete types1D[i];
  types1D.clear();
  b200::workspace_release();
}

void trans_type3D::init(const int ids[3]) {
  name = new char[1];
  name[0] = 0;
  is_set = false;
  prec = 0;
  for (int i = 0; i < 3; i++) {
    types[i] = ids[i];
    if (ids[i] < 0 || ids[i] >= (int)types1D.size()) {
      printf("Error in trans_type3D: invalid 1D transform type ID %d (was p3dfft setup() called?)\n", ids[i]);
      return;
    }
    int p = types1D[ids[i]]->prec;
    if (i == 0) prec = p;
    else if (p != prec) {
      printf("Error in trans_type3D: precisions of types don't match\n");
      return;
    }
  }
  is_set = true;
}

trans_type3D::trans_type3D(int types_IDs[3]) { init(types_IDs); }

trans_type3D::trans_type3D(gen_trans_type *types_[3]) {
  int ids[3] = {-1, -1, -1};
  for (int i = 0; i < 3; i++)
    for (size_t j = 0; j < types1D.size(); j++)
      if (types1D[j] == types_[i] || *types1D[j] == *types_[i]) {
        ids[i] = (int)j;
        break;
      }
  init(ids);
}

trans_type3D::trans_type3D(const trans_type3D &rhs) {
  name = new char[strlen(rhs.name) + 1];
  strcpy(name, rhs.name);
  prec = rhs.prec;
  is_set = rhs.is_set;
  for (int i = 0; i < 3; i++) types[i] = rhs.types[i];
}

trans_type3D::~trans_type3D() { delete[] name; }

void timer::init() {
  reorder_deriv = reorder_trans = reorder_out = reorder_in = trans_exec = trans_deriv = 0;
  packsend = packsend_trans = packsend_deriv = unpackrecv = alltoall = 0;
}

void timer::print(MPI_Comm comm) {
  for (size_t i = 0; i < stored_trans3D.size(); i++)
    if (stored_trans3D[i]->impl) b200::plan_collect_times(stored_trans3D[i]->impl);
  int rank, n;
  MPI_Comm_rank(comm, &rank);
  MPI_Comm_size(comm, &n);
  const char *names[11] = {"Reorder_deriv", "Reorder_trans", "Reorder_out", "Reorder_in", "Trans_exec", "Trans_deriv",
                           "Packsend", "Packsend_trans", "Packsend_deriv", "Unpackrecv", "Alltoall"};
  double vals[11] = {reorder_deriv, reorder_trans, reorder_out, reorder_in, trans_exec, trans_deriv,
                     packsend, packsend_trans, packsend_deriv, unpackrecv, alltoall};
  double sum[11], mn[11], mx[11];
  MPI_Reduce(vals, sum, 11, MPI_DOUBLE, MPI_SUM, 0, comm);
  MPI_Reduce(vals, mn, 11, MPI_DOUBLE, MPI_MIN, 0, comm);
  MPI_Reduce(vals, mx, 11, MPI_DOUBLE, MPI_MAX, 0, comm);
  if (rank == 0) {
    printf("TIMERS (avg/min/max), seconds of GPU stage time:\n");
    for (int i = 0; i < 11; i++) printf("%-16s %lg %lg %lg\n", names[i], sum[i] / n, mn[i], mx[i]);
  }
}

}  // namespace p3dfft
